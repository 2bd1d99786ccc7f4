"""Thin host wrappers over the C-ABI: torch tensors are the device-memory containers,
every function enqueues on torch's current CUDA stream and returns without syncing.

Nothing here computes: no torch math on the data path, no CPU fallback. A missing
`libdpl_b200.so` raises from `_lib.lib()`.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (BLOB_FIELDS, F_NSEG, F_PTR, F_SEGLEN, F_STAT, check, lib, plan_blobs)

_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launches():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of {dtype}, got "
                         f"{t.dtype} on {t.device} (contiguous={t.is_contiguous()})")


class BlobBatch:
    """The dpl_blob table of one batch: a list of contiguous float32 CUDA tensors whose
    leading dimension is the image index. Keeps the tensors alive until released."""

    def __init__(self, tensors, stat_index=None):
        self.tensors = list(tensors)
        n = len(self.tensors)
        table = np.zeros((n, BLOB_FIELDS), dtype=np.uint64)
        max_seg = 0
        for i, t in enumerate(self.tensors):
            _need(t, torch.float32, f"blob {i}")
            n_seg = t.shape[0] if t.dim() > 0 else 1
            seg_len = t.numel() // n_seg if n_seg else 0
            table[i, F_PTR] = t.data_ptr()
            table[i, F_NSEG] = n_seg
            table[i, F_SEGLEN] = seg_len
            table[i, F_STAT] = i if stat_index is None else stat_index[i]
            max_seg = max(max_seg, seg_len)
        self.n_segments, self.n_seg_tiles, self.n_flat_tiles = plan_blobs(table)
        self.host_table = table
        self.max_seg_len = max_seg
        self.n_blobs = n
        self.device = self.tensors[0].device if n else torch.device("cuda")
        # small (8 KB for ResNet-50) upload; through pinned memory it does not stall the host behind
        # the kernels already queued on the stream
        host = torch.from_numpy(table.view(np.int64))
        if self.device.type == "cuda":
            self._pinned = host.pin_memory()
            self.table = self._pinned.to(self.device, non_blocking=True)
        else:
            self.table = host.to(self.device)
        self.elements = int(sum(int(t.numel()) for t in self.tensors))

    def seg_slices(self):
        """(offset, n_seg) of every blob in the per-segment output arrays."""
        base = self.host_table[:, _lib.F_SEG_OUT_BASE].astype(np.int64)
        nseg = self.host_table[:, F_NSEG].astype(np.int64)
        return list(zip(base.tolist(), nseg.tolist()))


class Workspace:
    """Grow-only scratch buffer (uint8) reused across launches."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self.buf


def segstats(batch, seg_min, seg_max, seg_abssum=None, seg_nnz=None, blob_min=None, blob_max=None,
             workspace=None, ctas_per_sm=0):
    """K1. seg_*: per-segment outputs (float32/float32/float64/int64[n_segments]);
    blob_min/blob_max: running per-blob extrema (float32[n_stats]), updated in place."""
    _need(seg_min, torch.float32, "seg_min")
    _need(seg_max, torch.float32, "seg_max")
    _need(seg_abssum, torch.float64, "seg_abssum")
    _need(seg_nnz, torch.int64, "seg_nnz")
    _need(blob_min, torch.float32, "blob_min")
    _need(blob_max, torch.float32, "blob_max")
    assert seg_min.numel() >= batch.n_segments and seg_max.numel() >= batch.n_segments
    need = lib().dpl_segstats_scratch_bytes(batch.n_seg_tiles)
    ws = (workspace or Workspace(batch.device)).get(need)
    check(lib().dpl_segstats_f32(batch.table.data_ptr(), batch.n_blobs, batch.n_segments,
                                 batch.n_seg_tiles, seg_min.data_ptr(), seg_max.data_ptr(),
                                 _lib._ptr(seg_abssum), _lib._ptr(seg_nnz), _lib._ptr(blob_min),
                                 _lib._ptr(blob_max), ws.data_ptr(), ws.numel(), int(ctas_per_sm), _stream()),
          "dpl_segstats_f32")
    _count(2)


def absmax(blob_min, blob_max, data_max):
    _need(blob_min, torch.float32, "blob_min")
    _need(blob_max, torch.float32, "blob_max")
    _need(data_max, torch.float32, "data_max")
    check(lib().dpl_absmax_f32(blob_min.data_ptr(), blob_max.data_ptr(), data_max.data_ptr(),
                               data_max.numel(), _stream()), "dpl_absmax_f32")
    _count()


def hist_abs(batch, data_max, counts, bins, variant=0):
    """K2. counts: int64[n_stats, bins] accumulated in place (uint64 bit pattern)."""
    _need(data_max, torch.float32, "data_max")
    _need(counts, torch.int64, "counts")
    assert counts.numel() % bins == 0
    check(lib().dpl_hist_abs_f32(batch.table.data_ptr(), batch.n_blobs, batch.n_flat_tiles,
                                 data_max.data_ptr(), int(bins), counts.data_ptr(), int(variant),
                                 _stream()), "dpl_hist_abs_f32")
    _count()


def hist_percentile(counts, bins, threshold, data_max, blob_min, blob_max, clip, sel_bin=None):
    """K3. clip: float32[n_stats, 2]; sel_bin: int32[n_stats] or None."""
    _need(counts, torch.int64, "counts")
    _need(clip, torch.float32, "clip")
    _need(sel_bin, torch.int32, "sel_bin")
    n_stats = counts.numel() // bins
    check(lib().dpl_hist_percentile(counts.data_ptr(), n_stats, int(bins), float(threshold),
                                    data_max.data_ptr(), blob_min.data_ptr(), blob_max.data_ptr(),
                                    clip.data_ptr(), _lib._ptr(sel_bin), _stream()),
          "dpl_hist_percentile")
    _count()


def octav(batch, seg_abssum, seg_nnz, k_const, out_s, out_iters=None, max_iter=20, workspace=None):
    """K4. out_s: float32[n_segments]."""
    _need(seg_abssum, torch.float64, "seg_abssum")
    _need(seg_nnz, torch.int64, "seg_nnz")
    _need(out_s, torch.float32, "out_s")
    _need(out_iters, torch.int32, "out_iters")
    need = lib().dpl_octav_scratch_bytes(batch.max_seg_len, batch.n_segments)
    ws = (workspace or Workspace(batch.device)).get(need)
    check(lib().dpl_octav_f32(batch.table.data_ptr(), batch.n_blobs, batch.n_segments,
                              batch.max_seg_len, seg_abssum.data_ptr(), seg_nnz.data_ptr(),
                              float(k_const), int(max_iter), out_s.data_ptr(),
                              _lib._ptr(out_iters), ws.data_ptr(), ws.numel(), _stream()),
          "dpl_octav_f32")
    _count()


def fakequant(x, scale, zero_point=None, qlo=-128, qhi=127, axis=None, drop_prob=1.0, seed=0,
              out=None):
    """K5. scale: float32[C] (C = 1 per-tensor); axis: channel axis for per-channel."""
    _need(x, torch.float32, "x")
    _need(scale, torch.float32, "scale")
    _need(zero_point, torch.int32, "zero_point")
    y = torch.empty_like(x) if out is None else out
    _need(y, torch.float32, "out")
    c = scale.numel()
    if c == 1:
        inner = 1
    else:
        assert axis is not None and x.shape[axis] == c
        inner = 1
        for d in x.shape[axis + 1:]:
            inner *= d
    check(lib().dpl_fakequant_f32(x.data_ptr(), y.data_ptr(), x.numel(), scale.data_ptr(),
                                  _lib._ptr(zero_point), c, inner, int(qlo), int(qhi),
                                  float(drop_prob), int(seed) & (2 ** 64 - 1), _stream()),
          "dpl_fakequant_f32")
    _count()
    return y


def channel_sumdiff(a, b, channels, acc):
    """K7a. a, b: [n_img, channels, ...]; acc: float64[channels] accumulated in place."""
    _need(a, torch.float32, "a")
    _need(b, torch.float32, "b")
    _need(acc, torch.float64, "acc")
    assert a.shape == b.shape and a.shape[1] == channels
    n_img = a.shape[0]
    inner = a.numel() // (n_img * channels) if a.numel() else 1
    check(lib().dpl_channel_sumdiff_f32(a.data_ptr(), b.data_ptr(), n_img, channels, inner,
                                        acc.data_ptr(), _stream()), "dpl_channel_sumdiff_f32")
    _count()


def cosine3(a, b, out):
    """K7b. a, b: [n_seg, ...]; out: float64[n_seg, 3] (zeroed by the caller)."""
    _need(a, torch.float32, "a")
    _need(b, torch.float32, "b")
    _need(out, torch.float64, "out")
    n_seg = a.shape[0]
    check(lib().dpl_cosine3_f32(a.data_ptr(), b.data_ptr(), n_seg, a.numel() // max(n_seg, 1),
                                out.data_ptr(), _stream()), "dpl_cosine3_f32")
    _count()


def adaround_init(w, scale):
    """alpha0 and floor(w/s) for a weight [C_out, ...] with per-channel scale [C_out]
    (or a single scale)."""
    _need(w, torch.float32, "w")
    _need(scale, torch.float32, "scale")
    c = scale.numel()
    inner = w.numel() // c
    alpha = torch.empty_like(w)
    wfloor = torch.empty_like(w)
    check(lib().dpl_adaround_init_f32(w.data_ptr(), scale.data_ptr(), c, inner, alpha.data_ptr(),
                                      wfloor.data_ptr(), _stream()), "dpl_adaround_init_f32")
    _count()
    return alpha, wfloor


def adaround_weight(wfloor, alpha, scale, qmin, qmax, soft, out=None):
    c = scale.numel()
    inner = wfloor.numel() // c
    wq = torch.empty_like(wfloor) if out is None else out
    check(lib().dpl_adaround_weight_f32(wfloor.data_ptr(), alpha.data_ptr(), scale.data_ptr(), c,
                                        inner, float(qmin), float(qmax), 1 if soft else 0,
                                        wq.data_ptr(), _stream()), "dpl_adaround_weight_f32")
    _count()
    return wq


def adaround_step(grad_w, wfloor, scale, qmin, qmax, beta, alpha, m, v, step, reg_alpha=0.01,
                  lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, reg_out=None, sched=None):
    c = scale.numel()
    inner = wfloor.numel() // c
    check(lib().dpl_adaround_step_f32(grad_w.data_ptr(), wfloor.data_ptr(), scale.data_ptr(), c,
                                      inner, float(qmin), float(qmax), float(beta),
                                      float(reg_alpha), float(lr), float(b1), float(b2), float(eps),
                                      int(step), float(grad_scale), alpha.data_ptr(), m.data_ptr(),
                                      v.data_ptr(), _lib._ptr(reg_out), _lib._ptr(sched), _stream()),
          "dpl_adaround_step_f32")
    _count()


def adaround_step_peer(peer_grads, peer_words, rank, epoch, wfloor, scale, qmin, qmax, beta, alpha, m, v, step,
                       reg_alpha=0.01, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, reg_out=None, sched=None, error=None):
    """K6 step with the all-reduce of dL/dW inside. peer_grads / peer_words: lists of
    `world` device addresses (ints) in rank order; see dpl_adaround_step_peer_f32."""
    world = len(peer_grads)
    assert len(peer_words) == world and 0 <= rank < world
    c = scale.numel()
    inner = wfloor.numel() // c
    grads = (ctypes.c_void_p * world)(*[int(p) for p in peer_grads])
    words = (ctypes.c_void_p * world)(*[int(p) for p in peer_words])
    check(lib().dpl_adaround_step_peer_f32(grads, words, world, int(rank), int(epoch), wfloor.data_ptr(),
                                           scale.data_ptr(), c, inner, float(qmin), float(qmax), float(beta),
                                           float(reg_alpha), float(lr), float(b1), float(b2), float(eps),
                                           int(step), alpha.data_ptr(), m.data_ptr(), v.data_ptr(),
                                           _lib._ptr(reg_out), _lib._ptr(sched), _lib._ptr(error), _stream()),
          "dpl_adaround_step_peer_f32")
    _count()


def recon_act(o, relu, quant=None, prob=1.0, seed=0, out=None, seed_dev=None):
    """K6 epilogue forward. quant: None or (scale, qmin, qmax) per-tensor."""
    y = torch.empty_like(o) if out is None else out
    s, lo, hi = quant if quant else (1.0, 0.0, 0.0)
    check(lib().dpl_recon_act_f32(o.data_ptr(), y.data_ptr(), o.numel(), int(bool(relu)),
                                  int(quant is not None), float(s), float(lo), float(hi), float(prob),
                                  int(seed) & (2 ** 64 - 1), _lib._ptr(seed_dev), _stream()), "dpl_recon_act_f32")
    _count()
    return y


def recon_act_bwd(o, gy, relu, quant=None, prob=1.0, seed=0, out=None, seed_dev=None):
    go = torch.empty_like(o) if out is None else out
    s, lo, hi = quant if quant else (1.0, 0.0, 0.0)
    check(lib().dpl_recon_act_bwd_f32(o.data_ptr(), gy.data_ptr(), go.data_ptr(), o.numel(),
                                      int(bool(relu)), int(quant is not None), float(s), float(lo),
                                      float(hi), float(prob), int(seed) & (2 ** 64 - 1), _lib._ptr(seed_dev),
                                      _stream()),
          "dpl_recon_act_bwd_f32")
    _count()
    return go


def recon_loss(o, tgt, inv_count, loss_acc, relu, quant=None, prob=1.0, seed=0, out=None, seed_dev=None):
    """Accumulates the L2 loss into loss_acc (float64[1]) and returns dL/do."""
    go = torch.empty_like(o) if out is None else out
    s, lo, hi = quant if quant else (1.0, 0.0, 0.0)
    check(lib().dpl_recon_loss_f32(o.data_ptr(), tgt.data_ptr(), go.data_ptr(), o.numel(),
                                   int(bool(relu)), int(quant is not None), float(s), float(lo),
                                   float(hi), float(prob), int(seed) & (2 ** 64 - 1),
                                   float(inv_count), _lib._ptr(loss_acc), _lib._ptr(seed_dev), _stream()),
          "dpl_recon_loss_f32")
    _count()
    return go


def mix_drop(a, b, prob, seed, out=None):
    y = torch.empty_like(a) if out is None else out
    check(lib().dpl_mix_drop_f32(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), float(prob),
                                 int(seed) & (2 ** 64 - 1), _stream()), "dpl_mix_drop_f32")
    _count()
    return y


# ---- forward-engine operators (dpl_eltwise.cu) -------------------------------------------
def _rng(rng):
    """(ptr to this blob's running min, ptr to its running max) or (0, 0): `rng` = (blob_min, blob_max,
    index) with float32 CUDA arrays."""
    if rng is None:
        return 0, 0
    bmin, bmax, idx = rng
    return bmin.data_ptr() + 4 * int(idx), bmax.data_ptr() + 4 * int(idx)


def clip(x, lo, hi, out=None, rng=None):
    """y = min(max(x, lo), hi); Relu is clip(x, 0, inf). `rng`: fused range statistics of y."""
    _need(x, torch.float32, "x")
    y = torch.empty_like(x) if out is None else out
    _need(y, torch.float32, "out")
    check(lib().dpl_clip_f32(x.data_ptr(), y.data_ptr(), x.numel(), float(lo), float(hi), *_rng(rng), _stream()),
          "dpl_clip_f32")
    _count()
    return y


def add(a, b, out=None, out_relu=None, rng=None, rng_relu=None):
    """y = a + b (same shape); with `out_relu` also max(y, 0) from the same pass."""
    _need(a, torch.float32, "a")
    _need(b, torch.float32, "b")
    assert a.shape == b.shape
    y = torch.empty_like(a) if out is None else out
    _need(y, torch.float32, "out")
    _need(out_relu, torch.float32, "out_relu")
    check(lib().dpl_add_f32(a.data_ptr(), b.data_ptr(), y.data_ptr(), _lib._ptr(out_relu), a.numel(),
                            *_rng(rng), *_rng(rng_relu), _stream()), "dpl_add_f32")
    _count()
    return y


def maxpool2d(x, kernel, stride, pad_top, pad_left, ho, wo, out=None, rng=None):
    _need(x, torch.float32, "x")
    n, c, h, w = x.shape
    y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device) if out is None else out
    _need(y, torch.float32, "out")
    check(lib().dpl_maxpool2d_f32(x.data_ptr(), y.data_ptr(), n * c, h, w, int(kernel[0]), int(kernel[1]),
                                  int(stride[0]), int(stride[1]), int(pad_top), int(pad_left), int(ho), int(wo),
                                  *_rng(rng), _stream()), "dpl_maxpool2d_f32")
    _count()
    return y


def global_avgpool(x, out=None, rng=None):
    _need(x, torch.float32, "x")
    n, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(n * c, 1)
    y = torch.empty((n, c) + (1,) * (x.dim() - 2), dtype=torch.float32, device=x.device) if out is None else out
    _need(y, torch.float32, "out")
    check(lib().dpl_global_avgpool_f32(x.data_ptr(), y.data_ptr(), n * c, hw, *_rng(rng), _stream()),
          "dpl_global_avgpool_f32")
    _count()
    return y


class BlobArena:
    """One device slab, bump-allocated: every blob of a calibration forward lives in it.
    A resident `hist` job holds ~110 GB of blobs; taking them from torch's caching allocator
    costs one cudaMalloc per blob per batch on a cold process (measured: 21 s of a 22 s first
    call), the slab costs one."""

    ALIGN = 256

    def __init__(self, nbytes, device):
        self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.off = 0

    def reset(self):
        self.off = 0

    def alloc(self, shape):
        n = 4
        for d in shape:
            n *= int(d)
        end = self.off + n
        if end > self.buf.numel():
            raise MemoryError("BlobArena exhausted (%d of %d bytes)" % (end, self.buf.numel()))
        t = self.buf[self.off:end].view(torch.float32).view(tuple(int(d) for d in shape))
        self.off = (end + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        return t


class GemmUnsupported(RuntimeError):
    """The operands do not meet TMA's alignment rules; the caller keeps its other path."""


_gemm_err = {}


def gemm_tf32(a, a_major, lda, a_batch_stride, b, b_major, ldb, b_batch_stride, d, ldd, d_batch_stride,
              M, N, K, batch=1, fold_batch=False, split_k=1, bias=None, bias_mode=0, relu=False):
    """K6 dense tile (tcgen05 / TMA / TMEM). Raw-layout interface, see include/dpl_b200.h."""
    dev = d.device
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    st = lib().dpl_gemm_tf32(a.data_ptr(), int(a_major), int(lda), int(a_batch_stride), b.data_ptr(),
                             int(b_major), int(ldb), int(b_batch_stride), d.data_ptr(), int(ldd),
                             int(d_batch_stride), int(M), int(N), int(K), int(batch), int(bool(fold_batch)),
                             int(split_k), _lib._ptr(bias), int(bias_mode), int(bool(relu)),
                             flag.data_ptr(), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_gemm_tf32")
    _count()
    return d


def gemm_check_errors(device=None):
    """Synchronising check of the in-kernel timeout flag (tests / debugging)."""
    for dev, flag in _gemm_err.items():
        if device is None or dev == device:
            if int(flag.item()) != 0:
                flag.zero_()
                raise RuntimeError("dpl_gemm_tf32: a pipeline wait timed out inside the kernel")


def conv1x1_forward(x, w, bias=None, relu=False, out=None):
    """O[img][co][hw] = W[co][ci] x X[img][ci][hw] (+ bias[co], relu), NCHW, stride 1."""
    n, ci, hh, ww = x.shape
    co = w.shape[0]
    hw = hh * ww
    o = torch.empty((n, co, hh, ww), dtype=torch.float32, device=x.device) if out is None else out
    return gemm_tf32(w, 0, ci, 0, x, 1, hw, ci * hw, o, hw, co * hw, co, hw, ci, batch=n,
                     bias=bias, bias_mode=1 if bias is not None else 0, relu=relu)


def conv1x1_wgrad(go, x, split_k=None, out=None):
    """dW[co][ci] = sum_img dO[img][co][hw] x X[img][ci][hw]^T."""
    n, co, hh, ww = go.shape
    ci = x.shape[1]
    hw = hh * ww
    tiles = ((co + 127) // 128) * ((ci + 127) // 128)
    if split_k is None:
        split_k = max(1, min(n, 148 // max(tiles, 1)))
    dw = torch.zeros((co, ci), dtype=torch.float32, device=x.device) if out is None else out.zero_()
    return gemm_tf32(go, 0, hw, co * hw, x, 0, hw, ci * hw, dw, ci, 0, co, ci, hw, batch=n,
                     fold_batch=True, split_k=split_k)


def conv1x1_dgrad(go, w, out=None):
    """dX[img][ci][hw] = W[co][ci]^T x dO[img][co][hw]."""
    n, co, hh, ww = go.shape
    ci = w.shape[1]
    hw = hh * ww
    dx = torch.empty((n, ci, hh, ww), dtype=torch.float32, device=go.device) if out is None else out
    # W^T as a K-major operand (one tiny transposition): the K-major x MN-major form takes the persistent kernel
    _, wt = taps_layout(w.reshape(co, ci, 1, 1), forward=False, dgrad=True)
    return gemm_tf32(wt, 0, co, 0, go, 1, hw, co * hw, dx, hw, ci * hw, ci, hw, co, batch=n)


def linear_forward(x, w, bias=None, relu=False):
    """Y[n][out] = X[n][k] x W[out][k]^T (+ bias[out])."""
    nrow, k = x.shape
    out_f = w.shape[0]
    y = torch.empty((nrow, out_f), dtype=torch.float32, device=x.device)
    return gemm_tf32(x, 0, k, 0, w, 0, k, 0, y, out_f, 0, nrow, out_f, k, batch=1,
                     bias=bias, bias_mode=2 if bias is not None else 0, relu=relu)


def linear_wgrad(go, x):
    """dW[out][in] = dY[n][out]^T x X[n][in] (both operands MN-major: n is the K index)."""
    nrow, out_f = go.shape
    in_f = x.shape[1]
    dw = torch.empty((out_f, in_f), dtype=torch.float32, device=x.device)
    return gemm_tf32(go, 1, out_f, 0, x, 1, in_f, 0, dw, in_f, 0, out_f, in_f, nrow, batch=1)


def linear_dgrad(go, w):
    """dX[n][in] = dY[n][out] x W[out][in]."""
    nrow, out_f = go.shape
    in_f = w.shape[1]
    dx = torch.empty((nrow, in_f), dtype=torch.float32, device=go.device)
    return gemm_tf32(go, 0, out_f, 0, w, 1, in_f, 0, dx, in_f, 0, nrow, in_f, out_f, batch=1)


def recon_schedule(d_iter, d_sched, d_seeds, t_max, seed_base=0, rel_start=0.2, start_b=20.0, end_b=2.0,
                   b1=0.9, b2=0.999):
    """Device-side per-iteration scalars (beta, Adam bias corrections, mask seeds); increments
    d_iter. d_iter: int32[1], d_sched: float32[>=3], d_seeds: int64[n] (uint64 bit patterns)."""
    n = 0 if d_seeds is None else d_seeds.numel()
    check(lib().dpl_recon_schedule(d_iter.data_ptr(), d_sched.data_ptr(), _lib._ptr(d_seeds), n,
                                   float(t_max), float(rel_start), float(start_b), float(end_b),
                                   float(b1), float(b2), int(seed_base) & (2 ** 64 - 1), _stream()),
          "dpl_recon_schedule")
    _count()


def tf32_residual(x):
    """x - trunc_tf32(x): the low operand of the 3xTF32 product (weights: computed once)."""
    lo = torch.empty_like(x)
    check(lib().dpl_tf32_residual_f32(x.data_ptr(), lo.data_ptr(), x.numel(), _stream()),
          "dpl_tf32_residual_f32")
    _count()
    return lo


def tf32_split(x):
    """(hi, lo) = (RN_tf32(x), RN_tf32(x - hi)): the unbiased operand split of a WEIGHT for the 3xTF32 kernels
    (computed once). Pass `hi` where the kernels take the weight and `lo` as its residual: truncating the raw
    pattern instead shrinks every product by ~2^-22, which compounds over ResNet-50's 53 layers."""
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(lib().dpl_tf32_split_f32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()),
          "dpl_tf32_split_f32")
    _count()
    return hi, lo


def gemm_tf32x3(a, a_lo, a_major, lda, a_batch_stride, b, b_major, ldb, b_batch_stride, d, ldd,
                d_batch_stride, M, N, K, batch=1, bias=None, bias_mode=0, relu=False, d_relu=None, rng=None,
                rng_relu=None):
    dev = d.device
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    st = lib().dpl_gemm_tf32x3(a.data_ptr(), a_lo.data_ptr(), int(a_major), int(lda), int(a_batch_stride),
                               b.data_ptr(), int(b_major), int(ldb), int(b_batch_stride), d.data_ptr(),
                               int(ldd), int(d_batch_stride), int(M), int(N), int(K), int(batch),
                               _lib._ptr(bias), int(bias_mode), int(bool(relu)), _lib._ptr(d_relu),
                               *_rng(rng), *_rng(rng_relu), flag.data_ptr(), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_gemm_tf32x3")
    _count()
    return d


def conv1x1_forward_x3(x, w, w_lo, bias=None, relu=False, out=None, out_relu=None, rng=None, rng_relu=None):
    """fp32-accurate O[img][co][hw] = W[co][ci] x X[img][ci][hw] (+ bias[co]) on tensor cores;
    `out_relu` (same shape) also receives max(O, 0)."""
    n, ci, hh, ww = x.shape
    co = w.shape[0]
    hw = hh * ww
    o = torch.empty((n, co, hh, ww), dtype=torch.float32, device=x.device) if out is None else out
    _need(out_relu, torch.float32, "out_relu")
    return gemm_tf32x3(w, w_lo, 0, ci, 0, x, 1, hw, ci * hw, o, hw, co * hw, co, hw, ci, batch=n,
                       bias=bias, bias_mode=1 if bias is not None else 0, relu=relu, d_relu=out_relu, rng=rng,
                       rng_relu=rng_relu)


def conv1x1_px_forward_x3(x, w, w_lo, bias=None, out=None, out_relu=None, rng=None, rng_relu=None):
    """fp32-accurate 1x1 convolution, pixel-major tile (activations through TMEM), NCHW in / out."""
    _need(x, torch.float32, "x")
    _need(w, torch.float32, "w")
    _need(w_lo, torch.float32, "w_lo")
    n, ci, hh, ww = x.shape
    co = w.shape[0]
    o = torch.empty((n, co, hh, ww), dtype=torch.float32, device=x.device) if out is None else out
    _need(o, torch.float32, "out")
    _need(out_relu, torch.float32, "out_relu")
    dev = x.device
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    st = lib().dpl_conv1x1_px_tf32x3(x.data_ptr(), w.data_ptr(), w_lo.data_ptr(), o.data_ptr(), n, ci, co,
                                     hh * ww, _lib._ptr(bias), _lib._ptr(out_relu), *_rng(rng), *_rng(rng_relu),
                                     flag.data_ptr(), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_conv1x1_px_tf32x3")
    _count()
    return o


def linear_forward_x3(x, w, w_lo=None, bias=None, out=None, rng=None, split=None):
    """fp32-accurate Y = X W^T (+ bias) on the tap-table kernel with ONE tap: the rows of X are the "pixels"
    (TMEM lanes), W [out][k] is the tap's filter - so the Gemm layer gets the chunked accumulation and the
    unbiased operand split of the convolutions (dpl_x3p.cuh). `split` = tf32_split(w) if the caller caches it."""
    nrow, k = x.shape
    out_f = w.shape[0]
    y = torch.empty((nrow, out_f), dtype=torch.float32, device=x.device) if out is None else out
    w_hi, w_lo2 = split if split is not None else tf32_split(w)
    shifts = (ctypes.c_int * 1)(0)
    st = lib().dpl_conv_taps_tf32x3(x.data_ptr(), nrow, w_hi.data_ptr(), w_lo2.data_ptr(), y.data_ptr(), nrow, k,
                                    out_f, 1, 1, 1, 1, 0, 1, shifts, _lib._ptr(bias), 0, 0, *_rng(rng), 0, 0,
                                    _err_flag(x.device).data_ptr(), _stream())
    _unsupported(st, "dpl_conv_taps_tf32x3")
    _count()
    return y


def conv_taps_prepare(w):
    """Tap-major copy [kh*kw][co][ci] of a [co][ci][kh][kw] filter, split into its TF32 leading part
    (rounded to nearest) and residual (once per weight)."""
    co, ci = w.shape[0], w.shape[1]
    taps = w.permute(2, 3, 0, 1).reshape(-1, co, ci).contiguous()
    return tf32_split(taps)


conv3x3_prepare = conv_taps_prepare


class ConvPlan:
    """Staging geometry of dpl_conv_taps_tf32x3 for one (kernel, stride, input size)."""

    def __init__(self, n, h, w, ksize, stride):
        self.n, self.h, self.w, self.stride = n, h, w, stride
        if ksize == 3 and stride == 1:
            # compact zero border: ONE zero column per row and ONE zero row per image, shared with the next row /
            # image (pixel (h, w) at h * (W + 1) + w; column W of row h is also column -1 of row h + 1, the zero
            # row behind image i is also row -1 of image i + 1; rows before the first image are TMA's zero fill).
            # (H + 1)(W + 1) plane points per image instead of (H + 2)(W + 2): 1.31 -> 1.15 x the useful work
            # at 14 x 14, 1.65 -> 1.31 at 7 x 7.
            self.ho, self.wo, self.hp, self.wp, self.origin, self.planes = h, w, h + 1, w + 1, 0, 1
            self.shifts = [(kh - 1) * self.wp + (kw - 1) for kh in range(3) for kw in range(3)]
        elif ksize == 3 and stride == 2:
            self.ho, self.wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            self.hp, self.wp, self.origin, self.planes = self.ho + 1, self.wo + 1, 1, 4
            rows = n * self.hp * self.wp
            self.shifts = [(((kh + 1) & 1) * 2 + ((kw + 1) & 1)) * rows - (kh == 0) * self.wp - (kw == 0)
                           for kh in range(3) for kw in range(3)]
        elif ksize == 1 and stride in (1, 2):
            self.ho, self.wo = (h - 1) // stride + 1, (w - 1) // stride + 1
            self.hp, self.wp, self.origin, self.planes = self.ho, self.wo, 0, 1
            self.shifts = [0]
        else:
            raise GemmUnsupported("conv plan: kernel %d stride %d" % (ksize, stride))
        self.total_rows = self.planes * n * self.hp * self.wp
        self.c_shifts = (ctypes.c_int * len(self.shifts))(*self.shifts)


def conv3x3_plane_pitch(h, w):
    return (h + 1) * (w + 1)


def conv_taps_forward_x3(x, taps, taps_lo, ksize, stride, bias=None, relu=False, out=None, scratch=None,
                         out_relu=None, rng=None, rng_relu=None):
    """fp32-accurate 3x3 (stride 1 or 2, pad 1) or strided 1x1 convolution on the tensor cores:
    channel-last staging copy of x (dpl_pad_plane_f32), then the shifted-window GEMM."""
    n, ci, hh, ww = x.shape
    co = taps.shape[1]
    plan = ConvPlan(n, hh, ww, ksize, stride)
    need = plan.total_rows * ci
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.float32, device=x.device)
    check(lib().dpl_pad_plane_f32(x.data_ptr(), scratch.data_ptr(), n, ci, hh, ww, stride, plan.origin, plan.hp,
                                  plan.wp, plan.planes, _stream()), "dpl_pad_plane_f32")
    _count()
    y = torch.empty((n, co, plan.ho, plan.wo), dtype=torch.float32, device=x.device) if out is None else out
    dev = x.device
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    st = lib().dpl_conv_taps_tf32x3(scratch.data_ptr(), plan.total_rows, taps.data_ptr(), taps_lo.data_ptr(),
                                    y.data_ptr(), n, ci, co, plan.ho, plan.wo, plan.hp, plan.wp, plan.origin,
                                    len(plan.shifts), plan.c_shifts, _lib._ptr(bias), int(bool(relu)),
                                    _lib._ptr(out_relu), *_rng(rng), *_rng(rng_relu), flag.data_ptr(), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_conv_taps_tf32x3")
    _count()
    return y


def conv_direct_forward(x, w, bias, stride, pad, out=None, out_relu=None, rng=None, rng_relu=None):
    """Exact-fp32 direct convolution for few input channels (the stem), see dpl_conv_direct_f32."""
    _need(x, torch.float32, "x")
    _need(w, torch.float32, "w")
    n, c, hh, ww = x.shape
    co, _, kh, kw = w.shape
    ho, wo = (hh + 2 * pad - kh) // stride + 1, (ww + 2 * pad - kw) // stride + 1
    y = torch.empty((n, co, ho, wo), dtype=torch.float32, device=x.device) if out is None else out
    _need(y, torch.float32, "out")
    _need(out_relu, torch.float32, "out_relu")
    st = lib().dpl_conv_direct_f32(x.data_ptr(), w.data_ptr(), _lib._ptr(bias), y.data_ptr(), n, c, hh, ww, co, kh, kw,
                                   int(stride), int(pad), ho, wo, _lib._ptr(out_relu), *_rng(rng), *_rng(rng_relu),
                                   _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_conv_direct_f32")
    _count()
    return y


def dwconv2d_forward(x, w, bias, stride, pad, out=None, rng=None):
    """Depthwise convolution (group = channels), see dpl_dwconv2d_f32."""
    _need(x, torch.float32, "x")
    _need(w, torch.float32, "w")
    n, c, hh, ww = x.shape
    k = int(w.shape[2])
    ho, wo = (hh + 2 * pad - k) // stride + 1, (ww + 2 * pad - k) // stride + 1
    y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device) if out is None else out
    _need(y, torch.float32, "out")
    st = lib().dpl_dwconv2d_f32(x.data_ptr(), w.data_ptr(), _lib._ptr(bias), y.data_ptr(), n, c, hh, ww, k,
                                int(stride), int(pad), ho, wo, *_rng(rng), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_dwconv2d_f32")
    _count()
    return y


def conv_im2col_prepare(w):
    """[co][C][kh][kw] filter -> ([1][co][k_pad] row-major copy zero-padded to a multiple of 4, its
    TF32 residual, k_pad) for conv_im2col_forward_x3 (once per weight)."""
    co = w.shape[0]
    k = w[0].numel()
    k_pad = (k + 3) // 4 * 4
    w2 = torch.zeros((1, co, k_pad), dtype=torch.float32, device=w.device)
    w2[0, :, :k] = w.reshape(co, k)
    return tf32_split(w2) + (k_pad,)


def conv_im2col_forward_x3(x, prepared, kernel, stride, pad, bias=None, out=None, scratch=None, out_relu=None):
    """fp32-accurate convolution with few input channels (the 7x7 / stride 2 stem): im2col staging
    copy (dpl_im2col_f32), then the tap-table tensor-core kernel with a single tap."""
    w2, w2_lo, k_pad = prepared
    n, c, hh, ww = x.shape
    kh, kw = kernel
    co = w2.shape[1]
    ho, wo = (hh + 2 * pad - kh) // stride + 1, (ww + 2 * pad - kw) // stride + 1
    rows = n * ho * wo
    need = rows * k_pad
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.float32, device=x.device)
    check(lib().dpl_im2col_f32(x.data_ptr(), scratch.data_ptr(), n, c, hh, ww, kh, kw, stride, pad, ho, wo, k_pad,
                               _stream()), "dpl_im2col_f32")
    _count()
    y = torch.empty((n, co, ho, wo), dtype=torch.float32, device=x.device) if out is None else out
    dev = x.device
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    shifts = (ctypes.c_int * 1)(0)
    st = lib().dpl_conv_taps_tf32x3(scratch.data_ptr(), rows, w2.data_ptr(), w2_lo.data_ptr(), y.data_ptr(), n, k_pad,
                                    co, ho, wo, ho, wo, 0, 1, shifts, _lib._ptr(bias), 0, _lib._ptr(out_relu),
                                    0, 0, 0, 0, flag.data_ptr(), _stream())
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, "dpl_conv_taps_tf32x3")
    _count()
    return y


def conv3x3_forward_x3(x, taps, taps_lo, bias=None, relu=False, out=None, scratch=None):
    return conv_taps_forward_x3(x, taps, taps_lo, 3, 1, bias, relu, out, scratch)


# ---- K6 dense contraction for k x k / strided / depthwise convolutions (single-pass TF32) -------------------
def _err_flag(dev):
    flag = _gemm_err.get(dev)
    if flag is None:
        flag = _gemm_err[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    return flag


def _unsupported(st, what):
    if st == 10003:
        raise GemmUnsupported(lib().dpl_last_error().decode("utf-8", "replace"))
    check(st, what)


def _ints(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


class ReconConvPlan(ConvPlan):
    """Geometry of one learnable convolution on the tap-table kernels: the forward / weight-gradient taps
    (ConvPlan) plus the data-gradient taps per output-parity class.
    Supported: 3x3 pad 1 stride 1 / 2, 1x1 pad 0 stride 1 / 2 (dilation 1, no groups)."""

    def __init__(self, n, h, w, ksize, stride, pad):
        if (ksize, pad) not in ((3, 1), (1, 0)) or stride not in (1, 2):
            raise GemmUnsupported("recon conv plan: kernel %d stride %d pad %d" % (ksize, stride, pad))
        ConvPlan.__init__(self, n, h, w, ksize, stride)
        self.ksize, self.pad = ksize, pad
        self.q_total = n * self.hp * self.wp
        # dX[s j + a] = sum over kh with (a + pad - kh) % s == 0 of dY[j + (a + pad - kh) / s] W[kh]: in the
        # staging copy of dY (same origin / pitch as the forward's q index) that is a row shift of dh * Wp + dw
        self.dgrad = []
        for a in range(stride):
            for b in range(stride):
                shifts, taps = [], []
                for kh in range(ksize):
                    if (a + pad - kh) % stride:
                        continue
                    for kw in range(ksize):
                        if (b + pad - kw) % stride:
                            continue
                        dh, dw = (a + pad - kh) // stride, (b + pad - kw) // stride
                        shifts.append(dh * self.wp + dw)
                        taps.append(kh * ksize + kw)
                self.dgrad.append((a, b, shifts, taps))
        self.c_taps = _ints(range(len(self.shifts)))


def taps_layout(w, forward=True, dgrad=False, out_f=None, out_d=None):
    """w [co][ci][kh][kw] -> (wf [T][co][ci] or None, wd [T][ci][co] or None), see dpl_taps_layout_f32."""
    _need(w, torch.float32, "w")
    co, ci = w.shape[0], w.shape[1]
    t = w[0, 0].numel()
    wf = (torch.empty((t, co, ci), dtype=torch.float32, device=w.device) if out_f is None else out_f) \
        if forward else None
    wd = (torch.empty((t, ci, co), dtype=torch.float32, device=w.device) if out_d is None else out_d) \
        if dgrad else None
    check(lib().dpl_taps_layout_f32(w.data_ptr(), _lib._ptr(wf), _lib._ptr(wd), co, ci, t, _stream()),
          "dpl_taps_layout_f32")
    _count()
    return wf, wd


def recon_stage_input(x, plan, out=None):
    """Channel-last zero-bordered staging copy Xp [planes * n * Hp * Wp][ci] of the layer input."""
    n, ci, hh, ww = x.shape
    need = plan.total_rows * ci
    xp = torch.empty(need, dtype=torch.float32, device=x.device) if out is None or out.numel() < need else out
    check(lib().dpl_pad_plane_f32(x.data_ptr(), xp.data_ptr(), n, ci, hh, ww, plan.stride, plan.origin, plan.hp,
                                  plan.wp, plan.planes, _stream()), "dpl_pad_plane_f32")
    _count()
    return xp


def recon_stage_grad(go, plan, out=None):
    """Staging copy Gp [n * Hp * Wp][co] of the output gradient in the forward's q indexing (zero border)."""
    n, co, ho, wo = go.shape
    need = plan.q_total * co
    gp = torch.empty(need, dtype=torch.float32, device=go.device) if out is None or out.numel() < need else out
    check(lib().dpl_pad_plane_f32(go.data_ptr(), gp.data_ptr(), n, co, ho, wo, 1, plan.origin, plan.hp, plan.wp, 1,
                                  _stream()), "dpl_pad_plane_f32")
    _count()
    return gp


def recon_conv_forward(xp, plan, wf, bias=None, out=None):
    """Y = conv(X, W) from the staging copy xp and the tap-major filter wf [T][co][ci] (TF32, fp32 accumulate)."""
    t, co, ci = wf.shape
    y = torch.empty((plan.n, co, plan.ho, plan.wo), dtype=torch.float32, device=xp.device) if out is None else out
    st = lib().dpl_tap_conv_tf32(xp.data_ptr(), plan.total_rows, wf.data_ptr(), t, y.data_ptr(), plan.n, ci, co,
                                 plan.ho, plan.wo, plan.hp, plan.wp, plan.origin, 1, 0, 0, len(plan.shifts),
                                 plan.c_shifts, plan.c_taps, _lib._ptr(bias), _err_flag(xp.device).data_ptr(),
                                 _stream())
    _unsupported(st, "dpl_tap_conv_tf32")
    _count()
    return y


def recon_conv_wgrad(gp, xp, plan, co, ci, out=None):
    """dW [co][ci][k][k] from the staging copies of dY (gp) and of the layer input (xp)."""
    k = plan.ksize
    dw = torch.empty((co, ci, k, k), dtype=torch.float32, device=gp.device) if out is None else out
    st = lib().dpl_tap_wgrad_tf32(gp.data_ptr(), plan.q_total, xp.data_ptr(), plan.total_rows, dw.data_ptr(), co, ci,
                                  k * k, len(plan.shifts), plan.c_shifts, plan.c_taps,
                                  _err_flag(gp.device).data_ptr(), _stream())
    _unsupported(st, "dpl_tap_wgrad_tf32")
    _count()
    return dw


def recon_conv_dgrad(gp, plan, wd, out=None):
    """dX [n][ci][H][W] from the staging copy of dY and the filter in data-gradient layout wd [T][ci][co];
    one launch per output-parity class of a strided convolution."""
    t, ci, co = wd.shape
    dx = torch.empty((plan.n, ci, plan.h, plan.w), dtype=torch.float32, device=gp.device) if out is None else out
    if any(not taps for _, _, _, taps in plan.dgrad):
        dx.zero_()                      # classes no tap reaches (1x1 stride 2)
    for a, b, shifts, taps in plan.dgrad:
        if not taps:
            continue
        st = lib().dpl_tap_conv_tf32(gp.data_ptr(), plan.q_total, wd.data_ptr(), t, dx.data_ptr(), plan.n, co, ci,
                                     plan.h, plan.w, plan.hp, plan.wp, plan.origin, plan.stride, a, b, len(taps),
                                     _ints(shifts), _ints(taps), 0, _err_flag(gp.device).data_ptr(), _stream())
        _unsupported(st, "dpl_tap_conv_tf32")
        _count()
    return dx


def dwconv2d_wgrad(x, go, k, stride, pad, out=None):
    """Depthwise weight gradient dW [C][1][k][k], exact fp32."""
    n, c, hh, ww = x.shape
    ho, wo = go.shape[2], go.shape[3]
    gw = torch.empty((c, 1, k, k), dtype=torch.float32, device=x.device) if out is None else out
    st = lib().dpl_dwconv2d_wgrad_f32(x.data_ptr(), go.data_ptr(), gw.data_ptr(), n, c, hh, ww, k, int(stride),
                                      int(pad), ho, wo, _stream())
    _unsupported(st, "dpl_dwconv2d_wgrad_f32")
    _count()
    return gw


def dwconv2d_dgrad(go, w, in_hw, stride, pad, out=None):
    """Depthwise data gradient dX [n][C][H][W], exact fp32."""
    n, c, ho, wo = go.shape
    k = int(w.shape[2])
    hh, ww = in_hw
    gx = torch.empty((n, c, hh, ww), dtype=torch.float32, device=go.device) if out is None else out
    st = lib().dpl_dwconv2d_dgrad_f32(go.data_ptr(), w.data_ptr(), gx.data_ptr(), n, c, hh, ww, k, int(stride),
                                      int(pad), ho, wo, _stream())
    _unsupported(st, "dpl_dwconv2d_dgrad_f32")
    _count()
    return gx


def conv_im2col_wgrad(x, go, kernel, stride, pad, scratch=None):
    """Weight gradient of a convolution with very few input channels (the stem): im2col staging copy, then
    dW[co][k] = sum_img dY[img][co][px] x cols[img][px][k] on the tcgen05 tile (batch folded into K)."""
    n, c, hh, ww = x.shape
    kh, kw = kernel
    co, ho, wo = go.shape[1], go.shape[2], go.shape[3]
    k = c * kh * kw
    k_pad = (k + 3) // 4 * 4
    px = ho * wo
    need = n * px * k_pad
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.float32, device=x.device)
    st = lib().dpl_im2col_f32(x.data_ptr(), scratch.data_ptr(), n, c, hh, ww, kh, kw, int(stride), int(pad), ho, wo,
                              k_pad, _stream())
    _unsupported(st, "dpl_im2col_f32")
    _count()
    dw = torch.zeros((co, k_pad), dtype=torch.float32, device=x.device)
    tiles = ((co + 127) // 128) * ((k_pad + 127) // 128)
    split = max(1, min(n, 296 // max(tiles, 1)))
    gemm_tf32(go, 0, px, co * px, scratch, 1, k_pad, px * k_pad, dw, k_pad, 0, co, k_pad, px, batch=n,
              fold_batch=True, split_k=split)
    return dw[:, :k].reshape(co, c, kh, kw).contiguous()

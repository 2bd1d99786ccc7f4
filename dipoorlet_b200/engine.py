"""Batched blob producer: executes the (fp32 or Q/DQ) graph on the GPU for a whole batch
of calibration images and leaves every requested activation blob resident in HBM, where
the statistics kernels consume it in place.

Replaces the reference's per-sample `ort.InferenceSession.run` with every node output
promoted to a graph output and copied back to the host (dipoorlet/forward_net.py:195-216)
and the per-node sessions of ActivationCache.forward_subnet (forward_net.py:81-128).

1x1 / 3x3 / strided convolutions and Gemm layers run on libdpl_b200's tcgen05 tiles in 3xTF32 mode with chunked
accumulation (fp32-accurate products on the TF32 tensor cores, csrc/dpl_x3p.cuh); the few-channel stem
convolution and the depthwise convolutions of MobileNetV2 on exact-fp32 FMA kernels (dpl_conv_direct.cu);
Relu / Clip / Add (+ the Relu behind it) / MaxPool / GlobalAveragePool are libdpl_b200 streaming kernels
(dpl_eltwise.cu) - a ResNet-50 or MobileNetV2 forward issues no library kernel. Operators neither family
contains (ConvTranspose, grouped non-depthwise or dilated convolutions, rarely used element-wise ops) are still
issued through torch's CUDA ops (cuDNN with TF32 disabled), i.e. a library call playing the role onnxruntime's
CUDA EP plays in the reference. QuantizeLinear + DequantizeLinear pairs are ONE fused K5 launch.

Blob memory: when `engine.arena` is set (forward_net.CalibrationSession does) every node output is
bump-allocated from that one slab instead of torch's caching allocator.
"""
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import kernels as K

QDQ = ("QuantizeLinear", "DequantizeLinear")


def _sym_pads(pads, nd):
    lo, hi = list(pads[:nd]), list(pads[nd:])
    return lo == hi, lo, hi


class RangeSink:
    """Where the forward's kernels fold min / max of the blobs they write (fused pass-1 statistics)."""

    def __init__(self, blob_min, blob_max, names):
        self.blob_min, self.blob_max = blob_min, blob_max
        self.index = {n: i for i, n in enumerate(names)}
        self.covered = set()


class Engine:
    def __init__(self, onnx_graph, device=None, allow_tf32=False, _unit_test_cpu=False):
        self.g = onnx_graph
        self.device = torch.device(device if device is not None else "cuda")
        # `_unit_test_cpu` lets tests/ check the operator interpreter against the oracle on a
        # box without a GPU; no product entry point sets it, and Q/DQ (K5) still needs CUDA.
        if self.device.type != "cuda" and not _unit_test_cpu:
            raise RuntimeError("dipoorlet_b200.engine.Engine runs on a CUDA device only "
                               "(no CPU fallback on the product path)")
        self.allow_tf32 = allow_tf32
        self.params = {}
        self.zero_points = {}
        self._w_lo = {}            # tf32 residuals of the 1x1 / Gemm weights (3xTF32 operand)
        self._w_taps = {}          # tap-major 3x3 filters + residuals
        self._im2col = {}          # [co][C kh kw] stem filters + residuals
        self._pad_scratch = None   # zero-bordered input copy of the 3x3 convolution (grow-only)
        self.tc_conv3x3 = os.environ.get("DPL_ENGINE_CONV3X3", "1") != "0"
        self._tc_off = set()       # nodes the tensor-core tile could not address
        self.tensor_cores = os.environ.get("DPL_ENGINE_TCGEN05", "1") != "0"
        # 1x1 convolutions on the pixel-major tile with the activations through TMEM (dpl_conv1x1_px_tf32x3 ->
        # csrc/dpl_x3ts.cuh) instead of the channel-major tile with both operands in shared memory
        # (dpl_gemm_tf32x3 -> csrc/dpl_x3p.cuh): correct and tested, but measured slower on ResNet-50's
        # shapes (hist job 5 660 vs 7 510 images/s, profiles/r2_summary.md), so opt-in
        self.conv1x1_px = os.environ.get("DPL_ENGINE_CONV1X1_PX", "0") == "1"
        self.native_ops = os.environ.get("DPL_ENGINE_NATIVE_OPS", "1") != "0"
        # Relu blob written by the Conv epilogue (second store stream). Measured on B200, hist job: a loss
        # with the row-per-thread epilogue stores (convolutions 6.07 -> 6.66 ms per 64 images), a small gain
        # (7820 -> 7890 images/s) once the persistent 1x1 tile stores through its shared-memory transpose.
        self.fuse_conv_relu = os.environ.get("DPL_ENGINE_FUSE_CONV_RELU", "1") != "0"
        self.arena = None          # kernels.BlobArena: where node outputs are allocated
        self.refresh_initializers()
        self.nodes = list(onnx_graph.model.graph.nodes)
        self._fused_q = set()
        self._relu_after = None    # Add node name -> the Relu node fused with it
        # few-channel stem convolution through im2col staging + the single-tap tensor-core kernel: correct
        # and tested, measured slower than cuDNN's fp32 kernel (1.05 vs 0.81 ms per 64 images: the staging
        # copy is 475 MB), so opt-in until the gather moves into the kernel
        self.stem_im2col = os.environ.get("DPL_ENGINE_STEM_IM2COL", "0") == "1"
        # few-channel stem convolution on libdpl_b200's direct fp32 kernel (Relu blob from its epilogue)
        self.stem_direct = os.environ.get("DPL_ENGINE_STEM_DIRECT", "1") != "0"

    # ------------------------------------------------------------------ blob memory
    def _new(self, shape, like):
        """Output buffer for a node: from the arena when one is attached."""
        if self.arena is not None and like.is_cuda:
            try:
                return self.arena.alloc(shape)
            except MemoryError:
                pass
        return torch.empty(tuple(int(d) for d in shape), dtype=torch.float32, device=like.device)

    def _adopt(self, v):
        """Move a tensor produced by a library op into the arena (one extra copy)."""
        if (self.arena is None or not torch.is_tensor(v) or not v.is_cuda or v.dtype != torch.float32
                or v.untyped_storage().data_ptr() == self.arena.buf.untyped_storage().data_ptr()):
            return v
        try:
            o = self.arena.alloc(v.shape)
        except MemoryError:
            return v
        o.copy_(v)
        return o

    def _relu_out(self, node, like, env, always=False):
        """Buffer for the fused Relu blob behind `node`, or None when there is nothing to fuse.
        always=True: the producer's epilogue is not the bottleneck (direct stem convolution)."""
        if not self.native_ops or not (self.fuse_conv_relu or always):
            return None
        relu = self._fusable_relu(node)
        if relu is None or relu.output[0] in env:
            return None
        return self._new(like.shape, like)

    def _publish_relu(self, node, r, env):
        if r is not None:
            env[self._fusable_relu(node).output[0]] = r

    def _rng(self, name):
        """Fused-statistics target of output `name` (None when no sink is attached)."""
        st = getattr(self, "_stats", None)
        if st is None or name not in st.index:
            return None
        st.covered.add(name)
        return (st.blob_min, st.blob_max, st.index[name])

    def _uncover(self, node):
        """A fused kernel refused the node after _rng() registered its outputs: K1 must read them."""
        st = getattr(self, "_stats", None)
        if st is not None:
            st.covered.discard(node.output[0])
            relu = self._fusable_relu(node)
            if relu is not None:
                st.covered.discard(relu.output[0])

    def _fallback(self, node):
        """A libdpl_b200 kernel refused this node's operands (alignment / shape): say so once, route the node
        to the next path for the rest of the run and give its blobs' range statistics back to K1."""
        if node.name not in self._tc_off:
            from .utils import logger
            logger.warning("engine: %s (%s) falls back from its libdpl_b200 kernel: %s"
                           % (node.name, node.op_type, K.lib().dpl_last_error().decode("utf-8", "replace")))
        self._tc_off.add(node.name)
        self._uncover(node)

    def _rng_relu(self, node):
        relu = self._fusable_relu(node)
        return self._rng(relu.output[0]) if relu is not None else None

    def _native(self, x):
        return self.native_ops and torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 \
            and x.is_contiguous()

    def _fusable_relu(self, node):
        """The Relu that is the single consumer of this Conv's / Add's output: its blob is written
        by the producer's epilogue instead of a pass of its own."""
        if self._relu_after is None:
            self._relu_after = {}
            for n in self.nodes:
                if n.op_type not in ("Add", "Conv"):
                    continue
                cons = self.g.input_map.get(n.output[0], [])
                if len(cons) == 1 and cons[0].op_type == "Relu" and n.output[0] not in self.g.network_outputs:
                    self._relu_after[n.name] = cons[0]
        return self._relu_after.get(node.name)

    def refresh_initializers(self, names=None):
        """(Re)upload initializers — after bias correction / rounding rewrote weights."""
        inits = self.g.model.graph.initializers
        for name in (names if names is not None else inits):
            self._w_lo.pop(name, None)
            self._w_taps.pop(name, None)
            self._im2col.pop(name, None)
            arr = inits[name]
            if arr.dtype == np.float64:
                arr = arr.astype(np.float32)
            self.params[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
            if arr.dtype in (np.int8, np.uint8):  # zero points: K5 takes int32
                self.zero_points[name] = (self.params[name].reshape(-1).to(torch.int32).contiguous(),
                                          arr.dtype == np.uint8)

    # ------------------------------------------------------------------ execution
    def blob_names(self):
        """Names in the order the reference's statistics dicts are filled
        (forward_net.py:195-198,220-235): network inputs, then every node output in node
        order with the original network outputs moved to the end."""
        outs = [o for n in self.nodes for o in n.output if o and o not in self.g.network_outputs]
        return list(self.g.network_inputs) + outs + list(self.g.network_outputs)

    def run(self, feeds, want="all", start_after=None, cache=None, stats=None):
        """feeds: name -> [B, ...] float32 CUDA tensor. Returns OrderedDict name -> tensor
        for `want` ('all' = every blob of blob_names(), or an iterable of names).
        `cache` (dict) is consulted before computing a tensor and is not modified.
        `stats` (RangeSink): running per-blob min / max that the producing kernels update while they
        write a blob; the names they covered are added to stats.covered."""
        self._stats = stats
        torch.backends.cudnn.allow_tf32 = self.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.allow_tf32
        if os.environ.get("DPL_CUDNN_BENCHMARK"):   # let cuDNN time its fp32 algorithms per shape
            torch.backends.cudnn.benchmark = os.environ["DPL_CUDNN_BENCHMARK"] == "1"
        env = dict(feeds)
        if cache:
            for k, v in cache.items():
                env.setdefault(k, v)
        want_all = isinstance(want, str) and want == "all"
        wanted = None if want_all else set(want)
        need = None
        if not want_all:
            need = self._needed_nodes(wanted, env)
        remaining = {}
        if not want_all:
            for i in need:
                for t in self.nodes[i].input:
                    remaining[t] = remaining.get(t, 0) + 1
        with torch.no_grad():
            for i, node in enumerate(self.nodes):
                if need is not None and i not in need:
                    continue
                if all(o in env for o in node.output if o):
                    continue
                outs = self._exec(node, env)
                for o, v in zip(node.output, outs):
                    if o:
                        env[o] = self._adopt(v)
                if not want_all:
                    for t in node.input:
                        if t in remaining:
                            remaining[t] -= 1
                            if remaining[t] == 0 and t not in wanted and t not in feeds \
                                    and not (cache and t in cache):
                                env.pop(t, None)
        self._stats = None
        names = self.blob_names() if want_all else [w for w in want]
        return OrderedDict((n, env[n]) for n in names if n in env)

    def inputs_required(self, want, have):
        """Network inputs that must be fed to compute `want` when the tensors in `have` are
        already available (lets ActivationCache skip the host->device copy of the images)."""
        need = self._needed_nodes(set(want), {k: None for k in have}, strict=False)
        used = {t for i in need for t in self.nodes[i].input}
        return [n for n in self.g.network_inputs if (n in used or n in want) and n not in have]

    def _needed_nodes(self, wanted, env, strict=True):
        need, stack = set(), [w for w in wanted if w not in env]
        while stack:
            t = stack.pop()
            prod = self.g.output_map.get(t)
            if prod is None:
                if strict and t not in self.params and t not in env and t != "":
                    raise KeyError(f"tensor {t!r} has no producer and was not fed")
                continue
            idx = self.g.name_idx_map[prod.name]
            if idx in need:
                continue
            need.add(idx)
            for i in prod.input:
                if i and i not in env and i not in self.params:
                    stack.append(i)
        return need

    def _val(self, name, env):
        if name in env:
            return env[name]
        return self.params[name]

    def _tc_conv_ok(self, node, x, w, stride, dil, lo, hi):
        """1x1, stride 1, unpadded, ungrouped convolutions whose C_in and H*W are multiples of 4
        (TMA's 16-byte rule) run on libdpl_b200's 3xTF32 tensor-core GEMM."""
        return (self.tensor_cores and x.is_cuda and node.name not in self._tc_off and w.dim() == 4
                and list(w.shape[2:]) == [1, 1] and list(stride) == [1, 1] and list(dil) == [1, 1]
                and not any(lo) and not any(hi) and node.attrs.get("group", 1) == 1
                and x.is_contiguous() and x.shape[1] % 4 == 0 and (x.shape[2] * x.shape[3]) % 4 == 0)

    def _split(self, name, w2):
        """(hi, lo) TF32 split of a weight (K.tf32_split); cached for initializers only (a weight produced by
        a Q/DQ pair inside the graph is recomputed: it changes whenever its source does)."""
        if name not in self.params:
            return K.tf32_split(w2)
        pair = self._w_lo.get(name)
        if pair is None or pair[0].shape != w2.shape:
            pair = self._w_lo[name] = K.tf32_split(w2)
        return pair

    def _taps(self, name, w):
        """Tap-major copy + residual of a filter for dpl_conv_taps_tf32x3 (same caching rule)."""
        if name not in self.params:
            return K.conv_taps_prepare(w)
        t = self._w_taps.get(name)
        if t is None:
            t = self._w_taps[name] = K.conv_taps_prepare(w)
        return t

    def _tc_conv_taps(self, node, x, w, stride, dil, lo, hi):
        """(kernel, stride) when the convolution runs on dpl_conv_taps_tf32x3 — 3x3 / pad 1 with
        stride 1 or 2, or 1x1 / stride 2 / unpadded; ungrouped, C_in a multiple of 4 — else None."""
        if not (self.tensor_cores and self.tc_conv3x3 and x.is_cuda and node.name not in self._tc_off
                and w.dim() == 4 and list(dil) == [1, 1] and node.attrs.get("group", 1) == 1
                and x.is_contiguous() and x.shape[1] % 4 == 0 and x.shape[1] >= 16
                and stride[0] == stride[1] and w.shape[2] == w.shape[3]):
            return None
        k, st = int(w.shape[2]), int(stride[0])
        if k == 3 and st in (1, 2) and list(lo) == [1, 1] and list(hi) == [1, 1]:
            return k, st
        if k == 1 and st == 2 and not any(lo) and not any(hi):
            return k, st
        if k == 1 and st == 1 and not any(lo) and not any(hi) and (x.shape[2] * x.shape[3]) % 4:
            return k, st      # e.g. 7 x 7 maps: rows of 49 floats are not TMA-addressable per image
        return None

    def _host(self, name, env):
        """Small constant operands (Clip bounds, Reshape targets) read on the host without
        a device sync when they are initializers."""
        arr = self.g.model.graph.initializers.get(name)
        return arr if arr is not None else self._val(name, env).cpu().numpy()

    def _exec(self, node, env):
        op, a = node.op_type, node.attrs
        if op in ("Conv", "ConvTranspose", "MaxPool", "AveragePool") and a.get("auto_pad", "NOTSET") not in ("NOTSET", "", None):
            # graph.py's shape inference honours auto_pad; executing from `pads` alone would silently compute a
            # differently padded layer into buffers sized for SAME / VALID
            raise NotImplementedError("%s %s: auto_pad=%s is not supported by the engine (export with explicit pads)"
                                      % (op, node.name, a.get("auto_pad")))
        x = self._val(node.input[0], env) if node.input and node.input[0] else None
        if op == "Conv":
            w = self._val(node.input[1], env)
            b = self._val(node.input[2], env) if len(node.input) > 2 and node.input[2] else None
            nd = w.dim() - 2
            stride = a.get("strides", [1] * nd)
            dil = a.get("dilations", [1] * nd)
            sym, lo, hi = _sym_pads(a.get("pads", [0] * (2 * nd)), nd)
            if self._tc_conv_ok(node, x, w, stride, dil, lo, hi):
                try:    # 1x1 convolution as a 3xTF32 GEMM on the tcgen05 tile (fp32-accurate)
                    w2 = w.view(w.shape[0], w.shape[1])
                    out = self._new((x.shape[0], w.shape[0], x.shape[2], x.shape[3]), x)
                    r = self._relu_out(node, out, env)
                    w_hi, w_lo = self._split(node.input[1], w2)
                    if self.conv1x1_px:
                        y = K.conv1x1_px_forward_x3(x, w_hi, w_lo, b, out=out, out_relu=r,
                                                    rng=self._rng(node.output[0]),
                                                    rng_relu=self._rng_relu(node) if r is not None else None)
                    else:
                        y = K.conv1x1_forward_x3(x, w_hi, w_lo, b, out=out, out_relu=r,
                                                 rng=self._rng(node.output[0]),
                                                 rng_relu=self._rng_relu(node) if r is not None else None)
                    self._publish_relu(node, r, env)
                    return [y]
                except K.GemmUnsupported:
                    self._fallback(node)
            taps_cfg = self._tc_conv_taps(node, x, w, stride, dil, lo, hi)
            if taps_cfg is not None:
                try:
                    taps, taps_lo = self._taps(node.input[1], w)
                    plan = K.ConvPlan(x.shape[0], x.shape[2], x.shape[3], *taps_cfg)
                    need = plan.total_rows * x.shape[1]
                    if self._pad_scratch is None or self._pad_scratch.numel() < need:
                        self._pad_scratch = torch.empty(need, dtype=torch.float32, device=x.device)
                    out = self._new((x.shape[0], w.shape[0], plan.ho, plan.wo), x)
                    r = self._relu_out(node, out, env)
                    y = K.conv_taps_forward_x3(x, taps, taps_lo, taps_cfg[0], taps_cfg[1], b,
                                               scratch=self._pad_scratch, out=out, out_relu=r,
                                               rng=self._rng(node.output[0]),
                                               rng_relu=self._rng_relu(node) if r is not None else None)
                    self._publish_relu(node, r, env)
                    return [y]
                except K.GemmUnsupported:
                    self._fallback(node)
            if (self.stem_direct and x.is_cuda and node.name not in self._tc_off and w.dim() == 4 and x.dim() == 4
                    and x.is_contiguous() and w.is_contiguous() and a.get("group", 1) == x.shape[1] == w.shape[0]
                    and w.shape[1] == 1 and x.shape[1] > 1 and list(dil) == [1, 1] and sym and lo[0] == lo[1]
                    and stride[0] == stride[1] and w.shape[2] == w.shape[3] and int(w.shape[2]) in (3, 5)):
                try:   # depthwise convolution: one streaming pass
                    ho = (x.shape[2] + 2 * lo[0] - w.shape[2]) // stride[0] + 1
                    wo = (x.shape[3] + 2 * lo[1] - w.shape[3]) // stride[1] + 1
                    out = self._new((x.shape[0], w.shape[0], ho, wo), x)
                    return [K.dwconv2d_forward(x, w, b, int(stride[0]), int(lo[0]), out=out,
                                               rng=self._rng(node.output[0]))]
                except K.GemmUnsupported:
                    self._fallback(node)
            if (self.stem_direct and x.is_cuda and node.name not in self._tc_off and w.dim() == 4 and x.dim() == 4
                    and x.is_contiguous() and w.is_contiguous() and a.get("group", 1) == 1 and list(dil) == [1, 1]
                    and sym and lo[0] == lo[1] and stride[0] == stride[1] and w.shape[2] == w.shape[3]
                    and x.shape[1] <= 4 and (int(w.shape[2]), int(stride[0])) in ((7, 2), (5, 1), (5, 2), (3, 1), (3, 2))):
                try:   # few input channels (the stem): exact-fp32 direct convolution on the FMA pipe
                    ho = (x.shape[2] + 2 * lo[0] - w.shape[2]) // stride[0] + 1
                    wo = (x.shape[3] + 2 * lo[1] - w.shape[3]) // stride[1] + 1
                    out = self._new((x.shape[0], w.shape[0], ho, wo), x)
                    r = self._relu_out(node, out, env, always=True)
                    y = K.conv_direct_forward(x, w, b, int(stride[0]), int(lo[0]), out=out, out_relu=r,
                                              rng=self._rng(node.output[0]),
                                              rng_relu=self._rng_relu(node) if r is not None else None)
                    self._publish_relu(node, r, env)
                    return [y]
                except K.GemmUnsupported:
                    self._fallback(node)
            if (self.stem_im2col and self.tensor_cores and x.is_cuda and node.name not in self._tc_off and w.dim() == 4
                    and x.is_contiguous() and a.get("group", 1) == 1 and list(dil) == [1, 1] and sym
                    and lo[0] == lo[1] and stride[0] == stride[1] and x.shape[1] < 16
                    and w.shape[1] * w.shape[2] * w.shape[3] <= 256):
                try:   # few input channels (the stem): im2col staging + the single-tap tensor-core kernel
                    prep = self._im2col.get(node.input[1]) if node.input[1] in self.params else None
                    if prep is None:
                        prep = K.conv_im2col_prepare(w)
                        if node.input[1] in self.params:
                            self._im2col[node.input[1]] = prep
                    ho = (x.shape[2] + 2 * lo[0] - w.shape[2]) // stride[0] + 1
                    wo = (x.shape[3] + 2 * lo[1] - w.shape[3]) // stride[1] + 1
                    need = x.shape[0] * ho * wo * prep[2]
                    if self._pad_scratch is None or self._pad_scratch.numel() < need:
                        self._pad_scratch = torch.empty(need, dtype=torch.float32, device=x.device)
                    out = self._new((x.shape[0], w.shape[0], ho, wo), x)
                    r = self._relu_out(node, out, env)
                    y = K.conv_im2col_forward_x3(x, prep, (int(w.shape[2]), int(w.shape[3])), int(stride[0]), int(lo[0]),
                                                 b, out=out, scratch=self._pad_scratch, out_relu=r)
                    self._publish_relu(node, r, env)
                    return [y]
                except K.GemmUnsupported:
                    self._fallback(node)
            if not sym:
                x = F.pad(x, [p for i in reversed(range(nd)) for p in (lo[i], hi[i])])
                lo = [0] * nd
            fn = F.conv2d if nd == 2 else (F.conv1d if nd == 1 else F.conv3d)
            return [fn(x, w, b, stride, lo, dil, a.get("group", 1))]
        if op == "ConvTranspose":
            w = self._val(node.input[1], env)
            b = self._val(node.input[2], env) if len(node.input) > 2 and node.input[2] else None
            nd = w.dim() - 2
            pads = a.get("pads", [0] * (2 * nd))
            return [F.conv_transpose2d(x, w, b, a.get("strides", [1] * nd), pads[:nd],
                                       a.get("output_padding", [0] * nd), a.get("group", 1),
                                       a.get("dilations", [1] * nd))]
        if op == "Relu":
            if self._native(x):
                return [K.clip(x, 0.0, float("inf"), out=self._new(x.shape, x), rng=self._rng(node.output[0]))]
            return [torch.relu(x)]
        if op == "Clip":
            lo = a.get("min")
            hi = a.get("max")
            if len(node.input) > 1 and node.input[1]:
                lo = float(self._host(node.input[1], env).reshape(-1)[0])
            if len(node.input) > 2 and node.input[2]:
                hi = float(self._host(node.input[2], env).reshape(-1)[0])
            if self._native(x):
                return [K.clip(x, float("-inf") if lo is None else lo, float("inf") if hi is None else hi,
                               out=self._new(x.shape, x), rng=self._rng(node.output[0]))]
            return [torch.clamp(x, lo, hi)]
        if op in ("MaxPool", "AveragePool"):
            nd = x.dim() - 2
            k = a["kernel_shape"]
            stride = a.get("strides", [1] * nd)
            sym, lo, hi = _sym_pads(a.get("pads", [0] * (2 * nd)), nd)
            ceil_mode = bool(a.get("ceil_mode", 0))
            if op == "MaxPool" and nd == 2 and self._native(x) and list(a.get("dilations", [1, 1])) == [1, 1]:
                def osz(n, k_, s_, p0, p1):
                    v = n + p0 + p1 - k_
                    o = (-(-v // s_) if ceil_mode else v // s_) + 1
                    if ceil_mode and (o - 1) * s_ >= n + p0:   # the last window must start inside
                        o -= 1
                    return o
                ho, wo = osz(x.shape[2], k[0], stride[0], lo[0], hi[0]), osz(x.shape[3], k[1], stride[1], lo[1], hi[1])
                return [K.maxpool2d(x, k, stride, lo[0], lo[1], ho, wo,
                                    out=self._new((x.shape[0], x.shape[1], ho, wo), x), rng=self._rng(node.output[0]))]
            if op == "MaxPool":
                if not sym:
                    x = F.pad(x, [p for i in reversed(range(nd)) for p in (lo[i], hi[i])],
                              value=float("-inf"))
                    lo = [0] * nd
                return [F.max_pool2d(x, k, stride, lo, a.get("dilations", [1] * nd), ceil_mode)]
            if not sym:
                raise NotImplementedError("AveragePool with asymmetric pads")
            return [F.avg_pool2d(x, k, stride, lo, ceil_mode, bool(a.get("count_include_pad", 0)))]
        if op == "GlobalAveragePool":
            if self._native(x) and x.dim() >= 3:
                return [K.global_avgpool(x, out=self._new(tuple(x.shape[:2]) + (1,) * (x.dim() - 2), x),
                                         rng=self._rng(node.output[0]))]
            return [x.mean(dim=tuple(range(2, x.dim())), keepdim=True)]
        if op in ("Add", "Sub", "Mul", "Div"):
            y = self._val(node.input[1], env)
            if op == "Add" and self._native(x) and self._native(y) and x.shape == y.shape:
                relu = self._fusable_relu(node)
                if relu is not None and relu.output[0] not in env:
                    # the block's Add and the Relu behind it: both blobs from one read of the operands
                    r = self._new(x.shape, x)
                    out = K.add(x, y, out=self._new(x.shape, x), out_relu=r, rng=self._rng(node.output[0]),
                                rng_relu=self._rng(relu.output[0]))
                    env[relu.output[0]] = r
                    return [out]
                return [K.add(x, y, out=self._new(x.shape, x), rng=self._rng(node.output[0]))]
            return [{"Add": torch.add, "Sub": torch.sub, "Mul": torch.mul, "Div": torch.div}[op](x, y)]
        if op == "Flatten":
            ax = a.get("axis", 1)
            return [x.flatten(ax) if ax > 0 else x.reshape(1, -1)]
        if op == "Gemm":
            w = self._val(node.input[1], env)
            c = self._val(node.input[2], env) if len(node.input) > 2 and node.input[2] else None
            alpha, beta = a.get("alpha", 1.0), a.get("beta", 1.0)
            if a.get("transA", 0):
                x = x.t()
            if alpha == 1.0 and beta == 1.0 and a.get("transB", 0) and (c is None or c.dim() == 1):
                if (self.tensor_cores and x.is_cuda and node.name not in self._tc_off and x.dim() == 2
                        and x.shape[1] % 4 == 0 and x.is_contiguous()):
                    try:
                        return [K.linear_forward_x3(x, w, None, c, out=self._new((x.shape[0], w.shape[0]), x),
                                                    rng=self._rng(node.output[0]),
                                                    split=self._split(node.input[1], w))]
                    except K.GemmUnsupported:
                        self._fallback(node)
                return [F.linear(x, w, c)]
            y = alpha * (x @ (w.t() if a.get("transB", 0) else w))
            return [y if c is None else y + beta * c]
        if op == "MatMul":
            return [x @ self._val(node.input[1], env)]
        if op == "Reshape":
            tgt = [int(v) for v in self._host(node.input[1], env).reshape(-1)]
            declared = self.g.tensor_name_shape_map.get(node.input[0])
            if declared and tgt and declared[0] == tgt[0] and x.shape[0] != declared[0]:
                tgt[0] = x.shape[0]  # the leading dim is the image index
            tgt = [x.shape[i] if v == 0 else v for i, v in enumerate(tgt)]
            return [x.reshape(tgt)]
        if op == "Sigmoid":
            return [torch.sigmoid(x)]
        if op == "Identity" or op == "Dropout":
            return [x]
        if op == "Concat":
            return [torch.cat([self._val(i, env) for i in node.input], a["axis"])]
        if op == "Transpose":
            return [x.permute(a.get("perm", list(range(x.dim()))[::-1])).contiguous()]
        if op == "LeakyRelu":
            return [F.leaky_relu(x, a.get("alpha", 0.01))]
        if op == "PRelu":
            slope = self._val(node.input[1], env)
            return [torch.where(x >= 0, x, x * slope)]
        if op == "HardSigmoid":
            return [torch.clamp(a.get("alpha", 0.2) * x + a.get("beta", 0.5), 0, 1)]
        if op == "Softmax":
            return [torch.softmax(x, a.get("axis", -1))]
        if op == "ReduceMean":
            return [x.mean(dim=a.get("axes"), keepdim=bool(a.get("keepdims", 1)))]
        if op == "BatchNormalization":
            s, b, m, v = (self._val(i, env) for i in node.input[1:5])
            return [F.batch_norm(x, m, v, s, b, False, 0.0, a.get("epsilon", 1e-5))]
        if op == "QuantizeLinear":
            return [self._quantize(node, x, env)]
        if op == "DequantizeLinear":
            if node.input[0] in self._fused_q:
                return [x]  # K5 already produced the dequantised values
            scale = self._val(node.input[1], env)
            zp = self._val(node.input[2], env).float() if len(node.input) > 2 else 0.
            if scale.numel() > 1:
                shape = [1] * x.dim()
                shape[a.get("axis", 1)] = -1
                scale, zp = scale.reshape(shape), (zp.reshape(shape) if torch.is_tensor(zp) else zp)
            return [(x.float() - zp) * scale]
        raise NotImplementedError(f"engine: unsupported op {op} ({node.name})")

    def _quantize(self, node, x, env):
        """QuantizeLinear immediately undone by its DequantizeLinear (the only pattern
        quant_graph emits, quantize.py:209-231): one fused K5 pass; the `_q` tensor holds
        the already-dequantised float values."""
        scale = self._val(node.input[1], env).reshape(-1)
        zp, unsigned = None, False
        if len(node.input) > 2 and node.input[2]:
            zp, unsigned = self.zero_points[node.input[2]]
        qlo, qhi = (0, 255) if unsigned else (-128, 127)
        axis = node.attrs.get("axis", 1) if scale.numel() > 1 else None
        x = x if x.is_contiguous() else x.contiguous()
        self._fused_q.add(node.output[0])
        return K.fakequant(x, scale, zp, qlo, qhi, axis=axis)

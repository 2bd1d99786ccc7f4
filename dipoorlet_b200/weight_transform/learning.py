"""The learned-rounding optimisation loop shared by adaround and brecq/qdrop
(dipoorlet/weight_transform/adaround.py:119-144, brecq.py:158-200).

Reference: torch autograd + torch.optim.Adam under DistributedDataParallel, one Python
iteration = dozens of small kernels. Here: a fixed launch sequence per iteration, no
autograd graph, alpha/m/v updated in place by one fused kernel, gradient averaging over
ranks by one NCCL all-reduce of dL/dW per layer (DDP semantics)."""
import os

import numpy as np
import torch

from .. import dist_helper
from .. import kernels as K
from ..utils import logger


_PEER_POOL = {}   # device index -> (symmetric float32 buffer, its rendezvous handle)
TIMINGS = None    # set to a list by a benchmark: one (loop_start, loop_end, iterations, graphs_used) of CUDA events per call


def peer_layout(sizes, words=64):
    """Float offsets of the per-layer regions of the peer buffer: [(words_offset, n, slot)], total floats.
    A region is `words` arrival words followed by two gradient slots of `slot` = n rounded up to 4 floats, so
    that every slot starts on a 16-byte boundary (the step kernel reads float4)."""
    layout, need = [], 0
    for n in sizes:
        slot = (int(n) + 3) & ~3
        layout.append((need, int(n), slot))
        need += words + 2 * slot
    return layout, need


def peer_offsets(layout, li, epoch, words=64):
    """-> (float offset of layer li's gradient slot for `epoch`, float offset of its arrival words)."""
    off, _, slot = layout[li]
    return off + words + (epoch & 1) * slot, off


class PeerGradients:
    """dL/dW of every layer of the block in a buffer that all ranks have mapped over NVLink (torch symmetric
    memory = CUDA IPC), so that the step kernel can read the world's gradients itself instead of waiting for
    an NCCL all-reduce (K.adaround_step_peer). Per layer: two gradient slots used alternately
    and `world` arrival words. One buffer per process, grown when a block needs more, so that the IPC
    rendezvous is paid a few times per run and not per layer.
    Opt-in: DPL_PEER_ALLREDUCE=1, world > 1 on one NVLink domain."""

    WORDS = 64     # floats reserved in front of a layer's slots for its arrival words (world <= 8 used)

    def __init__(self, layers, dev):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist_helper.get_world_size(), dist_helper.get_rank()
        if self.world > 8:
            raise RuntimeError("DPL_PEER_ALLREDUCE supports at most 8 ranks (one NVLink domain)")
        self.epoch = 0
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        self.layout, need = peer_layout([layer.round_mask.numel() for layer in layers], self.WORDS)
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        pooled = _PEER_POOL.get(key)
        if pooled is None or pooled[0].numel() < need:
            # every rank sees the same layers, so every rank grows at the same call (symmetric allocation)
            buf = symm.empty(max(need + need // 2, 1 << 20), dtype=torch.float32, device=dev)
            pooled = (buf, symm.rendezvous(buf, torch.distributed.group.WORLD))
            _PEER_POOL[key] = pooled
        self.buf, self.hdl = pooled
        # the previous block's last step kernels may still be reading this buffer on a slower rank and writing
        # their arrival into ours: all ranks drain first, then the words are cleared, then everybody starts
        torch.cuda.synchronize(dev)
        self.hdl.barrier()
        for off, _, _ in self.layout:
            self.buf[off:off + self.WORDS].zero_()
        torch.cuda.synchronize(dev)
        self.hdl.barrier()

    def next_epoch(self):
        self.epoch += 1

    def slot(self, li):
        """This rank's gradient slot of layer `li` for the current epoch (flat float32 view)."""
        start, _ = peer_offsets(self.layout, li, self.epoch, self.WORDS)
        return self.buf[start:start + self.layout[li][1]]

    def pointers(self, li):
        """-> (gradient slot of the current epoch, arrival words) on every rank, as device addresses."""
        g, w = peer_offsets(self.layout, li, self.epoch, self.WORDS)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        inside = self.buf.data_ptr() - ptrs[self.rank]     # the tensor's offset in the (symmetric) allocation
        return [p + inside + g * 4 for p in ptrs], [p + inside + w * 4 for p in ptrs]

    def check(self):
        if int(self.error.item()):
            raise RuntimeError("peer all-reduce: a rank did not arrive within the kernel's time limit")


def _seed(base, it, layer):
    return (base * 1000003 + it * 131 + layer * 7 + 12345) & (2 ** 63 - 1)


def learning_round_mask(layers, q_in, tgt, reg, batch_size, max_epoch, fp_in=None, drop=False,
                        log_every=50, head="", seed=0):
    """layers: [AdaQLayer] (1 for adaround, <= 3 for a brecq block) applied in sequence.
    q_in / fp_in: block input from the quantised / fp graph, tgt: fp block output (after the
    trailing Relu when there is one), all float32 CUDA tensors [n, ...] resident in HBM.
    Returns the last mini-batch loss (as the reference logs it)."""
    # Numerics of the re-evaluation: the reference's F.conv2d runs with torch's default
    # cudnn.allow_tf32 = True and F.linear in true fp32. Here every convolution and the Gemm
    # layers run on libdpl_b200's single-pass TF32 tcgen05 kernels (depthwise / stem forward: exact fp32);
    # DPL_RECON_TF32=0 switches those kernels off and forces true fp32 on the library path (torch / cuDNN) -
    # a debugging aid for the update rule, used by tests/test_gpu_quant_kernels.py.
    tf32 = os.environ.get("DPL_RECON_TF32", "1") != "0"
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, os.environ.get("DPL_TCGEN05"))
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    if not tf32:
        os.environ["DPL_TCGEN05"] = "0"
    try:
        return _learn(layers, q_in, tgt, reg, batch_size, max_epoch, fp_in, drop, log_every, head, seed)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("DPL_TCGEN05", None)
        else:
            os.environ["DPL_TCGEN05"] = saved[2]


def _iteration(layers, x, t, reg_alpha, world, loss_acc, sched, seeds, peer=None):
    """One optimisation step on the mini-batch (x, t). Every per-iteration scalar (beta, Adam bias
    corrections, mask seeds) is read from device memory (`sched`, `seeds`, filled by the
    schedule kernel), so the same launch sequence serves every iteration — eagerly or as a
    replayed CUDA graph."""
    last = len(layers) - 1
    overlap = os.environ.get("DPL_OVERLAP_ALLREDUCE", "1") != "0"
    pending = []
    acts, outs, cfgs = [x], [], []
    for li, layer in enumerate(layers):
        w = layer.quant_weight(soft=True)
        o = layer.dense_forward(acts[-1], w)
        outs.append(o)
        cfg = layer.act_cfg(0)
        cfg["seed_dev"] = seeds[li:li + 1]
        cfgs.append(cfg)
        if li < last:
            acts.append(K.recon_act(o, **cfg))
    o = outs[-1]
    loss_acc.zero_()
    inv_count = float(o.shape[1]) / float(o.numel())     # sum over channels, mean over the rest
    go = K.recon_loss(o, t, inv_count, loss_acc, **cfgs[-1])
    for li in range(last, -1, -1):
        layer = layers[li]
        gx, gw = layer.dense_backward(acts[li], layer.w_soft, go, need_dx=li > 0)
        if peer is not None:               # the step kernel reads every rank's dL/dW over NVLink itself
            peer.slot(li).copy_(gw.reshape(-1))
            grads, words = peer.pointers(li)
            K.adaround_step_peer(grads, words, peer.rank, peer.epoch, layer.wfloor, layer.scale, layer.q_min,
                                 layer.q_max, 0.0, layer.round_mask, layer.m, layer.v, 1, reg_alpha=reg_alpha,
                                 sched=sched, error=peer.error)
        elif world > 1 and overlap:
            # SUM on NCCL's stream (the 1/world is folded into the step) while the earlier layers' backward runs on
            # ours: the <= 2.4 MB reductions are latency bound, so hide them instead of waiting three times
            pending.append((layer, gw, torch.distributed.all_reduce(gw, async_op=True)))
        else:
            if world > 1:
                torch.distributed.all_reduce(gw)   # SUM; the 1/world is folded into the step
            K.adaround_step(gw, layer.wfloor, layer.scale, layer.q_min, layer.q_max, 0.0, layer.round_mask,
                            layer.m, layer.v, 1, reg_alpha=reg_alpha, grad_scale=1.0 / world, sched=sched)
        if li > 0:
            go = K.recon_act_bwd(outs[li - 1], gx, **cfgs[li - 1])
    for layer, gw, work in pending:
        work.wait()                                # our stream waits for NCCL's, the host does not block
        K.adaround_step(gw, layer.wfloor, layer.scale, layer.q_min, layer.q_max, 0.0, layer.round_mask,
                        layer.m, layer.v, 1, reg_alpha=reg_alpha, grad_scale=1.0 / world, sched=sched)


def _learn(layers, q_in, tgt, reg, batch_size, max_epoch, fp_in, drop, log_every, head, seed):
    n = q_in.shape[0]
    n_batches = int(np.ceil(n / batch_size))
    world = dist_helper.get_world_size()
    rank0 = dist_helper.get_rank() == 0
    dev = q_in.device
    loss_acc = torch.zeros(1, dtype=torch.float64, device=dev)
    d_iter = torch.zeros(1, dtype=torch.int32, device=dev)
    sched = torch.zeros(4, dtype=torch.float32, device=dev)
    seeds = torch.zeros(len(layers) + 1, dtype=torch.int64, device=dev)
    t_max = float(reg.temp_anneal.t_max)
    ratio = 0.5 if drop else 1.0
    x_all = q_in if ratio >= 1.0 else torch.empty_like(q_in)

    peer = None
    if world > 1 and os.environ.get("DPL_PEER_ALLREDUCE", "0") == "1":
        peer = PeerGradients(layers, dev)

    def schedule():
        K.recon_schedule(d_iter, sched, seeds, t_max, seed_base=seed)

    # ---- one CUDA graph per mini-batch index, replayed every epoch -----------------------------------------
    # An iteration is ~15 launches per layer (weight build, re-layout, staging copies, contraction forward /
    # weight / data gradient, epilogues, fused step) and takes 0.1 - 0.5 ms of GPU time: issued eagerly from
    # Python it is launch bound (measured: 0.6 - 1.2 ms per ResNet-50 block iteration,
    # profiles/r2_finetune_*). Every per-iteration scalar lives in device memory (schedule kernel), the
    # mini-batches are fixed slices of buffers that stay in place for the whole run (x_all is rewritten in
    # place by the per-epoch QDrop mix), so the iteration on slice idx is captured ONCE, reading its slice
    # directly (no static-input copies), and replayed max_epoch times. The graphs share one memory pool.
    # Measured (profiles/r2_finetune_eager_vs_graph.json): the eager loop is GPU bound - the launches queue
    # ahead of the kernels - and replay gains 5 - 8 % per iteration, while capture costs about a second per
    # block (private pool allocation + instantiation). "auto": graphs when every one of them is replayed at
    # least 64 times (the CLI default --ada_epoch 5000 is); DPL_CUDA_GRAPH=0 / 1 force eager / graphs.
    # The peer (NVLink) gradient path passes a host-side epoch to its kernel and stays eager; NCCL inside a
    # captured iteration is opt-in (DPL_CUDA_GRAPH_NCCL=1).
    mode = os.environ.get("DPL_CUDA_GRAPH", "auto")
    nccl_ok = world == 1 or os.environ.get("DPL_CUDA_GRAPH_NCCL", "0") == "1"
    use_graph = (mode == "1" or (mode == "auto" and max_epoch >= 64)) and peer is None and nccl_ok and max_epoch > 1
    graphs = {}
    if use_graph:
        state = [(l.round_mask.clone(), l.m.clone(), l.v.clone()) for l in layers]
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # warm-up outside capture (lazy attributes, scratch buffers)
                if ratio < 1.0:
                    x_all.copy_(q_in)
                for _ in range(2):
                    schedule()
                    _iteration(layers, x_all[:batch_size], tgt[:batch_size], reg.alpha, world, loss_acc, sched,
                               seeds)
            torch.cuda.current_stream(dev).wait_stream(side)
            pool = None
            for idx in range(n_batches):
                st, ed = idx * batch_size, min((idx + 1) * batch_size, n)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    schedule()
                    _iteration(layers, x_all[st:ed], tgt[st:ed], reg.alpha, world, loss_acc, sched, seeds)
                pool = g.pool()
                graphs[idx] = g
        except Exception as e:   # capture not possible (e.g. an op that syncs): stay eager
            logger.warning("CUDA graph capture of the rounding loop failed (%s); running eagerly" % (e,))
            graphs = {}
        # the warm-up iterations must not count: restore alpha, Adam state and the iteration counter
        for l, (a, m, v) in zip(layers, state):
            l.round_mask.copy_(a)
            l.m.copy_(m)
            l.v.copy_(v)
        d_iter.zero_()

    if TIMINGS is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    for epoch in range(max_epoch):
        if ratio < 1.0:  # QDrop: a fresh Bernoulli mix of quantised and fp block inputs per epoch
            K.mix_drop(q_in, fp_in, ratio, _seed(seed, epoch, 991), out=x_all)
        for idx in range(n_batches):
            g = graphs.get(idx)
            if g is not None:
                g.replay()
            else:
                st, ed = idx * batch_size, min((idx + 1) * batch_size, n)
                schedule()
                if peer is not None:
                    peer.next_epoch()
                _iteration(layers, x_all[st:ed], tgt[st:ed], reg.alpha, world, loss_acc, sched, seeds, peer)
        if epoch % log_every == 0 and rank0:
            logger.info("Epoch: {:<5} L2 Loss: {:>10.3f} Beta: {:>3.3f}".format(
                epoch, float(loss_acc.item()), float(sched[0].item())))
    if TIMINGS is not None:
        ev1.record()
        TIMINGS.append((ev0, ev1, max_epoch * n_batches, bool(graphs)))
    loss = float(loss_acc.item()) if max_epoch > 0 else float("nan")
    if peer is not None:
        peer.check()
    reg.beta = float(sched[0].item()) if max_epoch > 0 else reg.beta
    if rank0:
        for layer in layers:
            h = reg.rectified_sigmoid(layer.round_mask)
            ceil_n = int((h + 1e-4 >= 1.0).sum().item())
            floor_n = int((h <= 1e-4).sum().item())
            total = h.numel()
            logger.info("{}Loss: {:>5.3f} Ceil: {:>5} Floor: {:>5} Total: {:>5} Ratio: {:>.3f}".format(
                head, loss, ceil_n, floor_n, total, (ceil_n + floor_n) / total))
    return loss

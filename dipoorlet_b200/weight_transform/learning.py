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


def _seed(base, it, layer):
    return (base * 1000003 + it * 131 + layer * 7 + 12345) & (2 ** 63 - 1)


def learning_round_mask(layers, q_in, tgt, reg, batch_size, max_epoch, fp_in=None, drop=False,
                        log_every=50, head="", seed=0):
    """layers: [AdaQLayer] (1 for adaround, <= 3 for a brecq block) applied in sequence.
    q_in / fp_in: block input from the quantised / fp graph, tgt: fp block output (after the
    trailing Relu when there is one), all float32 CUDA tensors [n, ...] resident in HBM.
    Returns the last mini-batch loss (as the reference logs it)."""
    # Numerics of the re-evaluation: the reference's F.conv2d runs with torch's default
    # cudnn.allow_tf32 = True and F.linear in true fp32. Same here: TF32 tensor
    # cores for the convolutions (tcgen05 tile for 1x1, cuDNN for the rest), fp32 for Gemm on
    # the library path. DPL_RECON_TF32=0 forces fp32 everywhere (and the tcgen05 tile off).
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


def _learn(layers, q_in, tgt, reg, batch_size, max_epoch, fp_in, drop, log_every, head, seed):
    n = q_in.shape[0]
    n_batches = int(np.ceil(n / batch_size))
    world = dist_helper.get_world_size()
    rank0 = dist_helper.get_rank() == 0
    loss_acc = torch.zeros(1, dtype=torch.float64, device=q_in.device)
    ratio = 0.5 if drop else 1.0
    cur_iter = 0
    last = len(layers) - 1
    for epoch in range(max_epoch):
        x_all = q_in
        if ratio < 1.0:  # QDrop: a fresh Bernoulli mix of quantised and fp block inputs per epoch
            x_all = K.mix_drop(q_in, fp_in, ratio, _seed(seed, epoch, 991))
        for idx in range(n_batches):
            st, ed = idx * batch_size, min((idx + 1) * batch_size, n)
            beta = reg.update(cur_iter)
            acts, outs, cfgs = [x_all[st:ed]], [], []
            for li, layer in enumerate(layers):
                w = layer.quant_weight(soft=True)
                o = layer.dense_forward(acts[-1], w)
                outs.append(o)
                cfgs.append(layer.act_cfg(_seed(seed, cur_iter, li)))
                if li < last:
                    acts.append(K.recon_act(o, **cfgs[-1]))
            o = outs[-1]
            loss_acc.zero_()
            inv_count = float(o.shape[1]) / float(o.numel())     # sum over channels, mean over the rest
            go = K.recon_loss(o, tgt[st:ed], inv_count, loss_acc, **cfgs[-1])
            for li in range(last, -1, -1):
                layer = layers[li]
                gx, gw = layer.dense_backward(acts[li], layer.w_soft, go, need_dx=li > 0)
                if world > 1:
                    torch.distributed.all_reduce(gw)   # SUM; the 1/world is folded into the step
                K.adaround_step(gw, layer.wfloor, layer.scale, layer.q_min, layer.q_max, beta,
                                layer.round_mask, layer.m, layer.v, cur_iter + 1, reg_alpha=reg.alpha,
                                grad_scale=1.0 / world)
                if li > 0:
                    go = K.recon_act_bwd(outs[li - 1], gx, **cfgs[li - 1])
            cur_iter += 1
        if epoch % log_every == 0 and rank0:
            logger.info("Epoch: {:<5} L2 Loss: {:>10.3f} Beta: {:>3.3f}".format(
                epoch, float(loss_acc.item()), reg.beta))
    loss = float(loss_acc.item()) if max_epoch > 0 else float("nan")
    if rank0:
        for layer in layers:
            h = reg.rectified_sigmoid(layer.round_mask)
            ceil_n = int((h + 1e-4 >= 1.0).sum().item())
            floor_n = int((h <= 1e-4).sum().item())
            total = h.numel()
            logger.info("{}Loss: {:>5.3f} Ceil: {:>5} Floor: {:>5} Total: {:>5} Ratio: {:>.3f}".format(
                head, loss, ceil_n, floor_n, total, (ceil_n + floor_n) / total))
    return loss

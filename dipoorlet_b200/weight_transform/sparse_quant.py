"""`--sparse` (dipoorlet/weight_transform/sparse_quant.py:19-130, sparse_quant_layer.py:9-175): per learnable layer,
finetune the float weight so that its PRUNED and QUANTISED version reproduces the fp layer output on the
quantised input, then deploy prune(quant(w)). Replaces adaround / brecq when given (weight_trans_base.py:55-66).

  pruning  `unstruction`: zero the int(rate * numel) smallest |w| (threshold = the largest of them, strict >);
           `nv24`: in every group of 4 along the input-channel axis (weights viewed O,H,W,I; Gemm rows as they
           lie), zero the 2 smallest |w|.
  quant    round(w / s) with a straight-through gradient, clamped per channel only (the per-tensor clamp of the
           reference discards its result, sparse_quant_layer.py:22-23 — mirrored), times s.
  loss     sum over channels, mean over the rest, of (layer(x_q) - fp_out)^2; Relu applied to both when one follows.
  update   SGD(lr 1e-3, momentum 0.9, weight decay 1e-4) on the weight ONLY — the bias is in no optimizer
           (sparse_quant.py:103) — cosine-annealed per epoch over --ada_epoch; gradients averaged over ranks.

SURVEY.md §8 f4: not the measured path and no BASELINE.json config; the loop runs on the device through torch's
autograd and optimizer exactly as the reference's does (no kernel of libdpl_b200 beyond the forward that fills the
activation caches). Pinned against the reference's own run: tests/golden/*/wt_sparse_*.npz."""
import copy

import numpy as np
import torch
import torch.nn.functional as F

from .. import dist_helper
from ..forward_net import ActivationCache
from ..platform_settings import platform_setting_table
from ..quantize import quant_graph
from ..utils import logger
from .adaround import quantised_input_name, shard
from .utils import LEARNABLE_LAYER_TYPES, follow_relu, get_quant_tensor, update_weight
from .weight_equalization import node_has_equalized


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        return v.round()

    @staticmethod
    def backward(ctx, g):
        return g


def keep_mask(weight, pattern, rate):
    """1.0 where the weight survives pruning (sparse_quant_layer.py:29-59)."""
    mag = weight.detach().abs()
    if pattern == "unstruction":
        n_prune = int(rate * mag.numel())
        cut = mag.min() - 1 if n_prune == 0 else torch.topk(mag.reshape(-1), n_prune, largest=False)[0].max()
        return (mag > cut).to(mag.dtype)
    if pattern == "nv24":
        grouped = mag.permute(0, 2, 3, 1) if mag.dim() == 4 else mag
        flat = grouped.reshape(-1, 4)
        drop = torch.argsort(flat, dim=1)[:, :2]
        mask = torch.ones_like(flat).scatter_(1, drop, 0).reshape(grouped.shape)
        return mask.permute(0, 3, 1, 2) if mag.dim() == 4 else mask
    raise ValueError(f"unknown sparse pattern {pattern!r}")


def prune_quant(weight, scale, q_min, q_max, per_channel, pattern, rate):
    q = _RoundSTE.apply(weight * keep_mask(weight, pattern, rate) / scale)
    if per_channel:
        q = torch.min(torch.max(q, q_min), q_max)
    return q * scale


class SparseQLayer:
    """One Conv / Gemm / ConvTranspose of the ONNX graph as a differentiable function of its float weight.
    ConvTranspose weights are held transposed ([C_out/g, C_in, ...]) as the reference holds them."""

    def __init__(self, node, weight, bias, scale, q_min, q_max, per_channel, relu_flag, pattern, rate, device):
        self.op, self.attrs = node.op_type, node.attrs
        w = torch.from_numpy(np.ascontiguousarray(weight)).to(device)
        self.weight = (w.transpose(0, 1) if self.op == "ConvTranspose" else w).clone().requires_grad_(True)
        self.bias = None if bias is None else torch.from_numpy(np.ascontiguousarray(bias)).to(device)
        self.q = (scale, q_min, q_max, per_channel, pattern, rate)
        self.relu_flag = relu_flag

    def deployed_weight(self):
        with torch.no_grad():
            w = prune_quant(self.weight, *self.q)
            return (w.transpose(0, 1) if self.op == "ConvTranspose" else w).contiguous().cpu().numpy()

    def __call__(self, x):
        w, a = prune_quant(self.weight, *self.q), self.attrs
        if self.op == "Gemm":
            y = F.linear(x, w, self.bias)
        else:
            nd = w.dim() - 2
            geometry = dict(stride=a.get("strides", [1] * nd), padding=list(a.get("pads", [0] * 2 * nd))[:nd],
                            dilation=a.get("dilations", [1] * nd), groups=a.get("group", 1))
            if self.op == "Conv":
                y = F.conv2d(x, w, self.bias, **geometry)
            else:
                y = F.conv_transpose2d(x, w.transpose(0, 1), self.bias, output_padding=a.get("output_padding", 0),
                                       **geometry)
        return F.relu(y) if self.relu_flag else y


def learning_sparse_quant(x_all, target, layer, batch_size, max_epoch):
    """sparse_quant.py:102-130. -> the last mini-batch loss."""
    opt = torch.optim.SGD([layer.weight], lr=0.001, momentum=0.9, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=opt, T_max=max_epoch)
    world, loss = dist_helper.get_world_size(), None
    for epoch in range(max_epoch):
        for st in range(0, x_all.shape[0], batch_size):
            out = layer(x_all[st:st + batch_size])
            loss = (out - target[st:st + batch_size]).pow(2.0).sum(1).mean()
            opt.zero_grad()
            loss.backward()
            if world > 1:                      # DistributedDataParallel semantics: the mean gradient over ranks
                torch.distributed.all_reduce(layer.weight.grad)
                layer.weight.grad /= world
            opt.step()
        sched.step()
        if epoch % 50 == 0 and dist_helper.get_rank() == 0:
            logger.info("Epoch: {:<4} L2 Loss: {:>10.6f}, LR: {:>10.6f}".format(epoch, float(loss.detach()), sched.get_last_lr()[0]))
    if dist_helper.get_rank() == 0 and loss is not None:
        logger.info("Loss: {:>10.6f}".format(float(loss.detach())))
    return None if loss is None else float(loss.detach())


def sparse_quant(graph_ori, graph, act_clip_val, weight_clip_val, args):
    dist_helper.barrier()
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_sq = copy.deepcopy(graph)
    rank_st, rank_ed, _ = shard(args)
    fp_cache = ActivationCache(graph_ori, args, rank_st, rank_ed)
    graph_q, _ = quant_graph(graph_sq, copy.deepcopy(clip_val), args)
    q_cache = ActivationCache(graph_q, args, rank_st, rank_ed)
    qw_param = platform_setting_table[args.deploy]['qw_params']
    per_channel = bool(qw_param.get('per_channel'))
    for node in graph_ori.graph.node:
        if node.name in args.skip_layers or node.op_type not in LEARNABLE_LAYER_TYPES:
            continue
        if getattr(args, "we", False) and node_has_equalized(graph, node):
            continue                         # an equalised pair cannot be mimicked layer by layer (sparse_quant.py:37-38)
        if dist_helper.get_rank() == 0:
            logger.info("sparse_quant for: {}".format(node.name))
        q_in = q_cache[quantised_input_name(graph_q, node.input[0])]
        fp_out = fp_cache[node.output[0]]
        weight = graph_sq.get_initializer(node.input[1])
        bias = graph_sq.get_initializer(node.input[2]) if len(node.input) == 3 else None
        wshape = list(weight.shape)
        if node.op_type == 'ConvTranspose':
            wshape[0], wshape[1] = wshape[1], wshape[0]
        scale, q_min, q_max = get_quant_tensor(wshape, qw_param, copy.deepcopy(clip_val[node.input[1]]), q_in.device)
        relu_flag = follow_relu(graph, node)
        target = torch.relu(fp_out) if relu_flag else fp_out
        layer = SparseQLayer(node, weight, bias, scale, q_min, q_max, per_channel, relu_flag,
                             args.pattern, args.sparse_rate, q_in.device)
        # the reference's F.conv2d runs with torch's default cudnn.allow_tf32 = True; the engine
        # that filled the caches switched it off for its fp32 forward
        saved = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            with torch.enable_grad():
                learning_sparse_quant(q_in, target, layer, args.ada_bs, args.ada_epoch)
        finally:
            torch.backends.cudnn.allow_tf32 = saved
        new_weight = layer.deployed_weight()
        update_weight(graph_sq, new_weight, node.input[1])
        update_weight(graph_q, new_weight, node.input[1])
        q_cache.update_initializers([node.input[1]], node)
        fp_cache.drop([node.output[0]])
        del layer, q_in, fp_out, target
    graph_sq.update_model()
    if dist_helper.get_rank() == 0:
        graph_sq.save_onnx_model('sparse_quant')
    return graph_sq

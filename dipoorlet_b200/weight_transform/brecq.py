"""BRECQ / QDrop: block-wise learned rounding (dipoorlet/weight_transform/brecq.py:20-155)."""
import copy

import numpy as np
import torch

from .. import dist_helper
from ..forward_net import ActivationCache
from ..platform_settings import platform_setting_table
from ..quantize import quant_graph
from ..utils import logger
from .ada_quant_layer import AdaQLayer, adaround_reg
from .adaround import quantised_input_name, shard
from .learning import learning_round_mask
from .weight_equalization import node_has_equalized
from .utils import (LEARNABLE_LAYER_TYPES, follow_relu, following_relu, get_block_from_first,
                    get_quant_tensor, update_weight)


def brecq(graph_ori, graph, act_clip_val, weight_clip_val, args):
    dist_helper.barrier()
    rank_st, rank_ed, num_per_rank = shard(args)
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_brecq = copy.deepcopy(graph)
    fp_cache = ActivationCache(graph_ori, args, rank_st, rank_ed)
    graph_q, _ = quant_graph(graph_brecq, copy.deepcopy(clip_val), args)
    q_cache = ActivationCache(graph_q, args, rank_st, rank_ed)
    qw_param = platform_setting_table[args.deploy]['qw_params']
    if qw_param.get('per_channel') and not qw_param.get('symmetric'):
        # the clamp limits then differ per channel (they depend on each channel's zero point); the K6 kernels
        # take one (q_min, q_max) pair per layer, so refuse instead of clamping every channel with channel 0's
        raise NotImplementedError("learned rounding with per-channel asymmetric weight quantisation (-D %s)"
                                  % args.deploy)
    qi_param = platform_setting_table[args.deploy]['qi_params']
    head = 'Qdrop' if args.drop is True else 'Brecq'
    already = []
    for node in graph_ori.graph.node:
        if node.name in args.skip_layers or node.op_type not in LEARNABLE_LAYER_TYPES \
                or node.name in already:
            continue
        block = get_block_from_first(graph, node, args)
        # an equalised layer cannot close a block: its output differs from graph_ori's (brecq.py:38-41)
        if getattr(args, "we", False) and node_has_equalized(graph, block[-1]):
            block = block[:-1]
            if not block:       # the reference would index an empty list here; skipping is the only sane reading
                continue
        if dist_helper.get_rank() == 0:
            logger.info("{} for: {}".format(head, ' '.join(n.name for n in block)))
        already.extend(n.name for n in block)
        q_in = q_cache[quantised_input_name(graph_q, block[0].input[0])]
        fp_in = fp_cache[block[0].input[0]]
        fp_out = fp_cache[block[-1].output[0]]
        total_iter = args.ada_epoch * len(block) * np.ceil(num_per_rank / args.ada_bs)
        reg = adaround_reg(total_iter)
        layers = []
        for n in block:
            weight = graph_brecq.get_initializer(n.input[1])
            bias = graph_brecq.get_initializer(n.input[2]) if len(n.input) == 3 else None
            wshape = list(weight.shape)
            if n.op_type == 'ConvTranspose':
                wshape[0], wshape[1] = wshape[1], wshape[0]
            scale, q_min, q_max = get_quant_tensor(wshape, qw_param, copy.deepcopy(clip_val[n.input[1]]),
                                                   q_in.device)
            relu_flag = follow_relu(graph, n)
            out_node = following_relu(graph, n) if relu_flag else n
            a_scale, a_min, a_max = get_quant_tensor(graph.get_tensor_shape(out_node.output[0]), qi_param,
                                                     copy.deepcopy(clip_val[out_node.output[0]]), q_in.device)
            qi = (float(a_scale.reshape(-1)[0].item()), float(a_min.reshape(-1)[0].item()),
                  float(a_max.reshape(-1)[0].item()))
            layers.append(AdaQLayer(n, weight, bias, scale, q_min.reshape(-1)[0].item(),
                                    q_max.reshape(-1)[0].item(), relu_flag, qi=qi,
                                    acti_quant=args.acti_quant, device=q_in.device))
        target = torch.relu(fp_out) if follow_relu(graph, block[-1]) else fp_out
        learning_round_mask(layers, q_in, target, reg, args.ada_bs, args.ada_epoch * len(block),
                            fp_in=fp_in, drop=args.drop, log_every=100, head="")
        for n, layer in zip(block, layers):
            new_weight = layer.hard_weight().cpu().numpy()
            update_weight(graph_brecq, new_weight, n.input[1])
            update_weight(graph_q, new_weight, n.input[1])
            q_cache.update_initializers([n.input[1]], n)
        fp_cache.drop([block[-1].output[0]])
        del layers, q_in, fp_in, fp_out, target
    graph_brecq.update_model()
    if dist_helper.get_rank() == 0:
        graph_brecq.save_onnx_model('brecq')
    return graph_brecq

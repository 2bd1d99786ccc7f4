"""AdaRound: per-layer learned rounding (dipoorlet/weight_transform/adaround.py:19-116)."""
import copy

import numpy as np
import torch

from .. import dist_helper
from ..forward_net import ActivationCache
from ..platform_settings import platform_setting_table
from ..quantize import QUANT_NODE_NAME_LIST, quant_graph
from ..utils import logger
from .ada_quant_layer import AdaQLayer, adaround_reg
from .learning import learning_round_mask
from .utils import LEARNABLE_LAYER_TYPES, follow_relu, get_quant_tensor, update_weight
from .weight_equalization import node_has_equalized


def quantised_input_name(graph_q, tensor):
    """`<tensor>_dq` when the tensor is fake-quantised in graph_q, i.e. when its first
    consumer there is the QuantizeLinear -> DequantizeLinear chain (adaround.py:46-50)."""
    first = graph_q.get_tensor_consumer(tensor)[0]
    if isinstance(first, str):
        return tensor
    second = graph_q.get_tensor_consumer(first.output[0])[0]
    if not isinstance(second, str) and second.op_type == QUANT_NODE_NAME_LIST[-1]:
        return second.output[0]
    return tensor


def shard(args):
    """adaround.py:25-27: n // world images per rank, contiguous."""
    world = dist_helper.get_world_size()
    per = args.data_num // world
    st = dist_helper.get_rank() * per
    return st, st + per, per


def adaround(graph_ori, graph, act_clip_val, weight_clip_val, args):
    dist_helper.barrier()
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_ada = copy.deepcopy(graph)
    rank_st, rank_ed, num_per_rank = shard(args)
    fp_cache = ActivationCache(graph_ori, args, rank_st, rank_ed)
    graph_q, _ = quant_graph(graph_ada, copy.deepcopy(clip_val), args)
    q_cache = ActivationCache(graph_q, args, rank_st, rank_ed)
    qw_param = platform_setting_table[args.deploy]['qw_params']
    if qw_param.get('per_channel') and not qw_param.get('symmetric'):
        # the clamp limits then differ per channel (they depend on each channel's zero point); the K6 kernels
        # take one (q_min, q_max) pair per layer, so refuse instead of clamping every channel with channel 0's
        raise NotImplementedError("learned rounding with per-channel asymmetric weight quantisation (-D %s)"
                                  % args.deploy)
    for node in graph_ori.graph.node:
        if node.name in args.skip_layers or node.op_type not in LEARNABLE_LAYER_TYPES:
            continue
        # an equalised layer's output no longer matches graph_ori's: it cannot be mimicked (adaround.py:35-36)
        if getattr(args, "we", False) and node_has_equalized(graph, node):
            continue
        if dist_helper.get_rank() == 0:
            logger.info("Adaround for: {}".format(node.name))
        q_in = q_cache[quantised_input_name(graph_q, node.input[0])]
        fp_out = fp_cache[node.output[0]]
        weight = graph_ada.get_initializer(node.input[1])
        bias = graph_ada.get_initializer(node.input[2]) if len(node.input) == 3 else None
        wshape = list(weight.shape)
        if node.op_type == 'ConvTranspose':
            wshape[0], wshape[1] = wshape[1], wshape[0]
        scale, q_min, q_max = get_quant_tensor(wshape, qw_param, copy.deepcopy(clip_val[node.input[1]]),
                                               q_in.device)
        relu_flag = follow_relu(graph, node)
        target = torch.relu(fp_out) if relu_flag else fp_out
        total_iter = args.ada_epoch * np.ceil(num_per_rank / args.ada_bs)
        reg = adaround_reg(total_iter)
        layer = AdaQLayer(node, weight, bias, scale, q_min.reshape(-1)[0].item(), q_max.reshape(-1)[0].item(),
                          relu_flag, qi=None, acti_quant=args.acti_quant, device=q_in.device)
        learning_round_mask([layer], q_in, target, reg, args.ada_bs, args.ada_epoch, log_every=50)
        new_weight = layer.hard_weight().cpu().numpy()
        update_weight(graph_ada, new_weight, node.input[1])
        update_weight(graph_q, new_weight, node.input[1])
        q_cache.update_initializers([node.input[1]], node)
        fp_cache.drop([node.output[0]])
        del layer, q_in, fp_out, target
    graph_ada.update_model()
    if dist_helper.get_rank() == 0:
        graph_ada.save_onnx_model('adaround')
    return graph_ada

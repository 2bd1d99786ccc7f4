"""Bias correction (dipoorlet/weight_transform/bias_correction.py:9-55): for every Conv /
Gemm in graph order, bias += mean over (N, H, W) of (fp_out - q_out), each layer seeing the
already corrected layers before it. The reduction is K7a on the device over tensors that
stay in HBM; the Q/DQ graph is built once and only the changed bias is re-uploaded (the
reference deep-copies and re-quantises the whole graph per layer)."""
import copy

import numpy as np
import torch

from .. import kernels as K
from ..forward_net import ActivationCache
from ..quantize import quant_graph
from ..utils import ONNXGraph, logger

BIAS_CORRECTION_NODE_TYPES = ['Conv', 'Gemm']


def bias_correction(graph, act_clip_val, weight_clip_val, args):
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_bc = ONNXGraph()
    graph_bc.copy_from(graph)
    graph_q, _ = quant_graph(graph_bc, copy.deepcopy(clip_val), args)
    fp_cache = ActivationCache(graph, args)         # rank 0 processes all N images (bias_correction.py:40)
    q_cache = ActivationCache(graph_q, args)
    n_img = args.data_num
    for node in graph.graph.node:
        if node.op_type not in BIAS_CORRECTION_NODE_TYPES:
            continue
        logger.info("Update bias for node: {}".format(node.name))
        out = node.output[0]
        fp_out, q_out = fp_cache[out], q_cache[out]
        channels = fp_out.shape[1]
        acc = torch.zeros(channels, dtype=torch.float64, device=fp_out.device)
        K.channel_sumdiff(fp_out, q_out, channels, acc)
        count = fp_out.numel() // channels
        bias_diff = (acc / count).to(torch.float32).cpu().numpy()
        if len(node.input) > 2:
            name = node.input[2]
            new_bias = (graph_bc.get_initializer(name) + bias_diff).astype(np.float32)
        else:
            name = node.name + '_bias'
            new_bias = bias_diff.astype(np.float32)
            for g in (graph_bc, graph_q):
                for n in g.graph.node:
                    if n.name == node.name:
                        n.input.append(name)
                g.tensor_name_shape_map[name] = list(new_bias.shape)
                g.input.append(name)
        for g in (graph_bc, graph_q):
            g.set_initializer(name, new_bias)
        q_node = next(n for n in graph_q.graph.node if n.name == node.name)
        q_cache.update_initializers([name], q_node)   # drops q_out and everything downstream of it
        del fp_out, q_out                             # fp tensors stay cached (LRU): the next layer starts there
    graph_bc.update_model()
    graph_bc.save_onnx_model('update_bias_model')
    return graph_bc

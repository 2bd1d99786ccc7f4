"""`--update_bn` (dipoorlet/weight_transform/update_bn.py:13-50): re-estimate the running statistics of every
BatchNormalization node from the QUANTISED model's activations, node by node in graph order, each node seeing
the nodes before it already updated. Per node, over the images in order,
    mean <- 0.9 mean + 0.1 mean_{N,H,W}(x_i),   var <- 0.9 var + 0.1 std_{N,H,W}(x_i)
(the reference feeds np.std, not the variance, into the variance slot, update_bn.py:17 — mirrored, the saved
model must be the one the reference saves). Writes update_bn_model.onnx.
SURVEY.md §8 f4: not on the measured hot path. The activations stay in HBM (ActivationCache); the per-image,
per-channel moments are one device reduction per node (torch.var_mean, float32) and the 0.9 / 0.1 recurrence
over images runs on the host in float32 exactly as the reference writes it. The Q/DQ graph is built once and
only the two changed initializers are re-uploaded (the reference re-quantises the graph per node).
The reference also re-calibrates inside this function and discards the result (update_bn.py:49-50 vs
weight_trans_base.py:42); the caller's own re-calibration is the one that counts, so that one is kept."""
import copy

import numpy as np
import torch

from ..forward_net import ActivationCache
from ..quantize import quant_graph
from ..utils import ONNXGraph, logger

MOMENTUM = 0.9


def channel_moments(x):
    """x [n, C, ...] on the device -> (mean, std) float32 NumPy [n, C]: per image, per channel, population std."""
    var, mean = torch.var_mean(x.reshape(x.shape[0], x.shape[1], -1), dim=2, unbiased=False)
    return mean.cpu().numpy().astype(np.float32), var.sqrt().cpu().numpy().astype(np.float32)


def fold_running_stats(running_mean, running_var, means, stds, momentum=MOMENTUM):
    """The reference's recurrence (update_bn.py:15-17), float32 throughout under NumPy 2 (Python floats are weak)."""
    running_mean = np.asarray(running_mean)
    running_var = np.asarray(running_var)
    for m, s in zip(means, stds):
        running_mean = momentum * running_mean + (1.0 - momentum) * m
        running_var = momentum * running_var + (1.0 - momentum) * s
    return running_mean, running_var


def update_bn(graph, act_clip_val, weight_clip_val, args):
    clip_val = dict(act_clip_val)
    clip_val.update(weight_clip_val)
    graph_bn = ONNXGraph()
    graph_bn.copy_from(graph)
    graph_q, _ = quant_graph(graph_bn, copy.deepcopy(clip_val), args)
    q_cache = ActivationCache(graph_q, args)          # rank 0 processes all N images (update_bn.py:41)
    for node in graph_bn.graph.node:
        if node.op_type != "BatchNormalization":
            continue
        logger.info("Update BN for node: {}".format(node.name))
        q_node = next(n for n in graph_q.graph.node if n.name == node.name)
        means, stds = channel_moments(q_cache[q_node.input[0]])
        mean_name, var_name = node.input[3], node.input[4]
        new_mean, new_var = fold_running_stats(graph_bn.get_initializer(mean_name), graph_bn.get_initializer(var_name),
                                               means, stds)
        for g in (graph_bn, graph_q):
            g.set_initializer(mean_name, new_mean)
            g.set_initializer(var_name, new_var)
        q_cache.update_initializers([mean_name, var_name], q_node)   # drops this node's output and what follows it
    graph_bn.update_model()
    graph_bn.save_onnx_model('update_bn_model')
    return graph_bn

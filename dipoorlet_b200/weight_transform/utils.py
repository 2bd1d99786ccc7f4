"""Helpers of the rounding finetune (dipoorlet/weight_transform/utils.py:7-65)."""
import numpy as np
import torch

from ..quantize import get_qnode_by_param

LEARNABLE_LAYER_TYPES = ['Conv', 'Gemm', 'ConvTranspose']
__all__ = ['LEARNABLE_LAYER_TYPES', 'follow_relu', 'following_relu', 'update_weight',
           'get_quant_tensor', 'get_block_from_first']


def follow_relu(graph, node):
    """True iff the layer's only consumer is a Relu (Clip / ReLU6 does not count)."""
    nxt = graph.get_tensor_consumer(node.output[0])
    return len(nxt) == 1 and not isinstance(nxt[0], str) and nxt[0].op_type == 'Relu'


def following_relu(graph, node):
    nxt = graph.get_tensor_consumer(node.output[0])
    assert nxt[0].op_type == 'Relu'
    return nxt[0]


def update_weight(graph, weight_tensor, weight_name):
    graph.set_initializer(weight_name, weight_tensor)


def get_quant_tensor(shape, param, value_range, device=None):
    """-> (scale, q_min, q_max): float32 CUDA tensors shaped [C, 1, ...] for per-channel
    parameters, 0-d otherwise (weight_transform/utils.py:29-51)."""
    q_nodes, q_min, q_max = get_qnode_by_param(param, 'tmp', shape, value_range)
    scale = dict(q_nodes.initializer)['tmp_scale']
    dev = device if device is not None else torch.device("cuda")
    if param.get('per_channel'):
        view = [shape[0]] + [1] * (len(shape) - 1)
        mk = lambda a: torch.from_numpy(np.array(a).astype(np.float32)).view(view).to(dev)  # noqa: E731
    else:
        mk = lambda a: torch.from_numpy(np.array(a).astype(np.float32)).to(dev)  # noqa: E731
    return mk(scale), mk(q_min), mk(q_max)


def get_block_from_first(graph, node, args):
    """Chain of learnable layers joined by single-consumer edges, Relu transparent, at most
    3 (weight_transform/utils.py:54-65): a ResNet bottleneck's conv1-conv2-conv3."""
    res = [node]
    while True:
        nxt = graph.get_tensor_consumer(node.output[0])
        if len(nxt) != 1 or isinstance(nxt[0], str) or \
                nxt[0].op_type not in LEARNABLE_LAYER_TYPES + ['Relu']:
            return res
        if nxt[0].op_type != 'Relu':
            res.append(nxt[0])
            if len(res) == 3:
                return res
        node = nxt[0]

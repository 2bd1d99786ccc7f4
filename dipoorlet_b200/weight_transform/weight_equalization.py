"""Cross-layer weight range equalisation, `--we` (dipoorlet/weight_transform/weight_equalization.py:10-101).

For every Conv whose only consumer chain is Conv -> [Relu | PRelu] -> Conv (or Conv -> Conv), the
output channels of the first layer and the matching input channels of the second are rescaled by
    s = r1 / sqrt(r1 * r2),   r1 = max|W1[channel]|,  r2 = max|W2[:, channel]|   (ranges < 1e-6 count as 0,
                                                                                 s = 1 when inf / NaN)
W1[channel] /= s, b1[channel] /= s, W2[:, channel] *= s, repeated until both weights move by less than
1e-4 in Frobenius norm (the update that would move them less is NOT applied, as in the reference).
Pure host arithmetic on the weights (a few MB): no kernel involved; the re-calibration that follows it
in `weight_calibration` is the GPU path. The per-channel Python loops of the reference are evaluated
here as float32 array expressions with the same operation order, hence bit-identical weights.
"""
import numpy as np

from ..utils import ONNXGraph, logger
from .utils import update_weight


def find_successor(cur_node, graph):
    """The Conv nodes fed by `cur_node` directly or through one Relu / PRelu; [] as soon as
    any consumer is something else (weight_equalization.py:10-31)."""
    result = []
    for node in graph.get_tensor_consumer(cur_node.output[0]):
        if isinstance(node, str):
            return []
        if node.op_type in ('Relu', 'PRelu'):
            for nxt in graph.get_tensor_consumer(node.output[0]):
                if not isinstance(nxt, str) and nxt.op_type == 'Conv':
                    result.append(nxt)
                else:
                    return []
        elif node.op_type == 'Conv':
            result.append(node)
        else:
            return []
    return result


def node_has_equalized(graph, node):
    return len(find_successor(node, graph)) == 1


def _equalize_once(w1, b1, w2):
    """One sweep over all groups / channels -> (new_w1, new_b1, new_w2)."""
    num_group = w1.shape[0] // w2.shape[1]
    gi, go = w1.shape[0] // num_group, w2.shape[0] // num_group
    n_in = w2.shape[1]                                   # channels handled per group (reference: range(shape[1]))
    new_w1, new_w2 = w1.copy(), w2.copy()
    new_b1 = None if b1 is None else b1.copy()
    a1 = np.abs(w1).reshape(w1.shape[0], -1).max(axis=1)                       # [C1]
    a2 = np.abs(w2).reshape(num_group, go, n_in, -1).max(axis=(1, 3))          # [group, n_in]
    for g in range(num_group):
        r1 = a1[g * gi:g * gi + n_in].astype(np.float32).copy()
        r2 = a2[g].astype(np.float32).copy()
        r1[r1 < 1e-6] = 0.
        r2[r2 < 1e-6] = 0.
        with np.errstate(divide='ignore', invalid='ignore'):
            s = r1 / np.sqrt(r1 * r2)
        s[np.isinf(s) | np.isnan(s)] = 1.0
        s = s.astype(np.float32)
        shape1 = (n_in,) + (1,) * (w1.ndim - 1)
        new_w1[g * gi:g * gi + n_in] /= s.reshape(shape1)
        shape2 = (1, n_in) + (1,) * (w2.ndim - 2)
        new_w2[g * go:(g + 1) * go] *= s.reshape(shape2)
        if new_b1 is not None:
            new_b1[g * gi:g * gi + n_in] /= s
    return new_w1, new_b1, new_w2


def converged(cur_weight, prev_weight, threshold=1e-4):
    norm_sum = np.linalg.norm(cur_weight[0] - prev_weight[0]) + np.linalg.norm(cur_weight[1] - prev_weight[1])
    return norm_sum < threshold


def weight_equalization(graph, args):
    graph_we = ONNXGraph()
    graph_we.copy_from(graph)
    for node in graph_we.graph.node:
        if node.op_type != 'Conv':
            continue
        succ = find_successor(node, graph_we)
        if len(succ) != 1:
            continue
        nxt = succ[0]
        it = 1
        while True:
            w1 = np.asarray(graph_we.get_initializer(node.input[1]))
            b1 = np.asarray(graph_we.get_initializer(node.input[2])) if len(node.input) == 3 else None
            w2 = np.asarray(graph_we.get_initializer(nxt.input[1]))
            logger.info('Cross Layer WE: {} --- {} Groups: {} Iter: {}'.format(
                node.name, nxt.name, w1.shape[0] // w2.shape[1], it))
            new_w1, new_b1, new_w2 = _equalize_once(w1, b1, w2)
            if converged([w1, w2], [new_w1, new_w2]):
                break
            it += 1
            update_weight(graph_we, new_w1, node.input[1])
            update_weight(graph_we, new_w2, nxt.input[1])
            if new_b1 is not None:
                update_weight(graph_we, new_b1, node.input[2])
            graph_we.update_model()
    graph_we.save_onnx_model('weight_equal_model')
    return graph_we

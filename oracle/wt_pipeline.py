"""Weight-transform pipelines of the reference restated on the plain-container graph:
bias correction, adaround and brecq driven layer by layer with oracle.forward for the
activations and oracle.adaround for the optimisation. The Q/DQ graph is an INPUT (the
tests pass the one built by dipoorlet_b200.quantize.quant_graph, which
tests/test_golden_reference.py proves identical to the reference's).

Follows dipoorlet/weight_transform/bias_correction.py:9-55, adaround.py:19-116,
brecq.py:20-155, weight_transform/utils.py:11-65, quantize.py:128-143.
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F

from . import adaround as OA
from . import forward as OF

LEARNABLE = ("Conv", "Gemm", "ConvTranspose")


def _maps(model):
    prod, cons = {}, {}
    for n in model.graph.nodes:
        for o in n.output:
            prod[o] = n
        for i in n.input:
            cons.setdefault(i, []).append(n)
    return prod, cons


def _tensor_for_all(model, images, name, in_name="input"):
    outs = []
    shape = [vi.shape for vi in model.graph.inputs if vi.name == in_name][0]
    for i in range(images.shape[0]):
        feeds = {in_name: images[i].reshape(shape)}
        if name == in_name:
            outs.append(feeds[in_name])
        else:
            outs.append(OF.forward_all(model, feeds)[name])
    return np.stack(outs)


def quantised_input_name(model_q, tensor):
    """adaround.py:46-50."""
    _, cons = _maps(model_q)
    first = cons[tensor][0]
    second = cons.get(first.output[0], [None])[0]
    if second is not None and second.op_type == "DequantizeLinear":
        return second.output[0]
    return tensor


def follow_relu(model, node):
    _, cons = _maps(model)
    nxt = cons.get(node.output[0], [])
    return len(nxt) == 1 and nxt[0].op_type == "Relu"


def weight_qparams(weight_range, shape):
    """Symmetric 8-bit per-channel (trt qw_params): quantize.py:128-143 + utils.py:38-45."""
    data_max = np.max(np.abs(np.stack([np.asarray(weight_range[0]), np.asarray(weight_range[1])])), axis=0)
    scale = np.array(data_max) / ([127] * len(data_max))
    scale = np.where(scale == 0, 1., scale)
    view = [shape[0]] + [1] * (len(shape) - 1)
    mk = lambda a: torch.from_numpy(np.array(a).astype(np.float32)).view(view)  # noqa: E731
    return mk(np.array(scale.tolist(), dtype=np.float32)), mk([-127] * shape[0]), mk([127] * shape[0])


def act_qparams(act_range):
    """Symmetric 8-bit per-tensor (trt qi_params)."""
    lo, hi = np.min(act_range[0]), np.max(act_range[1])
    scale = np.array(np.max(np.abs([lo, hi]), axis=0)) / [127]
    scale = np.where(scale == 0, 1., scale)
    mk = lambda a: torch.from_numpy(np.array(a).astype(np.float32))  # noqa: E731
    return mk(np.array(scale.tolist(), dtype=np.float32)), mk([-127]), mk([127])


def _layer(model_w, node, clip_val, relu_flag, qi=None, acti_quant=False):
    w = torch.from_numpy(model_w.graph.initializers[node.input[1]].copy())
    b = torch.from_numpy(model_w.graph.initializers[node.input[2]].copy()) if len(node.input) == 3 else None
    scale, q_min, q_max = weight_qparams(clip_val[node.input[1]], list(w.shape))
    attrs = dict(node.attrs)
    return OA.Layer(node.op_type, attrs, w, b, scale, q_min, q_max, relu_flag, qi, acti_quant)


def adaround(model_fp, model_q, images, clip_val, ada_bs, ada_epoch):
    """-> {weight name: rounded weight}; model_q's weights are updated in place like
    graph_q in the reference."""
    n = images.shape[0]
    out = {}
    for node in model_fp.graph.nodes:
        if node.op_type not in LEARNABLE:
            continue
        q_in = torch.from_numpy(_tensor_for_all(model_q, images, quantised_input_name(model_q, node.input[0])))
        fp_out = torch.from_numpy(_tensor_for_all(model_fp, images, node.output[0]))
        relu_flag = follow_relu(model_fp, node)
        tgt = F.relu(fp_out) if relu_flag else fp_out
        layer = _layer(model_q, node, clip_val, relu_flag)
        total_iter = ada_epoch * np.ceil(n / ada_bs)
        OA.learn([layer], q_in.squeeze(1), tgt.squeeze(1), total_iter, ada_bs, ada_epoch)
        new_w = layer.hard_weight().numpy()
        model_q.graph.initializers[node.input[1]] = new_w
        out[node.input[1]] = new_w
    return out


def block_from_first(model, node):
    """weight_transform/utils.py:54-65."""
    _, cons = _maps(model)
    res = [node]
    while True:
        nxt = cons.get(node.output[0], [])
        if len(nxt) != 1 or nxt[0].op_type not in LEARNABLE + ("Relu",):
            return res
        if nxt[0].op_type != "Relu":
            res.append(nxt[0])
            if len(res) == 3:
                return res
        node = nxt[0]


def brecq(model_fp, model_q, images, clip_val, ada_bs, ada_epoch, drop=False, generator=None):
    n = images.shape[0]
    _, cons = _maps(model_fp)
    out, already = {}, []
    for node in model_fp.graph.nodes:
        if node.op_type not in LEARNABLE or node.name in already:
            continue
        block = block_from_first(model_fp, node)
        already.extend(b.name for b in block)
        q_in = torch.from_numpy(_tensor_for_all(model_q, images, quantised_input_name(model_q, block[0].input[0])))
        fp_in = torch.from_numpy(_tensor_for_all(model_fp, images, block[0].input[0]))
        fp_out = torch.from_numpy(_tensor_for_all(model_fp, images, block[-1].output[0]))
        total_iter = ada_epoch * len(block) * np.ceil(n / ada_bs)
        layers = []
        for b in block:
            relu_flag = follow_relu(model_fp, b)
            out_t = cons[b.output[0]][0].output[0] if relu_flag else b.output[0]
            qi = act_qparams(copy.deepcopy(clip_val[out_t]))
            layers.append(_layer(model_q, b, clip_val, relu_flag, qi, acti_quant=drop))
        tgt = F.relu(fp_out) if follow_relu(model_fp, block[-1]) else fp_out
        OA.learn(layers, q_in.squeeze(1), tgt.squeeze(1), total_iter, ada_bs, ada_epoch * len(block),
                 fp_in=fp_in.squeeze(1), drop=drop, generator=generator)
        for b, layer in zip(block, layers):
            new_w = layer.hard_weight().numpy()
            model_q.graph.initializers[b.input[1]] = new_w
            out[b.input[1]] = new_w
    return out


def bias_correction(model_fp, build_q, images):
    """bias_correction.py:9-55. build_q(model) -> Q/DQ model of the bias-corrected model so
    far (the reference re-runs quant_graph per layer). -> {bias name: corrected bias}."""
    model_bc = copy.deepcopy(model_fp)
    out = {}
    for node in model_fp.graph.nodes:
        if node.op_type not in ("Conv", "Gemm"):
            continue
        model_q = build_q(model_bc)
        fp = _tensor_for_all(model_fp, images, node.output[0])
        q = _tensor_for_all(model_q, images, node.output[0])
        bias_diff = fp - q
        axis = (0, 2, 3) if node.op_type == "Conv" else (0)
        bias_diff = np.squeeze(bias_diff, axis=1).mean(axis=axis)
        bc_node = [m for m in model_bc.graph.nodes if m.name == node.name][0]
        if len(bc_node.input) > 2:
            name = bc_node.input[2]
            model_bc.graph.initializers[name] = model_bc.graph.initializers[name] + bias_diff
        else:
            name = node.name + "_bias"
            model_bc.graph.initializers[name] = bias_diff
            bc_node.input.append(name)
        out[name] = model_bc.graph.initializers[name]
    return out

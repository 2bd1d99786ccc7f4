"""CPU restatement of the reference's forward: what `ort_session.run` with every node
output promoted to a graph output returns (dipoorlet/forward_net.py:195-216), one image
per call, fp32. ONNXRuntime itself is third-party (requirements.txt:4, unpinned) and not
installable here, so the ONNX operator semantics (opset 13) are restated with torch CPU
float32 ops; QuantizeLinear / DequantizeLinear follow the ONNX operator spec
(round-half-even, saturation to the zero-point's integer type).

Works on the plain-container graph of dipoorlet_b200.onnx_lite (nodes / initializers /
inputs / outputs) but shares no code with dipoorlet_b200.engine.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def run_node(node, ins, attrs):
    op = node.op_type
    x = ins[0] if ins else None
    if op == "Conv":
        w, b = ins[1], (ins[2] if len(ins) > 2 else None)
        nd = w.dim() - 2
        pads = attrs.get("pads", [0] * (2 * nd))
        if list(pads[:nd]) != list(pads[nd:]):
            x = F.pad(x, [p for i in reversed(range(nd)) for p in (pads[i], pads[nd + i])])
            pads = [0] * (2 * nd)
        return F.conv2d(x, w, b, attrs.get("strides", [1] * nd), list(pads[:nd]),
                        attrs.get("dilations", [1] * nd), attrs.get("group", 1))
    if op == "ConvTranspose":
        w, b = ins[1], (ins[2] if len(ins) > 2 else None)
        nd = w.dim() - 2
        pads = attrs.get("pads", [0] * (2 * nd))
        return F.conv_transpose2d(x, w, b, attrs.get("strides", [1] * nd), list(pads[:nd]),
                                  attrs.get("output_padding", [0] * nd), attrs.get("group", 1),
                                  attrs.get("dilations", [1] * nd))
    if op == "Relu":
        return F.relu(x)
    if op == "Clip":
        lo = ins[1].item() if len(ins) > 1 and ins[1] is not None else attrs.get("min")
        hi = ins[2].item() if len(ins) > 2 and ins[2] is not None else attrs.get("max")
        return torch.clamp(x, lo, hi)
    if op == "MaxPool":
        nd = x.dim() - 2
        pads = attrs.get("pads", [0] * (2 * nd))
        return F.max_pool2d(x, attrs["kernel_shape"], attrs.get("strides", [1] * nd), list(pads[:nd]),
                            attrs.get("dilations", [1] * nd), bool(attrs.get("ceil_mode", 0)))
    if op == "AveragePool":
        nd = x.dim() - 2
        pads = attrs.get("pads", [0] * (2 * nd))
        return F.avg_pool2d(x, attrs["kernel_shape"], attrs.get("strides", [1] * nd), list(pads[:nd]),
                            bool(attrs.get("ceil_mode", 0)), bool(attrs.get("count_include_pad", 0)))
    if op == "GlobalAveragePool":
        return x.mean(dim=tuple(range(2, x.dim())), keepdim=True)
    if op == "Add":
        return x + ins[1]
    if op == "Mul":
        return x * ins[1]
    if op == "Sub":
        return x - ins[1]
    if op == "Flatten":
        ax = attrs.get("axis", 1)
        return x.reshape(int(np.prod(x.shape[:ax])) if ax else 1, -1)
    if op == "Gemm":
        a = x.t() if attrs.get("transA", 0) else x
        w = ins[1].t() if attrs.get("transB", 0) else ins[1]
        y = attrs.get("alpha", 1.0) * (a @ w)
        if len(ins) > 2:
            y = y + attrs.get("beta", 1.0) * ins[2]
        return y
    if op == "Reshape":
        tgt = [int(v) for v in ins[1].tolist()]
        tgt = [x.shape[i] if v == 0 else v for i, v in enumerate(tgt)]
        return x.reshape(tgt)
    if op == "Sigmoid":
        return torch.sigmoid(x)
    if op == "Identity":
        return x
    if op == "Concat":
        return torch.cat(ins, attrs["axis"])
    if op == "BatchNormalization":
        # ONNX opset 9-14, inference: (x - mean) / sqrt(var + eps) * scale + B, per channel (axis 1)
        shape = [1, -1] + [1] * (x.dim() - 2)
        scale, bias, mean, var = (t.reshape(shape) for t in ins[1:5])
        return (x - mean) / torch.sqrt(var + attrs.get("epsilon", 1e-5)) * scale + bias
    if op == "QuantizeLinear":
        scale, zp = ins[1], (ins[2] if len(ins) > 2 else torch.zeros((), dtype=torch.uint8))
        lo, hi = (0, 255) if zp.dtype == torch.uint8 else (-128, 127)
        if scale.numel() > 1:
            shape = [1] * x.dim()
            shape[attrs.get("axis", 1)] = -1
            scale, zp = scale.reshape(shape), zp.reshape(shape)
        q = torch.clamp(torch.round(x / scale) + zp.to(torch.float32), lo, hi)
        return q.to(zp.dtype)
    if op == "DequantizeLinear":
        scale, zp = ins[1], (ins[2] if len(ins) > 2 else torch.zeros((), dtype=x.dtype))
        if scale.numel() > 1:
            shape = [1] * x.dim()
            shape[attrs.get("axis", 1)] = -1
            scale, zp = scale.reshape(shape), zp.reshape(shape)
        return (x.to(torch.int32) - zp.to(torch.int32)).to(torch.float32) * scale
    raise NotImplementedError(f"oracle forward: op {op}")


def forward_all(model, feeds, threads=None):
    """One sample. feeds: name -> ndarray with the model's declared input shape.
    Returns OrderedDict name -> ndarray of every node output, in the order the reference's
    session lists them (non-network outputs in node order, network outputs last)."""
    if threads:
        torch.set_num_threads(threads)
    g = model.graph
    env = {k: _t(v) for k, v in g.initializers.items()}
    for k, v in feeds.items():
        env[k] = _t(np.asarray(v, dtype=np.float32))
    net_out = [o.name for o in g.outputs]
    produced = []
    with torch.no_grad():
        for node in g.nodes:
            ins = [env[i] if i else None for i in node.input]
            out = run_node(node, ins, node.attrs)
            env[node.output[0]] = out
            produced.append(node.output[0])
    order = [n for n in produced if n not in net_out] + net_out
    return OrderedDict((n, env[n].numpy()) for n in order)


def blobs_for_images(model, images_by_input, n, threads=None):
    """name -> list of per-image arrays, inputs first then node outputs: the dict the
    reference's statistics loops iterate over (forward_net.py:220-235)."""
    g = model.graph
    in_names = [vi.name for vi in g.inputs if vi.name not in g.initializers]
    shapes = {vi.name: vi.shape for vi in g.inputs}
    blobs = OrderedDict()
    for idx in range(n):
        feeds = {nm: np.asarray(images_by_input[nm][idx]).reshape(shapes[nm]) for nm in in_names}
        outs = forward_all(model, feeds, threads)
        for nm in in_names:
            blobs.setdefault(nm, []).append(feeds[nm])
        for nm, v in outs.items():
            blobs.setdefault(nm, []).append(v)
    return blobs

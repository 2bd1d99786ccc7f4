"""Graph IR of the calibrated network: the stand-in for the reference's `ONNXGraph`
(dipoorlet/utils.py:22-250), which wraps an `onnx.ModelProto`. Here the model is the
plain-container form decoded by onnx_lite; `ONNXGraph` exposes the attribute surface the
registry functions touch (graph.node, network_inputs/outputs, initializer,
get_tensor_shape, get_tensor_producer/consumer, get/set_initializer, copy_from,
update_model, save_onnx_model), so plugins written against the reference keep working.

Also here, because `onnxsim` / `onnx.shape_inference` are not available either
(dipoorlet/__main__.py:98-103): `simplify()` — Constant/Identity folding and BatchNorm
folding — and an analytic shape inference for the operator set of the two model families.
"""
import copy
import os

import numpy as np

from . import onnx_lite as ol
from .platform_settings import platform_setting_table

INPUT_TOKEN = "INPUT_TOKEN"
OUTPUT_TOKEN = "OUTPUT_TOKEN"


# ------------------------------------------------------------------ shape inference
def _conv_out(size, k, s, p0, p1, d, ceil_mode=False):
    eff = d * (k - 1) + 1
    num = size + p0 + p1 - eff
    return (-(-num // s) if ceil_mode else num // s) + 1


def _pads(attrs, nd, in_hw=None, kernel=None, strides=None, dil=None):
    auto = attrs.get("auto_pad", "NOTSET")
    if auto in ("NOTSET", "", None):
        p = attrs.get("pads", [0] * (2 * nd))
        return list(p[:nd]), list(p[nd:])
    if auto == "VALID":
        return [0] * nd, [0] * nd
    lo, hi = [], []
    for i in range(nd):
        out = -(-in_hw[i] // strides[i])
        total = max((out - 1) * strides[i] + dil[i] * (kernel[i] - 1) + 1 - in_hw[i], 0)
        a, b = total // 2, total - total // 2
        if auto == "SAME_LOWER":
            a, b = b, a
        lo.append(a)
        hi.append(b)
    return lo, hi


def infer_node_shape(node, shapes, consts):
    """Output shapes (lists of ints) of one node given its input shapes."""
    op, a = node.op_type, node.attrs
    x = shapes.get(node.input[0]) if node.input else None
    if op in ("Relu", "Clip", "Sigmoid", "Identity", "LeakyRelu", "PRelu", "Abs", "Tanh",
              "HardSigmoid", "Softmax", "QuantizeLinear", "DequantizeLinear", "BatchNormalization",
              "Dropout", "Reciprocal", "Sqrt", "Exp"):
        return [list(x)]
    if op in ("Add", "Mul", "Sub", "Div"):
        y = shapes[node.input[1]]
        return [list(np.broadcast_shapes(tuple(x), tuple(y)))]
    if op in ("Conv", "MaxPool", "AveragePool"):
        nd = len(x) - 2
        if op == "Conv":
            w = shapes[node.input[1]]
            kernel = list(a.get("kernel_shape", w[2:]))
            cout = w[0]
        else:
            kernel = list(a["kernel_shape"])
            cout = x[1]
        strides = list(a.get("strides", [1] * nd))
        dil = list(a.get("dilations", [1] * nd))
        lo, hi = _pads(a, nd, x[2:], kernel, strides, dil)
        ceil_mode = bool(a.get("ceil_mode", 0))
        sp = [_conv_out(x[2 + i], kernel[i], strides[i], lo[i], hi[i], dil[i], ceil_mode)
              for i in range(nd)]
        return [[x[0], cout] + sp]
    if op == "ConvTranspose":
        w = shapes[node.input[1]]
        nd = len(x) - 2
        kernel = list(a.get("kernel_shape", w[2:]))
        strides = list(a.get("strides", [1] * nd))
        dil = list(a.get("dilations", [1] * nd))
        p = a.get("pads", [0] * (2 * nd))
        op_ = a.get("output_padding", [0] * nd)
        g = a.get("group", 1)
        sp = [(x[2 + i] - 1) * strides[i] - p[i] - p[nd + i] + dil[i] * (kernel[i] - 1) + op_[i] + 1
              for i in range(nd)]
        return [[x[0], w[1] * g] + sp]
    if op == "GlobalAveragePool":
        return [[x[0], x[1]] + [1] * (len(x) - 2)]
    if op == "Flatten":
        ax = a.get("axis", 1)
        ax = ax + len(x) if ax < 0 else ax
        return [[int(np.prod(x[:ax])) if ax else 1, int(np.prod(x[ax:]))]]
    if op == "Gemm":
        w = shapes[node.input[1]]
        m = x[1] if a.get("transA", 0) else x[0]
        n = w[0] if a.get("transB", 0) else w[1]
        return [[m, n]]
    if op == "MatMul":
        w = shapes[node.input[1]]
        return [list(x[:-1]) + [w[-1]]]
    if op == "Reshape":
        tgt = consts.get(node.input[1])
        if tgt is None:
            raise ValueError(f"Reshape {node.name}: shape input must be a constant")
        tgt = [int(v) for v in np.asarray(tgt).reshape(-1)]
        out = [x[i] if v == 0 else v for i, v in enumerate(tgt)]
        if -1 in out:
            known = int(np.prod([v for v in out if v != -1])) or 1
            out[out.index(-1)] = int(np.prod(x)) // known
        return [out]
    if op == "Concat":
        ax = a["axis"]
        ins = [shapes[i] for i in node.input]
        ax = ax + len(ins[0]) if ax < 0 else ax
        out = list(ins[0])
        out[ax] = sum(s[ax] for s in ins)
        return [out]
    if op == "Transpose":
        perm = a.get("perm", list(range(len(x)))[::-1])
        return [[x[p] for p in perm]]
    if op in ("Squeeze", "Unsqueeze"):
        axes = a.get("axes")
        if axes is None and len(node.input) > 1:
            axes = [int(v) for v in np.asarray(consts[node.input[1]]).reshape(-1)]
        out = list(x)
        if op == "Squeeze":
            axes = [ax + len(x) if ax < 0 else ax for ax in (axes or [i for i, d in enumerate(x) if d == 1])]
            out = [d for i, d in enumerate(x) if i not in axes]
        else:
            for ax in sorted(ax + len(x) + len(axes) if ax < 0 else ax for ax in axes):
                out.insert(ax, 1)
        return [out]
    if op == "ReduceMean":
        axes = a.get("axes", list(range(len(x))))
        axes = [ax + len(x) if ax < 0 else ax for ax in axes]
        keep = a.get("keepdims", 1)
        return [[(1 if i in axes else d) for i, d in enumerate(x) if keep or i not in axes]]
    raise NotImplementedError(f"shape inference: unsupported op {op} ({node.name})")


# ------------------------------------------------------------------ simplifier
def simplify(model):
    """The part of onnxsim the two model families need (dipoorlet/__main__.py:101):
    Constant -> initializer, Identity elimination, BatchNormalization folded into the
    Conv that feeds it, dead initializers dropped. Returns a new Model."""
    m = copy.deepcopy(model)
    g = m.graph
    # Constant -> initializer
    kept = []
    for n in g.nodes:
        if n.op_type == "Constant" and "value" in n.attrs:
            g.initializers[n.output[0]] = np.asarray(n.attrs["value"])
        else:
            kept.append(n)
    g.nodes = kept
    # Identity elimination (duplicate an initializer, rewire an activation)
    out_names = {o.name for o in g.outputs}
    rename = {}
    kept = []
    for n in g.nodes:
        n.input = [rename.get(i, i) for i in n.input]
        if n.op_type == "Identity" and n.output[0] not in out_names:
            src = n.input[0]
            if src in g.initializers:
                g.initializers[n.output[0]] = g.initializers[src].copy()
            else:
                rename[n.output[0]] = src
        else:
            kept.append(n)
    g.nodes = kept
    # BatchNormalization folding: Conv -> BN with the Conv output used only by the BN
    uses = {}
    for n in g.nodes:
        for i in n.input:
            uses[i] = uses.get(i, 0) + 1
    producer = {o: n for n in g.nodes for o in n.output}
    kept, dropped = [], set()
    for n in g.nodes:
        if n.op_type == "BatchNormalization":
            conv = producer.get(n.input[0])
            ok = (conv is not None and conv.op_type == "Conv" and uses.get(n.input[0], 0) == 1 and
                  n.input[0] not in out_names and all(i in g.initializers for i in n.input[1:5]) and
                  conv.input[1] in g.initializers)
            if ok:
                scale, bias, mean, var = (g.initializers[i].astype(np.float64) for i in n.input[1:5])
                eps = float(n.attrs.get("epsilon", 1e-5))
                w = g.initializers[conv.input[1]].astype(np.float64)
                f = scale / np.sqrt(var + eps)
                b0 = (g.initializers[conv.input[2]].astype(np.float64)
                      if len(conv.input) > 2 and conv.input[2] else np.zeros_like(mean))
                g.initializers[conv.input[1]] = (w * f.reshape(-1, *[1] * (w.ndim - 1))).astype(np.float32)
                if len(conv.input) > 2 and conv.input[2]:
                    bname = conv.input[2]
                else:
                    # nodes may still be unnamed here (names are assigned after simplification): derive the new
                    # initializer's name from the Conv's output tensor, which is unique
                    bname = n.input[0] + "_bias"
                    while bname in g.initializers:
                        bname += "_"
                g.initializers[bname] = ((b0 - mean) * f + bias).astype(np.float32)
                if len(conv.input) > 2:
                    conv.input[2] = bname
                else:
                    conv.input.append(bname)
                conv.output[0] = n.output[0]
                producer[n.output[0]] = conv
                dropped.add(id(n))
                continue
        kept.append(n)
    g.nodes = kept
    used = {i for n in g.nodes for i in n.input} | out_names
    g.initializers = {k: v for k, v in g.initializers.items() if k in used}
    g.inputs = [vi for vi in g.inputs if vi.name not in g.initializers]
    g.value_info = []
    return m


# ------------------------------------------------------------------ graph wrapper
class _GraphView:
    """`onnx_graph.graph`: the reference reads `.node`, `.input`, `.output`,
    `.initializer`, `.name` off a GraphProto."""

    def __init__(self, g):
        self._g = g

    node = property(lambda self: self._g.nodes)
    input = property(lambda self: self._g.inputs)
    output = property(lambda self: self._g.outputs)
    name = property(lambda self: self._g.name)

    @property
    def initializer(self):
        return list(self._g.initializers.items())


class ONNXGraph:
    def __init__(self, model=None, output_dir="", deploy=None, model_type=None):
        self.model = model
        self.output_dir = output_dir
        self.deploy = deploy
        self.model_type = model_type
        self.initializer = {}           # name -> (ndarray, position)
        self.input_map = {}             # tensor -> consumer nodes
        self.output_map = {}            # tensor -> producer node
        self.network_inputs = []
        self.network_outputs = []
        self.tensor_name_shape_map = {}
        self.value_name_type_map = {}
        self.name_idx_map = {}
        self.input = []
        self.output = []
        if model is not None:
            self._rebuild(first=True)

    # -- construction ---------------------------------------------------------
    @property
    def graph(self):
        return _GraphView(self.model.graph)

    def _rebuild(self, first=False):
        g = self.model.graph
        for idx, node in enumerate(g.nodes):      # utils.py:49-52
            if node.name == "":
                node.name = f"{node.op_type}_{idx}"
        if first:
            for node in g.nodes:                   # utils.py:54-58
                if node.op_type == "Constant" and "value" in node.attrs:
                    g.initializers[node.output[0]] = np.asarray(node.attrs["value"])
        self.topologize_graph()
        self.prepare_initializer()
        self.set_index()
        self.get_inp_oup()
        self.get_shape_type()

    def prepare_initializer(self):
        self.initializer = {name: (arr, i) for i, (name, arr) in
                            enumerate(self.model.graph.initializers.items())}

    def topologize_graph(self):
        self.input_map, self.output_map = {}, {}
        for node in self.model.graph.nodes:
            for o in node.output:
                self.output_map[o] = node
            for i in node.input:
                self.input_map.setdefault(i, []).append(node)

    def set_index(self):
        self.name_idx_map = {n.name: i for i, n in enumerate(self.model.graph.nodes)}

    def index(self, node):
        return self.name_idx_map[node.name]

    def get_inp_oup(self):
        g = self.model.graph
        self.network_inputs = [vi.name for vi in g.inputs
                               if vi.name not in self.output_map and vi.name not in self.initializer]
        self.network_outputs = [vi.name for vi in g.outputs]
        self.input = list(self.network_inputs)
        self.output = list(self.network_outputs)
        for node in g.nodes:
            for i in node.input:
                if i in self.initializer and i not in self.input:
                    self.input.append(i)
            for o in node.output:
                if o not in self.output:
                    self.output.append(o)

    def get_shape_type(self):
        """Shapes of every tensor. The reference reads them from onnxsim's value_info
        (utils.py:88-117); here they are inferred analytically from the input shapes."""
        g = self.model.graph
        shapes, types = {}, {}
        for vi in g.inputs:
            if vi.name in self.network_inputs:
                shapes[vi.name] = list(vi.shape or [])
                types[vi.name] = vi.elem_type
        consts = {}
        for name, (arr, _) in self.initializer.items():
            shapes[name] = list(arr.shape)
            consts[name] = arr
        for node in g.nodes:
            try:
                outs = infer_node_shape(node, shapes, consts)
            except (KeyError, NotImplementedError, TypeError):
                continue
            for o, s in zip(node.output, outs):
                shapes[o] = [int(d) for d in s]
                types[o] = ol.INT8 if node.op_type == "QuantizeLinear" else ol.FLOAT
        for vi in list(g.outputs) + list(g.value_info):
            if vi.name not in shapes and vi.shape is not None:
                shapes[vi.name] = list(vi.shape)
                types[vi.name] = vi.elem_type
        for name in list(shapes):
            if name.endswith("_q") or name.endswith("_dq"):
                continue
            shapes.setdefault(name + "_q", shapes[name])
            if self.deploy is not None:
                key = "qw_params" if name in self.initializer else "qi_params"
                sym = platform_setting_table[self.deploy][key]["symmetric"]
                types[name + "_q"] = ol.INT8 if sym else ol.UINT8
                shapes.setdefault(name + "_dq", shapes[name])
                types[name + "_dq"] = ol.FLOAT
        self.tensor_name_shape_map = shapes
        self.value_name_type_map = types

    # -- queries ----------------------------------------------------------------
    def get_tensor_shape(self, name):
        return self.tensor_name_shape_map[name]

    def get_value_type(self, name):
        return self.value_name_type_map[name]

    def get_initializer(self, name):
        return self.initializer[name][0]

    def get_tensor_producer(self, name):
        return self.output_map.get(name, INPUT_TOKEN)

    def get_tensor_consumer(self, name):
        return self.input_map.get(name, [OUTPUT_TOKEN])

    def get_constant(self, name):
        arr = self.model.graph.initializers.get(name)
        return None if arr is None else arr.tolist()

    # -- mutation ---------------------------------------------------------------
    def set_initializer(self, name, value, raw=True):
        self.model.graph.initializers[name] = np.asarray(value)
        self._init_version = getattr(self, "_init_version", 0) + 1   # cached engines re-upload
        self.prepare_initializer()

    def del_initializer(self, name):
        self.initializer.pop(name, None)

    def remove_node_purely(self, node):
        self.model.graph.nodes.remove(node)

    def insert_node_purely(self, node, idx=0):
        self.model.graph.nodes.insert(idx, node)

    def insert_qnodes_purely(self, q_nodes, idx=0, node=None):
        """q_nodes: (nodes, initializers) of one Q/DQ pair, inserted before `node`."""
        nodes, inits = q_nodes.node, q_nodes.initializer
        if node is not None:
            idx = self.index(node)
        for n in reversed(nodes):
            self.model.graph.nodes.insert(idx, n)
        for name, arr in inits:
            self.model.graph.initializers[name] = arr
        self.set_index()

    def del_network_output(self, name):
        i = self.network_outputs.index(name)
        self.model.graph.outputs.pop(i)
        self.network_outputs.remove(name)

    def add_network_output(self, vi):
        self.model.graph.outputs.append(vi)
        self.network_outputs.append(vi.name)

    def update_model(self):
        self.set_index()
        self.prepare_initializer()

    def __deepcopy__(self, memo):
        new = ONNXGraph()
        new.copy_from(self)
        new.name_idx_map = dict(self.name_idx_map)
        return new

    def copy_from(self, src):
        self.model = copy.deepcopy(src.model)
        self.output_dir, self.deploy, self.model_type = src.output_dir, src.deploy, src.model_type
        self.topologize_graph()
        self.prepare_initializer()
        self.set_index()
        self.network_inputs = list(src.network_inputs)
        self.network_outputs = list(src.network_outputs)
        self.input, self.output = list(src.input), list(src.output)
        self.tensor_name_shape_map = copy.deepcopy(src.tensor_name_shape_map)
        self.value_name_type_map = dict(src.value_name_type_map)

    def save_onnx_model(self, name="tmp", size_threshold=2048):
        os.makedirs(self.output_dir, exist_ok=True)
        g = self.model.graph
        # Q/DQ graphs carry int8 intermediates: declare them so that other tools can load the file
        ol.save(self.model, os.path.join(self.output_dir, f"{name}.onnx"))


def load_graph(path, output_dir="", deploy=None, model_type=None, do_simplify=True):
    """onnx.load + onnxsim.simplify + ONNXGraph(...) (dipoorlet/__main__.py:98-103)."""
    model = ol.load(path)
    if do_simplify:
        model = simplify(model)
    return ONNXGraph(model, output_dir, deploy, model_type)

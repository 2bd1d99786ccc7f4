"""Stand-ins that let the UNMODIFIED reference package (/root/reference/dipoorlet) be
imported and executed in this container, where onnx / onnxruntime / onnxsim / termcolor
are not installable. Used only by oracle/gen_golden.py to produce tests/golden/*.

  onnx         plain-Python message objects with the attribute surface the reference
               touches (graph.node / initializer / input / output / value_info lists,
               helper.make_*, numpy_helper, TensorProto enums, load / save through
               dipoorlet_b200.onnx_lite's wire codec)
  onnxruntime  InferenceSession executing the graph with oracle.forward.run_node (torch
               CPU fp32; Q/DQ per the ONNX operator spec) — the one third-party piece of
               arithmetic that has to be restated
  onnxsim      simplify() = identity (the fixtures are built already simplified)
  torch glue   dist.* -> single process, .cuda() -> no-op, DDP -> pass-through, so that
               adaround / brecq run on the CPU

Everything else — statistics loops, clip search, JSON round trip, quant_graph,
get_qnode_by_param, ActivationCache, bias_correction, AdaQLayer, learning_round_mask,
profiling, deploy — is the reference's own code.
"""
import copy
import sys
import types

import numpy as np

_REGISTRY = {}


# ----------------------------------------------------------------------------- messages
class _Dim:
    def __init__(self, v):
        self.dim_value = int(v)
        self.dim_param = ""


class _Shape:
    def __init__(self, dims):
        self.dim = [_Dim(d) for d in dims]


class _TensorType:
    def __init__(self, elem_type, shape):
        self.elem_type = elem_type
        self.shape = _Shape(shape if shape is not None else [])


class _Type:
    def __init__(self, elem_type, shape):
        self.tensor_type = _TensorType(elem_type, shape)


class ValueInfoProto:
    def __init__(self, name="", elem_type=1, shape=None):
        self.name = name
        self.type = _Type(elem_type, shape)


class TensorProto:
    UNDEFINED, FLOAT, UINT8, INT8, UINT16, INT16, INT32, INT64, STRING, BOOL, FLOAT16, DOUBLE, \
        UINT32, UINT64 = range(14)

    def __init__(self, name="", array=None):
        self.name = name
        self.array = array

    @property
    def dims(self):
        return list(self.array.shape)

    @property
    def data_type(self):
        return _DT[self.array.dtype.type]


_NP = {TensorProto.FLOAT: np.float32, TensorProto.UINT8: np.uint8, TensorProto.INT8: np.int8,
       TensorProto.INT32: np.int32, TensorProto.INT64: np.int64, TensorProto.DOUBLE: np.float64,
       TensorProto.BOOL: np.bool_}
_DT = {v: k for k, v in _NP.items()}


class AttributeProto:
    def __init__(self, name, value):
        self.name = name
        self.value = value

    @property
    def t(self):
        return self.value if isinstance(self.value, TensorProto) else TensorProto("", np.asarray(self.value))


class NodeProto:
    def __init__(self, op_type, inputs, outputs, name="", attrs=None):
        self.op_type = op_type
        self.input = list(inputs)
        self.output = list(outputs)
        self.name = name
        self.attribute = [AttributeProto(k, v) for k, v in (attrs or {}).items()]

    def attrs(self):
        return {a.name: a.value for a in self.attribute}


class GraphProto:
    def __init__(self, nodes=(), name="g", inputs=(), outputs=(), initializer=(), value_info=()):
        self.node = list(nodes)
        self.name = name
        self.input = list(inputs)
        self.output = list(outputs)
        self.initializer = list(initializer)
        self.value_info = list(value_info)


class _Opset:
    def __init__(self, version=13, domain=""):
        self.version = version
        self.domain = domain


class ModelProto:
    def __init__(self, graph, opset_import=None, producer_name=""):
        self.graph = graph
        self.opset_import = opset_import or [_Opset()]
        self.producer_name = producer_name

    def SerializeToString(self):
        key = f"shim-model-{id(self)}-{len(_REGISTRY)}".encode()
        _REGISTRY[key] = copy.deepcopy(self)
        return key


# ----------------------------------------------------------------------------- onnx module
def _helper():
    h = types.ModuleType("onnx.helper")

    def make_tensor(name, data_type, dims, vals, raw=False):
        arr = np.asarray(vals).astype(_NP[data_type]).reshape(dims)
        return TensorProto(name, arr)

    def make_tensor_value_info(name, elem_type, shape):
        return ValueInfoProto(name, elem_type, shape)

    def make_node(op_type, inputs, outputs, name="", **attrs):
        return NodeProto(op_type, inputs, outputs, name, attrs)

    def make_graph(nodes, name, inputs, outputs, initializer=()):
        return GraphProto(nodes, name, inputs, outputs, initializer)

    def make_model(graph, producer_name="", opset_imports=None, **kw):
        return ModelProto(graph, opset_imports, producer_name)

    def get_attribute_value(attr):
        return attr.value

    h.make_tensor, h.make_tensor_value_info, h.make_node = make_tensor, make_tensor_value_info, make_node
    h.make_graph, h.make_model, h.get_attribute_value = make_graph, make_model, get_attribute_value
    return h


def _numpy_helper():
    m = types.ModuleType("onnx.numpy_helper")
    m.to_array = lambda t: t.array
    m.from_array = lambda arr, name="": TensorProto(name, np.asarray(arr))
    return m


def executed_shapes(model):
    """name -> shape of every node output, read off the tensors the oracle's own forward (oracle/forward.py, torch-CPU
    ops) produces for an all-zero input - independent of the product's analytic shape inference."""
    from oracle import forward as OF
    g = model.graph
    feeds = {v.name: np.zeros([int(d) for d in v.shape], dtype=np.float32) for v in g.inputs
             if v.name not in g.initializers}
    return {k: list(v.shape) for k, v in OF.forward_all(model, feeds).items()}


def from_lite(model):
    """dipoorlet_b200.onnx_lite.Model -> shim ModelProto. value_info (what onnxsim leaves in the file) is filled
    from the shapes of an executed forward, NOT from dipoorlet_b200.graph's shape inference, so that the
    graph-IR comparison against the reference can catch a shape-inference bug of the product."""
    g = model.graph
    shapes = executed_shapes(model)
    nodes = [NodeProto(n.op_type, n.input, n.output, n.name, dict(n.attrs)) for n in g.nodes]
    inits = [TensorProto(k, v.copy()) for k, v in g.initializers.items()]
    inputs = [ValueInfoProto(v.name, v.elem_type, v.shape) for v in g.inputs]
    outputs = [ValueInfoProto(v.name, v.elem_type, v.shape) for v in g.outputs]
    out_names = {v.name for v in g.outputs}
    vinfo = [ValueInfoProto(o, 1, shapes[o]) for n in g.nodes for o in n.output if o not in out_names]
    return ModelProto(GraphProto(nodes, g.name, inputs, outputs, inits, vinfo),
                      [_Opset(v, k) for k, v in model.opsets.items()])


def to_lite(mp):
    from dipoorlet_b200 import onnx_lite as ol
    g = ol.Graph(mp.graph.name)
    for n in mp.graph.node:
        attrs = {}
        for a in n.attribute:
            attrs[a.name] = a.value.array if isinstance(a.value, TensorProto) else a.value
        g.nodes.append(ol.Node(n.op_type, n.input, n.output, n.name, attrs))
    for t in mp.graph.initializer:
        g.initializers[t.name] = np.asarray(t.array)
    for src, dst in ((mp.graph.input, g.inputs), (mp.graph.output, g.outputs)):
        for v in src:
            dst.append(ol.ValueInfo(v.name, v.type.tensor_type.elem_type,
                                    [d.dim_value for d in v.type.tensor_type.shape.dim]))
    return ol.Model(g, opsets={o.domain: o.version for o in mp.opset_import})


def _onnx_module():
    m = types.ModuleType("onnx")
    m.ValueInfoProto, m.TensorProto, m.ModelProto = ValueInfoProto, TensorProto, ModelProto
    m.helper, m.numpy_helper = _helper(), _numpy_helper()
    ext = types.ModuleType("onnx.external_data_helper")
    ext.convert_model_to_external_data = lambda *a, **k: None
    m.external_data_helper = ext

    def save(model, path):
        from dipoorlet_b200 import onnx_lite as ol
        ol.save(to_lite(model), path)

    def load(path):
        from dipoorlet_b200 import onnx_lite as ol
        return from_lite(ol.load(path))

    m.save, m.load = save, load
    chk = types.ModuleType("onnx.checker")
    chk.check_model = lambda model: None
    chk.ValidationError = Exception
    m.checker = chk
    vc = types.ModuleType("onnx.version_converter")
    vc.convert_version = lambda model, v: model
    m.version_converter = vc
    return m, ext


# ----------------------------------------------------------------------------- onnxruntime
class _Out:
    def __init__(self, name):
        self.name = name


class InferenceSession:
    def __init__(self, model_bytes, providers=None):
        self.model = _REGISTRY[bytes(model_bytes)]

    def get_provider_options(self):
        return {"CUDAExecutionProvider": {}}

    def get_outputs(self):
        return [_Out(o.name) for o in self.model.graph.output]

    def run(self, outputs, feeds):
        import torch
        from oracle import forward as OF
        g = self.model.graph
        env = {t.name: torch.from_numpy(np.ascontiguousarray(t.array)) for t in g.initializer}
        for k, v in feeds.items():
            env[k] = torch.from_numpy(np.ascontiguousarray(np.asarray(v)))
        with torch.no_grad():
            for node in g.node:
                ins = [env[i] if i else None for i in node.input]
                env[node.output[0]] = OF.run_node(node, ins, node.attrs())
        return [env[o].numpy() for o in outputs]


def _ort_module():
    m = types.ModuleType("onnxruntime")
    m.InferenceSession = InferenceSession
    m.set_default_logger_severity = lambda level: None
    q = types.ModuleType("onnxruntime.quantization")
    oq = types.ModuleType("onnxruntime.quantization.onnx_quantizer")
    oq.ONNXQuantizer = object
    qu = types.ModuleType("onnxruntime.quantization.quant_utils")
    qu.QuantizationMode = types.SimpleNamespace(QLinearOps=0)
    qu.QuantType = types.SimpleNamespace(QInt8=0, QUInt8=1)
    q.onnx_quantizer, q.quant_utils = oq, qu
    m.quantization = q
    return m, q, oq, qu


# ----------------------------------------------------------------------------- install
def install(reference_root="/root/reference"):
    """Put the stand-ins into sys.modules, patch torch for single-process CPU execution and
    make `import dipoorlet` resolve to the reference checkout."""
    import torch
    import torch.distributed as dist
    onnx, ext = _onnx_module()
    sys.modules.update({"onnx": onnx, "onnx.helper": onnx.helper, "onnx.numpy_helper": onnx.numpy_helper,
                        "onnx.external_data_helper": ext, "onnx.checker": onnx.checker,
                        "onnx.version_converter": onnx.version_converter})
    ort, q, oq, qu = _ort_module()
    sys.modules.update({"onnxruntime": ort, "onnxruntime.quantization": q,
                        "onnxruntime.quantization.onnx_quantizer": oq,
                        "onnxruntime.quantization.quant_utils": qu})
    sim = types.ModuleType("onnxsim")
    sim.simplify = lambda model: (model, True)
    sys.modules["onnxsim"] = sim
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules["termcolor"] = tc
    # torch: one CPU process
    dist.get_rank = lambda *a, **k: 0
    dist.get_world_size = lambda *a, **k: 1
    dist.barrier = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: 0
    torch.cuda.device_count = lambda: 1

    class _DDP(torch.nn.Module):
        def __init__(self, module, device_ids=None, **kw):
            super().__init__()
            self.module = module

        def forward(self, *a, **k):
            return self.module(*a, **k)

    import torch.nn.parallel
    torch.nn.parallel.DistributedDataParallel = _DDP
    from importlib.machinery import ModuleSpec
    for name in ("onnx", "onnx.helper", "onnx.numpy_helper", "onnx.external_data_helper",
                 "onnx.checker", "onnx.version_converter", "onnxruntime", "onnxruntime.quantization",
                 "onnxruntime.quantization.onnx_quantizer", "onnxruntime.quantization.quant_utils",
                 "onnxsim", "termcolor"):
        sys.modules[name].__spec__ = ModuleSpec(name, None)  # importlib.util.find_spec() probes these
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return onnx

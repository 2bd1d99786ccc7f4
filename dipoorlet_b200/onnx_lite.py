"""Minimal ONNX (protobuf wire format) reader / writer.

The reference manipulates `onnx.ModelProto` objects (dipoorlet/utils.py:22-250); the
`onnx` package is not available on the build box, and the hot path only needs a handful
of messages, so the wire format is decoded directly into plain Python containers:

    Model(ir_version, opsets, producer, graph)
    Graph(name, nodes, initializers{name: ndarray}, inputs[ValueInfo], outputs, value_info)
    Node(name, op_type, input[], output[], attrs{name: python value}, domain)

Field numbers follow onnx/onnx.proto (ONNX IR v3+).
"""
import struct

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, UINT16, INT16, INT32, INT64, STRING, BOOL, FLOAT16, DOUBLE, UINT32, UINT64 = \
    range(1, 14)
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, UINT16: np.uint16, INT16: np.int16,
          INT32: np.int32, INT64: np.int64, BOOL: np.bool_, FLOAT16: np.float16,
          DOUBLE: np.float64, UINT32: np.uint32, UINT64: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}


class Node:
    __slots__ = ("name", "op_type", "input", "output", "attrs", "domain")

    def __init__(self, op_type, input, output, name="", attrs=None, domain=""):
        self.op_type = op_type
        self.input = list(input)
        self.output = list(output)
        self.name = name
        self.attrs = dict(attrs or {})
        self.domain = domain

    def __repr__(self):
        return f"Node({self.op_type} {self.name!r}: {self.input} -> {self.output})"

    def clone(self):
        attrs = {k: (v.copy() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v))
                 for k, v in self.attrs.items()}
        return Node(self.op_type, self.input, self.output, self.name, attrs, self.domain)


class ValueInfo:
    __slots__ = ("name", "elem_type", "shape")

    def __init__(self, name, elem_type=FLOAT, shape=None):
        self.name = name
        self.elem_type = elem_type
        self.shape = None if shape is None else list(shape)  # ints; 0 for symbolic dims

    def clone(self):
        return ValueInfo(self.name, self.elem_type, self.shape)


class Graph:
    def __init__(self, name="graph"):
        self.name = name
        self.nodes = []
        self.initializers = {}   # insertion-ordered: name -> ndarray
        self.inputs = []
        self.outputs = []
        self.value_info = []


class Model:
    def __init__(self, graph=None, ir_version=7, opsets=None, producer="dipoorlet_b200"):
        self.graph = graph or Graph()
        self.ir_version = ir_version
        self.opsets = dict(opsets or {"": 13})
        self.producer = producer


# --------------------------------------------------------------------------- decode
def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _fields(buf):
    """Yield (field_number, wire_type, value) over a message; value is an int for
    varint / fixed, a memoryview for length-delimited."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _packed_varints(v):
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


def _rep_int(acc, wt, v):
    if wt == 2:
        acc.extend(_packed_varints(v))
    else:
        acc.append(_signed(v))


def _rep_float(acc, wt, v, fmt="<f", size=4):
    if wt == 2:
        acc.extend(np.frombuffer(bytes(v), dtype=fmt).tolist())
    else:
        acc.append(struct.unpack(fmt, v)[0])


def _str(v):
    return bytes(v).decode("utf-8")


def _tensor(buf):
    dims, dtype, name, raw = [], FLOAT, "", None
    f32, i32, i64, f64, u64 = [], [], [], [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            _rep_int(dims, wt, v)
        elif fn == 2:
            dtype = v
        elif fn == 4:
            _rep_float(f32, wt, v)
        elif fn == 5:
            _rep_int(i32, wt, v)
        elif fn == 7:
            _rep_int(i64, wt, v)
        elif fn == 8:
            name = _str(v)
        elif fn == 9:
            raw = bytes(v)
        elif fn == 10:
            _rep_float(f64, wt, v, "<d", 8)
        elif fn == 11:
            _rep_int(u64, wt, v)
        elif fn == 14 and v == 1:
            raise ValueError(f"initializer {name!r} uses external data, which is not supported")
    if dtype not in _NP_OF:
        raise ValueError(f"unsupported tensor data type {dtype} for {name!r}")
    npdt = _NP_OF[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(npdt).newbyteorder("<")).astype(npdt)
    elif dtype == FLOAT:
        arr = np.asarray(f32, dtype=np.float32)
    elif dtype == DOUBLE:
        arr = np.asarray(f64, dtype=np.float64)
    elif dtype == INT64:
        arr = np.asarray(i64, dtype=np.int64)
    elif dtype in (UINT32, UINT64):
        arr = np.asarray(u64, dtype=npdt)
    elif dtype == FLOAT16:
        arr = np.asarray(i32, dtype=np.uint16).view(np.float16)
    else:
        arr = np.asarray(i32).astype(npdt)
    return name, arr.reshape(dims).copy()


def _attribute(buf):
    name, atype = "", 0
    f = i = s = t = None
    floats, ints, strings = [], [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            name = _str(v)
        elif fn == 20:
            atype = v
        elif fn == 2:
            f = struct.unpack("<f", v)[0]
        elif fn == 3:
            i = _signed(v)
        elif fn == 4:
            s = bytes(v)
        elif fn == 5:
            t = _tensor(v)[1]
        elif fn == 7:
            _rep_float(floats, wt, v)
        elif fn == 8:
            _rep_int(ints, wt, v)
        elif fn == 9:
            strings.append(bytes(v))
    if atype == 1 or (atype == 0 and f is not None):
        return name, float(f)
    if atype == 2 or (atype == 0 and i is not None):
        return name, int(i)
    if atype == 3 or (atype == 0 and s is not None):
        return name, s.decode("utf-8", "replace")
    if atype == 4 or (atype == 0 and t is not None):
        return name, t
    if atype == 6:
        return name, [float(x) for x in floats]
    if atype == 7:
        return name, [int(x) for x in ints]
    if atype == 8:
        return name, [x.decode("utf-8", "replace") for x in strings]
    if floats:
        return name, floats
    if ints:
        return name, ints
    return name, None


def _node(buf):
    n = Node("", [], [])
    for fn, wt, v in _fields(buf):
        if fn == 1:
            n.input.append(_str(v))
        elif fn == 2:
            n.output.append(_str(v))
        elif fn == 3:
            n.name = _str(v)
        elif fn == 4:
            n.op_type = _str(v)
        elif fn == 7:
            n.domain = _str(v)
        elif fn == 5:
            k, val = _attribute(v)
            n.attrs[k] = val
    return n


def _value_info(buf):
    vi = ValueInfo("", FLOAT, None)
    for fn, wt, v in _fields(buf):
        if fn == 1:
            vi.name = _str(v)
        elif fn == 2:  # TypeProto
            for fn2, _, v2 in _fields(v):
                if fn2 == 1:  # tensor_type
                    for fn3, _, v3 in _fields(v2):
                        if fn3 == 1:
                            vi.elem_type = v3
                        elif fn3 == 2:  # shape
                            shape = []
                            for fn4, _, v4 in _fields(v3):
                                if fn4 == 1:
                                    d = 0
                                    for fn5, _, v5 in _fields(v4):
                                        if fn5 == 1:
                                            d = _signed(v5)
                                    shape.append(d)
                            vi.shape = shape
    return vi


def _graph(buf):
    g = Graph()
    for fn, wt, v in _fields(buf):
        if fn == 1:
            g.nodes.append(_node(v))
        elif fn == 2:
            g.name = _str(v)
        elif fn == 5:
            name, arr = _tensor(v)
            g.initializers[name] = arr
        elif fn == 11:
            g.inputs.append(_value_info(v))
        elif fn == 12:
            g.outputs.append(_value_info(v))
        elif fn == 13:
            g.value_info.append(_value_info(v))
    return g


def loads(data):
    buf = memoryview(data)
    m = Model(Graph(), opsets={})
    for fn, wt, v in _fields(buf):
        if fn == 1:
            m.ir_version = v
        elif fn == 2:
            m.producer = _str(v)
        elif fn == 7:
            m.graph = _graph(v)
        elif fn == 8:
            dom, ver = "", 0
            for fn2, _, v2 in _fields(v):
                if fn2 == 1:
                    dom = _str(v2)
                elif fn2 == 2:
                    ver = _signed(v2)
            m.opsets[dom] = ver
    if not m.opsets:
        m.opsets = {"": 13}
    return m


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())


# --------------------------------------------------------------------------- encode
def _enc_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fn, wt):
    return _enc_varint((fn << 3) | wt)


def _ld(fn, payload):
    return _key(fn, 2) + _enc_varint(len(payload)) + payload


def _enc_str(fn, s):
    return _ld(fn, s.encode("utf-8"))


def _enc_int(fn, v):
    return _key(fn, 0) + _enc_varint(int(v))


def _enc_tensor(name, arr):
    arr = np.asarray(arr)
    if arr.dtype not in _DT_OF:
        raise ValueError(f"cannot serialise dtype {arr.dtype}")
    out = bytearray()
    if arr.ndim:
        out += _ld(1, b"".join(_enc_varint(int(d)) for d in arr.shape))
    out += _enc_int(2, _DT_OF[arr.dtype])
    out += _enc_str(8, name)
    out += _ld(9, np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())
    return bytes(out)


def _enc_attr(name, val):
    out = bytearray(_enc_str(1, name))
    if isinstance(val, bool):
        val = int(val)
    if isinstance(val, float):
        out += _key(2, 5) + struct.pack("<f", val) + _enc_int(20, 1)
    elif isinstance(val, (int, np.integer)):
        out += _enc_int(3, int(val)) + _enc_int(20, 2)
    elif isinstance(val, str):
        out += _ld(4, val.encode("utf-8")) + _enc_int(20, 3)
    elif isinstance(val, np.ndarray):
        out += _ld(5, _enc_tensor("", val)) + _enc_int(20, 4)
    elif isinstance(val, (list, tuple)):
        if val and isinstance(val[0], float):
            out += _ld(7, np.asarray(val, dtype="<f4").tobytes()) + _enc_int(20, 6)
        elif val and isinstance(val[0], str):
            for s in val:
                out += _ld(9, s.encode("utf-8"))
            out += _enc_int(20, 8)
        else:
            out += _ld(8, b"".join(_enc_varint(int(x)) for x in val)) + _enc_int(20, 7)
    else:
        raise ValueError(f"cannot serialise attribute {name}={val!r}")
    return bytes(out)


def _enc_node(n):
    out = bytearray()
    for s in n.input:
        out += _enc_str(1, s)
    for s in n.output:
        out += _enc_str(2, s)
    if n.name:
        out += _enc_str(3, n.name)
    out += _enc_str(4, n.op_type)
    for k, v in n.attrs.items():
        if v is not None:
            out += _ld(5, _enc_attr(k, v))
    if n.domain:
        out += _enc_str(7, n.domain)
    return bytes(out)


def _enc_value_info(vi):
    tt = bytearray(_enc_int(1, vi.elem_type))
    if vi.shape is not None:
        dims = b"".join(_ld(1, _enc_int(1, d)) for d in vi.shape)
        tt += _ld(2, dims)
    return _enc_str(1, vi.name) + _ld(2, _ld(1, bytes(tt)))


def dumps(model):
    g = model.graph
    gb = bytearray()
    for n in g.nodes:
        gb += _ld(1, _enc_node(n))
    gb += _enc_str(2, g.name)
    for name, arr in g.initializers.items():
        gb += _ld(5, _enc_tensor(name, arr))
    for vi in g.inputs:
        gb += _ld(11, _enc_value_info(vi))
    for vi in g.outputs:
        gb += _ld(12, _enc_value_info(vi))
    for vi in g.value_info:
        gb += _ld(13, _enc_value_info(vi))
    out = bytearray(_enc_int(1, model.ir_version))
    out += _enc_str(2, model.producer)
    out += _ld(7, bytes(gb))
    for dom, ver in model.opsets.items():
        out += _ld(8, _enc_str(1, dom) + _enc_int(2, ver))
    return bytes(out)


def save(model, path):
    with open(path, "wb") as f:
        f.write(dumps(model))

"""The C-ABI library builds, loads and exports every symbol include/dpl_b200.h declares
(no compute calls: there is no GPU in the CPU test run)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dpl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(dpl_built):
    from dipoorlet_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    handle = ctypes.CDLL(dpl_built)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in dpl_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) <= set(names)
    assert _lib.lib().dpl_version() == 100


def test_planner_is_pure_host_arithmetic(dpl_built):
    from dipoorlet_b200 import _lib
    t = np.zeros((4, _lib.BLOB_FIELDS), dtype=np.uint64)
    t[:, _lib.F_NSEG] = [3, 3, 3, 3]
    t[:, _lib.F_SEGLEN] = [1000, 802816, 0, 8192]
    segs, seg_tiles, flat_tiles = _lib.plan_blobs(t)
    assert segs == 12
    assert seg_tiles == 3 * 1 + 3 * 98 + 0 + 3 * 1
    assert flat_tiles == 1 + (3 * 802816 + 8191) // 8192 + 0 + 3
    assert t[:, _lib.F_SEG_OUT_BASE].tolist() == [0, 3, 6, 9]
    assert t[:, _lib.F_SEG_TILE_BEGIN].tolist() == [0, 3, 297, 297]
    assert t[:, _lib.F_FLAT_TILE_BEGIN].tolist() == [0, 1, 295, 295]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dipoorlet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.lib()
    except _lib.DplLibraryError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("a missing libdpl_b200.so must raise")


def test_engine_refuses_cpu():
    import pytest
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.engine import Engine
    from dipoorlet_b200.graph import ONNXGraph
    g = ONNXGraph(W.build_resnet50(blocks=[1], planes=(4,), stem=4, num_classes=3, image=16), "", "trt")
    with pytest.raises(RuntimeError):
        Engine(g, "cpu")


def _prototypes():
    """name -> list of C parameter type strings, parsed from the header."""
    text = open(os.path.join(ROOT, "include", "dpl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(dpl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if params in ([""], ["void"]) else params
    return protos


def test_ctypes_signatures_match_the_header(dpl_built):
    """Every prototype of include/dpl_b200.h against _lib.SIGNATURES: same number of parameters and the
    same class of type at each position (pointer / 32-bit int / 64-bit int / float / double / size_t), so
    that a changed C signature cannot silently shift the arguments of a ctypes call."""
    from dipoorlet_b200 import _lib

    def c_class(decl):
        decl = re.sub(r"\b[a-zA-Z_][a-zA-Z0-9_]*$", "", decl.strip()).strip()   # drop the parameter name
        if "*" in decl:
            return "ptr"
        base = decl.replace("const", "").replace("unsigned", "u").strip()
        return {"int": "i32", "float": "f32", "double": "f64", "size_t": "size", "uint64_t": "i64", "long long": "i64",
                "u long long": "i64", "int32_t": "i32", "uint32_t": "i32"}[base]

    def ct_class(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_int: "i32", ctypes.c_float: "f32", ctypes.c_double: "f64", ctypes.c_size_t: "size",
                ctypes.c_uint64: "i64", ctypes.c_longlong: "i64", ctypes.c_uint32: "i32"}[t]

    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        _, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (decl, t) in enumerate(zip(params, argtypes)):
            got, want = ct_class(t), c_class(decl)
            # size_t and uint64_t are the same width on this ABI; both are accepted for either
            assert got == want or {got, want} == {"size", "i64"}, (name, i, decl, t)

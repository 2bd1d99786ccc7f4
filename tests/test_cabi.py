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

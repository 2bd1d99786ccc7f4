"""Fuzz fixture for the quantisation-parameter arithmetic: the REFERENCE's own get_qnode_by_param
(dipoorlet/quantize.py:111-194, imported unmodified under oracle/ref_shim) on seeded random ranges for the
weight and activation parameter sets of ALL eight platforms of its platform_setting_table (symmetric /
asymmetric, per-tensor / per-channel, log_scale, dynamic_sym) -> tests/golden/qparam_fuzz.json.

    python oracle/gen_golden_qparams.py      # build container only; the fixture is committed
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402,F401
from oracle import ref_shim  # noqa: E402


def cases(rng):
    """(range as written to the fixture, shape) - scalars and per-channel arrays, incl. the corner cases."""
    out = []
    for _ in range(6):
        lo, hi = sorted(rng.normal(size=2).astype(np.float64) * 4)
        out.append(([float(lo), float(hi)], [1, 8, 4, 4]))
    out.append(([0.0, 3.5], [1, 8, 4, 4]))            # dynamic_sym trigger (min == 0)
    out.append(([0.0, 0.0], [1, 8, 4, 4]))            # zero scale -> 1
    out.append(([0.25, 2.0], [1, 8, 4, 4]))           # both positive: asym clamps min to 0
    out.append(([-3.0, -0.5], [1, 8, 4, 4]))          # both negative
    for c in (1, 5, 16):
        lo = -np.abs(rng.normal(size=c)) * 2
        hi = np.abs(rng.normal(size=c)) * 2
        if c > 1:
            lo[1], hi[1] = 0.0, 0.0                    # an all-zero channel
            lo[2] = 0.5                                # a positive minimum
        out.append(([lo.tolist(), hi.tolist()], [c, 3, 3, 3]))
    return out


def main():
    ref_shim.install()
    from dipoorlet.platform_settings import platform_setting_table
    from dipoorlet.quantize import get_qnode_by_param
    rng = np.random.default_rng(7)
    rows = []
    for platform, setting in platform_setting_table.items():
        for key in ("qw_params", "qi_params"):
            param = setting[key]
            for rng_val, shape in cases(rng):
                rr = [np.array(v, dtype=np.float64) if isinstance(v, list) else np.float64(v) for v in rng_val]
                try:
                    q_nodes, q_min, q_max = get_qnode_by_param(param, "t", shape, copy.deepcopy(rr))
                except Exception as e:   # recorded: the product must fail the same way or be a superset
                    rows.append({"platform": platform, "key": key, "range": rng_val, "shape": shape,
                                 "error": type(e).__name__})
                    continue
                inits = {t.name: t.array for t in q_nodes.initializer}
                rows.append({"platform": platform, "key": key, "range": rng_val, "shape": shape,
                             "scale": np.asarray(inits["t_scale"], dtype=np.float32).reshape(-1).tolist(),
                             "scale_dtype": str(np.asarray(inits["t_scale"]).dtype),
                             "zero_point": np.asarray(inits["t_zero_point"]).reshape(-1).astype(int).tolist(),
                             "zp_dtype": str(np.asarray(inits["t_zero_point"]).dtype),
                             "q_min": np.asarray(q_min).reshape(-1).astype(int).tolist(),
                             "q_max": np.asarray(q_max).reshape(-1).astype(int).tolist(),
                             "axis": [dict((a.name, a.value) for a in n.attribute).get("axis") for n in q_nodes.node]})
    out = os.path.join(ROOT, "tests", "golden", "qparam_fuzz.json")
    json.dump({"params": {p: {k: s[k] for k in ("qw_params", "qi_params")} for p, s in platform_setting_table.items()},
               "rows": rows}, open(out, "w"))
    print("wrote", out, len(rows), "cases,", sum("error" in r for r in rows), "errors")


if __name__ == "__main__":
    main()

"""quant_graph on EVERY platform of the reference's platform_setting_table (dipoorlet/quantize.py:20-108,
platform_settings.py): which tensors get Q/DQ pairs, node order and wiring, scale / zero-point values, for
the two small seeded models -> tests/golden/<model>/quant_graph_platforms.json. Runs the reference's own
quant_graph (imported unmodified under oracle/ref_shim) on the clip values already committed
(calibration.json minmax + weight_clip.npz).

    python oracle/gen_golden_quant_graphs.py      # build container only; the fixtures are committed
"""
import copy
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402,F401
from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD  # noqa: E402


def main():
    from dipoorlet_b200 import onnx_lite as ol
    ref_shim.install()
    import dipoorlet.utils as RU
    from dipoorlet.platform_settings import platform_setting_table
    from dipoorlet.quantize import quant_graph
    for mname in ("tiny_r50", "tiny_mbv2"):
        d = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(d, "model.onnx"))
        calib = json.load(open(os.path.join(d, "calibration.json")))
        gold_w = np.load(os.path.join(d, "weight_clip.npz"))
        out = {}
        for platform in platform_setting_table:
            tmp = tempfile.mkdtemp(prefix="dpl_gold_qg_")
            g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, platform, None)
            clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
            for key in gold_w.files:
                name, i = key.rsplit("|", 1)
                clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
            args = types.SimpleNamespace(deploy=platform, skip_layers=[], optim_transformer=False)
            gq, qlist = quant_graph(g, copy.deepcopy(clip), args)
            inits = {t.name: np.asarray(t.array) for t in gq.graph.initializer
                     if t.name.endswith("_scale") or t.name.endswith("_zero_point")}
            out[platform] = {
                "quant_node_list": [n.name for n in qlist],
                "nodes": [[n.op_type, n.name, list(n.input), list(n.output),
                           {a.name: a.value for a in n.attribute if a.name == "axis"}] for n in gq.graph.node],
                "network_outputs": list(gq.network_outputs),
                "qparams": {k: [str(v.dtype), np.asarray(v, dtype=np.float64).reshape(-1).tolist()]
                            for k, v in inits.items()}}
        json.dump(out, open(os.path.join(d, "quant_graph_platforms.json"), "w"))
        print(mname, {p: len(v["nodes"]) for p, v in out.items()})


if __name__ == "__main__":
    main()

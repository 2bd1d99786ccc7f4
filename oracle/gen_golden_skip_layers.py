"""Golden fixture for `--skip_layers`: the REFERENCE's own quant_graph (/root/reference/dipoorlet, unmodified, under
oracle/ref_shim) with two layers named in args.skip_layers (a Conv in the middle of the net and the final Gemm),
on the two small seeded models, platform trt -> tests/golden/<model>/quant_graph_skip_layers.json.

    python oracle/gen_golden_skip_layers.py        # build container only; the fixtures are committed
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
import torch  # noqa: E402,F401

from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD  # noqa: E402


def main():
    from dipoorlet_b200 import onnx_lite as ol
    ref_shim.install()
    import dipoorlet.utils as RU
    from dipoorlet.quantize import quant_graph
    for mname in ("tiny_r50", "tiny_mbv2"):
        d = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(d, "model.onnx"))
        calib = json.load(open(os.path.join(d, "calibration.json")))
        gold_w = np.load(os.path.join(d, "weight_clip.npz"))
        convs = [n.name for n in model.graph.nodes if n.op_type == "Conv"]
        skip = [convs[len(convs) // 2], [n.name for n in model.graph.nodes if n.op_type == "Gemm"][-1]]
        tmp = tempfile.mkdtemp(prefix="dpl_gold_skip_")
        g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, "trt", None)
        clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
        for key in gold_w.files:
            name, i = key.rsplit("|", 1)
            clip.setdefault(name, [None, None])[int(i)] = gold_w[key]
        args = types.SimpleNamespace(deploy="trt", skip_layers=skip, optim_transformer=False)
        gq, qlist = quant_graph(g, copy.deepcopy(clip), args)
        out = {"skip_layers": skip, "quant_node_list": [n.name for n in qlist],
               "nodes": [[n.op_type, n.name, list(n.input), list(n.output)] for n in gq.graph.node]}
        json.dump(out, open(os.path.join(d, "quant_graph_skip_layers.json"), "w"))
        print(mname, skip, len(out["nodes"]))


if __name__ == "__main__":
    main()

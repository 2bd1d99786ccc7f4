"""Golden fixture for the user-visible report: the lines the REFERENCE's own show_model_profiling_res,
show_model_ranges and weight_need_perchannel log (/root/reference/dipoorlet/profiling.py:199-260, unmodified, under
oracle/ref_shim) for the small seeded models, fed the committed clip values and profiling numbers ->
tests/golden/<model>/reports.json. Platforms: trt (per-channel weights) and snpe (per-tensor: the
"degradate by per layer" ranking is printed).

    python oracle/gen_golden_reports.py        # build container only; the fixtures are committed
"""
import json
import logging
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


class _Lines(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def main():
    from dipoorlet_b200 import onnx_lite as ol
    ref_shim.install()
    import dipoorlet.utils as RU
    from dipoorlet import profiling as RP
    from dipoorlet.quantize import quant_graph
    for mname in ("tiny_r50", "tiny_mbv2"):
        d = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(d, "model.onnx"))
        calib = json.load(open(os.path.join(d, "calibration.json")))
        gold_w = np.load(os.path.join(d, "weight_clip.npz"))
        prof = json.load(open(os.path.join(d, "profiling_bc.json")))
        out = {}
        for platform in ("trt", "snpe"):
            tmp = tempfile.mkdtemp(prefix="dpl_gold_rep_")
            g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, platform, None)
            act = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
            weight = {}
            for key in gold_w.files:
                name, i = key.rsplit("|", 1)
                weight.setdefault(name, [None, None])[int(i)] = gold_w[key]
            if platform == "snpe":      # per-tensor platform: load_clip_val collapses the arrays (utils.py:365-367)
                weight = {k: [np.min(v[0]), np.max(v[1])] for k, v in weight.items()}
            args = types.SimpleNamespace(deploy=platform, skip_layers=[], optim_transformer=False,
                                         skip_prof_layer=False)
            clip = dict(act)
            clip.update(weight)
            import copy
            _, qlist = quant_graph(g, copy.deepcopy(clip), args)
            layer = {t: prof["layer"].get(t, 0.5) for n in qlist for t in n.output}
            model_cos = {k: list(v) for k, v in prof["model"].items()}
            h = _Lines()
            RU.logger.addHandler(h)
            RU.logger.setLevel(logging.INFO)
            RP.show_model_profiling_res(g, layer, model_cos, qlist, args)
            RP.show_model_ranges(g, act, weight, args)
            RP.weight_need_perchannel(g, args)
            RU.logger.removeHandler(h)
            out[platform] = {"lines": h.lines, "layer": layer}
        json.dump(out, open(os.path.join(d, "reports.json"), "w"), indent=0)
        print(mname, {p: len(v["lines"]) for p, v in out.items()})


if __name__ == "__main__":
    main()

"""Fixture for the per-rank clip-value / profiling files and their rank-0 combine
(dipoorlet/utils.py:313-412): the reference's own save_clip_val / reduce_clip_val / load_clip_val /
save_profiling_res / reduce_profiling_res on seeded 3-rank inputs -> tests/golden/persistence.json
(inputs, every file's text, reduced values with their Python / NumPy type names).

    python oracle/gen_golden_persistence.py     # build container only; the fixture is committed
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


def main():
    ref_shim.install()
    import dipoorlet.utils as RU
    rng = np.random.default_rng(3)
    names = ["input", "conv1_out", "relu1_out", "fc_out"]
    wnames = {"conv1.weight": 4, "conv1.bias": 4, "fc.weight": 3}
    ranks = 3
    act_ranks = [{n: [np.float32(-abs(rng.normal())), np.float32(abs(rng.normal()) + 0.1)] for n in names}
                 for _ in range(ranks)]
    weight = {n: [(-np.abs(rng.normal(size=c))).astype(np.float32), np.abs(rng.normal(size=c)).astype(np.float32)]
              for n, c in wnames.items()}
    layer_ranks = [{n: np.float64(1 - abs(rng.normal()) * 1e-3) for n in names[1:]} for _ in range(ranks)]
    model_ranks = [{"fc_out": [np.float64(1 - abs(rng.normal()) * 1e-4), np.float64(1 - abs(rng.normal()) * 1e-3)]}
                   for _ in range(ranks)]
    out = {"ranks": ranks,
           "act_ranks": [{k: [float(v[0]), float(v[1])] for k, v in a.items()} for a in act_ranks],
           "weight": {k: [v[0].tolist(), v[1].tolist()] for k, v in weight.items()},
           "layer_ranks": [{k: float(v) for k, v in d.items()} for d in layer_ranks],
           "model_ranks": [{k: [float(v[0]), float(v[1])] for k, v in d.items()} for d in model_ranks],
           "cases": {}}
    for deploy in ("trt", "snpe"):            # per-channel weights / per-layer weights on reload
        for algo in ("minmax", "hist"):
            tmp = tempfile.mkdtemp(prefix="dpl_gold_persist_")
            args = types.SimpleNamespace(output_dir=tmp, act_quant=algo, deploy=deploy, model_type=None)
            for r in range(ranks):
                RU.save_clip_val(copy.deepcopy(act_ranks[r]), copy.deepcopy(weight), args,
                                 act_fname=f"act_clip_val.json.rank{r}", weight_fname=f"weight_clip_val.json.rank{r}")
            RU.reduce_clip_val(ranks, args)
            act, w = RU.load_clip_val(args)
            files = {f: open(os.path.join(tmp, f)).read() for f in sorted(os.listdir(tmp))}
            out["cases"][f"{deploy}|{algo}"] = {
                "files": files,
                "act": {k: [float(v[0]), float(v[1]), type(v[0]).__name__] for k, v in act.items()},
                "weight": {k: [np.asarray(v[0]).tolist(), np.asarray(v[1]).tolist(), type(v[0]).__name__,
                               list(np.asarray(v[0]).shape)] for k, v in w.items()}}
    # profiling files: the reference takes the rank from torch.distributed
    tmp = tempfile.mkdtemp(prefix="dpl_gold_persist_")
    args = types.SimpleNamespace(output_dir=tmp, model_type=None)
    real_rank = RU.dist.get_rank
    try:
        for r in range(ranks):
            RU.dist.get_rank = lambda r=r: r
            RU.save_profiling_res(copy.deepcopy(layer_ranks[r]), copy.deepcopy(model_ranks[r]), args)
    finally:
        RU.dist.get_rank = real_rank
    layer, model = RU.reduce_profiling_res(ranks, args)
    out["profiling"] = {"files": {f: open(os.path.join(tmp, f)).read() for f in sorted(os.listdir(tmp))},
                        "layer": layer, "model": model}
    path = os.path.join(ROOT, "tests", "golden", "persistence.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, sorted(out["cases"]), sorted(out["profiling"]["files"]))


if __name__ == "__main__":
    main()

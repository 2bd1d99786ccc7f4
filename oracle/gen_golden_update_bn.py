"""Golden fixture for `--update_bn`: runs the REFERENCE's own weight_calibration(update_bn=True)
(/root/reference/dipoorlet, unmodified, under oracle/ref_shim) on a small seeded pre-activation net whose
BatchNormalization nodes cannot be folded, and writes tests/golden/tiny_preact/: model.onnx, images.npy,
wt_update_bn.npz (the running mean / "var" initializers it rewrote — the reference feeds np.std into the
variance slot, update_bn.py:17) and wt_update_bn_clip.json (the clip values of the re-calibration that follows).

    python oracle/gen_golden_update_bn.py        # build container only; the fixtures are committed
"""
import json
import logging
import os
import shutil
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchvision  # noqa: E402,F401
import torch  # noqa: E402,F401

from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD, N_IMG, ADA_BS, ADA_EPOCH, jsonable_clip  # noqa: E402


def main():
    from dipoorlet_b200 import onnx_lite as ol, workloads as W
    ref_shim.install()
    import dipoorlet.tensor_cali as RTC
    import dipoorlet.utils as RU
    from dipoorlet.weight_transform import weight_calibration
    logging.getLogger("dipoorlet").setLevel(logging.WARNING)
    out = os.path.join(GOLD, "tiny_preact")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    model = W.build_preact_net(seed=13)
    ol.save(model, os.path.join(out, "model.onnx"))
    images = W.synthetic_images(N_IMG, (3, 32, 32), seed=23)
    np.save(os.path.join(out, "images.npy"), images)
    for algo in ("minmax", "hist"):
        tmp = tempfile.mkdtemp(prefix="dpl_gold_bn_")
        W.write_input_dir(images, os.path.join(tmp, "data"), "input")
        g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, "trt", None)
        args = types.SimpleNamespace(
            input_dir=os.path.join(tmp, "data"), output_dir=tmp, data_num=N_IMG, world_size=1, rank=0,
            local_rank=0, act_quant=algo, deploy="trt", bins=2048, threshold=0.99999,
            optim_transformer=False, skip_layers=[], bc=False, we=False, update_bn=True, adaround=False,
            brecq=False, drop=False, sparse=False, ada_bs=ADA_BS, ada_epoch=ADA_EPOCH, model=None,
            model_type=None, savefp=False, skip_prof_layer=False)
        act, w = RTC.tensor_calibration(g, args)
        graph, graph_ori, act2, w2 = weight_calibration(g, act, w, args)
        before = dict(model.graph.initializers)
        changed = {t.name: np.asarray(t.array) for t in graph.graph.initializer
                   if t.name not in before or not np.array_equal(before[t.name], np.asarray(t.array))}
        suffix = "" if algo == "minmax" else "_" + algo
        np.savez_compressed(os.path.join(out, f"wt_update_bn{suffix}.npz"), **changed)
        json.dump({"act_before": jsonable_clip(act), "act": jsonable_clip(act2), "weight": jsonable_clip(w2)},
                  open(os.path.join(out, f"wt_update_bn{suffix}_clip.json"), "w"), indent=1)
        shutil.rmtree(tmp)
        print(algo, "changed initializers:", sorted(changed))


if __name__ == "__main__":
    main()

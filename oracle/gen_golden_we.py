"""Golden fixture for `--we` (weight equalisation): runs the REFERENCE's own weight_calibration(we=True)
(/root/reference/dipoorlet, unmodified, under oracle/ref_shim) on the two small seeded models of
gen_golden.py and writes tests/golden/<model>/wt_we.npz (the initializers it changed) and
wt_we_clip.json (the activation clip values of the re-calibration that follows, minmax).

    python oracle/gen_golden_we.py        # build container only; the fixtures are committed
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
    for mname in ("tiny_r50", "tiny_mbv2"):
        out = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(out, "model.onnx"))
        images = np.load(os.path.join(out, "images.npy"))
        tmp = tempfile.mkdtemp(prefix="dpl_gold_we_")
        W.write_input_dir(images, os.path.join(tmp, "data"), "input")
        g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, "trt", None)
        args = types.SimpleNamespace(
            input_dir=os.path.join(tmp, "data"), output_dir=tmp, data_num=N_IMG, world_size=1, rank=0,
            local_rank=0, act_quant="minmax", deploy="trt", bins=2048, threshold=0.99999,
            optim_transformer=False, skip_layers=[], bc=False, we=True, update_bn=False, adaround=False,
            brecq=False, drop=False, sparse=False, ada_bs=ADA_BS, ada_epoch=ADA_EPOCH, model=None,
            model_type=None, savefp=False, skip_prof_layer=False)
        act, w = RTC.tensor_calibration(g, args)
        graph, graph_ori, act2, w2 = weight_calibration(g, act, w, args)
        before = dict(model.graph.initializers)
        changed = {t.name: np.asarray(t.array) for t in graph.graph.initializer
                   if t.name not in before or not np.array_equal(before[t.name], np.asarray(t.array))}
        np.savez_compressed(os.path.join(out, "wt_we.npz"), **changed)
        json.dump(jsonable_clip(act2), open(os.path.join(out, "wt_we_clip.json"), "w"), indent=1)
        shutil.rmtree(tmp)
        print(mname, "changed initializers:", len(changed))


if __name__ == "__main__":
    main()

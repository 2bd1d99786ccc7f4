"""Fixture for the vendor deploy writers (dipoorlet/deploy/deploy_{atlas,imx,magicmind,snpe,ti,rv,stpu}.py;
stpu also with --stpu_wg):
the reference's own calibrate -> save / reduce / load clip values -> to_deploy on the two small seeded
models, per platform -> tests/golden/<model>/deploy_vendors.json (clip-value file texts in, deploy file
texts out).

    python oracle/gen_golden_deploy.py       # build container only; the fixtures are committed
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
from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import GOLD, N_IMG  # noqa: E402

PLATFORMS = ("atlas", "imx", "magicmind", "snpe", "ti", "rv", "stpu")   # --stpu_wg crashes in the reference (NodeProto.get_attribute_value)


def main():
    from dipoorlet_b200 import onnx_lite as ol, workloads as W
    ref_shim.install()
    import dipoorlet.tensor_cali as RTC
    import dipoorlet.utils as RU
    from dipoorlet.deploy import to_deploy
    logging.getLogger("dipoorlet").setLevel(logging.WARNING)
    for mname in ("tiny_r50", "tiny_mbv2"):
        d = os.path.join(GOLD, mname)
        model = ol.load(os.path.join(d, "model.onnx"))
        images = np.load(os.path.join(d, "images.npy"))
        out = {}
        for case in PLATFORMS:
            platform, wg = case.split("+")[0], case.endswith("+wg")
            tmp = tempfile.mkdtemp(prefix="dpl_gold_deploy_")
            W.write_input_dir(images, os.path.join(tmp, "data"), "input")
            g = RU.ONNXGraph(ref_shim.from_lite(model), tmp, platform, None)
            args = types.SimpleNamespace(
                input_dir=os.path.join(tmp, "data"), output_dir=tmp, data_num=N_IMG, world_size=1, rank=0,
                local_rank=0, act_quant="minmax", deploy=platform, bins=2048, threshold=0.99999,
                optim_transformer=False, skip_layers=[], model=None, model_type=None, stpu_wg=wg)
            act, w = RTC.tensor_calibration(g, args)
            RU.save_clip_val(act, w, args, act_fname="act_clip_val.json.rank0", weight_fname="weight_clip_val.json.rank0")
            RU.reduce_clip_val(1, args)
            act, w = RU.load_clip_val(args)
            before = set(os.listdir(tmp))
            to_deploy(g, act, w, args)
            written = sorted(set(os.listdir(tmp)) - before)
            out[case] = {"act_clip_val": open(os.path.join(tmp, "act_clip_val.json")).read(),
                             "weight_clip_val": open(os.path.join(tmp, "weight_clip_val.json")).read(),
                             "files": {f: open(os.path.join(tmp, f)).read() for f in written}}
            shutil.rmtree(tmp)
        json.dump(out, open(os.path.join(d, "deploy_vendors.json"), "w"))
        print(mname, {p: sorted(v["files"]) for p, v in out.items()})


if __name__ == "__main__":
    main()

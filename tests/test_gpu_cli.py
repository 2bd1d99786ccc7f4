"""The drop-in CLI end to end on the GPU: `python -m dipoorlet_b200 -M model.onnx -I dir -N n -A .. -D trt`
on the fixture model and images, files compared with the ones the REFERENCE wrote for the
same inputs (tests/golden/*/calibration.json)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("mname,algo", [("tiny_r50", "minmax"), ("tiny_r50", "hist"), ("tiny_mbv2", "mse")])
def test_cli_writes_reference_files(dpl_built, tmp_path, mname, algo):
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.__main__ import main
    d = os.path.join(GOLD, mname)
    images = np.load(os.path.join(d, "images.npy"))
    W.write_input_dir(images, str(tmp_path / "data"), "input")
    out = str(tmp_path / "out")
    main(["-M", os.path.join(d, "model.onnx"), "-I", str(tmp_path / "data"), "-O", out, "-N", str(images.shape[0]),
          "-A", algo, "-D", "trt", "--bins", "2048"])
    for f in ("act_clip_val.json", "weight_clip_val.json", "trt_clip_val.json", "quant_model.onnx",
              "layer_res.json.rank0", "model_res.json.rank0"):
        assert os.path.exists(os.path.join(out, f)), f
    gold = json.load(open(os.path.join(d, "calibration.json")))[algo]
    got = json.load(open(os.path.join(out, "trt_clip_val.json")))["blob_range"]
    want = json.loads(gold["trt_clip_val_json"])["blob_range"]
    assert list(got) == list(want)
    act = json.load(open(os.path.join(out, "act_clip_val.json")))
    ref_act = gold["act"]
    for k in want:
        dm = max(abs(ref_act[k][0]), abs(ref_act[k][1]), 1e-12)
        tol = 2e-5 * dm + (1.5 * dm / 2048 if algo == "hist" else 0.0)   # GPU vs CPU forward rounding; +-1 bin
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])
        assert abs(act[k][0] - ref_act[k][0]) <= tol and abs(act[k][1] - ref_act[k][1]) <= tol, k
    cos = json.load(open(os.path.join(out, "model_res.json.rank0")))
    assert all(v[0] > 0.99 for v in cos.values())


def test_cli_adaround_and_bc(dpl_built, tmp_path):
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.__main__ import main
    d = os.path.join(GOLD, "tiny_r50")
    images = np.load(os.path.join(d, "images.npy"))
    W.write_input_dir(images, str(tmp_path / "data"), "input")
    out = str(tmp_path / "out")
    main(["-M", os.path.join(d, "model.onnx"), "-I", str(tmp_path / "data"), "-O", out, "-N", "8", "-A", "minmax",
          "-D", "trt", "--bc", "--adaround", "--ada_bs", "4", "--ada_epoch", "12"])
    assert os.path.exists(os.path.join(out, "update_bias_model.onnx"))
    m = ol.load(os.path.join(out, "adaround.onnx"))
    assert len(m.graph.nodes) == len(ol.load(os.path.join(d, "model.onnx")).graph.nodes)

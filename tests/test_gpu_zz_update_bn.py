"""`--update_bn` on the GPU against the running statistics the REFERENCE wrote for the same model and images
(tests/golden/tiny_preact, oracle/gen_golden_update_bn.py). Runs last among the GPU files: it is a §8 f4 widening,
not the measured path. The GPU forward and the reference's CPU forward round differently, so a few activations
land one quantisation step apart; on this model that moves a running mean by ~1e-5 (tests/test_host_logic.py
quantifies it), hence the tolerance."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_preact")


def test_update_bn_vs_reference(dpl_built, tmp_path):
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import find_clip_val_minmax_weight
    from dipoorlet_b200.weight_transform import weight_calibration
    model = ol.load(os.path.join(GOLD, "model.onnx"))
    images = np.load(os.path.join(GOLD, "images.npy"))
    gold = json.load(open(os.path.join(GOLD, "wt_update_bn_clip.json")))
    stats = np.load(os.path.join(GOLD, "wt_update_bn.npz"))
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=images.shape[0], deploy="trt",
                     act_quant="minmax", output_dir=str(tmp_path), calib_bs=8, update_bn=True)
    act = {k: [np.float32(v[0]), np.float32(v[1])] for k, v in gold["act_before"].items()}
    weight = find_clip_val_minmax_weight(graph, args)
    g2, _, act2, weight2 = weight_calibration(graph, act, weight, args)
    assert os.path.exists(os.path.join(str(tmp_path), "update_bn_model.onnx"))
    for name in stats.files:
        got = g2.get_initializer(name)
        assert np.allclose(got, stats[name], rtol=1e-3, atol=1e-4), (name, np.abs(got - stats[name]).max())
        assert not np.allclose(model.graph.initializers[name], stats[name], rtol=1e-2, atol=1e-3)
    assert list(act2) == list(gold["act"])
    for k, v in gold["act"].items():
        scale = max(abs(v[0]), abs(v[1]), 1e-3)
        assert abs(act2[k][0] - v[0]) <= 2e-3 * scale and abs(act2[k][1] - v[1]) <= 2e-3 * scale, (k, act2[k], v)

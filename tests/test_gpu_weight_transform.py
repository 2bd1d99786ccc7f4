"""The product's weight transforms and profiling on the GPU against the weights / cosines
the REFERENCE produced for the same model, images and hyper-parameters (tests/golden).
The GPU forward (cuDNN fp32) and the reference's CPU forward round differently, so learned
rounding may flip for weights whose alpha ends within ~1e-3 of zero: require >= 99.5 %
identical weights and every difference to be exactly one quantisation step."""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ADA_BS, ADA_EPOCH = 4, 12


def _setup(mname, tmp_path, **kw):
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    d = os.path.join(GOLD, mname)
    model = ol.load(os.path.join(d, "model.onnx"))
    images = np.load(os.path.join(d, "images.npy"))
    calib = json.load(open(os.path.join(d, "calibration.json")))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    act = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    weight = {}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        weight.setdefault(name, [None, None])[int(i)] = gold_w[key].astype(np.float64)
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=images.shape[0], deploy="trt",
                     output_dir=str(tmp_path), ada_bs=ADA_BS, ada_epoch=ADA_EPOCH, calib_bs=8, **kw)
    return d, model, graph, act, weight, args


def _check_rounded(graph_out, gold, scales_from):
    total = same = 0
    for k in gold.files:
        got, want = graph_out.get_initializer(k), gold[k]
        assert got.shape == want.shape, k
        diff = np.abs(got - want)
        step = np.abs(scales_from[k]).reshape([-1] + [1] * (want.ndim - 1))
        bad = diff > 1e-7
        assert np.all(np.abs(diff[bad] - np.broadcast_to(step, diff.shape)[bad]) <= 1e-5 * np.broadcast_to(step, diff.shape)[bad] + 1e-9), k
        total += want.size
        same += int((~bad).sum())
    return same / total


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
@pytest.mark.parametrize("algo", ["adaround", "brecq"])
def test_learned_rounding_vs_reference(dpl_built, mname, algo, tmp_path):
    from dipoorlet_b200.weight_transform import weight_calibration
    d, model, graph, act, weight, args = _setup(mname, tmp_path, **{algo: True})
    graph_wt, graph_ori, _, _ = weight_calibration(graph, act, copy.deepcopy(weight), args)
    gold = np.load(os.path.join(d, f"wt_{algo}.npz"))
    scales = {k: np.maximum(np.abs(weight[k][0]), np.abs(weight[k][1])) / 127 for k in gold.files}
    frac = _check_rounded(graph_wt, gold, scales)
    assert frac >= 0.995, frac
    assert os.path.exists(os.path.join(str(tmp_path), f"{algo}.onnx"))


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_qdrop_runs_and_rounds(dpl_built, mname, tmp_path):
    """QDrop uses its own Bernoulli stream (not torch's Philox), so only structural checks:
    every learnable weight lands on its quantisation grid."""
    from dipoorlet_b200.weight_transform import weight_calibration
    d, model, graph, act, weight, args = _setup(mname, tmp_path, brecq=True, drop=True)
    graph_wt, _, _, _ = weight_calibration(graph, act, copy.deepcopy(weight), args)
    for node in graph.graph.node:
        if node.op_type in ("Conv", "Gemm"):
            w = graph_wt.get_initializer(node.input[1])
            s = (np.maximum(np.abs(weight[node.input[1]][0]), np.abs(weight[node.input[1]][1])) / 127).astype(np.float32)
            s = np.where(s == 0, 1, s).reshape([-1] + [1] * (w.ndim - 1))
            q = w / s
            assert np.allclose(q, np.round(q), atol=1e-3), node.name
            assert np.abs(q).max() <= 127 + 1e-3


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_bias_correction_vs_reference(dpl_built, mname, tmp_path):
    from dipoorlet_b200.weight_transform import weight_calibration
    d, model, graph, act, weight, args = _setup(mname, tmp_path, bc=True)
    graph_wt, _, _, w2 = weight_calibration(graph, act, copy.deepcopy(weight), args)
    gold = np.load(os.path.join(d, "wt_bc.npz"))
    for k in gold.files:
        got, want = graph_wt.get_initializer(k), gold[k]
        tol = 2e-4 * max(np.abs(want).max(), 1e-3)
        assert np.allclose(got, want, rtol=0, atol=tol), (k, np.abs(got - want).max(), tol)


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_profiling_vs_reference(dpl_built, mname, tmp_path):
    from dipoorlet_b200.profiling import quantize_profiling_multipass
    d, model, graph, act, weight, args = _setup(mname, tmp_path)
    # profile the bias-corrected model the reference profiled (weights from the fixture)
    gold_bc = np.load(os.path.join(d, "wt_bc.npz"))
    from dipoorlet_b200.graph import ONNXGraph
    g2 = ONNXGraph()
    g2.copy_from(graph)
    for k in gold_bc.files:
        g2.set_initializer(k, gold_bc[k])
    from dipoorlet_b200.tensor_cali import find_clip_val_minmax_weight
    w2 = find_clip_val_minmax_weight(g2, args)
    layer, model_cos, _ = quantize_profiling_multipass(g2, graph, act, w2, args)
    gold = json.load(open(os.path.join(d, "profiling_bc.json")))
    assert list(layer) == list(gold["layer"])
    for k, v in gold["layer"].items():
        assert abs(float(layer[k]) - v) < 2e-4, (k, layer[k], v)
    for k, v in gold["model"].items():
        assert abs(float(model_cos[k][0]) - v[0]) < 2e-4 and abs(float(model_cos[k][1]) - v[1]) < 2e-4

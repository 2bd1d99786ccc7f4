"""The oracle's weight-transform restatement (bias correction, adaround, brecq) against the
weights the REFERENCE ITSELF produced (tests/golden/*/wt_*.npz): bit-exact on the CPU."""
import copy
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ADA_BS, ADA_EPOCH = 4, 12   # oracle/gen_golden.py


def _setup(mname, tmp_path):
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    d = os.path.join(GOLD, mname)
    model = ol.load(os.path.join(d, "model.onnx"))
    images = np.load(os.path.join(d, "images.npy"))
    calib = json.load(open(os.path.join(d, "calibration.json")))
    gold_w = np.load(os.path.join(d, "weight_clip.npz"))
    clip = {k: [np.float64(v[0]), np.float64(v[1])] for k, v in calib["minmax"]["act"].items()}
    for key in gold_w.files:
        name, i = key.rsplit("|", 1)
        clip.setdefault(name, [None, None])[int(i)] = gold_w[key].astype(np.float64)
    graph = ONNXGraph(model, str(tmp_path), "trt")
    args = make_args(input_dir="unused", data_num=images.shape[0], deploy="trt", output_dir=str(tmp_path))
    return d, model, images, clip, graph, args


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_oracle_adaround_equals_reference(mname, tmp_path):
    from dipoorlet_b200.quantize import quant_graph
    from oracle import wt_pipeline as WP
    d, model, images, clip, graph, args = _setup(mname, tmp_path)
    gq, _ = quant_graph(graph, copy.deepcopy(clip), args)
    got = WP.adaround(model, gq.model, images, clip, ADA_BS, ADA_EPOCH)
    gold = np.load(os.path.join(d, "wt_adaround.npz"))
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(got[k], gold[k]), (k, np.abs(got[k] - gold[k]).max())


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_oracle_brecq_equals_reference(mname, tmp_path):
    from dipoorlet_b200.quantize import quant_graph
    from oracle import wt_pipeline as WP
    d, model, images, clip, graph, args = _setup(mname, tmp_path)
    gq, _ = quant_graph(graph, copy.deepcopy(clip), args)
    got = WP.brecq(model, gq.model, images, clip, ADA_BS, ADA_EPOCH)
    gold = np.load(os.path.join(d, "wt_brecq.npz"))
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(got[k], gold[k]), (k, np.abs(got[k] - gold[k]).max())


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_oracle_bias_correction_equals_reference(mname, tmp_path):
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.quantize import quant_graph
    from oracle import wt_pipeline as WP
    d, model, images, clip, graph, args = _setup(mname, tmp_path)

    def build_q(m):
        g = ONNXGraph(copy.deepcopy(m), str(tmp_path), "trt")
        return quant_graph(g, copy.deepcopy(clip), args)[0].model

    got = WP.bias_correction(model, build_q, images)
    gold = np.load(os.path.join(d, "wt_bc.npz"))
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(got[k].astype(np.float32), gold[k].astype(np.float32)), \
            (k, np.abs(got[k] - gold[k]).max())


@pytest.mark.parametrize("mname", ["tiny_r50", "tiny_mbv2"])
def test_weight_equalization_equals_reference(mname, tmp_path):
    """`--we` (weight_equalization.py:36-94): the product's array-expression form must rewrite exactly
    the initializers the reference rewrote, bit for bit (tests/golden/*/wt_we.npz, produced by running
    the reference's own weight_calibration(we=True), oracle/gen_golden_we.py)."""
    from dipoorlet_b200.weight_transform.weight_equalization import node_has_equalized, weight_equalization
    d, model, images, clip, graph, args = _setup(mname, tmp_path)
    before = {k: np.asarray(v).copy() for k, v in model.graph.initializers.items()}
    g2 = weight_equalization(graph, args)
    assert os.path.exists(os.path.join(str(tmp_path), "weight_equal_model.onnx"))
    after = {k: np.asarray(g2.get_initializer(k)) for k in before}
    changed = {k: v for k, v in after.items() if not np.array_equal(before[k], v)}
    gold = np.load(os.path.join(d, "wt_we.npz"))
    assert sorted(changed) == sorted(gold.files)
    for k in gold.files:
        assert changed[k].dtype == gold[k].dtype and np.array_equal(changed[k], gold[k]), \
            (k, np.abs(changed[k] - gold[k]).max())
    # the input graph is untouched (the reference equalises a copy, :37-38)
    for k, v in before.items():
        assert np.array_equal(np.asarray(graph.get_initializer(k)), v)
    assert any(node_has_equalized(graph, n) for n in graph.graph.node if n.op_type == "Conv")

"""End-to-end calibration through the plugin API on a reduced ResNet-50, against the
oracle: (1) statistics computed on the SAME blobs (engine output copied to the host) must
be bit-exact / within 1e-5; (2) against the CPU-forward oracle the clip files must agree
up to the forward's rounding (different conv summation order)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 12
IMG = 64


@pytest.fixture(scope="module")
def setup(dpl_built, tmp_path_factory):
    import torch
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.engine import Engine
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    out = str(tmp_path_factory.mktemp("calib"))
    model = W.build_resnet50(blocks=[1, 1, 1, 1], width=16, num_classes=10, image=IMG)
    graph = ONNXGraph(model, out, "trt")
    images = W.synthetic_images(N, (3, IMG, IMG), seed=3)
    args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=N, deploy="trt",
                     output_dir=out, calib_bs=5)
    # "the same blobs": the session runs batches of calib_bs images, and cuDNN may pick a
    # different (differently rounding) algorithm per batch size, so reproduce its batching
    eng = Engine(graph, torch.device("cuda", 0))
    blobs = {}
    for b0 in range(0, N, args.calib_bs):
        part = eng.run({"input": torch.from_numpy(images[b0:b0 + args.calib_bs, 0]).cuda()}, want="all")
        for k, v in part.items():
            blobs.setdefault(k, []).extend(v[i].cpu().numpy()[None] for i in range(v.shape[0]))
    return dict(graph=graph, args=args, images=images, blobs=blobs, model=model, out=out)


def test_minmax_same_blobs(setup):
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = setup["args"]
    args.act_quant = "minmax"
    act, weight = tensor_calibration(setup["graph"], args)
    ref = O.clip_minmax(O.minmax_stats(setup["blobs"]))
    assert list(act) == list(ref)
    for k in ref:
        assert act[k][0] == ref[k][0] and act[k][1] == ref[k][1], k
        assert isinstance(act[k][0], np.float32)
    wref = O.weight_minmax({n: setup["graph"].get_initializer(n) for n in weight})
    for k in weight:
        assert np.array_equal(weight[k][0], wref[k][0]) and np.array_equal(weight[k][1], wref[k][1])


def test_hist_same_blobs_bit_exact(setup):
    from dipoorlet_b200 import forward_net as fwd
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = setup["args"]
    args.act_quant = "hist"
    act, _ = tensor_calibration(setup["graph"], args)
    sess = fwd._session(setup["graph"], args)
    mm = O.minmax_stats(setup["blobs"])
    hist = O.hist_stats(setup["blobs"], mm, 2048)
    ref, sel = O.clip_hist(mm, hist, 2048, args.threshold, return_bins=True)
    counts = sess.counts.cpu().numpy()
    for i, k in enumerate(ref):
        assert np.array_equal(counts[i], np.stack(hist[k]).sum(0)), k
        assert act[k][0] == ref[k][0] and act[k][1] == ref[k][1], k


def test_mse_same_blobs(setup):
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = setup["args"]
    args.act_quant = "mse"
    act, _ = tensor_calibration(setup["graph"], args)
    ref = O.clip_octav(O.octav_stats(setup["blobs"]))
    for k in ref:
        assert np.allclose(act[k][0], ref[k][0], rtol=1e-5, atol=0), (k, act[k], ref[k])
        assert np.allclose(act[k][1], ref[k][1], rtol=1e-5, atol=0), (k, act[k], ref[k])


@pytest.mark.parametrize("algo", ["minmax", "hist", "mse"])
def test_trt_file_vs_cpu_oracle(setup, algo):
    """Whole pipeline incl. the forward against the CPU oracle (torch-CPU forward + NumPy
    statistics). The two forwards round differently, so: minmax / mse within 1e-4
    relative, hist within one histogram bin (1/2048 of the range)."""
    import copy
    from dipoorlet_b200.deploy import to_deploy
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from dipoorlet_b200.utils import load_clip_val, save_clip_val
    from oracle import forward as OF
    from oracle import stats as O
    args = setup["args"]
    args.act_quant = algo
    act, weight = tensor_calibration(setup["graph"], args)
    save_clip_val(act, weight, args)
    act, weight = load_clip_val(args)
    to_deploy(setup["graph"], act, weight, args)
    got = json.load(open(os.path.join(setup["out"], "trt_clip_val.json")))["blob_range"]

    blobs = OF.blobs_for_images(setup["model"], {"input": setup["images"]}, N)
    mm = O.minmax_stats(blobs)
    if algo == "minmax":
        clip = O.clip_minmax(mm)
    elif algo == "hist":
        clip = O.clip_hist(mm, O.hist_stats(blobs, mm, 2048), 2048, args.threshold)
    else:
        clip = O.clip_octav(O.octav_stats(blobs))
    want = O.trt_blob_range(clip)
    assert list(got) == list(want)
    for k in want:
        tol = 1e-4 * max(abs(want[k]), 1e-12)
        if algo == "hist":  # the percentile may land one bin away
            tol += 1.5 * float(O.data_max_of(mm[k])) / 2048
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])


def test_weight_ranges_on_device_bit_exact(setup):
    """The per-channel weight min/max (basic_algorithm.py:72-91) come from one K1 launch over
    the engine's resident weights; they must equal the NumPy per-channel reduction exactly."""
    from dipoorlet_b200.platform_settings import LAYER_HAS_WEIGHT
    from dipoorlet_b200.tensor_cali import find_clip_val_minmax_weight
    from oracle import stats as O
    graph, args = setup["graph"], setup["args"]
    got = find_clip_val_minmax_weight(graph, args)
    weights = {}
    for node in graph.graph.node:
        if node.op_type in LAYER_HAS_WEIGHT:
            for name in node.input[1:]:
                weights.setdefault(name, graph.get_initializer(name))
    want = O.weight_minmax(weights)
    assert list(got) == list(want)
    for name in want:
        np.testing.assert_array_equal(got[name][0], np.asarray(want[name][0]))
        np.testing.assert_array_equal(got[name][1], np.asarray(want[name][1]))

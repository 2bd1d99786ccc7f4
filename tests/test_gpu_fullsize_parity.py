"""Full-size parity: the REAL ResNet-50 (224 x 224, 2048-wide, 123 blobs) and MobileNetV2 (101 blobs) of
BASELINE.json's configs through tensor_calibration on the GPU vs the CPU oracle (torch-CPU fp32 forward +
the reference's NumPy statistics: forward_net.py:192-342, basic_algorithm.py:13-69), 8 images.

Two comparisons per calibrator:
  same blobs   - the oracle's statistics on the blobs the GPU forward produced (copied to the host):
                 min / max and histogram counts bit-exact, the percentile bin and clip bit-exact, mse <= 1e-5;
  CPU forward  - against the oracle's own CPU forward: the clip file's relative error per blob (the two
                 forwards sum in different orders) and, for hist, how many blobs select another percentile bin.
The measured numbers are written to gpurun_out/fullsize_parity_<model>.json (copied to profiles/)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["r50", "mbv2"])
def full(request, dpl_built, tmp_path_factory):
    import torch
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.engine import Engine
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    from oracle import forward as OF
    name = request.param
    out = str(tmp_path_factory.mktemp("full_" + name))
    model = W.build_resnet50(seed=0) if name == "r50" else W.build_mobilenetv2(seed=0)
    graph = ONNXGraph(model, out, "trt")
    images = W.synthetic_images(N, seed=0)
    args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=N, deploy="trt", output_dir=out,
                     calib_bs=N, bins=2048)
    eng = Engine(graph, torch.device("cuda", 0))
    part = eng.run({"input": torch.from_numpy(images[:, 0]).cuda()}, want="all")
    gpu_blobs = {k: [v[i].cpu().numpy()[None] for i in range(v.shape[0])] for k, v in part.items()}
    del part, eng
    torch.cuda.empty_cache()
    cpu_blobs = OF.blobs_for_images(model, {"input": images}, N)
    report = {"model": name, "images": N, "blobs": len(cpu_blobs)}
    yield dict(name=name, graph=graph, args=args, gpu_blobs=gpu_blobs, cpu_blobs=cpu_blobs, report=report)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fullsize_parity_%s.json" % name), "w") as f:
        json.dump(report, f, indent=1)


def _rel(got, want):
    return abs(float(got) - float(want)) / max(abs(float(want)), 1e-30)


def _trt(clip):
    return {k: max(-float(v[0]), float(v[1])) for k, v in clip.items()}


def test_forward_blobs_close(full):
    """The GPU forward (3xTF32 tensor-core tiles, fp32 accumulate) against torch-CPU fp32, blob by blob."""
    worst, per_blob = 0.0, {}
    for k, cpu in full["cpu_blobs"].items():
        g, c = np.concatenate(full["gpu_blobs"][k]), np.concatenate(cpu).reshape(-1)
        err = np.abs(g.reshape(-1) - c).max() / max(np.abs(c).max(), 1e-30)
        per_blob[k] = float(err)
        worst = max(worst, float(err))
    full["report"]["forward_max_abs_err_over_blob_max"] = worst
    full["report"]["forward_err_per_blob"] = per_blob
    assert worst <= 2e-5, worst


def test_minmax(full):
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = full["args"]
    args.act_quant = "minmax"
    act, _ = tensor_calibration(full["graph"], args)
    same = O.clip_minmax(O.minmax_stats(full["gpu_blobs"]))
    assert list(act) == list(same)
    for k in same:
        assert act[k][0] == same[k][0] and act[k][1] == same[k][1], k
    got, want = _trt(act), _trt(O.clip_minmax(O.minmax_stats(full["cpu_blobs"])))
    rel = {k: _rel(got[k], want[k]) for k in want}
    full["report"]["minmax"] = {"same_blobs": "bit-exact", "cpu_forward_max_rel": max(rel.values()),
                                "cpu_forward_blobs_over_1e-5": sum(v > 1e-5 for v in rel.values()),
                                "worst_blob": max(rel, key=rel.get)}
    assert max(rel.values()) <= 1e-5, full["report"]["minmax"]


def test_hist(full):
    from dipoorlet_b200 import forward_net as fwd
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = full["args"]
    args.act_quant = "hist"
    act, _ = tensor_calibration(full["graph"], args)
    sess = fwd._session(full["graph"], args)
    counts = sess.counts.cpu().numpy()
    mm = O.minmax_stats(full["gpu_blobs"])
    hist = O.hist_stats(full["gpu_blobs"], mm, 2048)
    same, sel_same = O.clip_hist(mm, hist, 2048, args.threshold, return_bins=True)
    for i, k in enumerate(same):
        assert np.array_equal(counts[i], np.stack(hist[k]).sum(0)), k
        assert act[k][0] == same[k][0] and act[k][1] == same[k][1], k
    mmc = O.minmax_stats(full["cpu_blobs"])
    want_clip, sel_cpu = O.clip_hist(mmc, O.hist_stats(full["cpu_blobs"], mmc, 2048), 2048, args.threshold,
                                     return_bins=True)
    got, want = _trt(act), _trt(want_clip)
    flips = {k: int(sel_same[k]) - int(sel_cpu[k]) for k in sel_cpu if sel_same[k] != sel_cpu[k]}
    rel = {k: _rel(got[k], want[k]) for k in want}
    rel_same_bin = [rel[k] for k in rel if k not in flips]
    full["report"]["hist"] = {"same_blobs": "counts, percentile bin and clip bit-exact",
                              "cpu_forward_bin_flips": len(flips), "flips": flips,
                              "cpu_forward_max_rel_where_same_bin": max(rel_same_bin),
                              "cpu_forward_max_rel": max(rel.values())}
    assert all(abs(d) <= 1 for d in flips.values()), flips          # never more than the neighbouring bin
    assert max(rel_same_bin) <= 1e-5, full["report"]["hist"]


def test_mse(full):
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from oracle import stats as O
    args = full["args"]
    args.act_quant = "mse"
    act, _ = tensor_calibration(full["graph"], args)
    same = O.clip_octav(O.octav_stats(full["gpu_blobs"]))
    rs = max(max(_rel(act[k][0], same[k][0]), _rel(act[k][1], same[k][1])) for k in same)
    assert rs <= 1e-5, rs
    got, want = _trt(act), _trt(O.clip_octav(O.octav_stats(full["cpu_blobs"])))
    rel = {k: _rel(got[k], want[k]) for k in want}
    full["report"]["mse"] = {"same_blobs_max_rel": rs, "cpu_forward_max_rel": max(rel.values()),
                             "cpu_forward_blobs_over_1e-5": sum(v > 1e-5 for v in rel.values()),
                             "worst_blob": max(rel, key=rel.get)}
    assert max(rel.values()) <= 1e-5, full["report"]["mse"]

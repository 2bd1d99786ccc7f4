"""The forward engine's streaming operators (dpl_eltwise.cu) against torch on the same inputs —
bit-exact: max / add / compare are exact IEEE operations — and the blob arena of the
calibration session (same statistics with and without it)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 5, 4096, 8192 * 37 + 3])
def test_clip_and_relu_bit_exact(dpl_built, n):
    import torch
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(n, device="cuda", generator=g) * 4
    x[0] = float("nan")
    assert torch.equal(K.clip(x, 0.0, float("inf"))[1:], torch.relu(x)[1:])
    assert torch.isnan(K.clip(x, 0.0, float("inf"))[0])
    assert torch.equal(K.clip(x, 0.0, 6.0)[1:], torch.clamp(x, 0.0, 6.0)[1:])
    # misaligned views take the scalar path
    if n > 8:
        xv = x[1:]
        assert torch.equal(K.clip(xv.contiguous(), -1.0, 1.0), torch.clamp(xv, -1.0, 1.0))
        y = torch.empty(n + 1, device="cuda")[1:]
        K.clip(x[1:].clone(), 0.0, float("inf"), out=y[1:])
        assert torch.equal(y[1:], torch.relu(x[1:]))


@pytest.mark.parametrize("shape", [(3, 7, 5, 5), (8, 256, 56, 56), (2, 1000)])
def test_add_and_fused_relu_bit_exact(dpl_built, shape):
    import torch
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(shape, device="cuda", generator=g)
    b = torch.randn(shape, device="cuda", generator=g)
    assert torch.equal(K.add(a, b), a + b)
    r = torch.empty_like(a)
    y = K.add(a, b, out_relu=r)
    assert torch.equal(y, a + b) and torch.equal(r, torch.relu(a + b))


@pytest.mark.parametrize("cfg", [((4, 64, 112, 112), 3, 2, 1, False), ((3, 5, 15, 24), 3, 2, 1, False),
                                 ((2, 3, 8, 8), 3, 2, 1, False), ((2, 5, 13, 9), 2, 2, 0, False),
                                 ((2, 3, 14, 14), 3, 2, 0, True), ((1, 2, 7, 7), 3, 1, 1, False)])
def test_maxpool_bit_exact(dpl_built, cfg):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    shape, k, s, p, ceil_mode = cfg
    x = torch.randn(shape, device="cuda")
    want = F.max_pool2d(x, k, s, p, 1, ceil_mode)
    got = K.maxpool2d(x, (k, k), (s, s), p, p, want.shape[2], want.shape[3])
    assert torch.equal(got, want)
    if k == 3 and s == 2 and p == 1:        # the four-outputs-per-thread kernel: NaN propagates like torch's, range folds
        x[0, 0, 3, 5] = float("nan")
        x[-1, -1, 0, 0] = float("nan")
        want = F.max_pool2d(x, k, s, p, 1, ceil_mode)
        got = K.maxpool2d(x, (k, k), (s, s), p, p, want.shape[2], want.shape[3])
        assert torch.equal(torch.isnan(got), torch.isnan(want))
        assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))


def test_global_avgpool(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    x = torch.randn((16, 2048, 7, 7), device="cuda")
    got = K.global_avgpool(x)
    want = x.double().mean(dim=(2, 3), keepdim=True)
    assert got.shape == (16, 2048, 1, 1)
    assert (got.double() - want).abs().max().item() < 1e-6


def test_engine_native_ops_match_library_ops(dpl_built, monkeypatch):
    """The same reduced ResNet through the engine with libdpl_b200's streaming operators and with
    torch's: every blob identical except those downstream of GlobalAveragePool (summation order)."""
    import torch
    from dipoorlet_b200 import workloads as W
    from dipoorlet_b200.engine import Engine
    from dipoorlet_b200.graph import ONNXGraph
    model = W.build_resnet50(blocks=[1, 1, 1, 1], width=16, num_classes=10, image=64)
    graph = ONNXGraph(model, "/tmp/dpl_ops", "trt")
    x = torch.from_numpy(W.synthetic_images(5, (3, 64, 64), seed=3)[:, 0]).cuda()
    native = Engine(graph, torch.device("cuda", 0))
    native.fuse_conv_relu = True      # also cover the dual-output Conv epilogue
    plain = Engine(graph, torch.device("cuda", 0))
    plain.native_ops = False
    a = native.run({"input": x}, want="all")
    b = plain.run({"input": x}, want="all")
    assert list(a) == list(b)
    seen_gap = False
    for node in native.nodes:
        seen_gap |= node.op_type == "GlobalAveragePool"
        for o in node.output:
            if seen_gap:
                assert torch.allclose(a[o], b[o], rtol=1e-5, atol=1e-6), o
            else:
                assert torch.equal(a[o], b[o]), o


def test_fused_range_statistics_equal_k1(dpl_built):
    """per_image=False: the streaming kernels fold min / max of the blobs they write and K1 reads only
    the rest - the per-blob range (and so the histogram and the clip values) must be bit-identical to
    the K1-over-everything pass; and the single-kernel form of every fused operator against torch."""
    import torch
    from dipoorlet_b200 import forward_net as fwd, kernels as K, workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    x = torch.randn((3, 5, 9, 11), device="cuda") * 2
    y = torch.randn((3, 5, 9, 11), device="cuda")
    lo = torch.full((4,), float("inf"), device="cuda")
    hi = torch.full((4,), float("-inf"), device="cuda")
    r = torch.empty_like(x)
    K.clip(x, 0.0, 6.0, rng=(lo, hi, 0))
    s = K.add(x, y, out_relu=r, rng=(lo, hi, 1), rng_relu=(lo, hi, 2))
    g = K.global_avgpool(x, rng=(lo, hi, 3))
    want = [torch.clamp(x, 0, 6), s, r, g]
    assert lo.tolist() == [w.min().item() for w in want] and hi.tolist() == [w.max().item() for w in want]
    mp = K.maxpool2d(x, (3, 3), (2, 2), 1, 1, 5, 6, rng=(lo, hi, 0))     # folds into the running entry
    assert lo[0].item() == min(want[0].min().item(), mp.min().item())
    assert hi[0].item() == max(want[0].max().item(), mp.max().item())

    model = W.build_resnet50(blocks=[1, 1, 1, 1], width=16, num_classes=10, image=64)
    graph = ONNXGraph(model, "/tmp/dpl_fused", "trt")
    images = W.synthetic_images(10, (3, 64, 64), seed=6)
    res = {}
    for per_image in (True, False):
        args = make_args(input_dir=fwd.ArrayInput({"input": images[:, 0]}), data_num=10, deploy="trt",
                         act_quant="hist", output_dir="/tmp/dpl_fused", calib_bs=4)
        sess = fwd.CalibrationSession(graph, args)
        sess.run_minmax(per_image=per_image)
        assert (sess.seg_min is None) == (not per_image)
        sess.run_hist(2048)
        clip, _ = sess.percentile_clip(2048, 0.99999)
        res[per_image] = (sess.blob_min.cpu().numpy(), sess.blob_max.cpu().numpy(), sess.counts.cpu().numpy(),
                          clip.cpu().numpy())
    for u, v in zip(res[True], res[False]):
        assert np.array_equal(u, v)


def test_session_arena_resident_and_recompute_agree(dpl_built):
    """hist calibration with the blobs kept resident in one slab (pass 2 in place) and with the
    recycled one-batch slab (pass 2 recomputes): identical counts and clip values."""
    import torch
    from dipoorlet_b200 import forward_net as fwd, workloads as W
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.graph import ONNXGraph
    model = W.build_resnet50(blocks=[1, 1, 1, 1], width=16, num_classes=10, image=64)
    graph = ONNXGraph(model, "/tmp/dpl_arena", "trt")
    images = W.synthetic_images(10, (3, 64, 64), seed=5)
    res = {}
    for resident in (0, 1):
        args = make_args(input_dir=fwd.ArrayInput({"input": images[:, 0]}), data_num=10, deploy="trt",
                         act_quant="hist", output_dir="/tmp/dpl_arena", calib_bs=4, resident=resident)
        sess = fwd.CalibrationSession(graph, args)
        assert sess.keep_resident == bool(resident)
        sess.run_minmax()
        assert sess.arena is not None
        sess.run_hist(2048)
        clip, sel = sess.percentile_clip(2048, 0.99999)
        res[resident] = (sess.counts.cpu().numpy(), clip.cpu().numpy(), sess.seg_min.cpu().numpy())
    for u, v in zip(res[0], res[1]):
        assert np.array_equal(u, v)

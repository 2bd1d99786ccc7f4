"""tcgen05 TF32 GEMM (dpl_gemm_tf32) against float64 matmul of TF32-truncated operands
(kind::tf32 uses the top 19 bits of the fp32 pattern) and, loosely, plain fp32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(t):
    import torch
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _check(got, want64, scale):
    err = (got.double() - want64).abs().max().item()
    assert err <= 2e-5 * scale, (err, scale)


@pytest.mark.parametrize("shape", [(128, 128, 32), (200, 136, 100), (64, 1000, 2048), (300, 40, 36)])
def test_linear_kk(dpl_built, shape):
    import torch
    from dipoorlet_b200 import kernels as K
    m, n, k = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((m, k), device="cuda", generator=g)
    w = torch.randn((n, k), device="cuda", generator=g)
    b = torch.randn(n, device="cuda", generator=g)
    y = K.linear_forward(x, w, b, relu=True)
    K.gemm_check_errors()
    want = torch.relu(_tf32(x).double() @ _tf32(w).double().t() + b.double())
    _check(y, want, np.sqrt(k) * 4)
    assert torch.allclose(y, torch.relu(x @ w.t() + b), rtol=2e-2, atol=2e-2 * np.sqrt(k))
    go = torch.randn((m, n), device="cuda", generator=g)
    dw = K.linear_wgrad(go, x)
    dx = K.linear_dgrad(go, w)
    K.gemm_check_errors()
    _check(dw, _tf32(go).double().t() @ _tf32(x).double(), np.sqrt(m) * 4)
    _check(dx, _tf32(go).double() @ _tf32(w).double(), np.sqrt(n) * 4)


@pytest.mark.parametrize("dims", [(3, 64, 256, 56), (2, 96, 40, 28), (5, 256, 64, 14), (1, 8, 136, 12)])
def test_conv1x1_forward_wgrad_dgrad(dpl_built, dims):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, ci, co, hw = dims
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci), device="cuda", generator=g) * 0.1
    b = torch.randn(co, device="cuda", generator=g)
    go = torch.randn((n, co, hw, hw), device="cuda", generator=g)
    xt, wt, got = _tf32(x).double(), _tf32(w).double(), _tf32(go).double()
    o = K.conv1x1_forward(x, w, b)
    K.gemm_check_errors()
    want = torch.einsum("oc,nchw->nohw", wt, xt) + b.double().view(1, -1, 1, 1)
    _check(o, want, np.sqrt(ci))
    dw = K.conv1x1_wgrad(go, x)
    K.gemm_check_errors()
    want = torch.einsum("nohw,nchw->oc", got, xt)
    _check(dw, want, np.sqrt(n * hw * hw) * 4)
    dx = K.conv1x1_dgrad(go, w)
    K.gemm_check_errors()
    want = torch.einsum("oc,nohw->nchw", wt, got)
    _check(dx, want, np.sqrt(co))
    torch.backends.cudnn.allow_tf32 = False
    assert torch.allclose(o, F.conv2d(x, w.view(co, ci, 1, 1), b), rtol=2e-2, atol=2e-2)


def test_unsupported_alignment_is_reported(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    x = torch.randn((2, 8, 7, 7), device="cuda")      # hw = 49: row stride not a multiple of 16 bytes
    w = torch.randn((16, 8), device="cuda")
    with pytest.raises(K.GemmUnsupported):
        K.conv1x1_forward(x, w)


@pytest.mark.parametrize("dims", [(3, 64, 256, 56), (2, 96, 40, 28), (4, 1024, 256, 14), (1, 8, 136, 12),
                                  (64, 64, 256, 28), (40, 512, 128, 14), (7, 256, 520, 10), (300, 32, 24, 6)])
def test_conv1x1_forward_3xtf32_is_fp32_accurate(dpl_built, dims):
    """3xTF32 must sit at fp32 accuracy (the calibration forward is compared at 1e-5):
    error vs float64 no worse than ~2x the error of a plain fp32 convolution."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, ci, co, hw = dims
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci), device="cuda", generator=g) * 0.1
    b = torch.randn(co, device="cuda", generator=g)
    # K <= 512 runs the persistent kernel (several tiles per CTA above 148 tiles), else one tile per CTA
    out = torch.full((n, co, hw, hw), float("nan"), device="cuda")
    o = K.conv1x1_forward_x3(x, w, K.tf32_residual(w), b, out=out)
    K.gemm_check_errors()
    assert not torch.isnan(o).any()
    want = torch.einsum("oc,nchw->nohw", w.double(), x.double()) + b.double().view(1, -1, 1, 1)
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w.view(co, ci, 1, 1), b)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= max(3 * err32, 2e-6 * scale), (err, err32, scale)
    assert err <= 1e-5 * scale


@pytest.mark.parametrize("dims", [(3, 64, 256, 56), (2, 96, 40, 28), (4, 1024, 256, 14), (1, 8, 136, 12),
                                  (64, 64, 256, 28), (40, 512, 128, 14), (7, 256, 520, 10), (300, 32, 24, 6),
                                  (2, 2048, 512, 8), (3, 36, 70, 30)])
def test_conv1x1_pixel_major_3xtf32_is_fp32_accurate(dpl_built, dims):
    """The pixel-major tile (activations as the TMEM operand): fp32 accuracy, ragged pixel / channel
    tails (zero-filled by TMA, masked in the epilogue), and the fused Relu output."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, ci, co, hw = dims
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci), device="cuda", generator=g) * 0.1
    b = torch.randn(co, device="cuda", generator=g)
    out = torch.full((n, co, hw, hw), float("nan"), device="cuda")
    out_relu = torch.full((n, co, hw, hw), float("nan"), device="cuda")
    o = K.conv1x1_px_forward_x3(x, w, K.tf32_residual(w), b, out=out, out_relu=out_relu)
    K.gemm_check_errors()
    assert not torch.isnan(o).any()
    assert torch.equal(out_relu, torch.relu(o))
    want = torch.einsum("oc,nchw->nohw", w.double(), x.double()) + b.double().view(1, -1, 1, 1)
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w.view(co, ci, 1, 1), b)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # one accumulator for the leading term: the tensor core's truncating accumulation shows beyond
    # K = 1024 (5.8e-6 of the output scale at K = 2048, fp32 cuDNN: 1.7e-6)
    assert err <= max((3 if ci <= 1024 else 5) * err32, 2e-6 * scale), (err, err32, scale)
    assert err <= 1e-5 * scale
    # no bias, no relu output
    o2 = K.conv1x1_px_forward_x3(x, w, K.tf32_residual(w))
    K.gemm_check_errors()
    assert torch.allclose(o2, o - b.view(1, -1, 1, 1), rtol=0, atol=1e-5 * scale)


@pytest.mark.parametrize("cfg", [(4, 3, 64, 64, 7, 2, 3), (2, 3, 16, 33, 7, 2, 3), (3, 1, 8, 20, 5, 1, 2),
                                 (2, 4, 40, 18, 3, 2, 1)])
def test_conv_im2col_stem_3xtf32_is_fp32_accurate(dpl_built, cfg):
    """Few-channel convolution (ResNet's 7x7 / stride 2 stem) = im2col staging + the single-tap
    tensor-core kernel: fp32 accuracy against float64, incl. ragged K (147 -> 148) and borders."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, ci, co, hw, k, stride, pad = cfg
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.1
    b = torch.randn(co, device="cuda", generator=g)
    prep = K.conv_im2col_prepare(w)
    o = K.conv_im2col_forward_x3(x, prep, (k, k), stride, pad, b)
    K.gemm_check_errors()
    want = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)
    assert o.shape == want.shape
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w, b, stride=stride, padding=pad)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= max(3 * err32, 2e-6 * scale), (err, err32, scale)


@pytest.mark.parametrize("cfg", [(4, 3, 64, 64, 7, 2, 3), (2, 3, 16, 33, 7, 2, 3), (3, 1, 8, 20, 5, 1, 2),
                                 (2, 4, 40, 18, 3, 2, 1), (2, 3, 32, 37, 3, 2, 1), (1, 3, 70, 21, 3, 1, 1),
                                 (2, 2, 24, 30, 5, 2, 2)])
def test_conv_direct_stem_is_exact_fp32(dpl_built, cfg):
    """Direct few-channel convolution (the stem): fp32 FMA accumulation, so it must sit at the accuracy of
    cuDNN's fp32 kernel against float64; ragged tiles, channel tails (c_out = 70 spans two channel tiles,
    8 / 24 / 40 leave part of one empty), the fused Relu output and the fused range statistics."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, ci, co, hw, k, stride, pad = cfg
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn((n, ci, hw, hw + 3), device="cuda", generator=g)
    w = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.1
    b = torch.randn(co, device="cuda", generator=g)
    lo = torch.full((2,), float("inf"), device="cuda")
    hi = torch.full((2,), float("-inf"), device="cuda")
    want = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)
    out = torch.full(want.shape, float("nan"), device="cuda")
    out_relu = torch.full(want.shape, float("nan"), device="cuda")
    o = K.conv_direct_forward(x, w, b, stride, pad, out=out, out_relu=out_relu, rng=(lo, hi, 0), rng_relu=(lo, hi, 1))
    assert not torch.isnan(o).any()
    assert torch.equal(out_relu, torch.relu(o))
    assert lo.tolist() == [o.min().item(), out_relu.min().item()]
    assert hi.tolist() == [o.max().item(), out_relu.max().item()]
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w, b, stride=stride, padding=pad)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    assert err <= max(2 * err32, 1e-6 * want.abs().max().item()), (err, err32)
    o2 = K.conv_direct_forward(x, w, None, stride, pad)
    assert torch.allclose(o2, o - b.view(1, -1, 1, 1), rtol=0, atol=1e-5)


@pytest.mark.parametrize("cfg", [(3, 32, 56, 3, 1), (2, 96, 57, 3, 2), (2, 7, 9, 5, 1), (1, 16, 14, 5, 2)])
def test_depthwise_conv_is_exact_fp32(dpl_built, cfg):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    n, c, hw, k, stride = cfg
    pad = k // 2
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((n, c, hw, hw + 1), device="cuda", generator=g)
    w = torch.randn((c, 1, k, k), device="cuda", generator=g) * 0.3
    b = torch.randn(c, device="cuda", generator=g)
    lo = torch.full((1,), float("inf"), device="cuda")
    hi = torch.full((1,), float("-inf"), device="cuda")
    o = K.dwconv2d_forward(x, w, b, stride, pad, rng=(lo, hi, 0))
    want = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad, groups=c)
    assert o.shape == want.shape
    assert (o.double() - want).abs().max().item() <= 2e-6 * want.abs().max().item()
    assert lo.item() == o.min().item() and hi.item() == o.max().item()


def test_linear_forward_3xtf32(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((64, 2048), device="cuda", generator=g)
    w = torch.randn((1000, 2048), device="cuda", generator=g) * 0.02
    b = torch.randn(1000, device="cuda", generator=g)
    y = K.linear_forward_x3(x, w, K.tf32_residual(w), b)
    K.gemm_check_errors()
    want = x.double() @ w.double().t() + b.double()
    err = (y.double() - want).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), err


@pytest.mark.parametrize("n,ci,co,h,w", [(3, 64, 64, 56, 56), (2, 128, 128, 28, 28), (5, 256, 256, 14, 14),
                                         (4, 512, 512, 7, 7), (2, 48, 40, 9, 11), (1, 8, 200, 5, 3)])
def test_conv3x3_forward_3xtf32(dpl_built, n, ci, co, h, w):
    """Shifted-window 3x3 convolution on the tensor cores vs a float64 reference: within 3x the
    error of torch's own fp32 (non-TF32) convolution, i.e. fp32-accurate; every output element
    written (NaN pre-fill), borders exact (zero padding)."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + ci)
    x = torch.randn((n, ci, h, w), device="cuda", generator=g)
    wt = torch.randn((co, ci, 3, 3), device="cuda", generator=g) * 0.05
    b = torch.randn(co, device="cuda", generator=g)
    taps, taps_lo = K.conv3x3_prepare(wt)
    out = torch.full((n, co, h, w), float("nan"), device="cuda")
    o = K.conv3x3_forward_x3(x, taps, taps_lo, b, relu=False, out=out)
    K.gemm_check_errors()
    assert not torch.isnan(o).any()
    want = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, wt, b, padding=1)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= max(3 * err32, 2e-6 * scale), (err, err32, scale)
    o2 = K.conv3x3_forward_x3(x, taps, taps_lo, None, relu=True)
    K.gemm_check_errors()
    want2 = F.conv2d(x.double(), wt.double(), None, padding=1).clamp_min(0)
    assert (o2.double() - want2).abs().max().item() <= max(3 * err32, 2e-6 * scale)


@pytest.mark.parametrize("n,ci,co,h,w,k,stride", [(3, 128, 128, 56, 56, 3, 2), (4, 256, 256, 28, 28, 3, 2),
                                                  (2, 32, 48, 7, 9, 3, 2), (2, 64, 96, 15, 14, 3, 2),
                                                  (3, 256, 512, 56, 56, 1, 2), (5, 1024, 2048, 14, 14, 1, 2),
                                                  (2, 24, 40, 7, 5, 1, 2)])
def test_conv_strided_3xtf32(dpl_built, n, ci, co, h, w, k, stride):
    """Stride-2 3x3 (parity-plane split) and stride-2 1x1 (gather) convolutions on the same
    tap-table kernel, vs float64; odd sizes exercise the plane borders."""
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + ci + k)
    x = torch.randn((n, ci, h, w), device="cuda", generator=g)
    wt = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.05
    b = torch.randn(co, device="cuda", generator=g)
    taps, taps_lo = K.conv_taps_prepare(wt)
    pad = 1 if k == 3 else 0
    want = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=pad)
    out = torch.full(tuple(want.shape), float("nan"), device="cuda")
    o = K.conv_taps_forward_x3(x, taps, taps_lo, k, stride, b, out=out)
    K.gemm_check_errors()
    assert not torch.isnan(o).any()
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, wt, b, stride=stride, padding=pad)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= max(3 * err32, 2e-6 * scale), (err, err32, scale)

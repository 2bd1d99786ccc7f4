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


@pytest.mark.parametrize("dims", [(3, 64, 256, 56), (2, 96, 40, 28), (4, 1024, 256, 14), (1, 8, 136, 12)])
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
    o = K.conv1x1_forward_x3(x, w, K.tf32_residual(w), b)
    K.gemm_check_errors()
    want = torch.einsum("oc,nchw->nohw", w.double(), x.double()) + b.double().view(1, -1, 1, 1)
    torch.backends.cudnn.allow_tf32 = False
    ref32 = F.conv2d(x, w.view(co, ci, 1, 1), b)
    err = (o.double() - want).abs().max().item()
    err32 = (ref32.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= max(3 * err32, 2e-6 * scale), (err, err32, scale)
    assert err <= 1e-5 * scale


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

"""K6 contraction kernels for k x k / strided / depthwise convolutions (dpl_tap_conv_tf32, dpl_tap_wgrad_tf32,
dpl_dwconv2d_wgrad_f32 / _dgrad_f32, the im2col stem gradient) against torch in float64 on TF32-truncated
operands (kind::tf32 reads the top 19 bits of the fp32 pattern) — the forward, the weight gradient and the
data gradient that torch autograd derives from F.conv2d (ada_quant_layer.py:224-244)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(t):
    import torch
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _check(got, want64, scale, what):
    err = (got.double() - want64).abs().max().item()
    assert err <= 2e-5 * scale, (what, err, scale)


SHAPES = [  # n, ci, co, h, w, k, stride, pad
    (3, 64, 64, 56, 56, 3, 1, 1), (2, 128, 128, 28, 28, 3, 1, 1), (4, 256, 256, 14, 14, 3, 1, 1),
    (5, 512, 512, 7, 7, 3, 1, 1), (2, 48, 40, 9, 11, 3, 1, 1), (3, 128, 128, 56, 56, 3, 2, 1),
    (4, 256, 256, 28, 28, 3, 2, 1), (2, 32, 48, 7, 9, 3, 2, 1), (2, 64, 96, 15, 14, 3, 2, 1),
    (3, 256, 512, 56, 56, 1, 2, 0), (2, 24, 40, 7, 5, 1, 2, 0), (6, 512, 2048, 7, 7, 1, 1, 0),
    (3, 160, 960, 7, 7, 1, 1, 0), (64, 64, 64, 56, 56, 3, 1, 1),
]


@pytest.mark.parametrize("n,ci,co,h,w,k,stride,pad", SHAPES)
def test_tap_conv_forward_wgrad_dgrad(dpl_built, n, ci, co, h, w, k, stride, pad):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + ci + k)
    x = torch.randn((n, ci, h, w), device="cuda", generator=g)
    wt = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.05
    b = torch.randn(co, device="cuda", generator=g)
    plan = K.ReconConvPlan(n, h, w, k, stride, pad)
    wf, wd = K.taps_layout(wt, True, True)
    assert torch.equal(wf, wt.permute(2, 3, 0, 1).reshape(k * k, co, ci))
    assert torch.equal(wd, wt.permute(2, 3, 1, 0).reshape(k * k, ci, co))
    xp = K.recon_stage_input(x, plan)
    xd = _tf32(x).double().requires_grad_(True)
    wdd = _tf32(wt).double().requires_grad_(True)
    want = F.conv2d(xd, wdd, b.double(), stride=stride, padding=pad)
    out = torch.full(tuple(want.shape), float("nan"), device="cuda")
    y = K.recon_conv_forward(xp, plan, wf, b, out=out)
    K.gemm_check_errors()
    assert not torch.isnan(y).any()
    _check(y, want.detach(), np.sqrt(ci * k * k) * 0.3, "forward")
    go = torch.randn(tuple(want.shape), device="cuda", generator=g)
    gp = K.recon_stage_grad(go, plan)
    # weight gradient: both operands truncated to TF32
    (gw_want,) = torch.autograd.grad(F.conv2d(xd, wdd, None, stride=stride, padding=pad), wdd, _tf32(go).double())
    gw = K.recon_conv_wgrad(gp, xp, plan, co, ci, out=torch.full((co, ci, k, k), float("nan"), device="cuda"))
    K.gemm_check_errors()
    assert not torch.isnan(gw).any()
    _check(gw, gw_want, np.sqrt(n * want.shape[2] * want.shape[3]) * 4, "wgrad")
    (gx_want,) = torch.autograd.grad(F.conv2d(xd, wdd, None, stride=stride, padding=pad), xd, _tf32(go).double())
    gx = K.recon_conv_dgrad(gp, plan, wd, out=torch.full((n, ci, h, w), float("nan"), device="cuda"))
    K.gemm_check_errors()
    assert not torch.isnan(gx).any()
    _check(gx, gx_want, np.sqrt(co * k * k) * 0.3, "dgrad")


@pytest.mark.parametrize("n,c,h,w,k,stride", [(3, 32, 112, 112, 3, 1), (2, 96, 112, 112, 3, 2), (4, 144, 56, 56, 3, 2),
                                              (5, 960, 7, 7, 3, 1), (2, 40, 9, 11, 5, 1), (2, 24, 13, 10, 5, 2)])
def test_depthwise_gradients(dpl_built, n, c, h, w, k, stride):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(c + k)
    pad = k // 2
    x = torch.randn((n, c, h, w), device="cuda", generator=g)
    wt = torch.randn((c, 1, k, k), device="cuda", generator=g) * 0.3
    xd, wd = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    y = F.conv2d(xd, wd, None, stride=stride, padding=pad, groups=c)
    go = torch.randn(tuple(y.shape), device="cuda", generator=g)
    gx_want, gw_want = torch.autograd.grad(y, (xd, wd), go.double())
    gw = K.dwconv2d_wgrad(x, go, k, stride, pad)
    gx = K.dwconv2d_dgrad(go, wt, (h, w), stride, pad, out=torch.full((n, c, h, w), float("nan"), device="cuda"))
    assert not torch.isnan(gx).any()
    assert (gx.double() - gx_want).abs().max().item() <= 1e-5 * k
    assert (gw.double() - gw_want).abs().max().item() <= 2e-6 * gw_want.abs().max().item() + 1e-4


@pytest.mark.parametrize("n,c,co,h,w,k,stride,pad", [(4, 3, 64, 224, 224, 7, 2, 3), (3, 3, 32, 224, 224, 3, 2, 1),
                                                      (2, 3, 16, 33, 31, 3, 2, 1)])
def test_stem_wgrad(dpl_built, n, c, co, h, w, k, stride, pad):
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(co)
    x = torch.randn((n, c, h, w), device="cuda", generator=g)
    wt = torch.randn((co, c, k, k), device="cuda", generator=g) * 0.1
    xd, wd = _tf32(x).double(), _tf32(wt).double().requires_grad_(True)
    y = F.conv2d(xd, wd, None, stride=stride, padding=pad)
    go = torch.randn(tuple(y.shape), device="cuda", generator=g)
    if (y.shape[2] * y.shape[3]) % 4:
        pytest.skip("pixel count not a multiple of 4: the layer stays on its other path")
    (gw_want,) = torch.autograd.grad(y, wd, _tf32(go).double())
    gw = K.conv_im2col_wgrad(x, go, (k, k), stride, pad)
    K.gemm_check_errors()
    # K = n * Ho * Wo up to 50 176 products per output, split over CTAs that add their partial sums atomically
    err = (gw.double() - gw_want).abs().max().item()
    assert err <= 2e-4 * gw_want.abs().max().item(), err


def test_adaqlayer_native_kinds(dpl_built):
    """Every layer shape of ResNet-50 / MobileNetV2 is classified onto a libdpl_b200 contraction, and one
    forward + backward through AdaQLayer matches torch autograd (TF32 tolerance)."""
    import os
    import torch
    import torch.nn.functional as F
    from dipoorlet_b200.onnx_lite import Node
    from dipoorlet_b200.weight_transform.ada_quant_layer import AdaQLayer
    cases = [  # ci, co, k, stride, pad, groups, hw, expected kind
        (3, 64, 7, 2, 3, 1, 64, 'stem'), (64, 64, 1, 1, 0, 1, 16, 'c1x1'), (64, 64, 3, 1, 1, 1, 16, 'taps'),
        (128, 128, 3, 2, 1, 1, 16, 'taps'), (256, 512, 1, 2, 0, 1, 16, 'taps'), (512, 2048, 1, 1, 0, 1, 7, 'taps'),
        (3, 32, 3, 2, 1, 1, 64, 'stem'), (96, 96, 3, 2, 1, 96, 16, 'dw'), (144, 144, 3, 1, 1, 144, 8, 'dw'),
    ]
    os.environ["DPL_STRICT_NATIVE"] = "1"
    try:
        for ci, co, k, st, pd, grp, hw, kind in cases:
            g = torch.Generator(device="cuda").manual_seed(ci + co)
            node = Node("Conv", ["x", "w", "b"], ["y"], name="c",
                        attrs=dict(kernel_shape=[k, k], strides=[st, st], pads=[pd] * 4, dilations=[1, 1], group=grp))
            w = torch.randn((co, ci // grp, k, k), device="cuda", generator=g) * 0.1
            b = torch.randn(co, device="cuda", generator=g)
            scale = (w.abs().amax(dim=(1, 2, 3)) / 127).clamp_min(1e-8)
            layer = AdaQLayer(node, w.cpu().numpy(), b.cpu().numpy(), scale, -127, 127, False,
                              device=torch.device("cuda"))
            x = torch.randn((8, ci, hw, hw), device="cuda", generator=g)
            ws = layer.quant_weight(soft=True)
            y = layer.dense_forward(x, ws)
            assert layer._kind == kind, (layer._kind, kind)
            xd, wd = x.double().requires_grad_(True), ws.double().requires_grad_(True)
            want = F.conv2d(xd, wd, b.double(), stride=st, padding=pd, groups=grp)
            tol = 4e-3 * max(1.0, want.abs().max().item())
            assert (y.double() - want).abs().max().item() <= tol
            go = torch.randn(tuple(want.shape), device="cuda", generator=g)
            gx_want, gw_want = torch.autograd.grad(want, (xd, wd), go.double())
            need_dx = kind != 'stem'
            gx, gw = layer.dense_backward(x, ws, go, need_dx)
            assert (gw.double() - gw_want).abs().max().item() <= 4e-3 * max(1.0, gw_want.abs().max().item())
            if need_dx:
                assert (gx.double() - gx_want).abs().max().item() <= 4e-3 * max(1.0, gx_want.abs().max().item())
    finally:
        os.environ.pop("DPL_STRICT_NATIVE", None)

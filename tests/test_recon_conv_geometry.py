"""Host-side geometry of the K6 convolution kernels (kernels.ReconConvPlan: staging pitches, tap shifts, the
output-parity classes of the strided data gradient) against torch's convolution gradients, through a NumPy
emulation of the kernels' index arithmetic (tests/tap_emulation.py). No GPU needed."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import tap_emulation as E
from dipoorlet_b200 import kernels as K


@pytest.mark.parametrize("n,ci,co,h,w,k,s,p", [(2, 4, 8, 6, 6, 3, 1, 1), (2, 4, 8, 7, 5, 3, 2, 1), (2, 8, 4, 6, 8, 3, 2, 1),
                                                (1, 4, 4, 7, 7, 1, 2, 0), (2, 4, 8, 5, 5, 1, 1, 0), (2, 4, 4, 8, 6, 1, 2, 0)])
def test_plan_matches_autograd(n, ci, co, h, w, k, s, p):
    torch.manual_seed(0)
    x = torch.randn(n, ci, h, w, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(co, ci, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, wt, None, stride=s, padding=p)
    go = torch.randn_like(y)
    y.backward(go)
    plan = K.ReconConvPlan(n, h, w, k, s, p)
    xp = E.pad_plane(x.detach().numpy(), s, plan.origin, plan.hp, plan.wp, plan.planes)
    assert xp.shape[0] == plan.total_rows
    wf = wt.detach().numpy().transpose(2, 3, 0, 1).reshape(k * k, co, ci)
    wd = wt.detach().numpy().transpose(2, 3, 1, 0).reshape(k * k, ci, co)
    yy = np.full(y.shape, np.nan)
    E.tap_conv(xp, wf, n, co, plan.ho, plan.wo, plan.hp, plan.wp, plan.origin, 1, 0, 0, plan.shifts, range(k * k), yy)
    assert np.allclose(yy, y.detach().numpy())
    gp = E.pad_plane(go.numpy(), 1, plan.origin, plan.hp, plan.wp, 1)
    assert gp.shape[0] == plan.q_total
    dw = E.tap_wgrad(gp, xp, co, ci, k * k, plan.shifts, range(k * k)).reshape(co, ci, k, k)
    assert np.allclose(dw, wt.grad.numpy())
    dx = np.zeros(x.shape) if any(not t for *_, t in plan.dgrad) else np.full(x.shape, np.nan)
    for a, b, sh, tp in plan.dgrad:
        if tp:
            E.tap_conv(gp, wd, n, ci, h, w, plan.hp, plan.wp, plan.origin, s, a, b, sh, tp, dx)
    assert np.allclose(dx, x.grad.numpy())


def test_plan_rejects_other_geometries():
    for k, s, p in [(5, 1, 2), (3, 3, 1), (3, 1, 0), (7, 2, 3)]:
        with pytest.raises(K.GemmUnsupported):
            K.ReconConvPlan(1, 8, 8, k, s, p)

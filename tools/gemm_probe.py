"""Which dpl_gemm_tf32 operand layouts are right? (one-off debug aid)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402


def tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


g = torch.Generator(device="cuda").manual_seed(0)
for (n, ci, co, hw) in [(2, 64, 128, 16), (3, 96, 40, 28)]:
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci), device="cuda", generator=g) * 0.1
    go = torch.randn((n, co, hw, hw), device="cuda", generator=g)
    xt, wt, got = tf32(x).double(), tf32(w).double(), tf32(go).double()
    o = K.conv1x1_forward(x, w)
    e1 = (o.double() - torch.einsum("oc,nchw->nohw", wt, xt)).abs().max().item()
    dw = K.conv1x1_wgrad(go, x)
    e2 = (dw.double() - torch.einsum("nohw,nchw->oc", got, xt)).abs().max().item()
    dx = K.conv1x1_dgrad(go, w)
    e3 = (dx.double() - torch.einsum("oc,nohw->nchw", wt, got)).abs().max().item()
    K.gemm_check_errors()
    print((n, ci, co, hw), "fwd(K,MN) err", e1, " wgrad(K,K fold) err", e2, " dgrad(MN,MN) err", e3)

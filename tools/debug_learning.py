"""Diagnose the fused learning loop against torch autograd on the GPU (one-off debug aid)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K, onnx_lite as ol  # noqa: E402
from dipoorlet_b200.weight_transform.ada_quant_layer import AdaQLayer, adaround_reg  # noqa: E402
from oracle import adaround as OA  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
n, bs = 16, 8
x_fp = torch.randn((n, 6, 10, 10), device=dev, generator=g)
x = torch.round(x_fp / 0.1) * 0.1
w1 = torch.randn((8, 6, 3, 3), device=dev, generator=g) * 0.2
b1 = torch.randn(8, device=dev, generator=g) * 0.1
w2 = torch.randn((5, 8, 1, 1), device=dev, generator=g) * 0.3
a1 = {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]}
a2 = {"dilations": [1, 1], "group": 1, "kernel_shape": [1, 1], "pads": [0, 0, 0, 0], "strides": [2, 2]}
with torch.no_grad():
    h = torch.relu(F.conv2d(x_fp, w1, b1, padding=1))
    tgt = F.conv2d(h, w2, None, stride=2)
specs = [(a1, w1, b1, True), (a2, w2, None, False)]
ref, mine = [], []
for a, w, b, r in specs:
    scale = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous()
    s4 = scale.view(-1, 1, 1, 1)
    ref.append(OA.Layer("Conv", a, w, b, s4, torch.full_like(s4, -127), torch.full_like(s4, 127), r))
    mine.append(AdaQLayer(ol.Node("Conv", ["x", "w"], ["y"], "c", a), w, b, scale, -127, 127, r, device=dev))
# --- one iteration: compare every intermediate ---
xb, tb = x[:bs], tgt[:bs]
out = xb
for l in ref:
    out = l.forward(out)
loss = OA.l2_norm(out, tb)
loss.backward()
acts, outs = [xb], []
for li, L in enumerate(mine):
    w = L.quant_weight(True)
    print(li, "w_soft vs ref", (w - OA.quant_weight(ref[li].weight, ref[li].round_mask.detach(), ref[li].scale, ref[li].q_min, ref[li].q_max)).abs().max().item())
    o = L.dense_forward(acts[-1], w)
    print(li, "o contiguous", o.is_contiguous(), o.stride())
    outs.append(o)
    if li == 0:
        acts.append(K.recon_act(o, relu=True))
print("out diff", (outs[-1] - out.detach()).abs().max().item())
acc = torch.zeros(1, dtype=torch.float64, device=dev)
go = K.recon_loss(outs[-1], tb, float(outs[-1].shape[1]) / outs[-1].numel(), acc, relu=False)
print("loss", acc.item(), float(loss))
for li in (1, 0):
    L = mine[li]
    gx, gw = L.dense_backward(acts[li], L.w_soft, go, need_dx=li > 0)
    print(li, "gw contiguous", gw.is_contiguous(), "gx", None if gx is None else (gx.is_contiguous(), gx.stride()))
    # autograd's dL/dalpha = dL/dW * s * h'
    a = ref[li].round_mask.detach()
    sg = torch.sigmoid(a)
    dh = 1.2 * sg * (1 - sg)
    ga_mine = gw * ref[li].scale * dh
    print(li, "dL/dalpha rel diff", ((ga_mine - ref[li].round_mask.grad).abs().max() / ref[li].round_mask.grad.abs().max()).item(),
          "max|g|", ref[li].round_mask.grad.abs().max().item(), "min|g|", ref[li].round_mask.grad.abs().min().item())
    if li > 0:
        go = K.recon_act_bwd(outs[0], gx, relu=True)
# --- Adam step parity on a synthetic gradient ---
for li in (0, 1):
    L = mine[li]
    gw = torch.randn_like(L.weight) * 1e-3
    alpha0 = L.round_mask.clone()
    p = alpha0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p])
    for t in range(1, 4):
        wq = OA.quant_weight(L.weight, p, L.scale.view(-1, 1, 1, 1), torch.tensor(-127., device=dev), torch.tensor(127., device=dev))
        lossp = (wq * gw).sum() + OA.reg_loss(p, 5.0)
        opt.zero_grad()
        lossp.backward()
        opt.step()
        K.adaround_step(gw, L.wfloor, L.scale, -127, 127, 5.0, L.round_mask, L.m, L.v, t)
        print(li, "step", t, "alpha diff", (p.detach() - L.round_mask).abs().max().item())

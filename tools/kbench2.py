"""Microbenchmark of the streaming kernels around the statistics path (run on the GPU box):
K5 fake-quant, K6 elementwise (soft weight, fused d-alpha + Adam, epilogues, loss, QDrop mix),
K7 reductions and the forward engine's operators, each on buffers far larger than L2
(a 64-image batch of a 256 x 56 x 56 blob = 205.5 MB per tensor; the 126 MB L2 cannot hold an operand
between launches), CUDA events around 8 back-to-back launches, median of 10.

    python tools/kbench2.py [--out gpurun_out/kbench2.json] [--quick]

`bytes` is the ALGORITHMIC traffic of the kernel (every operand read once, every result
written once); `frac` is against MEASURED_PEAKS.json's hbm_gbs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402


REPS = 8   # launches per timed interval: these kernels run 60 - 400 us, an event pair adds a few us


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REPS):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REPS)
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kbench2.json")
    ap.add_argument("--quick", action="store_true", help="one launch per kernel (for ncu)")
    args = ap.parse_args()
    peak = 6545.0
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    global timed
    if args.quick:
        timed = lambda fn, iters=1, warm=0: (fn(), torch.cuda.synchronize(), (1.0, 1.0))[2]  # noqa: E731
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    n, c, h, w = 64, 256, 56, 56
    x = torch.randn((n, c, h, w), device=dev, generator=g)
    y = torch.randn((n, c, h, w), device=dev, generator=g)
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    nb = x.numel() * 4
    rows = []

    def row(name, traffic_bytes, fn, what):
        med, best = timed(fn)
        rows.append({"kernel": name, "what": what, "bytes": traffic_bytes, "ms": med, "ms_min": best,
                     "gbs": traffic_bytes / med / 1e6, "frac": traffic_bytes / med / 1e6 / peak})

    # ---- K5 fake-quant (Q/DQ pair) -----------------------------------------------------------
    s1 = torch.tensor([0.05], device=dev)
    sc = torch.rand(c, device=dev, generator=g) * 0.1 + 0.01
    row("K5 fakequant per-tensor", 2 * nb, lambda: K.fakequant(x, s1, None, -128, 127, out=o1), "read x, write y")
    row("K5 fakequant per-channel (axis 1)", 2 * nb, lambda: K.fakequant(x, sc, None, -128, 127, axis=1, out=o1),
        "read x, write y")
    row("K5 fakequant + QDrop (p = 0.5)", 2 * nb,
        lambda: K.fakequant(x, s1, None, -127, 127, drop_prob=0.5, seed=7, out=o1), "read x, write y")
    # ---- K7 reductions ---------------------------------------------------------------------
    acc = torch.zeros(c, dtype=torch.float64, device=dev)
    row("K7a channel_sumdiff", 2 * nb, lambda: K.channel_sumdiff(x, y, c, acc), "read a, b")
    cos = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    row("K7b cosine3", 2 * nb, lambda: K.cosine3(x, y, cos), "read a, b")
    # ---- K6 epilogues / loss / mix -------------------------------------------------------------
    row("K6 recon_act (relu + drop-fakequant)", 2 * nb,
        lambda: K.recon_act(x, True, (0.05, -127.0, 127.0), prob=0.5, seed=3, out=o1), "read o, write y")
    row("K6 recon_act_bwd", 3 * nb,
        lambda: K.recon_act_bwd(x, y, True, (0.05, -127.0, 127.0), prob=0.5, seed=3, out=o1), "read o, gy; write go")
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    row("K6 recon_loss (L2 + dL/do)", 3 * nb,
        lambda: K.recon_loss(x, y, 1.0 / (n * h * w), loss, True, None, out=o1), "read o, tgt; write go")
    row("K6 mix_drop (QDrop block input)", 3 * nb, lambda: K.mix_drop(x, y, 0.5, 11, out=o1), "read a, b; write y")
    # ---- K6 weight-shaped kernels (alpha has the weight's shape; 51.4 M elements here) ----------
    wshape = (c * 4, x.numel() // (c * 4))
    wv = x.view(wshape)
    scw = torch.rand(wshape[0], device=dev, generator=g) * 0.1 + 0.01
    alpha, wfloor = K.adaround_init(wv, scw)
    row("K6 adaround_weight (soft)", 3 * nb, lambda: K.adaround_weight(wfloor, alpha, scw, -127, 127, True, out=o1.view(wshape)),
        "read wfloor, alpha; write w_soft")
    m, v = torch.zeros_like(alpha), torch.zeros_like(alpha)
    gw = y.view(wshape)
    row("K6 adaround_step (d-alpha + reg + Adam)", 8 * nb,
        lambda: K.adaround_step(gw, wfloor, scw, -127, 127, 8.0, alpha, m, v, 5), "read grad, wfloor, alpha, m, v; write alpha, m, v")
    # ---- forward-engine operators --------------------------------------------------------------
    row("fwd Relu (dpl_clip_f32)", 2 * nb, lambda: K.clip(x, 0.0, float("inf"), out=o1), "read x, write y")
    row("fwd Add", 3 * nb, lambda: K.add(x, y, out=o1), "read a, b; write y")
    row("fwd Add + Relu (two blobs)", 4 * nb, lambda: K.add(x, y, out=o1, out_relu=o2), "read a, b; write y, relu(y)")
    xm = torch.randn((n, 64, 112, 112), device=dev, generator=g)
    om = torch.empty((n, 64, 56, 56), device=dev)
    row("fwd MaxPool 3x3 s2", xm.numel() * 4 + om.numel() * 4, lambda: K.maxpool2d(xm, (3, 3), (2, 2), 1, 1, 56, 56, out=om),
        "read x, write y")
    row("torch copy (for scale)", 2 * nb, lambda: o1.copy_(x), "read x, write y")
    res = {"peak_gbs": peak, "tensor_bytes": nb, "rows": rows}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for r in rows:
        print(f"{r['kernel']:42s} {r['ms']:8.3f} ms  {r['gbs']:8.1f} GB/s  {100 * r['frac']:5.1f}% of {peak:.0f}   ({r['what']})")


if __name__ == "__main__":
    main()

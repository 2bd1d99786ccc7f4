"""K6 contraction kernels (single-pass TF32, tcgen05) next to cuDNN with TF32 allowed - the reference's numerics
and library for this step (ada_quant_layer.py:224-244 under torch's default cudnn.allow_tf32 = True) - on the
ResNet-50 / MobileNetV2 layer shapes at --ada_bs 64: forward, weight gradient, data gradient, CUDA events.
The dpl times INCLUDE the channel-last staging copies the tap-table kernels need (pad_plane of x for the forward,
of dY for the gradients) and the filter re-layout."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402

B = int(os.environ.get("CONV_BATCH", "64"))
REPS = int(os.environ.get("REPS", "10"))
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
shapes = [  # ci, co, hw, k, stride, groups
    (64, 64, 56, 3, 1, 1), (128, 128, 28, 3, 1, 1), (256, 256, 14, 3, 1, 1), (512, 512, 7, 3, 1, 1),
    (128, 128, 56, 3, 2, 1), (256, 256, 28, 3, 2, 1), (256, 512, 56, 1, 2, 1), (64, 256, 56, 1, 1, 1),
    (256, 64, 56, 1, 1, 1), (1024, 256, 14, 1, 1, 1), (512, 2048, 7, 1, 1, 1), (96, 96, 112, 3, 2, 96),
    (144, 144, 56, 3, 1, 144), (960, 960, 7, 3, 1, 960),
]


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


g = torch.Generator(device="cuda").manual_seed(0)
rows = []
for ci, co, hw, k, st, grp in shapes:
    pad = k // 2
    x = torch.randn((B, ci, hw, hw), device="cuda", generator=g)
    w = torch.randn((co, ci // grp, k, k), device="cuda", generator=g) * 0.05
    y = F.conv2d(x, w, None, st, pad, 1, grp)
    go = torch.randn_like(y)
    flop = 2.0 * y.numel() * (ci // grp) * k * k
    r = {"shape": [B, ci, co, hw, k, st, grp], "gflop": round(flop / 1e9, 2)}
    r["cudnn_tf32_fwd_ms"] = timed(lambda: F.conv2d(x, w, None, st, pad, 1, grp))
    r["cudnn_tf32_wgrad_ms"] = timed(lambda: torch.ops.aten.convolution_backward(
        go, x, w, None, [st, st], [pad, pad], [1, 1], False, [0, 0], grp, [False, True, False]))
    r["cudnn_tf32_dgrad_ms"] = timed(lambda: torch.ops.aten.convolution_backward(
        go, x, w, None, [st, st], [pad, pad], [1, 1], False, [0, 0], grp, [True, False, False]))
    if grp == 1 and k == 1 and st == 1 and (hw * hw) % 4 == 0:
        w2 = w.view(co, ci)
        r["kind"] = "c1x1 (dpl_gemm_tf32)"
        r["dpl_fwd_ms"] = timed(lambda: K.conv1x1_forward(x, w2))
        r["dpl_wgrad_ms"] = timed(lambda: K.conv1x1_wgrad(go, x))
        r["dpl_dgrad_ms"] = timed(lambda: K.conv1x1_dgrad(go, w2))
    elif grp == 1:
        plan = K.ReconConvPlan(B, hw, hw, k, st, pad)
        bufs = {}

        def fwd():
            bufs["wf"], bufs["wd"] = K.taps_layout(w, True, True, bufs.get("wf"), bufs.get("wd"))
            bufs["xp"] = K.recon_stage_input(x, plan, bufs.get("xp"))
            return K.recon_conv_forward(bufs["xp"], plan, bufs["wf"])

        def wgrad():
            bufs["gp"] = K.recon_stage_grad(go, plan, bufs.get("gp"))
            return K.recon_conv_wgrad(bufs["gp"], bufs["xp"], plan, co, ci)

        r["kind"] = "taps (dpl_tap_conv_tf32 / dpl_tap_wgrad_tf32)"
        r["dpl_fwd_ms"] = timed(fwd)
        r["dpl_wgrad_ms"] = timed(wgrad)
        r["dpl_dgrad_ms"] = timed(lambda: K.recon_conv_dgrad(bufs["gp"], plan, bufs["wd"]))
    else:
        r["kind"] = "depthwise (exact fp32, FMA pipe)"
        r["dpl_fwd_ms"] = timed(lambda: K.dwconv2d_forward(x, w, None, st, pad))
        r["dpl_wgrad_ms"] = timed(lambda: K.dwconv2d_wgrad(x, go, k, st, pad))
        r["dpl_dgrad_ms"] = timed(lambda: K.dwconv2d_dgrad(go, w, (hw, hw), st, pad))
    K.gemm_check_errors()
    for key in list(r):
        if key.endswith("_ms"):
            r[key] = round(r[key], 4)
    r["dpl_fwd_tflops"] = round(flop / r["dpl_fwd_ms"] / 1e9, 1)
    r["dpl_wgrad_tflops"] = round(flop / r["dpl_wgrad_ms"] / 1e9, 1)
    rows.append(r)
    print(json.dumps(r))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/recon_conv_bench.json", "w"), indent=1)

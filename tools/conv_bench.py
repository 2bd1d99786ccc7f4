"""dpl_conv3x3_tf32x3 (tcgen05, 3xTF32) vs cuDNN fp32 / TF32 on ResNet-50's 3x3 shapes (batch 64)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


rows = []
n = int(os.environ.get("CONV_BATCH", 64))
only_1x1 = os.environ.get("CONV_ONLY_1X1") == "1"
for (ci, co, hw, k, stride) in [] if only_1x1 else [(64, 64, 56, 3, 1), (128, 128, 28, 3, 1), (256, 256, 14, 3, 1), (512, 512, 7, 3, 1),
                                (128, 128, 56, 3, 2), (256, 256, 28, 3, 2), (512, 512, 14, 3, 2),
                                (256, 512, 56, 1, 2), (512, 1024, 28, 1, 2), (1024, 2048, 14, 1, 2)]:
    x = torch.randn((n, ci, hw, hw), device="cuda")
    w = torch.randn((co, ci, k, k), device="cuda") * 0.05
    b = torch.randn(co, device="cuda")
    pad = 1 if k == 3 else 0
    ho = (hw - 1) // stride + 1
    flops = 2.0 * n * co * ci * k * k * ho * ho
    res = {"shape": [n, ci, co, hw, k, stride], "gflop": flops / 1e9}
    for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        res[name + "_ms"] = timed(lambda: F.conv2d(x, w, b, stride=stride, padding=pad))
    torch.backends.cudnn.allow_tf32 = False
    taps, taps_lo = K.conv_taps_prepare(w)
    o = torch.empty((n, co, ho, ho), device="cuda")
    scratch = torch.empty(K.ConvPlan(n, hw, hw, k, stride).total_rows * ci, device="cuda")
    res["dpl_ms"] = timed(lambda: K.conv_taps_forward_x3(x, taps, taps_lo, k, stride, b, out=o, scratch=scratch))
    K.gemm_check_errors()
    want = F.conv2d(x, w, b, stride=stride, padding=pad)
    res["max_abs_diff_vs_cudnn_fp32"] = (o - want).abs().max().item()
    res["dpl_fp32_equiv_tflops"] = flops / res["dpl_ms"] / 1e9
    res["dpl_tf32_mma_tflops"] = 3 * flops / res["dpl_ms"] / 1e9
    rows.append(res)
    print(json.dumps({k2: (round(v, 5) if isinstance(v, float) else v) for k2, v in res.items()}), flush=True)
# 1x1 stride 1: per-image MN-major GEMM (current engine path) vs the channel-last staged path
for (ci, co, hw) in [(64, 256, 56), (256, 64, 56), (128, 512, 28), (512, 128, 28), (256, 1024, 14),
                     (1024, 256, 14), (512, 2048, 7), (2048, 512, 7)]:
    x = torch.randn((n, ci, hw, hw), device="cuda")
    w = torch.randn((co, ci, 1, 1), device="cuda") * 0.05
    b = torch.randn(co, device="cuda")
    flops = 2.0 * n * co * ci * hw * hw
    io = 4.0 * n * (ci + co) * hw * hw
    res = {"shape": [n, ci, co, hw, 1, 1], "gflop": flops / 1e9, "io_mb": io / 1e6, "hbm_floor_ms": io / 6.4833e9}
    for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        res[name + "_ms"] = timed(lambda: F.conv2d(x, w, b))
    torch.backends.cudnn.allow_tf32 = False
    w2 = w.view(co, ci)
    w_lo = K.tf32_residual(w2)
    o = torch.empty((n, co, hw, hw), device="cuda")
    if (hw * hw) % 4 == 0:
        res["dpl_mn_ms"] = timed(lambda: K.conv1x1_forward_x3(x, w2, w_lo, b, out=o))
        res["dpl_px_ms"] = timed(lambda: K.conv1x1_px_forward_x3(x, w2, w_lo, b, out=o))
        K.gemm_check_errors()
        res["px_max_abs_diff_vs_cudnn_fp32"] = (o - F.conv2d(x, w, b)).abs().max().item()
    taps, taps_lo = K.conv_taps_prepare(w)
    scratch = torch.empty(K.ConvPlan(n, hw, hw, 1, 1).total_rows * ci, device="cuda")
    res["dpl_staged_ms"] = timed(lambda: K.conv_taps_forward_x3(x, taps, taps_lo, 1, 1, b, out=o, scratch=scratch))
    K.gemm_check_errors()
    res["max_abs_diff_vs_cudnn_fp32"] = (o - F.conv2d(x, w, b)).abs().max().item()
    rows.append(res)
    print(json.dumps({k2: (round(v, 5) if isinstance(v, float) else v) for k2, v in res.items()}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/conv_bench.json", "w"), indent=1)

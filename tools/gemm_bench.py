"""dpl_gemm_tf32 (tcgen05) vs cuDNN fp32 / TF32 on ResNet-50 1x1-conv shapes (batch 64)."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402


def timed(fn, iters=20, warm=3):
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
for (n, ci, co, hw) in [(64, 64, 256, 56), (64, 256, 64, 56), (64, 512, 128, 28), (64, 128, 512, 28),
                        (64, 1024, 256, 14), (64, 256, 1024, 14)]:
    x = torch.randn((n, ci, hw, hw), device="cuda")
    w = torch.randn((co, ci), device="cuda") * 0.05
    b = torch.randn(co, device="cuda")
    go = torch.randn((n, co, hw, hw), device="cuda")
    w4 = w.view(co, ci, 1, 1)
    flops = 2.0 * n * co * ci * hw * hw
    io_fwd = 4.0 * (x.numel() + n * co * hw * hw + w.numel())
    res = {"shape": [n, ci, co, hw], "gflop": flops / 1e9, "fwd_io_mb": io_fwd / 1e6}
    for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        res[name + "_fwd_ms"] = timed(lambda: F.conv2d(x, w4, b))
        res[name + "_wgrad_ms"] = timed(lambda: torch.ops.aten.convolution_backward(
            go, x, w4, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False]))
        res[name + "_dgrad_ms"] = timed(lambda: torch.ops.aten.convolution_backward(
            go, x, w4, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False]))
    o = torch.empty((n, co, hw, hw), device="cuda")
    dx = torch.empty_like(x)
    dw = torch.empty((co, ci), device="cuda")
    res["dpl_fwd_ms"] = timed(lambda: K.conv1x1_forward(x, w, b, out=o))
    res["dpl_wgrad_ms"] = timed(lambda: K.conv1x1_wgrad(go, x, out=dw))
    res["dpl_dgrad_ms"] = timed(lambda: K.conv1x1_dgrad(go, w, out=dx))
    K.gemm_check_errors()
    res["dpl_fwd_gbs"] = io_fwd / res["dpl_fwd_ms"] / 1e6
    res["dpl_fwd_tflops"] = flops / res["dpl_fwd_ms"] / 1e9
    rows.append(res)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)

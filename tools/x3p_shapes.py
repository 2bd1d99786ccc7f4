"""The calibration forward's main convolution shapes (ResNet-50, batch 128) on the chunked 3xTF32 kernel, timed
with CUDA events (and the target of the ncu capture of round 2): ms, TF32-MMA TFLOP/s (3 MMAs per product),
HBM GB/s of the algorithmic traffic (input + output + Relu copy)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402

B = int(os.environ.get("CONV_BATCH", "128"))
REPS = int(os.environ.get("REPS", "5"))
shapes = [  # ci, co, hw, k, stride
    (64, 256, 56, 1, 1), (256, 64, 56, 1, 1), (512, 128, 28, 1, 1), (1024, 256, 14, 1, 1), (256, 1024, 14, 1, 1),
    (2048, 512, 7, 1, 1), (64, 64, 56, 3, 1), (128, 128, 28, 3, 1), (256, 256, 14, 3, 1), (512, 512, 7, 3, 1),
    (256, 256, 28, 3, 2), (512, 1024, 28, 1, 2),
]
g = torch.Generator(device="cuda").manual_seed(0)
rows = []
for ci, co, hw, k, st in shapes:
    x = torch.randn((B, ci, hw, hw), device="cuda", generator=g).clamp_min(0)
    w = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.05
    b = torch.randn(co, device="cuda", generator=g)
    ho = (hw - 1) // st + 1
    y = torch.empty((B, co, ho, ho), device="cuda")
    y2 = torch.empty_like(y)
    if k == 1 and st == 1 and (hw * hw) % 4 == 0:
        hi, lo = K.tf32_split(w.view(co, ci))
        fn = lambda: K.conv1x1_forward_x3(x, hi, lo, b, out=y, out_relu=y2)  # noqa: E731
    else:
        taps, taps_lo = K.conv_taps_prepare(w)
        scratch = torch.empty(K.ReconConvPlan(B, hw, hw, k, st, k // 2).total_rows * ci, device="cuda")
        fn = lambda: K.conv_taps_forward_x3(x, taps, taps_lo, k, st, b, out=y, scratch=scratch, out_relu=y2)  # noqa: E731
    for _ in range(int(os.environ.get("WARMUP", "2"))):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    K.gemm_check_errors()
    ms = e0.elapsed_time(e1) / REPS
    flop = 2.0 * B * co * ho * ho * ci * k * k
    byts = 4.0 * (x.numel() + 2 * y.numel())
    rows.append({"shape": [B, ci, co, hw, k, st], "ms": round(ms, 4), "tf32_mma_tflops": round(3 * flop / ms / 1e9, 1),
                 "fp32_equiv_tflops": round(flop / ms / 1e9, 1), "hbm_gbs": round(byts / ms / 1e6, 1)})
    print(json.dumps(rows[-1]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/x3p_shapes.json", "w"), indent=1)

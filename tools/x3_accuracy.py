"""Error of the 3xTF32 kernels vs float64, next to cuDNN fp32's own error (same inputs)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(5)
print("DPL_X3_ALT =", os.environ.get("DPL_X3_ALT", "1"))
RN = os.environ.get("X3_RN", "0") == "1"   # operands pre-rounded to TF32 (round to nearest): isolates the accumulation error


def rn_tf32(t):
    u = t.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    r = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    r = torch.where(r >= 2 ** 31, r - 2 ** 32, r)
    return r.to(torch.int32).view(torch.float32)


for (n, ci, co, hw, k, relu_in) in [(8, 2048, 512, 14, 1, True), (8, 1024, 256, 14, 1, True), (8, 256, 1024, 14, 1, True),
                                    (8, 512, 512, 7, 3, True), (8, 256, 256, 14, 3, True), (8, 64, 64, 56, 3, True),
                                    (8, 512, 512, 7, 3, False)]:
    x = torch.randn((n, ci, hw, hw), device="cuda", generator=g)
    if relu_in:
        x = x.clamp_min(0)           # post-ReLU activations: same-sign partial sums (worst case for truncation)
    w = torch.randn((co, ci, k, k), device="cuda", generator=g) * 0.05
    if relu_in:
        w = w.abs()
    if RN:
        x, w = rn_tf32(x), rn_tf32(w)
    want = F.conv2d(x.double(), w.double(), None, padding=k // 2)
    ref32 = F.conv2d(x, w, None, padding=k // 2)
    if k == 1:
        w2 = w.view(co, ci)
        o = K.conv1x1_forward_x3(x, w2, K.tf32_residual(w2))
    else:
        taps, lo = K.conv_taps_prepare(w)
        o = K.conv_taps_forward_x3(x, taps, lo, 3, 1)
    K.gemm_check_errors()
    scale = want.abs().max().item()
    e = (o.double() - want)
    e32 = (ref32.double() - want)
    print("ci %4d co %4d hw %2d k %d %s: dpl max %.2e mean %.2e bias %+.2e | cudnn fp32 max %.2e mean %.2e bias %+.2e  (relative to max |y| = %.3g)" % (
        ci, co, hw, k, "pos" if relu_in else "mix", e.abs().max().item() / scale, e.abs().mean().item() / scale,
        e.mean().item() / scale, e32.abs().max().item() / scale, e32.abs().mean().item() / scale, e32.mean().item() / scale, scale))

import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K, _lib

n, ci, co, h, w = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 32, 64, 6, 6))]
x = torch.randn((n, ci, h, w), device="cuda")
wt = torch.randn((co, ci, 3, 3), device="cuda") * 0.05
taps, taps_lo = K.conv3x3_prepare(wt)
torch.cuda.synchronize(); print("prepare ok", flush=True)
pitch = K.conv3x3_plane_pitch(h, w)
xp = torch.empty(ci * n * pitch, device="cuda")
st = _lib.lib().dpl_pad_plane_f32(x.data_ptr(), xp.data_ptr(), n, ci, h, w, None)
torch.cuda.synchronize(); print("pad ok", st, flush=True)
ref = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(n * pitch, ci)
got = xp.view(n * pitch, ci)
print("pad equal", torch.equal(ref, got), flush=True)
y = K.conv3x3_forward_x3(x, taps, taps_lo, None)
torch.cuda.synchronize(); print("conv ok", flush=True)
K.gemm_check_errors()
want = F.conv2d(x.double(), wt.double(), None, padding=1)
print("max err", (y.double() - want).abs().max().item(), "scale", want.abs().max().item())

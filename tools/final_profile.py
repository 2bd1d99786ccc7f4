"""One launch of each headline kernel on L2-busting inputs, for `ncu --set full` (round-end capture):
K1 / K2 on a 16-image ResNet-50 blob batch, K5, the QDrop epilogue, Relu and the direct stem convolution."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402
from dipoorlet_b200.workloads import resnet50_blob_shapes  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
tensors = []
for i, shp in enumerate(resnet50_blob_shapes()):
    t = torch.randn((16,) + tuple(shp), device=dev, generator=g)
    tensors.append(torch.relu_(t) if i % 2 else t)
batch = K.BlobBatch(tensors)
n = batch.n_segments
smin, smax = torch.empty(n, device=dev), torch.empty(n, device=dev)
bmin = torch.full((batch.n_blobs,), float("inf"), device=dev)
bmax = torch.full((batch.n_blobs,), float("-inf"), device=dev)
K.segstats(batch, smin, smax, None, None, bmin, bmax)
dm = torch.empty(batch.n_blobs, device=dev)
K.absmax(bmin, bmax, dm)
counts = torch.zeros((batch.n_blobs, 2048), dtype=torch.int64, device=dev)
K.hist_abs(batch, dm, counts, 2048)
x = torch.randn((64, 256, 56, 56), device=dev, generator=g)
o = torch.empty_like(x)
K.fakequant(x, torch.tensor([0.05], device=dev), None, -128, 127, out=o)
K.recon_act(x, True, (0.05, -127.0, 127.0), prob=0.5, seed=3, out=o)
K.clip(x, 0.0, float("inf"), out=o)
img = torch.randn((64, 3, 224, 224), device=dev, generator=g)
w = torch.randn((64, 3, 7, 7), device=dev, generator=g) * 0.05
K.conv_direct_forward(img, w, None, 2, 3)
torch.cuda.synchronize()
print("done")

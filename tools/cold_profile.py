"""Where a COLD (first call in a fresh process) `-A hist` calibration spends its wall time:
cProfile of tensor_calibration on ResNet-50, 1024 images, from pinned host buffers, followed by a
second (warm) call for comparison. Usage: python tools/cold_profile.py [images] [calib_bs]"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import forward_net as fwd, workloads as W  # noqa: E402
from dipoorlet_b200.cli_args import make_args  # noqa: E402
from dipoorlet_b200.graph import ONNXGraph  # noqa: E402
from dipoorlet_b200.tensor_cali import tensor_calibration  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
algo = sys.argv[3] if len(sys.argv) > 3 else "hist"
graph = ONNXGraph(W.build_resnet50(seed=0), "/tmp/dpl_cold", "trt")
images = W.synthetic_images(n_img, seed=0)[:, 0]
t0 = time.perf_counter()
src = fwd.ArrayInput({"input": images})
torch.cuda.synchronize()
print("pin host images: %.3f s" % (time.perf_counter() - t0))
args = make_args(input_dir=src, data_num=n_img, deploy="trt", act_quant=algo, bins=2048,
                 output_dir="/tmp/dpl_cold", calib_bs=bs)
for label in ("cold", "warm"):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    tensor_calibration(graph, args)
    torch.cuda.synchronize()
    pr.disable()
    dt = time.perf_counter() - t0
    print("== %s call: %.3f s (%.0f images/s), reserved %.1f GB" %
          (label, dt, n_img / dt, torch.cuda.memory_reserved() / 1e9))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print("\n".join(s.getvalue().splitlines()[4:40]))

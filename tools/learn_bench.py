"""Iterations/s of the rounding loop, eager launch sequence vs CUDA-graph replay."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import onnx_lite as ol  # noqa: E402
from dipoorlet_b200.weight_transform.ada_quant_layer import AdaQLayer, adaround_reg  # noqa: E402
from dipoorlet_b200.weight_transform.learning import learning_round_mask  # noqa: E402

dev = torch.device("cuda")


def layer(ci, co, k, groups, relu=True):
    w = torch.randn((co, ci // groups, k, k), device=dev) * 0.1
    attrs = {"dilations": [1, 1], "group": groups, "kernel_shape": [k, k], "pads": [k // 2] * 4, "strides": [1, 1]}
    scale = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous()
    return AdaQLayer(ol.Node("Conv", ["x", "w"], ["y"], "c", attrs), w, None, scale, -127, 127, relu, device=dev)


cases = {
    "mbv2 depthwise 3x3, 96ch, 56x56": (lambda: [layer(96, 96, 3, 96)], (512, 96, 56, 56)),
    "mbv2 pointwise 1x1 96->24, 56x56": (lambda: [layer(96, 24, 1, 1, relu=False)], (512, 96, 56, 56)),
    "r50 bottleneck 256->64->64->256, 56x56": (lambda: [layer(256, 64, 1, 1), layer(64, 64, 3, 1), layer(64, 256, 1, 1)],
                                               (256, 256, 56, 56)),
}
rows = []
for name, (mk, xshape) in cases.items():
    x = torch.randn(xshape, device=dev)
    xq = torch.round(x / 0.05) * 0.05
    with torch.no_grad():
        ls = mk()
        t = x
        for l in ls:
            t = l.dense_forward(t, l.weight)
            if l.relu_flag:
                t = torch.relu(t)
    for mode in ("0", "1"):
        os.environ["DPL_CUDA_GRAPH"] = mode
        ls = mk()
        epochs = 100
        n_it = epochs * (xshape[0] // 64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        learning_round_mask(ls, xq, t, adaround_reg(n_it), 64, epochs, log_every=10 ** 9)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rows.append({"case": name, "cuda_graph": mode == "1", "iterations": n_it, "ms_per_iteration": 1e3 * dt / n_it})
        print(json.dumps(rows[-1]))
json.dump(rows, open("gpurun_out/learn_bench.json", "w"), indent=1)

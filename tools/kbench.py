"""Kernel microbenchmark on the real ResNet-50 blob shapes (run on the GPU box).

    python tools/kbench.py [--batch 32] [--out gpurun_out/kbench.json]
Times K1 / K2 (all variants) / K4 with CUDA events, L2 flushed by construction (the
batch is several GB), and prints achieved GB/s against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import kernels as K  # noqa: E402
from dipoorlet_b200.workloads import resnet50_blob_shapes  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/kbench.json")
    ap.add_argument("--quick", action="store_true", help="one data mode, one launch per kernel (for ncu)")
    args = ap.parse_args()
    peak = 6483.3
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda:0")
    shapes = resnet50_blob_shapes()
    res = {"batch": args.batch, "peak_gbs": peak, "rows": []}
    if args.quick:
        global timed
        timed = lambda fn, iters=1, warm=0: (fn(), torch.cuda.synchronize(), (1.0, 1.0))[2]  # noqa: E731
    for mode in (("mixed",) if args.quick else ("mixed", "dense", "relu")):
        g = torch.Generator(device=dev).manual_seed(0)
        tensors = []
        for i, shp in enumerate(shapes):
            t = torch.randn((args.batch,) + tuple(shp), device=dev, generator=g)
            if mode == "relu" or (mode == "mixed" and i % 2 == 1):
                t = torch.relu_(t)
            tensors.append(t.contiguous())
        batch = K.BlobBatch(tensors)
        nbytes = batch.elements * 4
        n = batch.n_segments
        smin = torch.empty(n, dtype=torch.float32, device=dev)
        smax = torch.empty_like(smin)
        ssum = torch.empty(n, dtype=torch.float64, device=dev)
        snnz = torch.empty(n, dtype=torch.int64, device=dev)
        bmin = torch.full((batch.n_blobs,), float("inf"), device=dev)
        bmax = torch.full((batch.n_blobs,), float("-inf"), device=dev)
        ws = K.Workspace(dev)
        med, best = timed(lambda: K.segstats(batch, smin, smax, ssum, snnz, bmin, bmax, ws))
        res["rows"].append({"kernel": "K1 segstats", "mode": mode, "ms": med, "ms_min": best,
                            "gbs": nbytes / med / 1e6, "frac": nbytes / med / 1e6 / peak})
        dm = torch.empty(batch.n_blobs, device=dev)
        K.absmax(bmin, bmax, dm)
        for variant in ((4, 7) if args.quick else (1, 4, 5, 7, 3, 2)):
            counts = torch.zeros((batch.n_blobs, 2048), dtype=torch.int64, device=dev)
            med, best = timed(lambda: K.hist_abs(batch, dm, counts, 2048, variant=variant))
            res["rows"].append({"kernel": f"K2 hist v{variant}", "mode": mode, "ms": med,
                                "ms_min": best, "gbs": nbytes / med / 1e6,
                                "frac": nbytes / med / 1e6 / peak})
        s = torch.empty(n, dtype=torch.float32, device=dev)
        iters = torch.empty(n, dtype=torch.int32, device=dev)
        med, best = timed(lambda: K.octav(batch, ssum, snnz, 1 / 4 ** 8 / 3, s, iters, workspace=ws),
                          iters=3, warm=1)
        res["rows"].append({"kernel": "K4 octav", "mode": mode, "ms": med, "ms_min": best,
                            "gbs": nbytes / med / 1e6, "frac": nbytes / med / 1e6 / peak,
                            "mean_iters": float(iters.float().mean().item())})
        # plain read bandwidth reference: torch.sum over the same bytes, one launch per blob
        med, best = timed(lambda: [t.sum() for t in tensors], iters=5)
        res["rows"].append({"kernel": "torch.sum per blob (123 launches)", "mode": mode, "ms": med,
                            "ms_min": best, "gbs": nbytes / med / 1e6,
                            "frac": nbytes / med / 1e6 / peak})
        del tensors, batch
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for r in res["rows"]:
        print(f"{r['mode']:6s} {r['kernel']:36s} {r['ms']:9.3f} ms  {r['gbs']:8.1f} GB/s  "
              f"{100 * r['frac']:5.1f}% of {peak:.0f}" + (f"  iters={r['mean_iters']:.1f}" if "mean_iters" in r else ""))


if __name__ == "__main__":
    main()

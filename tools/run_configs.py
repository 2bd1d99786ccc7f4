"""Run BASELINE.json's configurations end to end through the plugin API on synthetic data
and print one JSON line per config (wall time incl. host work, CUDA-synchronised).

    python tools/run_configs.py [--configs 2 3 4 5] [--ada-epoch 20] [--images N]

Config numbering follows BASELINE.json `configs` (0-based index + 1):
  2  ResNet-50, 1024 img, -A hist --bins 2048 -D trt
  3  ResNet-50, 1024 img, -A mse --bc -D trt
  4  MobileNetV2, 512 img, -A minmax --adaround -D trt     (--ada_epoch reduced, stated)
  5  ResNet-50, 1024 img per GPU, -A hist --brecq --drop -D trt (--ada_epoch reduced, stated)
"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dipoorlet_b200 import dist_helper, forward_net as fwd, kernels as K, workloads as W  # noqa: E402
from dipoorlet_b200.cli_args import make_args  # noqa: E402
from dipoorlet_b200.deploy import to_deploy  # noqa: E402
from dipoorlet_b200.graph import ONNXGraph  # noqa: E402
from dipoorlet_b200.profiling import quantize_profiling_multipass  # noqa: E402
from dipoorlet_b200.tensor_cali import tensor_calibration  # noqa: E402
from dipoorlet_b200.utils import load_clip_val, save_clip_val  # noqa: E402
from dipoorlet_b200.weight_transform import weight_calibration  # noqa: E402


def run(cfg, n_img, ada_epoch, out):
    rank, local_rank, world = dist_helper.init_from_env()
    kw = {2: dict(model="r50", act_quant="hist"), 3: dict(model="r50", act_quant="mse", bc=True),
          4: dict(model="mbv2", act_quant="minmax", adaround=True),
          5: dict(model="r50", act_quant="hist", brecq=True, drop=True)}[cfg]
    mname = kw.pop("model")
    model = W.build_resnet50(seed=0) if mname == "r50" else W.build_mobilenetv2(seed=0)
    tmp = [tempfile.mkdtemp(prefix=f"dpl_cfg{cfg}_") if rank == 0 else None]
    if world > 1:   # one output directory for all ranks (the CLI gets it from -O)
        torch.distributed.broadcast_object_list(tmp, src=0)
    tmp = tmp[0]
    graph = ONNXGraph(model, tmp, "trt")
    images = W.synthetic_images(n_img, seed=0, start=rank * n_img)[:, 0]
    args = make_args(input_dir=fwd.ArrayInput({"input": images}, start=rank * n_img), data_num=n_img * world,
                     deploy="trt", output_dir=tmp, bins=2048, ada_bs=64, ada_epoch=ada_epoch, calib_bs=32,
                     rank=rank, local_rank=local_rank, world_size=world, **kw)
    torch.cuda.synchronize()
    t = {}
    l0 = K.launches()
    t0 = time.perf_counter()
    act, weight = tensor_calibration(graph, args)
    torch.cuda.synchronize()
    t["calibration_s"] = time.perf_counter() - t0
    if rank == 0:
        save_clip_val(copy.deepcopy(act), copy.deepcopy(weight), args)
    dist_helper.barrier()
    act, weight = load_clip_val(args)
    t1 = time.perf_counter()
    g2, g_ori, act, weight = weight_calibration(graph, act, weight, args)
    torch.cuda.synchronize()
    t["weight_transform_s"] = time.perf_counter() - t1
    t2 = time.perf_counter()
    layer, mcos, _ = quantize_profiling_multipass(g2, g_ori, copy.deepcopy(act), copy.deepcopy(weight), args)
    torch.cuda.synchronize()
    t["profiling_s"] = time.perf_counter() - t2
    if rank == 0:
        to_deploy(g2, act, weight, args)
    total = time.perf_counter() - t0
    if rank == 0:
        line = {"config": cfg, "model": mname, "images_per_gpu": n_img, "n_gpus": world, "flags": kw,
                "ada_epoch": ada_epoch if (kw.get("adaround") or kw.get("brecq")) else None,
                "total_s": total, "calibration_images_per_s": n_img * world / t["calibration_s"], **t,
                "gpu_launches": K.launches() - l0,
                "model_output_cos": {k: [float(v[0]), float(v[1])] for k, v in mcos.items()},
                "min_layer_cos": float(min(layer.values())) if layer else None,
                "trt_file": os.path.join(tmp, "trt_clip_val.json")}
        print(json.dumps(line))
        if out:
            with open(out, "a") as f:
                f.write(json.dumps(line) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, nargs="+", default=[2, 3, 4, 5])
    ap.add_argument("--ada-epoch", type=int, default=20)
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/configs.jsonl")
    a = ap.parse_args()
    for cfg in a.configs:
        n = a.images or (512 if cfg == 4 else 1024)
        run(cfg, n, a.ada_epoch, a.out)
        fwd._SESSIONS.clear()
        torch.cuda.empty_cache()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

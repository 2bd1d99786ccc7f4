"""Round-2 verification of DPL_PEER_ALLREDUCE=1 on N GPUs of one box:

    gpurun --gpus 2 -- 'timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tools/peer_step_check.py'

Runs adaround on the small ResNet fixture twice per rank — NCCL all-reduce + step, then the peer step — and
compares the rounded weights (for 2 ranks the sums are commutative: identical; for more ranks NCCL's
reduction order may differ from rank order in the last bit) and the time per layer. Rank 0 prints one JSON line."""
import copy
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    torch.distributed.init_process_group("nccl")
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.cli_args import make_args
    from dipoorlet_b200.forward_net import ArrayInput
    from dipoorlet_b200.graph import ONNXGraph
    from dipoorlet_b200.tensor_cali import tensor_calibration
    from dipoorlet_b200.weight_transform import weight_calibration
    d = os.path.join(ROOT, "tests", "golden", "tiny_r50")
    model = ol.load(os.path.join(d, "model.onnx"))
    images = np.load(os.path.join(d, "images.npy"))
    out = {}
    for mode in ("0", "1"):
        os.environ["DPL_PEER_ALLREDUCE"] = mode
        tmp = f"/tmp/dpl_peer_{mode}_{rank}"
        os.makedirs(tmp, exist_ok=True)
        graph = ONNXGraph(model, tmp, "trt")
        args = make_args(input_dir=ArrayInput({"input": images[:, 0]}), data_num=images.shape[0], deploy="trt",
                         act_quant="minmax", output_dir=tmp, calib_bs=8, adaround=True, ada_bs=2, ada_epoch=40,
                         rank=rank, local_rank=rank, world_size=world)
        act, weight = tensor_calibration(graph, args)
        torch.cuda.synchronize()
        t0 = time.time()
        g2, _, _, _ = weight_calibration(graph, act, copy.deepcopy(weight), args)
        torch.cuda.synchronize()
        out[mode] = (time.time() - t0, {n.input[1]: g2.get_initializer(n.input[1]) for n in graph.graph.node
                                        if n.op_type in ("Conv", "Gemm")})
    same = sum(int((out["0"][1][k] == out["1"][1][k]).sum()) for k in out["0"][1])
    total = sum(v.size for v in out["0"][1].values())
    if rank == 0:
        print(json.dumps({"world": world, "seconds_nccl": round(out["0"][0], 3), "seconds_peer": round(out["1"][0], 3),
                          "identical_weights": same, "weights": total}))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

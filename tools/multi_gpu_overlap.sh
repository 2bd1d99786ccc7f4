#!/bin/bash
# NCCL all-reduce of dL/dW overlapped with the earlier layers' backward (default) vs issued in line.
N=${1:-2}; BLOCKS=${2:-1,7,12,18,20,21}; EPOCH=${3:-4}
mkdir -p gpurun_out
for ov in 1 0; do
  DPL_OVERLAP_ALLREDUCE=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --workload finetune --gpus $N --ft-blocks "$BLOCKS" --ft-epoch $EPOCH --steps 2 --warmup 1 \
    2> gpurun_out/ft_n${N}_ov$ov.err | tail -1 > gpurun_out/ft_n${N}_ov$ov.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ft_n${N}_ov$ov.json"))
    print("overlap=$ov N=$N", round(d["value"],1), "it/s", "identical:", d.get("replicas_bit_identical"), [round(b["loop_ms_per_iteration"],3) for b in d["per_block"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/ft_n${N}_ov$ov.err").read()[-1500:])
PY
done

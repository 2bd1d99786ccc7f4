#!/bin/bash
# Rounding finetune on N GPUs of one box (BASELINE.json configs[4] shape: brecq + drop, 1 rank per GPU, the weight
# gradients averaged per iteration): NCCL all-reduce per layer vs the peer (NVLink) step kernel.
#   gpurun --gpus N -- 'bash tools/multi_gpu_finetune.sh N [blocks] [epoch]'
N=${1:-2}; BLOCKS=${2:-1,7,12,18,20,21}; EPOCH=${3:-4}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --workload finetune --gpus $N --ft-blocks "$BLOCKS" --ft-epoch $EPOCH --steps 2 --warmup 1 \
    2> gpurun_out/ft_n${N}_$name.err | tail -1 > gpurun_out/ft_n${N}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ft_n${N}_$name.json"))
    print("$name N=$N", round(d["value"],1), "it/s", "identical:", d.get("replicas_bit_identical"), [round(b["loop_ms_per_iteration"],3) for b in d["per_block"]])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/ft_n${N}_$name.err").read()[-1500:])
PY
}
run nccl DPL_PEER_ALLREDUCE=0
run peer DPL_PEER_ALLREDUCE=1
if [ "$N" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_step_check.py 2>gpurun_out/peer_check.err | tail -1
fi

#!/bin/bash
# Round-2 measurement pass on one GPU: ncu traffic of the K2 launch of the bench command, the bench lines of
# the three calibrators and of the two finetune workloads (ours and the reference's torch loop).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:hist -s 6 -c 1 -o gpurun_out/prof_hist_bench -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_hist_bench.log 2>&1
for a in hist mse minmax; do
  timeout 600 python bench.py --algo $a --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$a.err | tail -1 > gpurun_out/bench_$a.json
done
timeout 900 python bench.py --workload finetune --ft-model mbv2 --ft-algo adaround --ft-epoch 2 --steps 2 --warmup 1 2>gpurun_out/ft_mbv2.err | tail -1 > gpurun_out/ft_mbv2_ours.json
timeout 900 python bench.py --workload finetune --ft-model mbv2 --ft-algo adaround --ft-epoch 2 --steps 2 --warmup 1 --impl reference 2>gpurun_out/ft_mbv2_ref.err | tail -1 > gpurun_out/ft_mbv2_ref.json
timeout 600 python bench.py --workload finetune --steps 2 --warmup 1 2>gpurun_out/ft_r50.err | tail -1 > gpurun_out/ft_r50_ours.json
python - <<'PY'
import json
for f in ["bench_hist","bench_mse","bench_minmax","ft_mbv2_ours","ft_mbv2_ref","ft_r50_ours"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), d["unit"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY

#!/bin/bash
# One GPU lease, many measurements (each gpurun call is charged a ~10 min minimum).
# Every step has its own timeout and writes straight into gpurun_out/ so that a hang in
# one step loses nothing else.  Usage: tools/gpu_session.sh [steps...]  (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS="${@:-tests kbench bench ref launches ncu smoke}"
for s in $STEPS; do
  echo "=== $s $(date +%T)" | tee -a gpurun_out/session.log
  case $s in
    tests)   timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; grep -E 'passed|failed|FAILED|Error' gpurun_out/pytest_gpu.log | tail -30 ;;
    kbench)  timeout 300 python tools/kbench.py --batch 32 > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log | tail -40 ;;
    bench)   timeout 600 python bench.py --steps 3 --warmup 3 --hist-variant ${HIST_VARIANT:-0} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
    ref)     timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
                 --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --images 64 --no-cpu-baseline \
                 > gpurun_out/launches_run.log 2>&1; tail -3 gpurun_out/launches_run.log ;;
    ncu)     timeout 600 ncu --set full --clock-control none --import-source on \
                 -k regex:'hist_lc|hist_lanecol|segstats_tiles|octav' -c 8 -f -o gpurun_out/prof_r1 \
                 python tools/kbench.py --batch 16 --quick > gpurun_out/ncu_run.log 2>&1; tail -3 gpurun_out/ncu_run.log ;;
    configs) timeout 1200 python tools/run_configs.py --configs ${CONFIGS:-2 3 4 5} --ada-epoch ${ADA_EPOCH:-10} > gpurun_out/configs.log 2>&1; tail -8 gpurun_out/configs.log ;;
    benchcb) DPL_CUDNN_BENCHMARK=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cudnn_benchmark.json 2> gpurun_out/bench_cb.err; tail -c 1500 gpurun_out/bench_cudnn_benchmark.json; tail -3 gpurun_out/bench_cb.err ;;
    batch)   for b in ${BATCHES:-32 128}; do timeout 600 python bench.py --steps 2 --warmup 3 --batch $b --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch', d['config']['forward_batch'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'hist frac', round(d['roofline']['frac'],3))"; done ;;
    ncuhist) timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hist_lc3|segstats_tiles' -c 2 -f -o gpurun_out/prof_hist_b32 \
                 python tools/kbench.py --batch 32 --quick > gpurun_out/ncu_hist_run.log 2>&1; tail -2 gpurun_out/ncu_hist_run.log ;;
    convbench) timeout 300 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; tail -6 gpurun_out/conv_bench.log ;;
    convtest) timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k 'conv3x3 or conv_strided or 3xtf32' > gpurun_out/conv_test.log 2>&1; tail -15 gpurun_out/conv_test.log ;;
    engprof) timeout 300 python tools/engine_profile.py 64 > gpurun_out/engine_profile.log 2>&1; tail -22 gpurun_out/engine_profile.log
             DPL_ENGINE_CONV3X3=0 PROFILE_TAG=_cudnn3x3 timeout 300 python tools/engine_profile.py 64 > gpurun_out/engine_profile_cudnn3x3.log 2>&1; tail -22 gpurun_out/engine_profile_cudnn3x3.log ;;
    smoke)   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ;;
  esac
done
echo "=== done $(date +%T)" | tee -a gpurun_out/session.log

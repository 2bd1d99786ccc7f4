#!/bin/bash
# First GPU call of the next round: verifies and measures the three paths that were written after the round-1
# GPU budget was spent (all opt-in, none on by default). One GPU, bounded by timeouts; results under gpurun_out/.
#
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# 1. conv_taps_ts_tf32x3_kernel (DPL_TAPS_TS=1): parity tests of the tap-table convolutions, then the shape sweep
#    next to the default kernel;
# 2. dpl_adaround_step_peer_f32: world-1 equivalence and the two-streams emulation (tests gated by
#    DPL_TEST_EXPERIMENTAL=1); the 2-GPU check is tools/peer_step_check.py under gpurun --gpus 2;
# 3. the GPU tests added without hardware (--update_bn, --sparse) and the bench with / without DPL_TAPS_TS.
mkdir -p gpurun_out
{
  echo "== taps TS parity"; DPL_TAPS_TS=1 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward_ops.py -q -m gpu -x 2>&1 | tail -5
  echo "== conv sweep default"; CONV_BATCH=128 timeout 200 python tools/conv_bench.py 2>&1 | head -10
  echo "== conv sweep TS"; DPL_TAPS_TS=1 CONV_BATCH=128 timeout 200 python tools/conv_bench.py 2>&1 | head -10
  echo "== peer step"; DPL_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_peer_step.py -q -m gpu 2>&1 | tail -5
  echo "== late tests"; timeout 300 python -m pytest tests/test_gpu_zz_update_bn.py tests/test_gpu_zz_sparse.py -q -m gpu 2>&1 | tail -5
  for ts in 0 1; do echo "== bench DPL_TAPS_TS=$ts"; DPL_TAPS_TS=$ts timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1; done
} > gpurun_out/round2_first_call.log 2>&1
tail -40 gpurun_out/round2_first_call.log

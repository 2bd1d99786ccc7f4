#!/bin/bash
# DPL_X3_CHUNK sweep: accuracy of the 3xTF32 forward kernels (tools/x3_accuracy.py) and the hist job rate.
mkdir -p gpurun_out
for c in 1 2; do
  echo "== DPL_X3_CHUNK=$c"
  DPL_X3_CHUNK=$c python tools/x3_accuracy.py 2>&1 | tail -7 | cut -c1-110
  DPL_X3_CHUNK=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'])"
done

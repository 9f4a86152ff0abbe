#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for mode in 1 0; do
  export DCAE_F16_PAIR=$mode
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "f16x3" -x > gpurun_out/ops_f16pair$mode.log 2>&1; echo "pair=$mode ops exit $?"; tail -n 3 gpurun_out/ops_f16pair$mode.log | cut -c1-300
  timeout 300 python tools/gemm_knobs.py quick f16x3 > gpurun_out/knobs_f16pair$mode.log 2>&1; echo "pair=$mode knobs exit $?"; cat gpurun_out/knobs_f16pair$mode.log | cut -c1-200
done
export DCAE_F16_PAIR=1
for bn in 256 160; do DCAE_TC_BN=$bn python - <<'PY'
import os, sys
sys.argv = ["x", "quick", "f16x3"]
PY
done
timeout 300 python -m pytest tests/test_gpu_slice_loop.py -q -m gpu -s -k "f16x3" 2>&1 | grep -E "^\[slice|passed|failed" | cut -c1-250
timeout 300 python bench.py --no-cpu-baseline --steps 10 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PAIR bench', round(d['value']), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in d['kernel_families'].items()})"

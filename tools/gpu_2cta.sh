#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
export DCAE_TC_2CTA=1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tf32" -x > gpurun_out/ops_2cta.log 2>&1; echo "ops_2cta exit $?"; tail -n 15 gpurun_out/ops_2cta.log | cut -c1-300
timeout 300 python tools/gemm_knobs.py quick > gpurun_out/knobs_2cta.log 2>&1; echo "knobs exit $?"; cat gpurun_out/knobs_2cta.log | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_slice_loop.py -q -m gpu -s -k "tf32x3 or invariance or kodak" > gpurun_out/loop_2cta.log 2>&1; echo "loop_2cta exit $?"; grep -E "^\[|passed|failed" gpurun_out/loop_2cta.log | cut -c1-300
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_2cta.log 2>&1; echo "bench exit $?"; python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/bench_2cta.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('2cta bench', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), {k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in d['kernel_families'].items()}, d['clocks'])
except Exception as e: print('no bench line', e)
PY

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_slice_loop.py -x -q -s -k "fast_math or range_coder" 2>&1 | tail -n 12
python bench.py --math f16 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c2_f16.json 2> gpurun_out/bench_c2_f16.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_c2_f16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_f16.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],3),v['launches_per_step']) for k,v in d['kernel_families'].items()})
PY
python tools/layer_times.py f16 2>&1 | tail -n 50 > gpurun_out/layer_times_f16.log; head -30 gpurun_out/layer_times_f16.log

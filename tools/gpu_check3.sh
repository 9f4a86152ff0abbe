#!/bin/bash
# after the batched weight packing: everything that packs weights (slice loop, modules, training, drop-in) + smoke + the training bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r2_smoke.log
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "bench c4 rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c4.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "phases_ms", "torch_gpu_baseline"):
        print(k, json.dumps(d.get(k))[:300])
except Exception as e:
    print("no line", e)
P

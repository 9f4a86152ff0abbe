#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_transforms.py -q -m gpu -s -k "container" > gpurun_out/r2_container.log 2>&1; echo "container rc=$?"; grep -E "container:|passed|failed|Error" gpurun_out/r2_container.log | tail -n 4
for b in 1 4; do
timeout 600 python bench.py --config 6 --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_c6_b$b.json 2> gpurun_out/r2_bench_c6_b$b.err; echo "b=$b rc=$?"
python - <<P
import json
d = json.loads(open("gpurun_out/r2_bench_c6_b$b.json").read().strip().splitlines()[-1])
print("B=$b", "ms", round(d["ms_per_step"], 2), "img/s", round(d["value"], 1), "launches", d["gpu_launches_per_step"], "stages", {k: round(v, 2) for k, v in d["stages_ms"].items()})
P
done

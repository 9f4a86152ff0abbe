#!/bin/bash
# round 2, first GPU pass: parity tests, smoke, the default bench line, sanitizers.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_summary.txt
python bench.py > gpurun_out/r2_bench_c2.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_summary.txt
cat gpurun_out/r2_summary.txt
tail -n 15 gpurun_out/r2_pytest.log
tail -n 4 gpurun_out/r2_smoke.log
tail -n 1 gpurun_out/r2_bench_c2.log | cut -c1-1500

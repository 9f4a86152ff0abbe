#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the reference arm, the default bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/final_smoke.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; tail -n 1 gpurun_out/final_bench_ref.json | cut -c1-260
python bench.py --gpus 1 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -n 1 gpurun_out/final_bench.json | cut -c1-260

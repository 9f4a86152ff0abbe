#!/bin/bash
# quick iteration: kernel-3 tests + micro, then the whole GPU suite
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_gc.py -x -q -s > gpurun_out/r2_gc_tests.log 2>&1; echo "gc tests rc=$?"
tail -n 12 gpurun_out/r2_gc_tests.log
python tools/gc_micro.py 2>&1 | tee gpurun_out/r2_gc_micro.log
python -m pytest tests -m gpu -q -s -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r2_pytest.log

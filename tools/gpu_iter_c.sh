#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_gc.py -x -q -s 2>&1 | tail -n 6
python tools/gc_micro.py 2>&1 | tee gpurun_out/r2_gc_micro.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6

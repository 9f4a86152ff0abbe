#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_training.py tests/test_gpu_reference_dropin.py -x -q -s 2>&1 | tail -n 30

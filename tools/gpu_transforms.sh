#!/bin/bash
# new transform-stack kernels: op-level + stack-level GPU tests, then the untouched op tests (the ReLU epilogue touched every GEMM kernel)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_transforms.py -q -m gpu -s > gpurun_out/r2_transforms.log 2>&1; echo "transforms rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_transforms.log | tail -n 3
grep -E "vs the|FAILED|Error|error" gpurun_out/r2_transforms.log | head -n 60
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r2_ops_after_relu.log 2>&1; echo "ops rc=$?"; tail -n 2 gpurun_out/r2_ops_after_relu.log

#!/bin/bash
# A/B of the single-CTA and pair GEMM kernels: correctness (op tests) + tile sweep
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for mode in 0 1; do
  export DCAE_TC_2CTA=$mode
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tf32" -x > gpurun_out/ops_ab$mode.log 2>&1; echo "2cta=$mode ops exit $?"; tail -n 3 gpurun_out/ops_ab$mode.log | cut -c1-300
  timeout 300 python tools/gemm_knobs.py quick > gpurun_out/knobs_ab$mode.log 2>&1; echo "2cta=$mode knobs exit $?"; cat gpurun_out/knobs_ab$mode.log | cut -c1-200
done

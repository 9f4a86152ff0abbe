#!/bin/bash
# One gpurun call: staged GPU checks, each in its own process (a CUDA fault in one stage cannot poison the next).
#   gpurun --timeout 1500 -- bash tools/gpu_check.sh
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run gc        300 python -m pytest tests/test_gpu_gc.py -q -x -m gpu
run ops_simt  400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not tf32"
run loop_fp32 600 python -m pytest tests/test_gpu_slice_loop.py -q -m gpu -s -k "fp32 or training or validation"
run ops_tc    400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tf32 or f16x3"
run loop_tc   600 python -m pytest tests/test_gpu_slice_loop.py -q -m gpu -s -k "not (fp32 or training or validation)"
run modules   300 python -m pytest tests/test_gpu_modules.py -q -m gpu
run smoke     300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_tc  600 python bench.py
run bench_fp32 600 python bench.py --steps 2 --warmup 3 --math fp32 --no-cpu-baseline
run bench_tf32x3 600 python bench.py --math tf32x3 --no-cpu-baseline
run bench_tf32 600 python bench.py --steps 20 --warmup 3 --math tf32 --no-cpu-baseline
exit 0

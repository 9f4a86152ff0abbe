#!/bin/bash
# ncu --set full of the non-GEMM kernels of the slice loop (dwconv, layernorm, spatial gate), after a plain run.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gc-micro-mb 256"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dwconv|layernorm|spatial|channel_stats" -s 60 -c 10 -o gpurun_out/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_other.log
exit 0

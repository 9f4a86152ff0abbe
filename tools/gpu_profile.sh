#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then launch list + full sets.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gc-micro-mb 256"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 990 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 330 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gc_fused -s 20 -c 2 -o gpurun_out/prof_gc $CMD > gpurun_out/ncu_gc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dict_attention -s 15 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_*.log
exit 0

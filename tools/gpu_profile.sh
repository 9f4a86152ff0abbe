#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then launch list + full sets.
# One lane and a short warm-up so that launch indices are predictable; a step is 110 GEMM, 5 attention, 5 kernel-3 and
# 47 element-wise launches (+ 7 transposes / reduction at load and store).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-gpu-baseline --gc-micro-mb 256 --lanes 1"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_all.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 700 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gc_eval_hot -s 30 -c 2 -o gpurun_out/prof_gc $CMD > gpurun_out/ncu_gc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dict_attention -s 31 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dwconv|layernorm|spatial|channel_stats|nchw" -s 300 -c 12 -o gpurun_out/prof_other $CMD > gpurun_out/ncu_other.log 2>&1
tail -n 1 gpurun_out/plain.log | cut -c1-300
tail -n 2 gpurun_out/ncu_*.log | cut -c1-200
exit 0

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/transform_micro.py > gpurun_out/transform_micro_plain.log 2>&1 || { cat gpurun_out/transform_micro_plain.log; exit 1; }
cat gpurun_out/transform_micro_plain.log
# third forward = the timed one: 15 window-attention launches per forward -> skip 30, capture the first 4 (hd 8, 16, 16, 32)
ncu --set full --clock-control none --import-source on -k regex:window_attention -s 30 -c 4 -o gpurun_out/prof_window_attention python tools/transform_micro.py > gpurun_out/ncu_winattn.log 2>&1; tail -n 1 gpurun_out/ncu_winattn.log
# single-CTA GEMM launches of the third forward: level-1 layers first (conv 3->96, RBB 1x1 / 3x3 / 1x1, qkv, proj, fc1, fc2)
ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3_kernel -s 160 -c 12 -o gpurun_out/prof_transform_gemm python tools/transform_micro.py > gpurun_out/ncu_tgemm.log 2>&1; tail -n 1 gpurun_out/ncu_tgemm.log
python tools/ncu_summary.py gpurun_out/prof_window_attention.ncu-rep gpurun_out/prof_window_attention_summary.csv
python tools/ncu_summary.py gpurun_out/prof_transform_gemm.ncu-rep gpurun_out/prof_transform_gemm_summary.csv
cat gpurun_out/prof_window_attention_summary.csv | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/transform_micro.py > gpurun_out/transform_micro_plain.log 2>&1 || { cat gpurun_out/transform_micro_plain.log; exit 1; }
cat gpurun_out/transform_micro_plain.log
ncu --set full --clock-control none --import-source on -k regex:window_attention -s 30 -c 4 -o gpurun_out/prof_window_attention_mma python tools/transform_micro.py > gpurun_out/ncu_winattn.log 2>&1; tail -n 1 gpurun_out/ncu_winattn.log
python tools/ncu_summary.py gpurun_out/prof_window_attention_mma.ncu-rep gpurun_out/prof_window_attention_mma_summary.csv
cut -d, -f1,2,3,5,8,9,10,11,13,15,16,17,19 gpurun_out/prof_window_attention_mma_summary.csv

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/attn_micro.py > gpurun_out/attn_micro_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dict_attention_f16 -s 10 -c 1 -o gpurun_out/prof_attn_r02 python tools/attn_micro.py > gpurun_out/ncu_attn_r02.log 2>&1
cat gpurun_out/attn_micro_plain.log; tail -n 2 gpurun_out/ncu_attn_r02.log

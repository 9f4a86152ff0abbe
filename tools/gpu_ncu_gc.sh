#!/bin/bash
# ncu --set full of kernel 3 alone (1 GiB microbenchmark: the roofline_gc figure of bench.py)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/gc_micro.py > gpurun_out/gc_micro_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gc_fused -s 6 -c 2 -o gpurun_out/prof_gc_r02 python tools/gc_micro.py > gpurun_out/ncu_gc_r02.log 2>&1
tail -n 4 gpurun_out/gc_micro_plain.log; tail -n 3 gpurun_out/ncu_gc_r02.log
python -m pytest tests/test_gpu_reference_dropin.py -q -s -x 2>&1 | tail -n 15

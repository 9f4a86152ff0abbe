#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --config 6 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_c6_n8.json 2> gpurun_out/r2_bench_c6_n8.err; echo "rc=$?"
tail -n 1 gpurun_out/r2_bench_c6_n8.json | cut -c1-700; tail -n 3 gpurun_out/r2_bench_c6_n8.err

#!/bin/bash
# builder-run lines at 8 GPUs: configs #2 (e2e scaling diagnostics), #3, #5, #4.  gpurun --gpus 8 -- 'bash tools/gpu_bench_n8.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29521 --config 2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "c2 rc=$?"
run 29522 --config 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "c3 rc=$?"
run 29523 --config 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 rc=$?"
run 29524 --config 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
for c in 2 3 5 4; do tail -n 1 gpurun_out/bench_c${c}_n$N.json | cut -c1-300; tail -n 2 gpurun_out/bench_c${c}_n$N.err | cut -c1-300; done

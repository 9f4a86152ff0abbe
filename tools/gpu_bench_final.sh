#!/bin/bash
# final builder-run lines at one GPU (the code as committed): configs 2, 3, 5, 4 and the fast mode
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for c in 2 3 5 4; do
  python bench.py --config $c > gpurun_out/bench_c${c}_n1.json 2> gpurun_out/bench_c${c}_n1.err; echo "config $c rc=$?"
done
python bench.py --math f16 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_c2_f16.json 2> gpurun_out/bench_c2_f16.err; echo "f16 rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_c2_reference_arm.json 2> gpurun_out/bench_c2_reference_arm.err; echo "ref rc=$?"
for f in bench_c2_n1 bench_c3_n1 bench_c5_n1 bench_c4_n1 bench_c2_f16 bench_c2_reference_arm; do tail -n 1 gpurun_out/$f.json | cut -c1-240; done

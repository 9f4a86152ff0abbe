#!/bin/bash
# builder-run bench lines for the north star's "larger images" (BASELINE configs #3 and #5) + the headline, one GPU
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for c in 3 5 2; do
  python bench.py --config $c > gpurun_out/bench_c${c}_n1.json 2> gpurun_out/bench_c${c}_n1.err; echo "config $c rc=$?"
  tail -n 1 gpurun_out/bench_c${c}_n1.json | cut -c1-400
  tail -n 3 gpurun_out/bench_c${c}_n1.err
done

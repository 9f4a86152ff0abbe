#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck, synccheck and initcheck over the default-path
# kernels on the 2 x 7 x 9 golden slice loop.  Logs -> gpurun_out/sanitize_*.log (summaries copied to profiles/r02/).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck synccheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_$tool.log python tools/sanitize_case.py > gpurun_out/sanitize_${tool}_stdout.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_plain.log
  tail -n 3 gpurun_out/sanitize_$tool.log
done
tail -n 6 gpurun_out/sanitize_plain.log
exit 0

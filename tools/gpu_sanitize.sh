#!/bin/bash
# compute-sanitizer evidence over the default-path kernels on the 2 x 7 x 9 golden slice loop.
# ONE tool per gpurun call (B200_PROFILING.md: several sanitizer tools in one call have left a GPU unusable):
#   gpurun -- 'bash tools/gpu_sanitize.sh memcheck'      (then synccheck, racecheck, initcheck in later calls)
# Logs -> gpurun_out/sanitize_<tool>.log (summaries copied to profiles/r02/).
TOOL=${1:-memcheck}
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitize_$TOOL.log python tools/sanitize_case.py > gpurun_out/sanitize_${TOOL}_stdout.log 2>&1
echo "$TOOL rc=$?" | tee -a gpurun_out/sanitize_plain.log
tail -n 5 gpurun_out/sanitize_$TOOL.log
tail -n 2 gpurun_out/sanitize_${TOOL}_stdout.log
exit 0

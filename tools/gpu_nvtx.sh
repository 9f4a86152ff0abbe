#!/bin/bash
# DCAE_NVTX=1: operator families as NVTX ranges; ncu selects the launches of one family by range name
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
DCAE_NVTX=1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/nvtx_plain.log 2>&1; echo "plain rc=$?"; tail -n 1 gpurun_out/nvtx_plain.log
DCAE_NVTX=1 ncu --nvtx --nvtx-include "attention/" --metrics gpu__time_duration.sum --clock-control none -c 6 --csv --log-file gpurun_out/nvtx_attention_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/nvtx_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/nvtx_attention_launches.csv | cut -d, -f5,15 | head -n 8

#!/bin/bash
# role counters of the K = 640 layers in pair mode (why does M = 256 x BN = 224 not beat single-CTA BN = 128 there?)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
DCAE_F16_PAIR=1 DCAE_TC_BN=224 timeout 300 python tools/f16_dbg.py 2> gpurun_out/f16_dbg_pair224.log >/dev/null; echo "rc=$?"
grep -A 100 "=== step 2" gpurun_out/f16_dbg_pair224.log | grep "N=640 K=640" | head -n 6
DCAE_F16_PAIR=1 DCAE_TC_BN=256 timeout 300 python tools/f16_dbg.py 2> gpurun_out/f16_dbg_pair256.log >/dev/null; echo "rc=$?"
grep -A 100 "=== step 2" gpurun_out/f16_dbg_pair256.log | grep "N=640 K=640" | head -n 4

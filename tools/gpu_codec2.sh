#!/bin/bash
# transform stacks after the planes data flow + window attention v2: all transform tests, bench config 6, ncu launch list of it
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_transforms.py -q -m gpu -s > gpurun_out/r2_transforms.log 2>&1; echo "transforms rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_transforms.log | tail -n 3
grep -E "vs the|accelerated stack|free-running|codec:|FAILED|Error|assert" gpurun_out/r2_transforms.log | head -n 40
timeout 900 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2_bench_c6.json 2> gpurun_out/r2_bench_c6.err; echo "bench c6 rc=$?"; tail -n 5 gpurun_out/r2_bench_c6.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c6.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "gpu_launches_per_step", "stages_ms", "kernel_families", "roofline", "torch_gpu_baseline", "clocks"):
        print(k, json.dumps(d.get(k))[:400])
except Exception as e:
    print("no line", e)
P
if [ "$1" = "ncu" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_c6.csv python bench.py --config 6 --batch 4 --steps 1 --warmup 1 --warmup-seconds 0 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_c6.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_bench_c6.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = r[kn].split("(")[0].split("<")[0]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:60s} n={v[0]:5d} total={v[1]/1e6:9.3f} ms  {100*v[1]/tot:5.1f}%")
P
fi

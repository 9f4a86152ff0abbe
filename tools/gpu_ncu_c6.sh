#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CMD="python bench.py --config 6 --batch 4 --steps 1 --warmup 1 --warmup-seconds 0 --no-cpu-baseline --no-gpu-baseline"
$CMD > gpurun_out/c6_b4_plain.json 2> gpurun_out/c6_b4_plain.err || { tail -n 5 gpurun_out/c6_b4_plain.err; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_c6_b4.csv $CMD > gpurun_out/ncu_c6.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections, re
rows = list(csv.reader(open("gpurun_out/launches_bench_c6_b4.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = re.sub(r"^void ", "", r[kn]).split("(")[0]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k[:70]:70s} n={v[0]:5d} {v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f}%")
P

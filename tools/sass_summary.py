"""Static per-kernel SASS instruction counts of libdcae_b200.so -> profiles/r02/sass_summary.txt
(tcgen05.mma = UTCHMMA, .2CTA = cta_group::2; tcgen05.ld / st = LDTM / STTM; TMA = UTMALDG / UTMASTG;
tcgen05.commit = UTCBAR; mbarrier = SYNCS; HMMA = warp-level mma.sync, the window attention).   python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass_summary.txt")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "dcae_b200", "libdcae_b200.so")], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU", "HMMA"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line) if cur else None
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
with open(out_path, "w") as f:
    f.write(__doc__.strip().splitlines()[0] + "\n")
    f.write(f"{'kernel':96s} {'total':>6s} " + " ".join(f"{p:>12s}" for p in pats) + "\n")
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), names):
        if c["total"] == 0:
            continue
        tot.update(c)
        name = name.replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dcae::", "")[:94]
        f.write(f"{name:96s} {c['total']:6d} " + " ".join(f"{c[p]:12d}" for p in pats) + "\n")
    f.write(f"{'ALL KERNELS':96s} {tot['total']:6d} " + " ".join(f"{tot[p]:12d}" for p in pats) + "\n")
print(open(out_path).read().splitlines()[-1])

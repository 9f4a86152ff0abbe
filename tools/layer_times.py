"""Per-launch device time of one slice-loop step at config #2 (CUDA events around every op)."""
import os, sys, csv, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200 import _lib
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
math = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B, h, w = int(os.environ.get("LT_B", 16)), 32, 48
eng = EntropySliceLoop(init_entropy_params(0, "lively"), math=math, lanes=1)
g = torch.Generator().manual_seed(1)
x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
for _ in range(3):
    eng.forward(*x)
lib = _lib.load()
ms = (C.c_double * 4)(); work = (C.c_double * 4)(); cnt = (C.c_int64 * 4)()
lib.dcae_profile_start()
eng.forward(*x)
path = os.path.join(ROOT, "gpurun_out", f"layer_times_{math}.csv")
lib.dcae_profile_dump(path.encode(), ms, work, cnt)
rows = list(csv.DictReader(open(path)))
T = B * h * w
print("family totals ms:", [round(v, 3) for v in ms])
# slice 2 GEMMs in order
gemms = [r for r in rows if r["family"] == "0"]
per = len(gemms) // 5
names = ["x_trans", "msa_s", "in0", "out0", "in1", "out1", "in2", "out2", "proj", "q_trans", "linear", "fc1", "fc2", "output_trans",
         "cc1", "mean2", "scale2", "mean3", "scale3", "lrp1y", "lrp2", "lrp3"]
for sl in (0, 4):
    print(f"--- slice {sl}")
    for n, r in zip(names, gemms[sl * per:(sl + 1) * per]):
        fl = float(r["work"]); t = float(r["ms"])
        print(f"{n:13s} {t * 1e3:8.1f} us  {fl / (t * 1e-3) / 1e12 * 3:7.1f} TF-MMA  (N*K = {fl / (2 * T):.0f})")
others = [r for r in rows if r["family"] == "3"]
print("other ops (first slice):", [round(float(r["ms"]) * 1e3, 1) for r in others[:20]])

# idle time between consecutive ops of the (single-lane, single-stream) step: start of op i+1 minus end of op i
gaps = [float(b["start_ms"]) - (float(a["start_ms"]) + float(a["ms"])) for a, b in zip(rows, rows[1:])]
span = float(rows[-1]["start_ms"]) + float(rows[-1]["ms"])
print(f"step span {span:.3f} ms, sum of op times {sum(float(r['ms']) for r in rows):.3f} ms, sum of gaps {sum(gaps):.3f} ms "
      f"(mean {1e3 * sum(gaps) / len(gaps):.2f} us, max {1e3 * max(gaps):.1f} us over {len(gaps)} transitions)")

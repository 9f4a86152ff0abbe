"""A/B of EntropySliceLoop lanes in ONE process (same box, same thermal state): alternating timed blocks."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
B, h, w = 16, 32, 48
params = init_entropy_params(0, "lively")
engs = {l: EntropySliceLoop(params, math="f16x3", lanes=l) for l in (1, 2, 4)}
g = torch.Generator().manual_seed(1)
x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
outs = {l: e.forward(*x) for l, e in engs.items()}
for l, e in engs.items():
    for _ in range(3):
        e.forward(*x, out=outs[l])
res = {l: [] for l in engs}
for rnd in range(8):
    for l, e in engs.items():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            e.forward(*x, out=outs[l])
        e1.record()
        torch.cuda.synchronize()
        res[l].append(e0.elapsed_time(e1) / 10)
for l, v in res.items():
    print(f"lanes={l}: median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}  all {[round(t, 2) for t in v]}")

"""Host-side enqueue time of one slice-loop step (174 launches + their tensor-map encodes) vs its device time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
B, h, w = 16, 32, 48
params = init_entropy_params(0, "lively")
g = torch.Generator().manual_seed(1)
x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for lanes in (1, 2):
    eng = EntropySliceLoop(params, math="f16x3", lanes=lanes)
    out = eng.forward(*x)
    for _ in range(3):
        eng.forward(*x, out=out)
    torch.cuda.synchronize()
    for reuse in (True, False):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):                       # 3 steps = 522 launches: below the launch-queue depth
                eng.forward(*x, out=out) if reuse else eng.forward(*x)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            ts.append(((t1 - t0) / 3 * 1e3, (t2 - t0) / 3 * 1e3))
        print(f"lanes={lanes} out_reuse={reuse}: host enqueue {min(t[0] for t in ts):.2f} ms/step, wall {min(t[1] for t in ts):.2f} ms/step")

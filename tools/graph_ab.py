"""Does replaying the slice loop as a CUDA graph beat stream launches?  (config #2, f16x3)"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
SHAPES = [(16, 32, 48), (1, 16, 16), (1, 88, 128), (4, 16, 16)]
params = init_entropy_params(0, "lively")
g = torch.Generator().manual_seed(1)
def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
import itertools
for (B, h, w), lanes in itertools.product(SHAPES, (1, 2)):
    if lanes == 2 and B == 1:
        continue
    x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
    eng = EntropySliceLoop(params, math="f16x3", lanes=lanes)
    out = eng.forward(*x)
    for _ in range(20): eng.forward(*x, out=out)
    torch.cuda.synchronize()
    ref = {k: v.clone() for k, v in out.items()}
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.forward(*x, out=out)
    for v in out.values(): v.zero_()
    gr.replay(); torch.cuda.synchronize()
    same = all(torch.equal(out[k], ref[k]) for k in ("y_hat", "means", "scales", "likelihoods"))
    a = [timeit(lambda: eng.forward(*x, out=out)) for _ in range(5)]
    b = [timeit(gr.replay) for _ in range(5)]
    print(f"B={B} {h}x{w} lanes={lanes}: stream {statistics.median(a):.3f} ms  graph {statistics.median(b):.3f} ms  identical={same}")

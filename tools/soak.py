"""Determinism soak (compute-sanitizer is closed on this pool): the hand-rolled mbarrier / TMEM protocols of the
tcgen05 kernels run N times on the same inputs -- slice loop at config #2 with two lanes + side stream, compress path,
whole codec -- and every output must be the same BITS as in run 0.  A race or a missed barrier shows up as a differing
run (or as the 2-second barrier trap)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200 import DCAECodec
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
from dcae_b200.transforms import init_transform_params
N = int(os.environ.get("SOAK_N", 300))
P = init_entropy_params(0, "lively")
g = torch.Generator().manual_seed(1)
lat = [4 * torch.randn(16, 320, 32, 48, generator=g).cuda(), torch.randn(16, 320, 32, 48, generator=g).cuda(), torch.randn(16, 320, 32, 48, generator=g).cuda()]
for lanes in (2, 1):
    eng = EntropySliceLoop(P, lanes=lanes)
    ref = {k: v.clone() for k, v in eng.forward(*lat, want_symbols=True).items()}
    bad = 0
    t0 = time.perf_counter()
    for i in range(N):
        out = eng.forward(*lat, want_symbols=True)
        if i % 10 == 9 or i == N - 1:                    # comparing costs more than a step: every 10th run
            bad += sum(int(not torch.equal(out[k], ref[k])) for k in ref)
    torch.cuda.synchronize()
    print(f"slice loop config #2, lanes={lanes}: {N} runs in {time.perf_counter() - t0:.1f} s, differing outputs: {bad}", flush=True)
    assert bad == 0
PC = dict(P); PC.update(init_transform_params(0))
codec = DCAECodec(PC)
x = torch.rand(4, 3, 512, 768, generator=torch.Generator().manual_seed(2)).cuda()
ref = codec.forward(x)
bad = 0
for i in range(N // 3):
    out = codec.forward(x)
    bad += int(not torch.equal(out["x_hat"], ref["x_hat"])) + int(not torch.equal(out["likelihoods"]["y"], ref["likelihoods"]["y"])) + int(not torch.equal(out["para"]["y"], ref["para"]["y"]))
print(f"whole codec, 4 x 768x512: {N // 3} runs, differing outputs: {bad}", flush=True)
assert bad == 0
print("soak ok")

"""The reduced-precision fast mode (math="f16": hi planes only, one MMA pass) on the whole model: speed and measured error
against the fp32-parity mode on the same weights and images (secondary figure, never the headline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200 import DCAECodec
from dcae_b200.params import init_entropy_params
from dcae_b200.transforms import init_transform_params
P = dict(init_entropy_params(0, "lively")); P.update(init_transform_params(0))
x = torch.rand(16, 3, 512, 768, generator=torch.Generator().manual_seed(1234)).cuda()
res = {}
for math in ("f16x3", "f16"):
    codec = DCAECodec(P, math=math)
    for _ in range(3):
        out = codec.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = codec.forward(x)
    e1.record(); torch.cuda.synchronize()
    res[math] = (e0.elapsed_time(e1) / 5, out["para"]["y"].clone(), out["x_hat"].clone(), float(-torch.log2(out["likelihoods"]["y"]).sum() / (16 * 512 * 768)))
    del codec
    torch.cuda.empty_cache()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
(t3, y3, x3, b3), (t1, y1, x1, b1) = res["f16x3"], res["f16"]
print(f"whole codec 16 x 768x512: f16x3 {t3:.1f} ms ({16e3 / t3:.0f} images/s), f16 {t1:.1f} ms ({16e3 / t1:.0f} images/s)")
print(f"f16 vs f16x3: y {rel(y1, y3):.2e}, x_hat max abs {float((x1 - x3).abs().max()):.3e} / mse {float(((x1 - x3) ** 2).mean()):.2e}, bpp(y) {b1:.4f} vs {b3:.4f}")

"""One forward of the analysis transform g_a (and g_s) on the library at B x 768x512: the command the ncu captures of the
transform-stack kernels use (profiles/r02/prof_window_attention_summary.csv, prof_transform_gemm_summary.csv)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200.transforms import TransformStack, init_transform_params
B = int(os.environ.get("TM_B", 4))
ga = TransformStack("g_a", init_transform_params(0, ("g_a",)))
x = torch.rand(B, 3, 512, 768, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    y = ga(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ga(x); e1.record(); torch.cuda.synchronize()
print(f"g_a B={B}: {e0.elapsed_time(e1):.2f} ms, y {tuple(y.shape)} finite {bool(torch.isfinite(y).all())}")

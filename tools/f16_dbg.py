"""Role counters of every f16x3 GEMM launch of ONE slice-loop step at config #2 (DCAE_F16_DBG=1 -> stderr)."""
import os, sys
os.environ["DCAE_F16_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
B, h, w = 16, 32, 48
eng = EntropySliceLoop(init_entropy_params(0, "lively"), math="f16x3", lanes=1)
g = torch.Generator().manual_seed(1)
x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
sys.stderr.write("=== step 1 (cold)\n")
eng.forward(*x)
torch.cuda.synchronize()
sys.stderr.write("=== step 2\n")
eng.forward(*x)
torch.cuda.synchronize()

"""Kernel 1 alone at config #2's size (T = 24 576): CUDA-event time per launch, planes out as in the slice loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dcae_b200 import _lib
from dcae_b200.weights import f16_weight_planes
lib = _lib.load()
dev = torch.device("cuda:0")
T = int(os.environ.get("ATT_T", 24576))
g = torch.Generator().manual_seed(1)
q = torch.randn(T, 640, generator=g).to(dev)
Kh = (torch.randn(20, 128, 32, generator=g) * 0.5).to(dev)
Vh = torch.randn(20, 128, 32, generator=g).to(dev)
sc = (torch.rand(20, generator=g) + 0.5).to(dev)
s = _lib.current_stream(dev)
k2d = Kh.permute(1, 0, 2).reshape(128, -1).contiguous()
vt2d = Vh.permute(0, 2, 1).reshape(-1, 128).contiguous()
kv = _lib.DictKV(Kh.data_ptr(), Vh.data_ptr(), None, None, None, None, sc.data_ptr())
k_hi, k_lo, _, kv.k_descale = f16_weight_planes(lib, k2d, 1, s)
v_hi, v_lo, _, kv.v_descale = f16_weight_planes(lib, vt2d, 1, s)
kv.K16_hi, kv.K16_lo, kv.Vt16_hi, kv.Vt16_lo = k_hi.data_ptr(), k_lo.data_ptr(), v_hi.data_ptr(), v_lo.data_ptr()
q_hi = q.half(); q_lo = (q - q_hi.float()).half()
q16 = _lib.Planes(q_hi.data_ptr(), q_lo.data_ptr(), 640)
o_hi = torch.empty(T, 640, dtype=torch.float16, device=dev); o_lo = torch.empty_like(o_hi)
o16 = _lib.Planes(o_hi.data_ptr(), o_lo.data_ptr(), 640)
def run():
    _lib.check(lib.dcae_op_dict_attention(None, 640, q16, kv, T, None, 640, o16, _lib.MATH["f16x3"], s))
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"dict_attention_f16_kernel T={T}: {us:.1f} us per launch, {327680.0 * T / us / 1e6:.1f} TFLOP/s algorithmic")

"""Per-operator device time of ONE g_a + g_s forward at B x 768x512 (library-side CUDA events around every op, each op timed
alone) with the operator's shape recorded on the host side: where the transform stacks' time goes, per level."""
import os, sys, csv, ctypes as C, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200 import _lib
from dcae_b200.transforms import LibKernels, TransformStack, init_transform_params
B = int(os.environ.get("TM_B", 16))
log = []
K = LibKernels("cuda:0")
for name in ("gemm", "layernorm", "window_attention", "dwconv_glu", "space_to_depth", "depth_to_space", "to_tokens", "to_nchw"):
    fn = getattr(K, name)
    def wrap(*a, _fn=fn, _name=name, **k):
        act = a[0]
        if _name == "gemm":
            pg = a[1]
            log.append((_name, act.T, pg.N, pg.K, pg.taps))
        elif _name in ("to_tokens",):
            log.append((_name, act.shape[0] * act.shape[2] * act.shape[3], act.shape[1], 0, 0))
        else:
            log.append((_name, act.T, act.ld, 0, 0))
        return _fn(*a, **k)
    setattr(K, name, wrap)
ga = TransformStack("g_a", init_transform_params(0, ("g_a",)), kernels=K)
gs = TransformStack("g_s", init_transform_params(0, ("g_s",)), kernels=K)
x = torch.rand(B, 3, 512, 768, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    y = ga(x); gs(y)
torch.cuda.synchronize()
lib = _lib.load()
ms = (C.c_double * 4)(); work = (C.c_double * 4)(); cnt = (C.c_int64 * 4)()
log.clear()
lib.dcae_profile_start()
y = ga(x); gs(y)
path = os.path.join(ROOT, "gpurun_out", "codec_layer_times.csv")
lib.dcae_profile_dump(path.encode(), ms, work, cnt)
rows = list(csv.DictReader(open(path)))
assert len(rows) == len(log), (len(rows), len(log))
print(f"g_a + g_s at B={B}: families ms gemm {ms[0]:.2f} attention {ms[1]:.2f} other {ms[3]:.2f}; {len(rows)} ops")
agg = collections.OrderedDict()
for r, l in zip(rows, log):
    key = l
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += float(r["ms"]); a[2] += float(r["work"])
print(f"{'op':18s} {'T':>9s} {'N/ld':>6s} {'K':>6s} taps   n   total ms   avg us   TFLOP/s(alg)   GB/s(min traffic)")
for (name, T, N, Kk, taps), (n, t, w) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = w / (t * 1e-3) / 1e12 if t > 0 and name == "gemm" else 0.0
    # minimal traffic of a GEMM: A planes (4 B/elem over pad64 K/taps) + output 4 B/elem
    gb = (T * ((Kk // max(taps, 1) + 63) // 64 * 64) * 4 + T * N * 4) * n / (t * 1e-3) / 1e9 if name == "gemm" else 0.0
    print(f"{name:18s} {T:9d} {N:6d} {Kk:6d} {taps:4d} {n:3d} {t:10.3f} {1e3 * t / n:8.1f} {tf:10.1f} {gb:12.0f}")

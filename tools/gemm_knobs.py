"""Sensitivity sweep of the tcgen05 GEMM (GPU box): pipeline depth, split cost, store cost, chunk length."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cabi as K
from dcae_b200 import _lib

B, h, w = 16, 32, 48
T = B * h * w
lib = _lib.load()
dev = torch.device("cuda:0")
KNOBS = ("DCAE_TC_BN", "DCAE_TC_STAGES", "DCAE_TC_NOSPLIT", "DCAE_TC_NOSTORE", "DCAE_TC_CHUNK", "DCAE_TC_EPI")


def bench(Kd, N, math, taps=1, reps=10, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["DCAE_TC_" + k.upper()] = str(v)
    C = Kd // taps
    buf = torch.randn(T, C, device=dev)
    wt = torch.randn(N, Kd, device=dev) * 0.02
    hi, lo = K.split_weight(wt)
    out = torch.empty(T, N, device=dev)
    from dcae_b200.weights import f16_weight_planes
    nbytes = lib.dcae_planes_bytes(T, C)
    planes = torch.empty(nbytes + 128, dtype=torch.uint8, device=dev)
    a = _lib.Operand(buf.data_ptr(), C, 0, C, 0, 0, taps, B, h, w, (planes.data_ptr() + 127) // 128 * 128, nbytes)
    h16, l16, K16, descale = f16_weight_planes(lib, wt, taps, _lib.current_stream(dev))
    W = _lib.Weight(wt.data_ptr(), hi.data_ptr(), lo.data_ptr(), N, Kd, h16.data_ptr(), l16.data_ptr(), K16, descale)
    e = _lib.Epilogue(); e.out, e.out_ld = out.data_ptr(), N
    s = _lib.current_stream(dev)
    for _ in range(3):
        _lib.check(lib.dcae_op_gemm(a, W, e, _lib.MATH[math], s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        lib.dcae_op_gemm(a, W, e, _lib.MATH[math], s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * T * N * Kd / (ms * 1e-3) / 1e12 * (3 if math in ("tf32x3", "f16x3") else 1)
    print(json.dumps({"K": Kd, "N": N, "taps": taps, "math": math, **env, "ms": round(ms, 4), "mma_TFLOPs": round(tf, 1)}), flush=True)


QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
for math in (sys.argv[2:] or ("tf32x3", "tf32")):
    for Kd, N, taps, bn in [(640, 640, 1, 160), (640, 640, 1, 128), (2560, 640, 1, 160), (640, 2560, 1, 256), (8640, 672, 9, 224),
                            (2016, 128, 9, 128), (1152, 64, 9, 64), (640, 320, 1, 160)]:
        bench(Kd, N, math, taps, bn=bn)
        if len(sys.argv) > 1 and sys.argv[1] == "ragged":
            if N == 640:
                os.environ["DCAE_F16_PAIR"] = "1"; bench(Kd, N, math, taps, bn=256); bench(Kd, N, math, taps, bn=160); bench(Kd, N, math, taps, bn=128)
                os.environ["DCAE_F16_PAIR"] = "0"; bench(Kd, N, math, taps, bn=256); bench(Kd, N, math, taps, bn=128)
            continue
        if len(sys.argv) > 1 and sys.argv[1] == "f16":
            bench(Kd, N, math, taps, bn=bn, nostore=1)
            for st in (2, 3):
                bench(Kd, N, math, taps, bn=bn, stages=st)
            for b2 in (64, 96, 128, 192):
                if N % b2 == 0 and b2 != bn:
                    bench(Kd, N, math, taps, bn=b2)
            continue
        if len(sys.argv) > 1 and sys.argv[1] == "epi":
            for v in (1, 2, 3):
                bench(Kd, N, math, taps, bn=bn, epi=v)
            bench(Kd, N, math, taps, bn=bn, nostore=1)
            continue
        if QUICK:
            continue
        for st in (1, 2, 3):
            bench(Kd, N, math, taps, bn=bn, stages=st)
        bench(Kd, N, math, taps, bn=bn, nostore=1)
        if math == "tf32x3":
            bench(Kd, N, math, taps, bn=bn, nosplit=1)
            bench(Kd, N, math, taps, bn=bn, nosplit=1, nostore=1)
            bench(Kd, N, math, taps, bn=bn, chunk=1000)
            bench(Kd, N, math, taps, bn=bn, chunk=4)

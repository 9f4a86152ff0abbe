"""Does polling NVML from a thread perturb a 20-step timed loop?  Variants: no polling / clock / power / reasons / all."""
import os, sys, time, threading, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml as nv
from dcae_b200.entropy_model import EntropySliceLoop
from dcae_b200.params import init_entropy_params
B, h, w = 16, 32, 48
eng = EntropySliceLoop(init_entropy_params(0, "lively"), math="f16x3", lanes=1)
g = torch.Generator().manual_seed(1)
x = [4 * torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda(), torch.randn(B, 320, h, w, generator=g).cuda()]
out = eng.forward(*x)
for _ in range(60):
    eng.forward(*x, out=out)
torch.cuda.synchronize()
nv.nvmlInit(); hd = nv.nvmlDeviceGetHandleByIndex(0)
calls = {"clock": lambda: nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM), "power": lambda: nv.nvmlDeviceGetPowerUsage(hd),
         "reasons": lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(hd)}
for n, f in calls.items():
    t0 = time.perf_counter(); [f() for _ in range(5)]; print(f"idle cost of {n}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/call")
def run(which, period):
    stop = threading.Event(); durs = []
    def poll():
        while not stop.is_set():
            t0 = time.perf_counter()
            for n in which: calls[n]()
            durs.append((time.perf_counter() - t0) * 1e3)
            stop.wait(period)
    th = threading.Thread(target=poll, daemon=True)
    if which: th.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.forward(*x, out=out)
    e1.record(); torch.cuda.synchronize()
    stop.set()
    if which: th.join()
    return e0.elapsed_time(e1) / 20, (max(durs) if durs else 0.0)
for which, period in [((), 0), (("clock",), 0.1), (("power",), 0.1), (("reasons",), 0.1), (("clock", "power", "reasons"), 0.1), (("clock", "power", "reasons"), 0.02), ((), 0)]:
    r = [run(which, period) for _ in range(6)]
    print(f"poll {'+'.join(which) or 'none':22s} every {period:4.2f}s: ms/step {[round(a, 2) for a, _ in r]}  max call {max(b for _, b in r):.1f} ms")

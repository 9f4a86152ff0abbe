"""Kernel 3 alone on a 1 GiB footprint (the roofline_gc microbenchmark of bench.py, without the rest of the bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dcae_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
peaks = bench.load_peaks()
for variant, lik in (("compress", "fast"), ("forward", "fast"), ("compress", "reference"), ("forward", "reference")):
    for _ in range(2):
        r = bench.gc_microbench(dev, lib, 1024, peaks, variant=variant, lik_math=lik)
    print(f"gc micro [{variant:8s} {lik:9s}]: {r['ms']:.4f} ms  {r['achieved']:.0f} GB/s  frac {r['frac']:.3f}")

"""Shared helpers for the test-suite."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["slice_loop_b2_7x9", "slice_loop_b1_8x12", "slice_loop_b1_16x16"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype in (np.uint8, np.int16):
            a = a.astype(np.int32)
        out[k] = torch.from_numpy(a) if a.ndim else a
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Per-tensor relative error: max|a-b| / max|b| (the norm all fp32 tolerances in tests/ use)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def mismatch_rate(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a != b).double().mean())

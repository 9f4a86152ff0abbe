"""Golden outputs of the reference's OWN transform modules (build container only).

    python tests/golden/make_golden_transforms.py        # writes tests/golden/transforms.npz

What runs: `DCAE().g_a / g_s / h_a / h_z_s1 / h_z_s2` of the UNMODIFIED `/root/reference/models/dcae.py`
(:541-582, loaded by oracle/reference_loader.py) in torch fp32 on the CPU, with the weights of
`dcae_b200.transforms.init_transform_params(0)` (which the GPU box regenerates bit for bit) on the seeded inputs of
`transform_golden_input()` below (also regenerated, so only the OUTPUTS are stored, as fp32).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# stack -> input shape: the smallest grids every Swin block of the stack accepts (larger than its window)
GOLDEN_SHAPES = {"g_a": (1, 3, 128, 128), "g_s": (1, 320, 8, 8), "h_a": (1, 320, 16, 16), "h_z_s1": (1, 192, 4, 4), "h_z_s2": (1, 192, 4, 4)}


def transform_golden_input(stack: str) -> torch.Tensor:
    g = torch.Generator().manual_seed(77 + sorted(GOLDEN_SHAPES).index(stack))
    shape = GOLDEN_SHAPES[stack]
    return torch.rand(shape, generator=g) if stack == "g_a" else torch.randn(shape, generator=g)


def main():
    from dcae_b200.transforms import STACKS, init_transform_params
    from oracle.reference_loader import load_reference_dcae_module
    ref = load_reference_dcae_module()
    torch.manual_seed(0)
    net = ref.DCAE().eval()
    missing, unexpected = torch.nn.Module.load_state_dict(net, init_transform_params(0), strict=False)
    assert not unexpected and not [m for m in missing if m.split(".")[0] in STACKS]
    out = {}
    torch.set_num_threads(1)          # one fixed reduction order
    with torch.no_grad():
        for stack in STACKS:
            out[stack] = getattr(net, stack)(transform_golden_input(stack)).numpy().astype(np.float32)
            print(stack, GOLDEN_SHAPES[stack], "->", out[stack].shape, float(np.abs(out[stack]).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "transforms.npz"), **out)


if __name__ == "__main__":
    main()

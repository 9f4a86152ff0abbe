"""Module-level drop-ins for the reference's own loop text.

`DCAE.forward / compress / decompress` (dcae.py:638-670, 713-753, 878-906) call four kinds of sub-modules inside
the slice loop: `self.dt_cross_attention[i](query, dt)`, `self.cc_mean_transforms[i](support)`,
`self.cc_scale_transforms[i](support)`, `self.lrp_transforms[i](lrp_support)` and `self.gaussian_conditional`.
`accelerate(net)` replaces exactly those attributes of a reference `DCAE` instance with modules that run on
libdcae_b200.so, so the reference's own Python loop runs unchanged on the CUDA kernels.  (The whole-loop object,
`EntropySliceLoop`, is faster: it keeps everything token-major between modules and fuses the first conv layer of
the three stacks; this file is for maintainers who want to keep `dcae.py` as it is.)
"""
from __future__ import annotations

import torch

from . import _lib
from .entropy_model import EntropySliceLoop
from .gaussian_conditional import GaussianConditional
from .params import M_LATENT, NUM_SLICES, SLICE_CH

HOT_PREFIXES = ("dt", "dt_cross_attention", "cc_mean_transforms", "cc_scale_transforms", "lrp_transforms")


class _LoopModule(torch.nn.Module):
    def __init__(self, loop: EntropySliceLoop, i: int):
        super().__init__()
        self._loop, self.i = [loop], i          # in a list: the engine is not a sub-module / parameter owner

    @property
    def loop(self) -> EntropySliceLoop:
        return self._loop[0]


class DictCrossAttention(_LoopModule):
    """`dt_cross_attention[i]`: forward(x [B, 640 + 64 i, h, w], dt) -> [B, 320, h, w]  (dcae.py:479-509).
    The dictionary is the one the engine was built with; `dt` is accepted for signature compatibility."""

    def forward(self, x: torch.Tensor, dt: torch.Tensor = None) -> torch.Tensor:
        return self.loop.module_dca(self.i, x)


class ConvStack(_LoopModule):
    """`cc_mean_transforms[i]` (which=0), `cc_scale_transforms[i]` (1), `lrp_transforms[i]` (2): forward(x) -> [B, 64, h, w]."""

    def __init__(self, loop: EntropySliceLoop, i: int, which: int):
        super().__init__(loop, i)
        self.which = which

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.loop.module_conv(self.i, self.which, x)


def accelerate(net: torch.nn.Module, device="cuda:0", math: str = "f16x3", state_dict=None) -> EntropySliceLoop:
    """Swap the hot-path sub-modules of a reference `DCAE` instance in place; returns the engine (also usable as the
    whole-loop object).  Call again after `load_state_dict` or an optimizer step: weights are packed at build time.
    `state_dict`: take the weights from here instead of `net.state_dict()` (same keys)."""
    sd = net.state_dict() if state_dict is None else state_dict
    hot = {k: v.detach() for k, v in sd.items() if k.split(".")[0] in HOT_PREFIXES}
    table = getattr(net.gaussian_conditional, "scale_table", None)
    loop = EntropySliceLoop(hot, device=device, math=math, lanes=1,
                            scale_table=table if table is not None and table.numel() else None)
    for i in range(NUM_SLICES):
        net.dt_cross_attention[i] = DictCrossAttention(loop, i)
        net.cc_mean_transforms[i] = ConvStack(loop, i, 0)
        net.cc_scale_transforms[i] = ConvStack(loop, i, 1)
        net.lrp_transforms[i] = ConvStack(loop, i, 2)
    fast = GaussianConditional(None).to(device)
    fast.load_state_dict(net.gaussian_conditional.state_dict(), strict=False)
    if fast.scale_table.numel() == 0:
        fast.update_scale_table(loop.scale_table)
    net.gaussian_conditional = fast
    return loop


__all__ = ["DictCrossAttention", "ConvStack", "accelerate", "HOT_PREFIXES", "M_LATENT", "SLICE_CH"]

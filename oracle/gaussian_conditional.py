"""ORACLE (test infrastructure, never imported by the product package `dcae_b200`).

CPU restatement of compressai's `GaussianConditional` / `LowerBound` as the reference uses
them on the entropy-model hot path.  compressai is a third-party, un-vendored and unpinned
dependency of the reference (`/root/reference/README.md:30`), absent from this image, so the
algorithm is restated from

  * the reference's in-tree copy of the likelihood math, `/root/reference/models/dcae.py:839-857`
    (`_likelihood`, `_standardized_cumulative`) -- PINNED: `tests/test_oracle_vs_reference.py`
    checks `likelihood()` below bit-for-bit against that code run from `/root/reference`;
  * the scale table and STE definitions at `dcae.py:28-30, 54-58`;
  * the call sites `dcae.py:657-659` (forward), `:738-740` (compress), `:891-896` (decompress);
  * compressai's published semantics for quantize / dequantize / build_indexes / update
    (entropy_models.py upstream; summarised in SURVEY.md §8c).  Those four are simple enough
    (round-half-even, int cast, a 63-step threshold count) that the call sites fully determine
    them; the CDF-table construction (`update`) is the only part that is "parity unpinned" in
    the strict sense: no golden tables exist in the reference and compressai cannot be run here.

All float math is torch-CPU fp32 in the reference's op order (no fused multiply-add, same
constant rounding), because likelihoods are compared at 1e-5 and symbols/indexes bit-exactly.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS) -> torch.Tensor:
    """dcae.py:54-55."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class LowerBoundFunction(torch.autograd.Function):
    """compressai `ops.bound_ops.LowerBoundFunction` (published source; the package is absent here): forward
    `torch.max(x, bound)`, backward passes the gradient where `x >= bound` OR the gradient pushes x up (`grad < 0`)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through_if = (x >= bound) | (grad_output < 0)
        return pass_through_if * grad_output, None


def lower_bound(x: torch.Tensor, bound: float) -> torch.Tensor:
    """compressai LowerBound.forward: torch.max(x, bound) (dcae.py:846 restates it for scales); under autograd with
    LowerBound's own backward rule (the training step, train.py:165-179, differentiates through both bounds)."""
    b = torch.tensor(bound, dtype=x.dtype, device=x.device)
    if torch.is_grad_enabled() and x.requires_grad:
        return LowerBoundFunction.apply(x, b)
    return torch.max(x, b)


def standardized_cumulative(inputs: torch.Tensor) -> torch.Tensor:
    """dcae.py:853-857: 0.5 * erfc(-(2**-0.5) * x), constant rounded to fp32 by torch."""
    half = float(0.5)
    const = float(-(2 ** -0.5))
    return half * torch.erfc(const * inputs)


def likelihood(inputs: torch.Tensor, scales: torch.Tensor, means: torch.Tensor | None,
               scale_bound: float = SCALES_MIN) -> torch.Tensor:
    """dcae.py:839-851 (compressai GaussianConditional._likelihood)."""
    half = float(0.5)
    values = inputs - means if means is not None else inputs
    scales = lower_bound(scales, scale_bound)
    values = torch.abs(values)
    upper = standardized_cumulative((half - values) / scales)
    lower = standardized_cumulative((-half - values) / scales)
    return upper - lower


def quantize(inputs: torch.Tensor, mode: str, means: torch.Tensor | None = None,
             noise: torch.Tensor | None = None) -> torch.Tensor:
    """compressai EntropyModel.quantize, as called at dcae.py:739 ("symbols") and inside forward.

    noise: `inputs + U(-1/2, 1/2)`; the oracle takes the noise tensor so tests are repeatable.
    dequantize: round(inputs - means) + means.   symbols: int32(round(inputs - means)).
    torch.round is round-half-to-even.
    """
    if mode not in ("noise", "dequantize", "symbols"):
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    if mode == "noise":
        if noise is None:
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
        return inputs + noise
    outputs = inputs.clone()
    if means is not None:
        outputs -= means
    outputs = torch.round(outputs)
    if mode == "dequantize":
        if means is not None:
            outputs += means
        return outputs
    return outputs.int()


def dequantize(inputs: torch.Tensor, means: torch.Tensor | None = None,
               dtype: torch.dtype = torch.float) -> torch.Tensor:
    """compressai EntropyModel.dequantize, call site dcae.py:896."""
    if means is not None:
        outputs = inputs.type_as(means)
        outputs = outputs + means
    else:
        outputs = inputs.type(dtype)
    return outputs


def build_indexes(scales: torch.Tensor, scale_table: torch.Tensor,
                  scale_bound: float = SCALES_MIN) -> torch.Tensor:
    """compressai GaussianConditional.build_indexes, call sites dcae.py:738, :891.

    idx = (len(table)-1) - sum_{t in table[:-1]} [max(scale, bound) <= t]
    """
    s = lower_bound(scales, scale_bound)
    indexes = s.new_full(s.size(), len(scale_table) - 1).int()
    for t in scale_table[:-1]:
        indexes -= (s <= t).int()
    return indexes


def ste_round(x: torch.Tensor) -> torch.Tensor:
    """dcae.py:57-58."""
    return torch.round(x) - x.detach() + x


# ----------------------------------------------------------------------------------------------
# CDF tables (compressai GaussianConditional.update / _pmf_to_cdf / _CXX.pmf_to_quantized_cdf).
# Parity unpinned (see module docstring); used only so that the `quantized_cdf/cdf_length/offset`
# attributes exist with the right shapes and so tests can check that symbols fall
# inside the coder's table range.
# ----------------------------------------------------------------------------------------------
def _standardized_quantile(q: float) -> float:
    from statistics import NormalDist
    return NormalDist().inv_cdf(q)


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """Restatement of compressai's C++ `pmf_to_quantized_cdf` (rans interface)."""
    cdf = [0] * (len(pmf) + 1)
    for i, p in enumerate(pmf):
        cdf[i + 1] = int(math.floor(float(p) * (1 << precision) + 0.5))   # std::round (half away from zero; p >= 0), not Python's half-to-even
    total = sum(cdf)
    cdf = [((1 << precision) * c) // total for c in cdf]          # integer renormalisation (floor)
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]                                       # std::partial_sum
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq = 1 << 32
            best_steal = -1
            for j in range(len(cdf) - 1):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq = freq
                    best_steal = j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    return cdf


def build_cdf_tables(scale_table: torch.Tensor, tail_mass: float = 1e-9, precision: int = 16):
    """compressai GaussianConditional.update()."""
    multiplier = -_standardized_quantile(tail_mass / 2)
    pmf_center = torch.ceil(scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length).item())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None])
    samples = samples.float()
    samples_scale = scale_table.unsqueeze(1).float()
    upper = standardized_cumulative((0.5 - samples) / samples_scale)
    lower = standardized_cumulative((-0.5 - samples) / samples_scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    quantized_cdf = torch.zeros(len(pmf_length), max_length + 2, dtype=torch.int32)
    for i in range(len(pmf_length)):
        n = int(pmf_length[i])
        prob = torch.cat((pmf[i, :n], tail[i]), dim=0).tolist()
        c = pmf_to_quantized_cdf(prob, precision)
        quantized_cdf[i, : len(c)] = torch.tensor(c, dtype=torch.int32)
    return quantized_cdf, -pmf_center, pmf_length + 2


class GaussianConditionalOracle(nn.Module):
    """Object with the compressai `GaussianConditional` surface the reference touches
    (dcae.py:614, 619, 657, 718-720, 738-739, 891, 896)."""

    def __init__(self, scale_table=None, scale_bound: float = SCALES_MIN, tail_mass: float = 1e-9,
                 likelihood_bound: float = 1e-9, entropy_coder_precision: int = 16):
        super().__init__()
        self.scale_bound = float(scale_bound)
        self.tail_mass = float(tail_mass)
        self.likelihood_bound = float(likelihood_bound)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.register_buffer("scale_table",
                             torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table is not None
                             else torch.Tensor())

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32).clone()
        q, off, ln = build_cdf_tables(self.scale_table, self.tail_mass, self.entropy_coder_precision)
        self._quantized_cdf, self._offset, self._cdf_length = q, off, ln
        return True

    def quantize(self, inputs, mode, means=None, noise=None):
        return quantize(inputs, mode, means, noise)

    def dequantize(self, inputs, means=None, dtype=torch.float):
        return dequantize(inputs, means, dtype)

    def build_indexes(self, scales):
        return build_indexes(scales, self.scale_table, self.scale_bound)

    def forward(self, inputs, scales, means=None, training=None, noise=None):
        if training is None:
            training = self.training
        outputs = quantize(inputs, "noise" if training else "dequantize", means, noise)
        lik = likelihood(outputs, scales, means, self.scale_bound)
        lik = lower_bound(lik, self.likelihood_bound)
        return outputs, lik

"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of compressai's `EntropyBottleneck` as the reference uses it for
the hyper-latent z (/root/reference/models/dcae.py:629-633 forward, :705-706 compress + decompress, :861 decompress,
`update()` through `CompressionModel.update`, `aux_loss` at train.py:177).

compressai is a third-party dependency that is absent from /root/reference (README.md:30, unpinned) and from this
image, so this follows its PUBLISHED source (entropy_models.py, 1.2.x: factorised prior of Balle et al. 2018 with
filters (3, 3, 3, 3), init_scale 10, tail_mass 1e-9; likelihood = sigmoid(upper) - sigmoid(lower)).  **Parity
unpinned**: the reference holds no golden vector for z; the checks are the density's own properties (the pmf sums to
1 - tail mass, the CDF tables are monotone and cover the quantiles) and round trips through the range coder.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

FILTERS = (3, 3, 3, 3)
INIT_SCALE = 10.0
TAIL_MASS = 1e-9


def init_params(channels: int, seed: int = 0, trained_like: bool = False) -> Dict[str, torch.Tensor]:
    """compressai's initialisation (matrices = log(expm1(1 / scale / fan)), biases U(-1/2, 1/2), factors 0, quantiles
    [-10, 0, 10]); trained_like additionally perturbs everything so that channels differ and the tanh terms are live."""
    g = torch.Generator().manual_seed(seed)
    filters = (1,) + FILTERS + (1,)
    scale = INIT_SCALE ** (1 / (len(FILTERS) + 1))
    p = {}
    for i in range(len(FILTERS) + 1):
        init = math.log(math.expm1(1 / scale / filters[i + 1]))
        p[f"_matrix{i}"] = torch.full((channels, filters[i + 1], filters[i]), init)
        p[f"_bias{i}"] = torch.rand(channels, filters[i + 1], 1, generator=g) - 0.5
        if i < len(FILTERS):
            p[f"_factor{i}"] = torch.zeros(channels, filters[i + 1], 1)
    p["quantiles"] = torch.tensor([-INIT_SCALE, 0.0, INIT_SCALE]).repeat(channels, 1, 1)
    if trained_like:
        for k in list(p):
            if k.startswith("_matrix"):
                p[k] = p[k] + 0.5 * torch.randn(p[k].shape, generator=g)
            elif k.startswith("_factor"):
                p[k] = 0.8 * torch.randn(p[k].shape, generator=g)
        med = 3.0 * torch.randn(channels, generator=g)
        width = 2.0 + 8.0 * torch.rand(channels, generator=g)
        p["quantiles"] = torch.stack([med - width, med, med + width], dim=1).unsqueeze(1)
    return p


def logits_cumulative(p: Dict[str, torch.Tensor], inputs: torch.Tensor) -> torch.Tensor:
    """inputs [C, 1, n] -> [C, 1, n]."""
    logits = inputs
    for i in range(len(FILTERS) + 1):
        logits = torch.matmul(F.softplus(p[f"_matrix{i}"]), logits)
        logits = logits + p[f"_bias{i}"]
        if i < len(FILTERS):
            logits = logits + torch.tanh(p[f"_factor{i}"]) * torch.tanh(logits)
    return logits


def likelihood(p, inputs):
    lower = logits_cumulative(p, inputs - 0.5)
    upper = logits_cumulative(p, inputs + 0.5)
    return torch.sigmoid(upper) - torch.sigmoid(lower), lower, upper


def medians(p):
    return p["quantiles"][:, :, 1:2]


def forward(p, x: torch.Tensor, noise: torch.Tensor | None = None, lik_bound: float = 1e-9):
    """EntropyBottleneck.forward: x [B, C, h, w] -> (outputs, likelihood); eval mode unless a noise tensor is given."""
    perm = x.transpose(0, 1).contiguous()
    shape = perm.shape
    values = perm.reshape(shape[0], 1, -1)
    med = medians(p)
    if noise is None:
        outputs = torch.round(values - med) + med
    else:
        outputs = values + noise.transpose(0, 1).reshape(shape[0], 1, -1)
    lik, _, _ = likelihood(p, outputs)
    lik = torch.max(lik, torch.tensor(lik_bound))
    back = lambda t: t.reshape(shape).transpose(0, 1).contiguous()      # noqa: E731
    return back(outputs), back(lik)


def symbols(p, x):
    return torch.round(x - medians(p).reshape(1, -1, 1, 1)).int()


def build_tables(p, pmf_to_quantized_cdf, tail_mass: float = TAIL_MASS, precision: int = 16):
    """EntropyBottleneck.update(): -> (_quantized_cdf [C, L + 2], _offset [C], _cdf_length [C])."""
    q = p["quantiles"]
    med = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
    pmf_start = med - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    pmf, lower, upper = likelihood(p, samples)
    pmf = pmf[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = torch.zeros(len(pmf_length), max_length + 2, dtype=torch.int32)
    for i in range(len(pmf_length)):
        n = int(pmf_length[i])
        row = pmf_to_quantized_cdf(torch.cat((pmf[i, :n], tail[i])).tolist(), precision)
        cdf[i, : len(row)] = torch.tensor(row, dtype=torch.int32)
    return cdf, -minima, pmf_length + 2


def aux_loss(p):
    """EntropyBottleneck.loss(): sum |logits_cumulative(quantiles) - target|."""
    target = math.log(2 / TAIL_MASS - 1)
    logits = logits_cumulative(p, p["quantiles"])
    return torch.abs(logits - torch.tensor([-target, 0.0, target])).sum()

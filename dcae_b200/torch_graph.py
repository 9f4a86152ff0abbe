"""The hot path's operator graph in differentiable PyTorch ops -- used ONLY to DIFFERENTIATE (training config #4).

The library's kernels are forward kernels.  `dcae_b200.training` runs the forward pass of the slice loop on them and,
in backward, recomputes the same graph here under autograd (gradient checkpointing with a fast forward) -- the plan of
SURVEY section 7 ("forward kernels + PyTorch-op backward via autograd.Function recompute first").  Nothing here is
called on an inference path.  Each function names the reference lines whose forward it mirrors; parameters arrive as a
dict with the reference's state-dict keys (prefix stripped), so gradients map one to one onto reference parameters.
Runs on the tensors' device (CUDA in practice); kernel 3's forward AND analytic backward stay on the library
(`GaussianConditionalFunction`, dcae_gc_fused / dcae_gc_backward).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .params import DICT_DIM, HEAD_DIM, HEAD_NUM, NUM_SLICES

P = Dict[str, torch.Tensor]


def sub(params: P, prefix: str) -> P:
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def _nchw(x):          # [B, h, w, C] -> [B, C, h, w]
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def _conv_with_dw(x, p: P, pre: str):
    """GELU -> 1x1 -> GELU -> depthwise 3x3 -> GELU -> 1x1 (dcae.py:399-414 behind the nn.GELU of :421-423); NCHW."""
    x = F.conv2d(F.gelu(x), p[pre + "in_trans.weight"], p[pre + "in_trans.bias"])
    x = F.conv2d(F.gelu(x), p[pre + "dw_conv.weight"], p[pre + "dw_conv.bias"], padding=1, groups=x.shape[1])
    return F.conv2d(F.gelu(x), p[pre + "out_trans.weight"], p[pre + "out_trans.bias"])


def multi_scale_aggregation(x, p: P):
    """dcae.py:435-448 (+ DenseBlock :416-433, SpatialAttentionModule :386-397); x: [B, h, w, 640] -> same."""
    s = F.conv2d(_nchw(x), p["msa.s.weight"], p["msa.s.bias"])
    outs = [s]
    for j in range(3):
        outs.append(_conv_with_dw(outs[-1], p, f"msa.dense.conv_layers.{j}.1."))
    s_out = F.conv2d(torch.cat(outs, dim=1), p["msa.dense.proj.weight"], p["msa.dense.proj.bias"])
    stats = torch.cat([s_out.mean(dim=1, keepdim=True), s_out.max(dim=1, keepdim=True)[0]], dim=1)
    gate = torch.sigmoid(F.conv2d(stats, p["msa.spatial_atte.conv1.weight"], padding=3))
    return _nhwc(s_out * gate)


def conv_glu(x, p: P):
    """ConvolutionalGLU, dcae.py:312-328; x: [B, h, w, 640]."""
    a, v = F.linear(x, p["mlp.fc1.weight"], p["mlp.fc1.bias"]).chunk(2, dim=-1)
    a = _nhwc(F.conv2d(_nchw(a), p["mlp.dwconv.dwconv.weight"], p["mlp.dwconv.dwconv.bias"], padding=1, groups=a.shape[-1]))
    return F.linear(F.gelu(a) * v, p["mlp.fc2.weight"], p["mlp.fc2.bias"])


def dictionary_cross_attention(x, dt, p: P):
    """MutiScaleDictionaryCrossAttentionGLU.forward, dcae.py:479-509.  x: [B, C, h, w]; dt: [128, 640] (the reference
    repeats it over the batch, :625 -- K and V are batch invariant)."""
    B, _, h, w = x.shape
    x = F.linear(_nhwc(x), p["x_trans.weight"], p["x_trans.bias"])
    x = multi_scale_aggregation(F.layer_norm(x, (DICT_DIM,), p["ln_scale.weight"], p["ln_scale.bias"]), p) + x * p["res_scale_1.scale"]
    shortcut = x
    q = F.linear(F.layer_norm(x, (DICT_DIM,), p["lnx.weight"], p["lnx.bias"]), p["q_trans.weight"], p["q_trans.bias"])
    q = q.reshape(B, h * w, HEAD_NUM, HEAD_DIM).transpose(1, 2)                       # b e n c
    d = F.layer_norm(dt, (DICT_DIM,), p["dict_ln.weight"], p["dict_ln.bias"])
    k = F.linear(d, p["k.weight"], p["k.bias"]).reshape(-1, HEAD_NUM, HEAD_DIM).transpose(0, 1)   # e d c
    v = d.reshape(-1, HEAD_NUM, HEAD_DIM).transpose(0, 1)
    sim = torch.einsum("benc,edc->bend", q, k) * p["scale"].reshape(1, HEAD_NUM, 1, 1)
    out = torch.einsum("bend,edc->benc", torch.softmax(sim, dim=-1), v)
    out = out.transpose(1, 2).reshape(B, h, w, DICT_DIM)
    out = F.linear(out, p["linear.weight"], p["linear.bias"]) + shortcut * p["res_scale_2.scale"]
    out = conv_glu(F.layer_norm(out, (DICT_DIM,), p["ln_mlp.weight"], p["ln_mlp.bias"]), p) + out * p["res_scale_3.scale"]
    return _nchw(F.linear(out, p["output_trans.0.weight"], p["output_trans.0.bias"]))


def conv_stack(x, p: P):
    """cc_mean / cc_scale / lrp transforms, dcae.py:584-611: conv3x3 GELU conv3x3 GELU conv3x3."""
    x = F.gelu(F.conv2d(x, p["0.weight"], p["0.bias"], padding=1))
    x = F.gelu(F.conv2d(x, p["2.weight"], p["2.bias"], padding=1))
    return F.conv2d(x, p["4.weight"], p["4.bias"], padding=1)


def ste_round(x):
    """dcae.py:57-58."""
    return torch.round(x) - x.detach() + x


def slice_loop(params: P, y, latent_scales, latent_means, gaussian_conditional, noise: Optional[torch.Tensor] = None):
    """DCAE.forward slice loop, dcae.py:638-670.  `gaussian_conditional(y_slice, scale, mu, noise_slice) -> likelihood`
    (training mode when a noise tensor is given).  -> y_hat, means, scales, likelihoods, each [B, 320, h, w]."""
    y_hat_slices: List[torch.Tensor] = []
    mus, scales, liks = [], [], []
    nz = noise.chunk(NUM_SLICES, 1) if noise is not None else [None] * NUM_SLICES
    for i, y_slice in enumerate(y.chunk(NUM_SLICES, 1)):
        query = torch.cat([latent_scales, latent_means] + y_hat_slices, dim=1)
        dict_info = dictionary_cross_attention(query, params["dt"], sub(params, f"dt_cross_attention.{i}."))
        support = torch.cat([query, dict_info], dim=1)
        mu = conv_stack(support, sub(params, f"cc_mean_transforms.{i}."))
        scale = conv_stack(support, sub(params, f"cc_scale_transforms.{i}."))
        liks.append(gaussian_conditional(y_slice, scale, mu, nz[i]))
        y_hat_slice = ste_round(y_slice - mu) + mu
        lrp = conv_stack(torch.cat([support, y_hat_slice], dim=1), sub(params, f"lrp_transforms.{i}."))
        y_hat_slices.append(y_hat_slice + 0.5 * torch.tanh(lrp))
        mus.append(mu)
        scales.append(scale)
    cat = lambda ts: torch.cat(ts, dim=1)      # noqa: E731
    return cat(y_hat_slices), cat(mus), cat(scales), cat(liks)

"""ORACLE (test infrastructure, never imported by the product package `dcae_b200`).

Functional torch-CPU fp32 restatement of the reference's entropy-model hot path: the channel-slice
loop of `DCAE.forward / compress / decompress` and the modules it calls.  It operates on a flat
state dict with the reference's own key names (see `dcae_b200/params.py`), so the same weights
can be loaded into the unmodified reference modules.

PINNED: `tests/test_oracle_vs_reference.py` runs this file against the reference's own classes
loaded from `/root/reference/models/dcae.py` (possible in the build container) and
`tests/golden/*.npz` holds outputs of the reference itself (generator: `tests/golden/make_golden.py`)
that this oracle must reproduce on the GPU box, where `/root/reference` does not exist.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from . import gaussian_conditional as gc

NUM_SLICES = 5
HEAD_NUM = 20
LN_EPS = 1e-5

Params = Dict[str, torch.Tensor]


def _sub(params: Params, prefix: str) -> Params:
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def _ln(x, p, name):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], LN_EPS)


def _lin(x, p, name):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _nhwc(x):   # 'b c h w -> b h w c'
    return x.permute(0, 2, 3, 1)


def _nchw(x):   # 'b h w c -> b c h w'
    return x.permute(0, 3, 1, 2)


def conv_with_dw(x, p, prefix):
    """models/dcae.py:399-414 ConvWithDW: 1x1 -> GELU -> depthwise 3x3 -> GELU -> 1x1 (NCHW)."""
    c = x.shape[1]
    x = F.conv2d(x, p[prefix + "in_trans.weight"], p[prefix + "in_trans.bias"])
    x = F.gelu(x)
    x = F.conv2d(x, p[prefix + "dw_conv.weight"], p[prefix + "dw_conv.bias"], padding=1, groups=c)
    x = F.gelu(x)
    return F.conv2d(x, p[prefix + "out_trans.weight"], p[prefix + "out_trans.bias"])


def dense_block(x, p, prefix):
    """models/dcae.py:416-433 DenseBlock: 3 x (GELU -> ConvWithDW) chained, concat of all 4, 1x1 proj."""
    outs = [x]
    for j in range(3):
        outs.append(conv_with_dw(F.gelu(outs[-1]), p, f"{prefix}conv_layers.{j}.1."))
    return F.conv2d(torch.cat(outs, dim=1), p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def spatial_attention(x, p, prefix):
    """models/dcae.py:386-397: sigmoid(conv7x7([mean_c(x), max_c(x)])), no bias."""
    avg = torch.mean(x, dim=1, keepdim=True)
    mx, _ = torch.max(x, dim=1, keepdim=True)
    return torch.sigmoid(F.conv2d(torch.cat([avg, mx], dim=1), p[prefix + "conv1.weight"], None, padding=3))


def multi_scale_aggregation(x_nhwc, p, prefix):
    """models/dcae.py:435-448."""
    x = _nchw(x_nhwc)
    s = F.conv2d(x, p[prefix + "s.weight"], p[prefix + "s.bias"])
    s_out = dense_block(s, p, prefix + "dense.")
    return _nhwc(s_out * spatial_attention(s_out, p, prefix + "spatial_atte."))


def convolutional_glu(x_nhwc, p, prefix):
    """models/dcae.py:312-328 ConvolutionalGLU with DWConv :300-310."""
    x, v = _lin(x_nhwc, p, prefix + "fc1").chunk(2, dim=-1)
    c = x.shape[-1]
    x = _nhwc(F.conv2d(_nchw(x), p[prefix + "dwconv.dwconv.weight"], p[prefix + "dwconv.dwconv.bias"],
                       padding=1, groups=c))
    return _lin(F.gelu(x) * v, p, prefix + "fc2")


def dictionary_cross_attention(x: torch.Tensor, dt: torch.Tensor, p: Params,
                               taps: Optional[dict] = None) -> torch.Tensor:
    """models/dcae.py:479-509 MutiScaleDictionaryCrossAttentionGLU.forward.

    x [B, Cq, H, W] NCHW, dt [128, 640] (the reference repeats it over the batch, :625; K and V are
    batch invariant so the oracle keeps one copy), p = state dict of one `dt_cross_attention.i`.
    Returns [B, 320, H, W].  `taps`, if given, receives intermediate token-major tensors.
    """
    B, C, H, W = x.shape
    e = HEAD_NUM
    x = _lin(_nhwc(x), p, "x_trans")                                        # :481-482
    x0 = x
    x = multi_scale_aggregation(_ln(x, p, "ln_scale"), p, "msa.") + x * p["res_scale_1.scale"]   # :484
    shortcut = x                                                            # :486
    q = _lin(_ln(x, p, "lnx"), p, "q_trans")                                # :487-488
    q = q.reshape(B, H * W, e, -1).permute(0, 2, 1, 3)                      # 'b (hw) (e c) -> b e (hw) c'
    d = _ln(dt, p, "dict_ln")                                               # :492
    k = _lin(d, p, "k")                                                     # :493
    k = k.reshape(-1, e, k.shape[-1] // e).permute(1, 0, 2)                 # 'n (e c) -> e n c'
    v = d.reshape(-1, e, d.shape[-1] // e).permute(1, 0, 2)                 # V = LN(dt), un-projected :495
    sim = torch.einsum("benc,edc->bend", q, k) * p["scale"]                 # :497-498 per-head learned scale
    probs = torch.softmax(sim, dim=-1)                                      # :499
    o = torch.einsum("bend,edc->benc", probs, v)                            # :500
    o = o.permute(0, 2, 1, 3).reshape(B, H, W, -1)                          # :501
    x2 = _lin(o, p, "linear") + shortcut * p["res_scale_2.scale"]           # :503
    x3 = convolutional_glu(_ln(x2, p, "ln_mlp"), p, "mlp.") + x2 * p["res_scale_3.scale"]   # :505
    out = _lin(x3, p, "output_trans.0")                                     # :507
    if taps is not None:
        taps.update(x0=x0, x1=shortcut, q=q, attn=o, x2=x2, x3=x3)
    return _nchw(out)                                                       # :508


def conv_stack(x: torch.Tensor, p: Params) -> torch.Tensor:
    """models/dcae.py:584-611: conv3x3 -> GELU -> conv3x3 -> GELU -> conv3x3, stride 1, pad 1."""
    x = F.gelu(F.conv2d(x, p["0.weight"], p["0.bias"], padding=1))
    x = F.gelu(F.conv2d(x, p["2.weight"], p["2.bias"], padding=1))
    return F.conv2d(x, p["4.weight"], p["4.bias"], padding=1)


class SliceLoopOracle:
    """The slice loop of DCAE.forward (dcae.py:638-670), compress (:713-753), decompress (:878-906)."""

    def __init__(self, params: Params, scale_table: Optional[torch.Tensor] = None):
        self.params = params
        self.dt = params["dt"]
        self.dca = [_sub(params, f"dt_cross_attention.{i}.") for i in range(NUM_SLICES)]
        self.cc_mean = [_sub(params, f"cc_mean_transforms.{i}.") for i in range(NUM_SLICES)]
        self.cc_scale = [_sub(params, f"cc_scale_transforms.{i}.") for i in range(NUM_SLICES)]
        self.lrp = [_sub(params, f"lrp_transforms.{i}.") for i in range(NUM_SLICES)]
        self.scale_table = gc.get_scale_table() if scale_table is None else scale_table

    # -- one slice: the entropy parameters (dcae.py:644-655 / :728-736 / :879-888) ------------------
    def slice_params(self, i, latent_scales, latent_means, y_hat_slices: List[torch.Tensor]):
        query = torch.cat([latent_scales, latent_means] + y_hat_slices, dim=1)
        dict_info = dictionary_cross_attention(query, self.dt, self.dca[i])
        support = torch.cat([query, dict_info], dim=1)
        mu = conv_stack(support, self.cc_mean[i])
        scale = conv_stack(support, self.cc_scale[i])
        return support, mu, scale

    def slice_lrp(self, i, support, y_hat_slice):
        """dcae.py:661-664."""
        lrp = conv_stack(torch.cat([support, y_hat_slice], dim=1), self.lrp[i])
        return y_hat_slice + 0.5 * torch.tanh(lrp)

    @torch.no_grad()
    def forward(self, y, latent_scales, latent_means, noise: Optional[torch.Tensor] = None):
        """dcae.py:638-670, eval mode (noise=None) or training mode with an explicit noise tensor.
        Returns y_hat, means, scales, y_likelihoods, each [B, 320, h, w]."""
        y_hat_slices, mus, scales, liks = [], [], [], []
        for i, y_slice in enumerate(y.chunk(NUM_SLICES, 1)):
            support, mu, scale = self.slice_params(i, latent_scales, latent_means, y_hat_slices)
            mus.append(mu)
            scales.append(scale)
            if noise is None:
                outputs = gc.quantize(y_slice, "dequantize", mu)
            else:
                outputs = gc.quantize(y_slice, "noise", mu, noise.chunk(NUM_SLICES, 1)[i])
            lik = gc.lower_bound(gc.likelihood(outputs, scale, mu), 1e-9)          # :657
            liks.append(lik)
            y_hat_slice = gc.ste_round(y_slice - mu) + mu                          # :659
            y_hat_slices.append(self.slice_lrp(i, support, y_hat_slice))
        cat = lambda ts: torch.cat(ts, dim=1)
        return cat(y_hat_slices), cat(mus), cat(scales), cat(liks)

    @torch.no_grad()
    def compress(self, y, latent_scales, latent_means):
        """dcae.py:727-753. Returns (symbols, indexes) int32 [5, B, 64, h, w] in the reference's
        slice-major coder order (`symbols_list.extend(...)` :742-743), y_hat, means, scales."""
        y_hat_slices, syms, idxs, mus, scales = [], [], [], [], []
        for i, y_slice in enumerate(y.chunk(NUM_SLICES, 1)):
            support, mu, scale = self.slice_params(i, latent_scales, latent_means, y_hat_slices)
            index = gc.build_indexes(scale, self.scale_table)                      # :738
            y_q = gc.quantize(y_slice, "symbols", mu)                              # :739
            y_hat_slice = y_q + mu                                                 # :740
            syms.append(y_q)
            idxs.append(index)
            mus.append(mu)
            scales.append(scale)
            y_hat_slices.append(self.slice_lrp(i, support, y_hat_slice))
        return (torch.stack(syms), torch.stack(idxs), torch.cat(y_hat_slices, 1),
                torch.cat(mus, 1), torch.cat(scales, 1))

    @torch.no_grad()
    def decompress(self, latent_scales, latent_means,
                   decode_slice: Callable[[int, torch.Tensor], torch.Tensor]):
        """dcae.py:878-906. `decode_slice(i, indexes[B,64,h,w]) -> symbols` plays the rANS decoder
        (:893).  Returns y_hat and the per-slice indexes that were handed to the decoder."""
        y_hat_slices, idxs = [], []
        for i in range(NUM_SLICES):
            support, mu, scale = self.slice_params(i, latent_scales, latent_means, y_hat_slices)
            index = gc.build_indexes(scale, self.scale_table)                      # :891
            rv = decode_slice(i, index)
            y_hat_slice = gc.dequantize(rv, mu)                                    # :896
            idxs.append(index)
            y_hat_slices.append(self.slice_lrp(i, support, y_hat_slice))
        return torch.cat(y_hat_slices, 1), torch.stack(idxs)


def bits_per_pixel(likelihoods: torch.Tensor, num_pixels: int) -> torch.Tensor:
    """train.py:82-85: sum(log(lik)) / (-ln 2 * num_pixels)."""
    import math
    return torch.log(likelihoods).sum() / (-math.log(2) * num_pixels)

"""Drop-in for compressai's `GaussianConditional` as the reference uses it:
`net.gaussian_conditional = dcae_b200.GaussianConditional(None)` keeps `DCAE.forward / compress /
decompress` (dcae.py:614, 619, 657, 718-720, 738-739, 891, 896) working, with every tensor op running in
ONE launch of kernel 3 (dcae_gc_fused) instead of ~140 elementwise launches.

State-dict keys match the reference's expectations (dcae.py:680-685): `_quantized_cdf`, `_offset`,
`_cdf_length`, `scale_table`.  The CDF tables themselves are coder-side data; they
are built on the host by `update_scale_table` following compressai's published `update()`.
"""
from __future__ import annotations

from statistics import NormalDist
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


class _Bound(nn.Module):
    """Key-compatible stand-in for compressai's `LowerBound` sub-module (one buffer, `bound`); the clamp itself
    happens inside kernel 3."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class GaussianConditional(nn.Module):
    def __init__(self, scale_table=None, scale_bound: float = 0.11, tail_mass: float = 1e-9,
                 likelihood_bound: float = 1e-9, entropy_coder_precision: int = 16, likelihood_math: str = "fast"):
        super().__init__()
        if likelihood_math not in _lib.GC_LIK:
            raise ValueError(f"likelihood_math must be one of {list(_lib.GC_LIK)}")
        self.likelihood_math = likelihood_math      # "reference": the reference's op order, bit-identical to torch-CUDA (test mode)
        self.tail_mass = float(tail_mass)
        self.entropy_coder_precision = int(entropy_coder_precision)
        # compressai's published module also owns `scale_bound`, `lower_bound_scale.bound` and
        # `likelihood_lower_bound.bound` (its source is not in /root/reference; the reference itself only touches the four
        # table buffers, dcae.py:680-685): registered so that a checkpoint saved from the real class loads key for key
        self.likelihood_lower_bound = _Bound(likelihood_bound)
        self.lower_bound_scale = _Bound(scale_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.register_buffer("scale_table", torch.Tensor(tuple(float(s) for s in scale_table))
                             if scale_table is not None else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self._bounds = (float(scale_bound), float(likelihood_bound))     # host copies: kernel arguments, no device read per call

    @property
    def likelihood_bound(self) -> float:
        return self._bounds[1]

    _TABLE_BUFFERS = ("_quantized_cdf", "_offset", "_cdf_length", "scale_table")

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """The coder tables change size with the checkpoint (`update()` / baked tables): resize the four buffers to
        the incoming shapes first, which is what the reference does with `update_registered_buffers` before it loads
        (dcae.py:679-687).  Keys of compressai's own state dict that this class does not hold are the caller's
        business (`strict=False`)."""
        for name in self._TABLE_BUFFERS:
            src = state_dict.get(prefix + name)
            if src is not None:
                cur = getattr(self, name)
                if tuple(cur.shape) != tuple(src.shape):
                    setattr(self, name, torch.empty(src.shape, dtype=cur.dtype, device=cur.device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        sb, lb = state_dict.get(prefix + "scale_bound"), state_dict.get(prefix + "likelihood_lower_bound.bound")
        self._bounds = (float(sb) if sb is not None and sb.numel() == 1 else self._bounds[0],
                        float(lb) if lb is not None and lb.numel() == 1 else self._bounds[1])

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    # ---- coder tables (host, once per checkpoint) ------------------------------------------------
    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32).clone().to(device)
        self.update()
        return True

    def update(self):
        """compressai GaussianConditional.update(): pmf of N(0, s) on integers, tail mass, 16-bit CDF."""
        table = self.scale_table.detach().cpu().float()
        multiplier = -NormalDist().inv_cdf(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(pmf_length.max())
        samples = (torch.arange(max_length).int() - pmf_center[:, None]).abs().float()
        s = table.unsqueeze(1)
        c = float(-(2 ** -0.5))
        upper = 0.5 * torch.erfc(c * ((0.5 - samples) / s))
        lower = 0.5 * torch.erfc(c * ((-0.5 - samples) / s))
        pmf, tail = upper - lower, 2 * lower[:, :1]
        q = torch.zeros(len(pmf_length), max_length + 2, dtype=torch.int32)
        for i in range(len(pmf_length)):
            n = int(pmf_length[i])
            cdf = _pmf_to_quantized_cdf(torch.cat((pmf[i, :n], tail[i])).tolist(), self.entropy_coder_precision)
            q[i, : len(cdf)] = torch.tensor(cdf, dtype=torch.int32)
        dev = self.scale_table.device
        self._quantized_cdf, self._offset, self._cdf_length = q.to(dev), (-pmf_center).to(dev), (pmf_length + 2).to(dev)

    # ---- kernel 3 --------------------------------------------------------------------------------
    def _run(self, mode, inputs, means, scales=None, noise=None, sym_in=None, want=("y_hat",)):
        ref = means if means is not None else inputs
        if not ref.is_cuda:
            raise _lib.DcaeError("dcae_b200.GaussianConditional runs on CUDA tensors only (no CPU fallback)")
        lib = _lib.load()
        shape = ref.shape
        if ref.numel() == 0:    # empty input: nothing to launch
            dts = {"y_hat": torch.float32, "lik": torch.float32, "sym": torch.int32, "idx": torch.int32}
            return {k: torch.empty(shape, dtype=dts[k], device=ref.device) for k in want}
        if means is None:
            means = torch.zeros_like(ref, dtype=torch.float32)

        def rows(t, dtype=torch.float32):
            """View t as (rows, inner) with contiguous inner; batch-strided NCHW slices stay zero-copy."""
            if t is None:
                return None, 0
            if t.dtype != dtype:
                t = t.to(dtype)
            if t.dim() >= 2 and t[0].is_contiguous() and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
                return t, t.stride(0) if t.shape[0] > 1 else t[0].numel()
            t = t.contiguous()
            return t, t[0].numel() if t.dim() >= 1 and t.shape[0] > 0 else 0

        n_rows = shape[0] if len(shape) >= 2 else 1
        inner = ref.numel() // max(n_rows, 1)
        pad = (-inner) % 4
        if pad or len(shape) < 2:   # odd sizes: fall back to one flat, padded row per tensor
            return self._run_flat(mode, inputs, means, scales, noise, sym_in, want)
        a = _lib.GcArgs()
        keep = []
        for name, t, dt in (("y", inputs, torch.float32), ("mu", means, torch.float32), ("scale", scales, torch.float32),
                            ("noise", noise, torch.float32), ("sym_in", sym_in, torch.int32)):
            tt, ld = rows(t, dt)
            keep.append(tt)
            setattr(a, name, _lib.ptr(tt))
            setattr(a, name + "_ld", ld)
        table = self.scale_table
        if "idx" in want:
            if table.numel() == 0:
                raise _lib.DcaeError("build_indexes needs a scale table: call update_scale_table() / net.update() first")
            table = table.to(ref.device, torch.float32).contiguous()
            keep.append(table)
            a.scale_table, a.n_table = table.data_ptr(), table.numel()
        a.scale_bound, a.lik_bound, a.mode = self._bounds[0], self._bounds[1], mode
        a.rows, a.inner = n_rows, inner
        a.lik_math = _lib.GC_LIK[self.likelihood_math]
        outs = {}
        for name, dt in (("y_hat", torch.float32), ("lik", torch.float32), ("sym", torch.int32), ("idx", torch.int32)):
            if name in want:
                o = torch.empty(shape, dtype=dt, device=ref.device)
                outs[name] = o
                setattr(a, name, o.data_ptr())
                setattr(a, name + "_ld", inner)
        if ref.numel():
            with torch.cuda.device(ref.device):
                _lib.check(lib.dcae_gc_fused(a, _lib.current_stream(ref.device)), "dcae_gc_fused")
        return outs

    def _run_flat(self, mode, inputs, means, scales, noise, sym_in, want):
        shape = (means if means is not None else inputs).shape
        n = means.numel()
        padn = (n + 3) // 4 * 4

        def flat(t, dt=torch.float32, fill=0):
            if t is None:
                return None
            o = torch.full((1, padn), fill, dtype=dt, device=means.device)
            o[0, :n] = t.reshape(-1).to(dt)
            return o

        outs = self._run(mode, flat(inputs), flat(means), flat(scales, fill=1), flat(noise), flat(sym_in, torch.int32), want)
        return {k: v[0, :n].reshape(shape) for k, v in outs.items()}

    def quantize(self, inputs, mode, means=None, noise: Optional[torch.Tensor] = None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            if noise is None:
                noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
            return inputs + noise
        if mode == "dequantize":
            return self._run(_lib.GC_EVAL, inputs, means, want=("y_hat",))["y_hat"]
        return self._run(_lib.GC_EVAL, inputs, means, want=("sym",))["sym"]

    def dequantize(self, inputs, means=None, dtype=torch.float):
        if means is None:
            return inputs.type(dtype)
        sym = inputs.to(means.device)
        if sym.dtype != torch.int32:
            sym = sym.round().to(torch.int32)   # decoder output arrives as a float tensor (dcae.py:894)
        return self._run(_lib.GC_DECODE, None, means, sym_in=sym, want=("y_hat",))["y_hat"]

    def build_indexes(self, scales):
        zeros = torch.zeros_like(scales, dtype=torch.float32)
        return self._run(_lib.GC_EVAL, None, zeros, scales=scales, want=("idx",))["idx"]

    def forward(self, inputs, scales, means=None, training=None, noise: Optional[torch.Tensor] = None):
        """-> (outputs, likelihood) like compressai (dcae.py:657); eval: outputs = round(x - mu) + mu.  Under autograd
        the likelihood is a `training.GaussianLikelihoodFunction` node (kernel 3 forward, dcae_gc_backward backward)."""
        if training is None:
            training = self.training
        if training and noise is None:
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
        if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (inputs, scales, means)):
            from .training import gaussian_likelihood
            mu = means if means is not None else torch.zeros_like(inputs)
            lik = gaussian_likelihood(inputs, scales, mu, noise if training else None, self._bounds[0], self._bounds[1], self.likelihood_math)
            outputs = inputs + noise if training else (torch.round(inputs - mu) + mu)      # torch ops: differentiable like compressai's
            return outputs, lik
        if training:
            o = self._run(_lib.GC_NOISE, inputs, means, scales=scales, noise=noise, want=("lik",))
            return inputs + noise, o["lik"]
        o = self._run(_lib.GC_EVAL, inputs, means, scales=scales, want=("y_hat", "lik"))
        return o["y_hat"], o["lik"]

    def fused(self, inputs, scales, means):
        """Everything compress() needs from one launch: (symbols, indexes, y_hat, likelihood)."""
        o = self._run(_lib.GC_EVAL, inputs, means, scales=scales, want=("y_hat", "lik", "sym", "idx"))
        return o["sym"], o["idx"], o["y_hat"], o["lik"]


def _pmf_to_quantized_cdf(pmf, precision: int = 16):
    """compressai `_CXX.pmf_to_quantized_cdf`: native, in libdcae_rans.so next to the coder that consumes the tables
    (dcae_pmf_to_quantized_cdf, include/dcae_rans.h)."""
    from .ans import pmf_to_quantized_cdf
    return pmf_to_quantized_cdf(pmf, precision)

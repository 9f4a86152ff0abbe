"""Pack a reference state dict (keys of `dcae_b200.params`, i.e. `/root/reference/models/dcae.py`
checkpoints) into the device layouts `dcae_slice_weights` expects (include/dcae_b200.h).

One-time work per weight load; torch is used only to move/reshape tensors.  The arithmetic that
produces derived tensors (TF32 split, K = k(LN(dt))) runs through the library's own kernels.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _lib
from .params import (CC_HID1, CC_HID2, DICT_DIM, DICT_NUM, HEAD_DIM, HEAD_NUM, M_LATENT, NUM_SLICES,
                     SLICE_CH, cq, cs)


def f16_weight_planes(lib, w: torch.Tensor, taps: int, stream, amax: float | None = None):
    """fp16 hi/lo planes of w * 2^e (e chosen so max|w| 2^e lands in [8192, 16384): w_lo stays a normal fp16),
    each tap's channel run padded to a multiple of 64.  Returns (hi, lo, K16, descale = 2^-e).
    amax: max|w| when the caller already has it (one batched reduction + one sync for many weights, see PackedWeights)."""
    import math
    N, K = w.shape
    kc = K // taps
    Kp = (kc + 63) // 64 * 64
    if amax is None:
        amax = float(w.abs().max())
    e = math.floor(math.log2(16384.0 / amax)) if amax > 0 else 0
    e = max(min(e, 24), -24)
    hi = torch.empty(N, taps * Kp, dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.dcae_split_f16_weight(w.data_ptr(), N, taps, kc, float(2.0 ** e), hi.data_ptr(), lo.data_ptr(), stream),
               "dcae_split_f16_weight")
    return hi, lo, taps * Kp, float(2.0 ** -e)


class PackedWeights:
    """Owns every packed device tensor of the 5 slices and the ctypes array handed to the C ABI."""

    def __init__(self, params: Dict[str, torch.Tensor], device: torch.device, split_tf32: bool = True, math: str = "all"):
        """math: which operand formats to prepare -- "all", or the engine's math mode: the TF32 hi/lo split only for the
        tf32 modes, the fp16 planes only for the f16 modes (a repack per optimizer step should not pay for both)."""
        self.device = device
        self._keep: List[torch.Tensor] = []
        self.split = split_tf32
        self.want_tf32 = math in ("all", "tf32x3", "tf32")
        self.want_f16 = math in ("all", "f16x3", "f16")
        self.lib = _lib.load()
        self.array = (_lib.SliceWeights * NUM_SLICES)()
        self._pending: list = []          # (weight tensor, taps, key): fp16 planes made in one batch, see _flush_f16
        with torch.cuda.device(device):
            dt = params["dt"].to(device, torch.float32).contiguous()
            for i in range(NUM_SLICES):
                self._pack_slice(i, params, dt, self.array[i])
            self._flush_f16()
            torch.cuda.synchronize(device)

    def _flush_f16(self) -> None:
        """The scaled fp16 planes of every queued weight.  Their power-of-two scales need max|w|: ONE batched reduction and
        one device->host copy for all 125 weights (a repack per optimizer step used to pay 125 synchronising `abs().max()`).
        The queue is keyed by the fp32 weight's device pointer (or ("k" / "v", slice) for the dictionary side): ctypes
        copies a struct on assignment, so the planes are written through VIEWS of the fields of `self.array` afterwards."""
        if not self._pending:
            return
        import ctypes as C
        s = _lib.current_stream(self.device)
        amax = torch.stack(torch._foreach_norm([w for w, _, _ in self._pending], float("inf"))).tolist()
        made = {}
        for (w, taps, key), a in zip(self._pending, amax):
            h16, l16, K16, descale = f16_weight_planes(self.lib, w, taps, s, amax=a)
            self._keep += [h16, l16]
            made[key] = (h16.data_ptr(), l16.data_ptr(), K16, descale)
        self._pending = []

        def fill(view):
            m = made.get(view.w)
            if m is not None:
                view.w16_hi, view.w16_lo, view.K16, view.descale = m

        for i, W in enumerate(self.array):
            for name, typ in _lib.SliceWeights._fields_:
                if typ is _lib.Weight:
                    fill(getattr(W, name))                      # a view into W's memory, not a copy
                elif isinstance(typ, type) and issubclass(typ, C.Array) and typ._type_ is _lib.Weight:
                    arr = getattr(W, name)
                    for j in range(len(arr)):
                        fill(arr[j])
            kv = W.kv
            if ("k", i) in made:
                kv.K16_hi, kv.K16_lo, _, kv.k_descale = made[("k", i)]
                kv.Vt16_hi, kv.Vt16_lo, _, kv.v_descale = made[("v", i)]

    # ---- helpers --------------------------------------------------------------------------------
    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().to(self.device, torch.float32).contiguous()
        self._keep.append(t)
        return t

    def _vec(self, t: torch.Tensor) -> int:
        return self._dev(t.reshape(-1)).data_ptr()

    def _weight(self, w2d: torch.Tensor, taps: int = 1) -> _lib.Weight:
        """[N, K] row-major fp32 + its TF32 hi/lo split + its scaled fp16 hi/lo planes (all made by the library)."""
        w = self._dev(w2d)
        N, K = w.shape
        out = _lib.Weight()
        out.w, out.N, out.K = w.data_ptr(), N, K
        if self.split:
            s = _lib.current_stream(self.device)
            if self.want_tf32:
                hi, lo = torch.empty_like(w), torch.empty_like(w)
                self._keep += [hi, lo]
                _lib.check(self.lib.dcae_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), s), "dcae_split_tf32")
                out.w_hi, out.w_lo = hi.data_ptr(), lo.data_ptr()
            if self.want_f16:
                self._pending.append((w, taps, w.data_ptr()))
        return out

    @staticmethod
    def _conv3x3_to_gemm(w: torch.Tensor, perm: torch.Tensor | None = None) -> torch.Tensor:
        """[N, C, 3, 3] -> [N, 9*C] with K ordered tap-major (tap = 3*ky + kx), channels permuted."""
        if perm is not None:
            w = w[:, perm]
        N, Cc = w.shape[:2]
        return w.permute(0, 2, 3, 1).reshape(N, 9 * Cc)

    @staticmethod
    def _support_perm(i: int) -> torch.Tensor:
        """Library support order [dict_info | latent_scales | latent_means | y_hat_0..] expressed as
        indices into the reference order [latent_scales, latent_means, y_hat_0.., dict_info] (dcae.py:645-647)."""
        q = cq(i)
        return torch.cat([torch.arange(q, q + M_LATENT), torch.arange(0, q)])

    def _dw(self, w: torch.Tensor) -> int:
        """[C, 1, 3, 3] -> [9, C]."""
        return self._vec(w.reshape(w.shape[0], 9).t().contiguous())

    # ---- one slice ------------------------------------------------------------------------------
    def _pack_slice(self, i: int, P: Dict[str, torch.Tensor], dt: torch.Tensor, W: _lib.SliceWeights) -> None:
        p = f"dt_cross_attention.{i}."
        g = lambda k: P[p + k]
        W.x_trans, W.x_trans_b = self._weight(g("x_trans.weight")), self._vec(g("x_trans.bias"))
        W.ln_scale_g, W.ln_scale_b = self._vec(g("ln_scale.weight")), self._vec(g("ln_scale.bias"))
        W.msa_s, W.msa_s_b = self._weight(g("msa.s.weight").reshape(DICT_DIM, DICT_DIM)), self._vec(g("msa.s.bias"))
        for j in range(3):
            q = f"msa.dense.conv_layers.{j}.1."
            W.dense_in[j] = self._weight(g(q + "in_trans.weight").reshape(DICT_DIM, DICT_DIM))
            W.dense_in_b[j] = self._vec(g(q + "in_trans.bias"))
            W.dense_dw[j] = self._dw(g(q + "dw_conv.weight"))
            W.dense_dw_b[j] = self._vec(g(q + "dw_conv.bias"))
            W.dense_out[j] = self._weight(g(q + "out_trans.weight").reshape(DICT_DIM, DICT_DIM))
            W.dense_out_b[j] = self._vec(g(q + "out_trans.bias"))
        W.dense_proj = self._weight(g("msa.dense.proj.weight").reshape(DICT_DIM, 4 * DICT_DIM))
        W.dense_proj_b = self._vec(g("msa.dense.proj.bias"))
        W.spatial_w7 = self._vec(g("msa.spatial_atte.conv1.weight"))
        for r in (1, 2, 3):
            setattr(W, f"res_scale_{r}", self._vec(g(f"res_scale_{r}.scale")))
        W.lnx_g, W.lnx_b = self._vec(g("lnx.weight")), self._vec(g("lnx.bias"))
        W.q_trans, W.q_trans_b = self._weight(g("q_trans.weight")), self._vec(g("q_trans.bias"))
        W.kv = self.dictionary_kv(dt, g("dict_ln.weight"), g("dict_ln.bias"), g("k.weight"), g("k.bias"), g("scale"), slice_index=i)
        W.linear, W.linear_b = self._weight(g("linear.weight")), self._vec(g("linear.bias"))
        W.ln_mlp_g, W.ln_mlp_b = self._vec(g("ln_mlp.weight")), self._vec(g("ln_mlp.bias"))
        W.fc1, W.fc1_b = self._weight(g("mlp.fc1.weight")), self._vec(g("mlp.fc1.bias"))
        W.mlp_dw, W.mlp_dw_b = self._dw(g("mlp.dwconv.dwconv.weight")), self._vec(g("mlp.dwconv.dwconv.bias"))
        W.fc2, W.fc2_b = self._weight(g("mlp.fc2.weight")), self._vec(g("mlp.fc2.bias"))
        W.output_trans, W.output_trans_b = self._weight(g("output_trans.0.weight")), self._vec(g("output_trans.0.bias"))

        perm = self._support_perm(i)
        c_sup = cs(i)
        mean, scale, lrp = (lambda k, f=f: P[f"{f}.{i}.{k}"] for f in
                            ("cc_mean_transforms", "cc_scale_transforms", "lrp_transforms"))
        lrp0 = lrp("0.weight")
        cc1 = torch.cat([self._conv3x3_to_gemm(mean("0.weight"), perm),
                         self._conv3x3_to_gemm(scale("0.weight"), perm),
                         self._conv3x3_to_gemm(lrp0[:, :c_sup], perm)], dim=0)
        W.cc1 = self._weight(cc1, taps=9)
        W.cc1_b = self._vec(torch.cat([mean("0.bias"), scale("0.bias"), mean("0.bias").new_zeros(CC_HID1)]))
        W.lrp1y = self._weight(self._conv3x3_to_gemm(lrp0[:, c_sup:]), taps=9)
        W.lrp1_b = self._vec(lrp("0.bias"))
        W.mean2, W.mean2_b = self._weight(self._conv3x3_to_gemm(mean("2.weight")), taps=9), self._vec(mean("2.bias"))
        W.scale2, W.scale2_b = self._weight(self._conv3x3_to_gemm(scale("2.weight")), taps=9), self._vec(scale("2.bias"))
        W.lrp2, W.lrp2_b = self._weight(self._conv3x3_to_gemm(lrp("2.weight")), taps=9), self._vec(lrp("2.bias"))
        W.mean3, W.mean3_b = self._weight(self._conv3x3_to_gemm(mean("4.weight")), taps=9), self._vec(mean("4.bias"))
        W.scale3, W.scale3_b = self._weight(self._conv3x3_to_gemm(scale("4.weight")), taps=9), self._vec(scale("4.bias"))
        W.lrp3, W.lrp3_b = self._weight(self._conv3x3_to_gemm(lrp("4.weight")), taps=9), self._vec(lrp("4.bias"))

    def dictionary_kv(self, dt, ln_w, ln_b, k_w, k_b, head_scale, slice_index: int = 0) -> _lib.DictKV:
        """K = k(dict_ln(dt)), V = dict_ln(dt), per head [20, 128, 32] (dcae.py:492-495); batch invariant,
        so computed once here with the library's LayerNorm + fp32 GEMM; plus the TF32 hi/lo splits of K and of
        V transposed per head ([20, 32, 128]) that the tcgen05 attention kernel consumes."""
        lib, s = self.lib, _lib.current_stream(self.device)
        dt = dt.to(self.device, torch.float32).contiguous()
        d = torch.empty(DICT_NUM, DICT_DIM, device=self.device)
        gam, bet = self._dev(ln_w), self._dev(ln_b)
        _lib.check(lib.dcae_op_layernorm(dt.data_ptr(), DICT_DIM, gam.data_ptr(), bet.data_ptr(), DICT_DIM, DICT_NUM,
                                         d.data_ptr(), DICT_DIM, None, s), "dcae_op_layernorm")
        kw, kb = self._dev(k_w), self._dev(k_b)
        k = torch.empty(DICT_NUM, DICT_DIM, device=self.device)
        a = _lib.Operand(d.data_ptr(), DICT_DIM, 0, DICT_DIM, 0, 0, 1, 1, 1, DICT_NUM, None, 0)
        w = _lib.Weight(kw.data_ptr(), None, None, DICT_DIM, DICT_DIM)
        e = _lib.Epilogue()
        e.bias, e.out, e.out_ld = kb.data_ptr(), k.data_ptr(), DICT_DIM
        _lib.check(lib.dcae_op_gemm(a, w, e, _lib.MATH["fp32"], s), "dcae_op_gemm(k)")
        # 'n (e c) -> e n c'
        Kh = k.reshape(DICT_NUM, HEAD_NUM, HEAD_DIM).permute(1, 0, 2).contiguous()
        Vh = d.reshape(DICT_NUM, HEAD_NUM, HEAD_DIM).permute(1, 0, 2).contiguous()
        Vt = Vh.transpose(1, 2).contiguous()          # [20, 32, 128]
        kv = _lib.DictKV()
        kv.Kh, kv.Vh, kv.head_scale = Kh.data_ptr(), Vh.data_ptr(), self._vec(head_scale)
        self._keep += [Kh, Vh, Vt]
        for src, hi_name, lo_name in ((Kh, "Kh_hi", "Kh_lo"), (Vt, "Vt_hi", "Vt_lo")):
            hi, lo = torch.empty_like(src), torch.empty_like(src)
            _lib.check(lib.dcae_split_tf32(src.data_ptr(), hi.data_ptr(), lo.data_ptr(), src.numel(), s), "dcae_split_tf32")
            self._keep += [hi, lo]
            setattr(kv, hi_name, hi.data_ptr())
            setattr(kv, lo_name, lo.data_ptr())
        # fp16 planes for the f16x3 attention kernel: K as the [128, 640] matrix itself, V transposed [640, 128]
        kc, vt = k.contiguous(), d.t().contiguous()
        self._keep += [kc, vt]
        self._pending += [(kc, 1, ("k", slice_index)), (vt, 1, ("v", slice_index))]
        return kv

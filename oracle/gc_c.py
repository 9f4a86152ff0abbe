"""TEST INFRASTRUCTURE ONLY.  ctypes front end of oracle/gc_oracle.c (built by __graft_entry__.build() into
oracle/_build/libgc_oracle.so): the plain-C restatement of the GaussianConditional element math, on torch CPU tensors."""
from __future__ import annotations

import ctypes as C
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgc_oracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def forward_eval(y, scale, mu, scale_bound=0.11, lik_bound=1e-9):
    """-> (y_hat, likelihood), both shaped like y."""
    y, mu, scale = (t.contiguous().float() for t in (y, mu, scale))
    y_hat, lik = torch.empty_like(y), torch.empty_like(y)
    _load().gc_forward_eval(_p(y), _p(mu), _p(scale), C.c_int64(y.numel()), C.c_float(scale_bound), C.c_float(lik_bound), _p(y_hat), _p(lik))
    return y_hat, lik


def symbols(y, mu):
    y, mu = y.contiguous().float(), mu.contiguous().float()
    out = torch.empty(y.shape, dtype=torch.int32)
    _load().gc_symbols(_p(y), _p(mu), C.c_int64(y.numel()), _p(out))
    return out


def build_indexes(scale, table, scale_bound=0.11):
    scale, table = scale.contiguous().float(), table.contiguous().float()
    out = torch.empty(scale.shape, dtype=torch.int32)
    _load().gc_build_indexes(_p(scale), C.c_int64(scale.numel()), _p(table), C.c_int32(table.numel()), C.c_float(scale_bound), _p(out))
    return out


def dequantize(sym, mu):
    sym, mu = sym.contiguous().int(), mu.contiguous().float()
    out = torch.empty_like(mu)
    _load().gc_dequantize(_p(sym), _p(mu), C.c_int64(mu.numel()), _p(out))
    return out

"""CPU: the plain-C restatement of kernel 3's element math (oracle/gc_oracle.c, glibc erfcf) against the torch oracle
(oracle/gaussian_conditional.py, pinned bit for bit against the reference's in-tree likelihood): symbols, indexes and
y_hat bit-exact; likelihood within 1e-6 relative + 2 ulp(0.5) absolute (two erfc implementations, each <= 1-2 ulp)."""
import math

import pytest
import torch

from oracle import gaussian_conditional as orc


@pytest.fixture(scope="module")
def gcc():
    import __graft_entry__ as ge
    ge.build()
    from oracle import gc_c
    assert gc_c.available(), "oracle/_build/libgc_oracle.so was not built"
    return gc_c


def _inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    y = 4 * torch.randn(n, generator=g)
    mu = 2 * torch.randn(n, generator=g)
    scale = torch.exp(torch.empty(n).uniform_(-3.0, 5.7, generator=g))
    return y, mu, scale


def test_c_oracle_matches_the_torch_oracle(gcc):
    y, mu, scale = _inputs(200_000, 4321)
    table = orc.get_scale_table()
    assert torch.equal(gcc.symbols(y, mu), orc.quantize(y, "symbols", mu))
    assert torch.equal(gcc.build_indexes(scale, table), orc.build_indexes(scale, table))
    y_hat, lik = gcc.forward_eval(y, scale, mu)
    want_hat = orc.quantize(y, "dequantize", mu)
    assert torch.equal(y_hat, want_hat)
    want = orc.lower_bound(orc.likelihood(want_hat, scale, mu), 1e-9)
    err = (lik.double() - want.double()).abs()
    assert bool((err <= 1e-6 * want.double() + 2 * 5.96e-8).all()), float((err / want.double()).max())
    assert torch.equal(gcc.dequantize(orc.quantize(y, "symbols", mu), mu), want_hat)


def test_c_oracle_edge_vectors(gcc):
    """half-to-even ties, scales at / next to every table entry and the clamp, NaN propagation."""
    table = orc.get_scale_table()
    ties = torch.tensor([0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 3.5])
    assert gcc.symbols(ties, torch.zeros_like(ties)).tolist() == [0, 0, 2, -2, 2, -2, 4]
    near = torch.cat([table, torch.nextafter(table, torch.tensor(math.inf)), torch.nextafter(table, torch.tensor(-math.inf)),
                      torch.tensor([-1.0, 0.0, 0.11, 256.0, 1e4, math.inf, math.nan])])
    assert torch.equal(gcc.build_indexes(near, table), orc.build_indexes(near, table))
    y = torch.tensor([0.0, 0.3, 40.0, math.nan])
    mu = torch.zeros(4)
    sc = torch.tensor([1.0, -5.0, 0.2, 1.0])
    _, lik = gcc.forward_eval(y, sc, mu)
    want = orc.lower_bound(orc.likelihood(orc.quantize(y, "dequantize", mu), sc, mu), 1e-9)
    assert math.isnan(float(lik[3])) and math.isnan(float(want[3]))
    assert torch.allclose(lik[:3], want[:3], rtol=1e-6, atol=1.2e-7) and float(lik[2]) == pytest.approx(1e-9)

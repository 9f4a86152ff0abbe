"""Fit and check of kernel 3's fast likelihood (dcae_b200/csrc/gc_kernel.cu, `gaussian_likelihood_fast`).

    lik = 1/2 erfc(a) - 1/2 erfc(b),  a = (v - 1/2) k,  b = (v + 1/2) k,  k = 1 / (s sqrt 2),  v = |out - mu|   (dcae.py:839-857)

The reference evaluates the two erfc in fp32 and subtracts: for large scales both are close to 1 and the difference
carries the rounding noise of its operands (up to ~4e-5 relative at s = 256).  The kernel instead writes, for x >= 0,
erfc(x) = 2^P(z) with ONE polynomial in the centred variable z = x - C over [0, XMAX] and uses

    a >= 0:  lik = 1/2 2^P(a) (1 - 2^(P(b) - P(a))),   P(b) - P(a) = (b - a) Q_a(b)       [Q_a = Horner intermediates of P at a]
    a <  0:  lik = 1/2 [(1 - 2^P(|a|)) + (1 - 2^P(b))]                                     [P(b) = P(|a|) + (b - |a|) Q_|a|(b)]

so the difference never cancels (b - a is k itself).  P has no constant-term constraint in z; accuracy near x = 0
comes from the fit weight.  This script fits the coefficients (weighted minimax by Lawson iteration in a Chebyshev
basis), evaluates the whole formula in emulated fp32 (Horner with FMA) against mpmath / fp64 and prints the
coefficients for the kernel."""
import sys

import numpy as np
from scipy import special

XMAX = 5.5
CEN = 2.75
LOG2E = 1.4426950408889634


def target(x):
    """log2 erfc(x), x >= 0, accurate in the tails (erfcx)."""
    return -x * x * LOG2E + np.log2(special.erfcx(x))


def weight(x):
    """required accuracy of the exponent: 1 below 4.2, relaxing beyond (see the range argument in DESIGN.md)."""
    w = np.ones_like(x)
    far = x > 4.2
    # E(x) / 2e-9 is the share of erfc(x) in a likelihood at the floor: tolerance grows with its inverse
    w[far] = np.minimum(1.0, special.erfc(x[far]) / 2e-9 * 2.0)
    return np.maximum(w, 1e-6)


def fit(deg, n=6000, iters=80):
    k = np.arange(n)
    x = 0.5 * XMAX * (1 - np.cos(np.pi * (k + 0.5) / n))
    z = (x - CEN) / CEN
    y, w = target(x), weight(x)
    V = np.polynomial.chebyshev.chebvander(z, deg)
    lam = np.ones(n)
    for _ in range(iters):
        ww = w * np.sqrt(lam)
        c = np.linalg.lstsq(V * ww[:, None], y * ww, rcond=None)[0]
        err = np.abs((V @ c - y) * w)
        lam = lam * (err / err.max() + 1e-3)
        lam /= lam.mean()
    # Chebyshev in z/CEN -> monomial in z = x - CEN
    mono = np.polynomial.chebyshev.cheb2poly(c)
    mono = mono / CEN ** np.arange(deg + 1)
    return mono, err.max()


ERF_SMALL_MAX = 0.5
ERF_SMALL_DEG = 3


def fit_erf_small(deg=ERF_SMALL_DEG, n=4000, xmax=ERF_SMALL_MAX):
    """erf(x) = x T(x^2) on [0, xmax]: least squares in Chebyshev nodes on the relative error (T ~ 1.128 .. 0.84)."""
    k = np.arange(n)
    x = 0.5 * xmax * (1 - np.cos(np.pi * (k + 0.5) / n)) + 1e-12
    t = x * x
    y = special.erf(x) / x
    V = np.vander(t, deg + 1, increasing=True)
    lam = np.ones(n)
    for _ in range(60):
        ww = np.sqrt(lam) / y
        c = np.linalg.lstsq(V * ww[:, None], y * ww, rcond=None)[0]
        err = np.abs(V @ c - y) / y
        lam = lam * (err / err.max() + 1e-3)
        lam /= lam.mean()
    return c, err.max()


f32 = np.float32
ERF_SMALL = None


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def ex2(x):
    return np.exp2(x.astype(np.float64)).astype(np.float32)          # ex2.approx.ftz: 2 ulp; modelled as correctly rounded


EXPM1_DEG = 7


def expm1_2_neg(u):
    """-(2^u - 1) for u <= 0, fp32: polynomial for u > -1, 1 - ex2(u) below."""
    ln2 = np.log(2.0)
    coef = [ln2 ** (j + 1) / np.math.factorial(j + 1) for j in range(EXPM1_DEG)] if hasattr(np, "math") else None
    import math
    coef = [ln2 ** (j + 1) / math.factorial(j + 1) for j in range(EXPM1_DEG)]
    r = np.full_like(u, f32(coef[-1]))
    for cj in coef[-2::-1]:
        r = fma(r, u, np.full_like(u, f32(cj)))
    small = (-(r * u)).astype(np.float32)          # -(u c1 + u^2 c2 + ...)
    big = (f32(1.0) - ex2(u)).astype(np.float32)
    return np.where(u > f32(-1.0), small, big)


def erf_small(x, tc):
    t = (x * x).astype(np.float32)
    r = np.full_like(x, f32(tc[-1]))
    for cj in tc[-2::-1]:
        r = fma(r, t, np.full_like(x, f32(cj)))
    return (x * r).astype(np.float32)


def lik_fast(v, s, coef):
    """The kernel's arithmetic in emulated fp32.  v = |out - mu| >= 0, s >= 0.11."""
    v, s = v.astype(np.float32), s.astype(np.float32)
    k = (f32(0.70710678118654752440) / s).astype(np.float32)          # one rounding (rcp + Newton in the kernel)
    a = ((v - f32(0.5)) * k).astype(np.float32)
    b = ((v + f32(0.5)) * k).astype(np.float32)
    neg = a < 0
    ap = np.minimum(np.abs(a), f32(XMAX))
    bp = np.minimum(b, f32(XMAX))
    d = np.where(neg, (f32(2.0) * v * k).astype(np.float32), k)
    d = np.where(b > f32(XMAX), (bp - ap).astype(np.float32), d)      # clamped: b' - a' directly (large step, no cancellation)
    za, zb = (ap - f32(CEN)).astype(np.float32), (bp - f32(CEN)).astype(np.float32)
    c = [f32(x) for x in coef]
    n = len(c) - 1
    # Horner at za, keeping the intermediates q_{n-1} .. q_0 (coefficients of Q), then Horner of Q at zb
    q = [np.full_like(za, c[n])]
    for j in range(n - 1, -1, -1):
        q.append(fma(q[-1], za, np.full_like(za, c[j])))
    pa = q[-1]
    Q = q[0]
    for j in range(1, n):
        Q = fma(Q, zb, q[j])
    delta = (d * Q).astype(np.float32)
    pa = np.minimum(pa, f32(0.0))
    pb = np.minimum((pa + delta).astype(np.float32), f32(0.0))
    ea = ex2(pa)
    x1 = expm1_2_neg(np.minimum(delta, f32(0.0)))
    case1 = (ea * x1).astype(np.float32)                                           # a >= 0
    case2 = ((f32(1.0) - ea) + (f32(1.0) - ex2(pb))).astype(np.float32)            # a < 0, b >= ERF_SMALL_MAX: erf(|a|) + erf(b), plain
    big = np.where(neg, case2, case1)
    # a < 0 and b small: both error functions from the odd polynomial (relative accuracy near 0)
    lim = f32(ERF_SMALL_MAX)
    sm = (erf_small(np.minimum(ap, lim), ERF_SMALL) + erf_small(np.minimum(bp, lim), ERF_SMALL)).astype(np.float32)
    tot = np.where(neg & (b < lim), sm, big)
    return (f32(0.5) * tot).astype(np.float32)


def lik_ref32(v, s):
    """the reference formula in fp32 (dcae.py:839-857), numpy erfc in fp64 rounded to fp32 per op."""
    v, s = v.astype(np.float32), s.astype(np.float32)
    c = f32(-(2 ** -0.5))
    up = f32(0.5) * special.erfc((c * ((f32(0.5) - v) / s)).astype(np.float64)).astype(np.float32)
    lo = f32(0.5) * special.erfc((c * ((f32(-0.5) - v) / s)).astype(np.float64)).astype(np.float32)
    return (up - lo).astype(np.float32)


def lik_exact(v, s):
    import mpmath as mp
    mp.mp.dps = 40
    out = np.empty(len(v))
    for i, (vv, ss) in enumerate(zip(v.astype(np.float32), s.astype(np.float32))):
        vv, ss = mp.mpf(float(vv)), mp.mpf(float(ss))
        k = 1 / (ss * mp.sqrt(2))
        out[i] = float((mp.erfc((vv - mp.mpf(0.5)) * k) - mp.erfc((vv + mp.mpf(0.5)) * k)) / 2)
    return out


def samples(n, seed=0):
    g = np.random.default_rng(seed)
    s = np.exp(g.uniform(np.log(0.11), np.log(300.0), n))
    mode = g.integers(0, 3, n)
    v = np.where(mode == 0, np.rint(np.abs(g.standard_normal(n)) * s * 1.5),                 # eval mode: integers, typical
                 np.where(mode == 1, np.abs(g.standard_normal(n)) * s * 2.5,                   # noise mode: continuous
                          np.rint(g.uniform(0, 7, n) * s)))                                    # tails down to the floor
    return v.astype(np.float32), s.astype(np.float32)


if __name__ == "__main__":
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    coef, e = fit(deg)
    print(f"degree {deg}: weighted minimax error of the exponent {e:.3e}")
    tc, te = fit_erf_small()
    globals()["ERF_SMALL"] = tc
    print(f"erf(x) = x T(x^2) on [0, {ERF_SMALL_MAX}], degree {ERF_SMALL_DEG} in x^2: max relative error {te:.3e}")
    v, s = samples(n)
    exact = lik_exact(v, s)
    keep = exact >= 1e-9
    fast, ref = lik_fast(v, s, coef).astype(np.float64), lik_ref32(v, s).astype(np.float64)
    rf = np.abs(fast - exact)[keep] / exact[keep]
    rr = np.abs(ref - exact)[keep] / exact[keep]
    print(f"{keep.sum()} samples above the 1e-9 floor; relative error vs exact (mpmath):")
    print(f"  fast formula (emulated fp32): max {rf.max():.3e}  p99.9 {np.quantile(rf, 0.999):.3e}  median {np.median(rf):.3e}")
    print(f"  reference fp32 formula      : max {rr.max():.3e}  p99.9 {np.quantile(rr, 0.999):.3e}  median {np.median(rr):.3e}")
    i = np.argmax(rf)
    print(f"  worst fast: v={v[keep][i]} s={s[keep][i]} exact={exact[keep][i]:.6e}")
    below = ~keep
    if below.any():
        print(f"  below the floor: max fast value {fast[below].max():.3e} (must stay < ~1.1e-9 so that the floor applies)")
    bb = (v[keep] + 0.5) * 0.70710678 / s[keep]
    for nm, m in (("a >= 0", v[keep] >= 0.5), ("a < 0, b small", (v[keep] < 0.5) & (bb < ERF_SMALL_MAX)), ("a < 0, b large", (v[keep] < 0.5) & (bb >= ERF_SMALL_MAX))):
        if m.any():
            print(f"    {nm:14s}: n={m.sum():6d} fast max {rf[m].max():.3e} p99.9 {np.quantile(rf[m], 0.999):.3e} | ref max {rr[m].max():.3e}")
    print("erf_small T coefficients (in x^2, increasing):")
    print(", ".join(f"{float(np.float32(cj))!r}f" for cj in tc))
    print("coefficients (z = x - %.2f, increasing degree):" % CEN)
    print(", ".join(f"{float(np.float32(cj))!r}f" for cj in coef))

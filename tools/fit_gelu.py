"""Fit and check of gelu_fast (dcae_b200/csrc/common.cuh): erfc(t) = 2^(-t R(t)), R degree 7; prints the coefficients."""
import numpy as np
from scipy import special
import torch
# fit R(t) with erfc(t) = exp2(-t*R(t)),  t in [0, TMAX]
TMAX = 4.3   # erfc(4.3)=1.2e-9 ; |x|=6.08
def target(t): return -np.log2(special.erfc(t)) / t
def fit(deg, n=4000):
    k = np.arange(n); t = 0.5*TMAX*(1-np.cos(np.pi*(k+0.5)/n)); t = np.maximum(t,1e-9)
    y = target(t)
    E = special.erfc(t)
    # error in E = E*ln2*t*dR  -> weight = E*t
    wgt = E*t*np.log(2)
    V = np.vander(t, deg+1, increasing=True)
    c = np.linalg.lstsq(V*wgt[:,None], y*wgt, rcond=None)[0]
    # iterate reweighting (Lawson) for minimax
    lam = np.ones(n)
    for it in range(60):
        c = np.linalg.lstsq(V*(wgt*np.sqrt(lam))[:,None], y*wgt*np.sqrt(lam), rcond=None)[0]
        err = np.abs((V@c - y)*wgt)
        lam = lam*(err/err.max()+1e-3); lam/=lam.mean()
    return c
def gelu32(x, c):
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x)*np.float32(0.70710678118654752440), np.float32(TMAX))
    r = np.float32(c[-1])*np.ones_like(t)
    for ci in c[-2::-1]:
        r = (r*t + np.float32(ci)).astype(np.float32)
    E = np.exp2((-(t*r)).astype(np.float32)).astype(np.float32)
    hx = np.float32(0.5)*x
    he = (hx*E).astype(np.float32)
    return np.where(x>=0, (x-he).astype(np.float32), he)
xs = np.concatenate([np.linspace(-12,12,2000001), np.random.default_rng(0).normal(size=2000000)*2]).astype(np.float32)
exact = 0.5*xs.astype(np.float64)*special.erfc(-xs.astype(np.float64)/np.sqrt(2))
tg = torch.nn.functional.gelu(torch.from_numpy(xs)).numpy()
et = np.abs(tg-exact); print("torch fp32 gelu: max abs err %.3e, max err/max(1,|x|) %.3e"%(et.max(), (et/np.maximum(1,np.abs(xs))).max()))
for deg in (5,6,7,8):
    c = fit(deg)
    g = gelu32(xs, c)
    e = np.abs(g-exact)
    m=np.abs(xs)<0.5
    print('   small-x max rel err %.3e'%(e[m]/np.maximum(np.abs(exact[m]),1e-30)).max())
    print(deg, "max abs err %.3e  rel-to-max(1,|x|) %.3e  vs torch max %.3e"%(e.max(), (e/np.maximum(1,np.abs(xs))).max(), np.abs(g-tg).max()))
    if deg==7: print([float(np.float32(v)) for v in c])


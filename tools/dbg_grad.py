import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import rel_err
from oracle.reference_loader import build_reference_net, inject_latents
from dcae_b200.params import init_entropy_params
from dcae_b200.training import EntropyModel
from dcae_b200 import torch_graph
from oracle import gaussian_conditional as ogc
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
P = init_entropy_params(7, "lively")
B, h, w = 2, 8, 12
pixels = B*h*w*256
net = build_reference_net(P).cuda().train()
model = EntropyModel(P, device="cuda:0", math="f16x3", likelihood_math="reference").train()
orig = net.gaussian_conditional.forward
gen = torch.Generator().manual_seed(21)
y = (4 * torch.randn(B, 320, h, w, generator=gen)).cuda()
ls, lm = torch.randn(B, 320, h, w, generator=gen).cuda(), torch.randn(B, 320, h, w, generator=gen).cuda()
noise = torch.empty(B, 320, h, w).uniform_(-0.5, 0.5, generator=gen).cuda()
for which in ("bpp", "mse", "mu", "sc"):
    chunks = iter(noise.chunk(5, 1))
    net.gaussian_conditional.forward = lambda inp, sc, means=None: orig(inp, sc, means, training=True, noise=next(chunks))
    lr = [t.clone().requires_grad_(True) for t in (y, ls, lm)]
    out = net(inject_latents(net, *lr))
    lo = [t.clone().requires_grad_(True) for t in (y, ls, lm)]
    o = model(*lo, noise=noise)
    # pure torch graph (product's recompute graph) with oracle GC
    lt = [t.clone().requires_grad_(True) for t in (y, ls, lm)]
    Pc = {k: v.cuda() for k, v in P.items()}
    def gc(ys, sc, mu, nz):
        return ogc.lower_bound(ogc.likelihood(ys + nz, sc, mu), 1e-9)
    yh_t, mu_t, sc_t, lik_t = torch_graph.slice_loop(Pc, *lt, gc, noise)
    def L(lik, yh, mu, sc):
        if which == "bpp": return torch.log(lik).sum() / (-math.log(2) * pixels)
        if which == "mse": return torch.mean((yh - y) ** 2)
        if which == "mu": return (mu * mu).mean()
        return (sc * sc).mean()
    gr = torch.autograd.grad(L(out["likelihoods"]["y"], out["x_hat"], out["para"]["means"], out["para"]["scales"]), lr)
    go = torch.autograd.grad(L(o["likelihoods"], o["y_hat"], o["means"], o["scales"]), lo)
    gt = torch.autograd.grad(L(lik_t, yh_t, mu_t, sc_t), lt)
    print(which, "ours vs ref:", [f"{rel_err(a, b):.2e}" for a, b in zip(go, gr)], " torch_graph vs ref:", [f"{rel_err(a, b):.2e}" for a, b in zip(gt, gr)])
    if which == "bpp":
        d = (go[0] - gr[0]).abs(); i = d.argmax(); idx = torch.unravel_index(i, d.shape)
        print("  worst elem", [int(v) for v in idx], "ours", float(go[0].flatten()[i]), "ref", float(gr[0].flatten()[i]), "lik ref", float(out["likelihoods"]["y"].flatten()[i]), "lik ours", float(o["likelihoods"].flatten()[i]),
              "scale", float(out["para"]["scales"].flatten()[i]), "v", float((y+noise-out["para"]["means"]).flatten()[i]))

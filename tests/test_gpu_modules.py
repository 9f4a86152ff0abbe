"""Module-level drop-ins: `net.dt_cross_attention[i]`, `cc_mean/cc_scale/lrp_transforms[i]` swapped for
dcae_b200 modules, checked against the oracle restatement of the same reference modules, and the reference's own
loop text (dcae.py:638-670, restated here line by line over the swapped attributes) against the oracle loop."""
import pytest
import torch

from oracle import entropy_model as orc
from oracle import gaussian_conditional as ogc

from _util import rel_err

pytestmark = pytest.mark.gpu

TOL = {"f16x3": 1e-5, "fp32": 4e-6}


def _sub(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


@pytest.fixture(scope="module")
def params():
    from dcae_b200.params import init_entropy_params
    return init_entropy_params(7, "lively")


class _Net(torch.nn.Module):
    """Just the attributes DCAE.forward touches inside the slice loop."""

    def __init__(self):
        super().__init__()
        ident = lambda: torch.nn.ModuleList(torch.nn.Identity() for _ in range(5))      # noqa: E731
        self.dt_cross_attention, self.cc_mean_transforms = ident(), ident()
        self.cc_scale_transforms, self.lrp_transforms = ident(), ident()
        from dcae_b200.gaussian_conditional import GaussianConditional
        self.gaussian_conditional = GaussianConditional(None)


@pytest.mark.parametrize("math", ["f16x3", "fp32"])
def test_swapped_modules_match_the_oracle_modules(math, params):
    from dcae_b200.modules import accelerate
    net = _Net()
    accelerate(net, math=math, state_dict=params)
    gen = torch.Generator().manual_seed(5)
    B, h, w = 2, 7, 9
    for i in (0, 3):
        cq, cs = 640 + 64 * i, 960 + 64 * i
        x = torch.randn(B, cq, h, w, generator=gen)
        want = orc.dictionary_cross_attention(x, params["dt"], _sub(params, f"dt_cross_attention.{i}."))
        got = net.dt_cross_attention[i](x.cuda(), params["dt"].cuda().unsqueeze(0).repeat(B, 1, 1))
        assert got.shape == (B, 320, h, w) and rel_err(got.cpu(), want) < TOL[math]
        sup = torch.randn(B, cs + 64, h, w, generator=gen)
        for which, name in ((0, "cc_mean_transforms"), (1, "cc_scale_transforms"), (2, "lrp_transforms")):
            xin = sup if which == 2 else sup[:, :cs].contiguous()
            want = orc.conv_stack(xin, _sub(params, f"{name}.{i}."))
            got = getattr(net, name)[i](xin.cuda())
            assert got.shape == (B, 64, h, w) and rel_err(got.cpu(), want) < TOL[math], (name, i)


def test_reference_loop_text_runs_on_the_swapped_modules(params):
    """dcae.py:638-670 restated over the swapped attributes; teacher-forced per slice against the oracle loop."""
    from dcae_b200.modules import accelerate
    net = _Net()
    accelerate(net, math="f16x3", state_dict=params)
    gen = torch.Generator().manual_seed(6)
    B, h, w = 1, 8, 12
    y = 4 * torch.randn(B, 320, h, w, generator=gen)
    ls, lm = torch.randn(B, 320, h, w, generator=gen), torch.randn(B, 320, h, w, generator=gen)
    o = orc.SliceLoopOracle(params)
    want_yhat, want_mu, want_scale, _ = o.forward(y, ls, lm)[:4]
    dt = params["dt"].cuda().unsqueeze(0)
    y_hat_slices = []
    for i, y_slice in enumerate(y.cuda().chunk(5, 1)):
        ref_prev = [want_yhat[:, 64 * j:64 * j + 64].cuda() for j in range(i)]          # teacher forcing
        query = torch.cat([ls.cuda(), lm.cuda()] + ref_prev, dim=1)                    # dcae.py:645
        dict_info = net.dt_cross_attention[i](query, dt)                               # :646
        support = torch.cat([query, dict_info], dim=1)                                 # :647
        mu = net.cc_mean_transforms[i](support)[:, :, :h, :w]                          # :649-651
        scale = net.cc_scale_transforms[i](support)[:, :, :h, :w]                      # :653-655
        _, lik = net.gaussian_conditional(y_slice, scale, mu)                          # :657
        y_hat_slice = ogc.ste_round(y_slice - mu) + mu                                 # :659
        lrp = net.lrp_transforms[i](torch.cat([support, y_hat_slice], dim=1))          # :661-662
        y_hat_slice = y_hat_slice + 0.5 * torch.tanh(lrp)                              # :663-664
        y_hat_slices.append(y_hat_slice)
        sl = slice(64 * i, 64 * i + 64)
        assert rel_err(mu.cpu(), want_mu[:, sl]) < 1e-5 and rel_err(scale.cpu(), want_scale[:, sl]) < 1e-5, i
        flips = (torch.round(y_slice - mu).cpu() != torch.round(y[:, sl] - want_mu[:, sl])).float().mean()
        if float(flips) == 0.0:
            assert rel_err(y_hat_slice.cpu(), want_yhat[:, sl]) < 1e-5, i
        assert bool(((lik >= 1e-9) & (lik <= 1)).all())

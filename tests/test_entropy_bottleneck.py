"""The hyper-latent entropy model (first part): `dcae_b200.EntropyBottleneck` against the oracle restatement
of compressai's published EntropyBottleneck, its own density properties, and the range coder.  CPU part: tables and
host logic; GPU part (marked): the fused device pass."""
import numpy as np
import pytest
import torch

from oracle import entropy_bottleneck as oeb
from oracle import gaussian_conditional as ogc

C = 192


@pytest.fixture(scope="module")
def params():
    return oeb.init_params(C, seed=3, trained_like=True)


def _module(params, device="cpu"):
    import __graft_entry__ as ge
    ge.build()
    from dcae_b200.entropy_bottleneck import EntropyBottleneck
    eb = EntropyBottleneck(C)
    missing, unexpected = eb.load_state_dict(params, strict=False)
    assert not unexpected and set(missing) <= {"target", "_offset", "_quantized_cdf", "_cdf_length"}
    return eb.to(device)


def test_density_is_a_density(params):
    """Implementation-independent: the per-channel network is a CDF in logit space -- monotone, and its unit-bin masses
    between the outer quantiles plus the two tails sum to 1."""
    v = torch.linspace(-40, 40, 801).reshape(1, 1, -1).repeat(C, 1, 1)
    logits = oeb.logits_cumulative(params, v)
    assert bool((logits[:, 0, 1:] > logits[:, 0, :-1]).all())
    q, off, ln = oeb.build_tables(params, ogc.pmf_to_quantized_cdf)
    for c in (0, 17, 191):
        row = q[c, : int(ln[c])]
        assert int(row[0]) == 0 and int(row[-1]) == 65536 and bool((row.diff() > 0).all())
    n = int(ln.max()) - 2
    samples = torch.arange(n)[None, :] + (oeb.medians(params)[:, 0, 0] + off)[:, None, None]
    pmf, lower, upper = oeb.likelihood(params, samples)
    for c in (0, 17, 191):
        k = int(ln[c]) - 2
        total = pmf[c, 0, :k].sum() + torch.sigmoid(lower[c, 0, 0]) + torch.sigmoid(-upper[c, 0, k - 1])
        assert abs(float(total) - 1.0) < 1e-5


def test_module_tables_keys_and_aux_loss_match_the_restatement(params):
    eb = _module(params)
    assert eb.update() is True and eb.update() is False
    q, off, ln = oeb.build_tables(params, ogc.pmf_to_quantized_cdf)
    assert torch.equal(eb.quantized_cdf, q) and torch.equal(eb.offset, off) and torch.equal(eb.cdf_length, ln)
    keys = set(eb.state_dict())
    assert {"_matrix0", "_matrix4", "_bias4", "_factor3", "quantiles", "target", "_offset", "_quantized_cdf", "_cdf_length"} <= keys and len(keys) == 19
    assert abs(float(eb.loss().detach()) - float(oeb.aux_loss(params))) < 1e-3
    assert tuple(eb._get_medians().shape) == (C, 1, 1)
    # a baked checkpoint (tables inside) loads into a fresh module: buffers resize
    eb2 = _module(params)
    eb2.load_state_dict(eb.state_dict())
    assert torch.equal(eb2.quantized_cdf, q)


def test_strings_round_trip_on_the_host_coder(params):
    """compress() / decompress() semantics without a GPU: symbols -> one stream per image -> symbols, channel = CDF index."""
    from dcae_b200 import ans
    eb = _module(params)
    eb.update()
    g = torch.Generator().manual_seed(5)
    med = oeb.medians(params).reshape(1, C, 1, 1)
    z = med + 4.0 * torch.randn(2, C, 6, 8, generator=g)
    z[0, 3, 2, 2] += 500.0                                   # far outside the table: bypass path
    sym = oeb.symbols(params, z)
    idx = torch.arange(C, dtype=torch.int32).reshape(-1, 1).expand(C, 48).reshape(-1).numpy()
    q, ln, off = eb.quantized_cdf, eb.cdf_length, eb.offset
    for i in range(2):
        enc = ans.BufferedRansEncoder()
        enc.encode_with_indexes(sym[i].reshape(-1).numpy(), idx, q, ln, off)
        s = enc.flush()
        dec = ans.RansDecoder()
        dec.set_stream(s)
        assert np.array_equal(dec.decode_array(idx, q, ln, off), sym[i].reshape(-1).numpy())


@pytest.mark.gpu
def test_device_pass_matches_the_restatement(params):
    eb = _module(params, "cuda").eval()
    eb.update()
    g = torch.Generator().manual_seed(9)
    med = oeb.medians(params).reshape(1, C, 1, 1)
    z = med + 6.0 * torch.randn(3, C, 8, 12, generator=g) * torch.rand(1, C, 1, 1, generator=g)
    want_out, want_lik = oeb.forward(params, z)
    out, lik = eb(z.cuda())
    assert torch.equal(out.cpu(), want_out)
    # sigmoid(upper) - sigmoid(lower): a difference of two rounded values in (0, 1) (1 ulp = 6e-8 each) of logits whose own
    # rounding (|logit| up to ~20, matmul vs FMA order) moves a sigmoid by a few 1e-7: 1e-5 relative + 5e-7 absolute
    def close(got, want):
        err = (got.double() - want.double()).abs()
        print(f"\nEntropyBottleneck likelihood vs the restatement: max abs {float(err.max()):.2e}, max rel {float((err / want.double()).max()):.2e}")
        return bool((err <= 1e-5 * want.double() + 5e-7).all())
    assert close(lik.cpu(), want_lik)
    # training mode with a given noise tensor
    noise = torch.empty_like(z).uniform_(-0.5, 0.5, generator=g)
    out_n, lik_n = eb(z.cuda(), training=True, noise=noise.cuda())
    want_out_n, want_lik_n = oeb.forward(params, z, noise)
    assert torch.allclose(out_n.cpu(), want_out_n) and close(lik_n.cpu(), want_lik_n)
    # compress -> decompress: z_hat of dcae.py:706 equals the forward pass's quantised value, bit for bit
    strings = eb.compress(z.cuda())
    assert len(strings) == 3 and all(isinstance(s, bytes) for s in strings)
    z_hat = eb.decompress(strings, z.shape[-2:])
    assert torch.equal(z_hat, out)
    assert torch.equal(z_hat.cpu(), oeb.symbols(params, z).float() + med)
    bits = float(-torch.log2(lik.double()).sum())
    assert abs(sum(len(s) for s in strings) * 8 - bits) / bits < 0.05
    with pytest.raises(Exception):
        eb(z.cuda().requires_grad_())


@pytest.mark.gpu
def test_full_reference_model_with_both_entropy_models_on_the_library(lively_params):
    """The REAL `DCAE` class (oracle/_ref): `accelerate(net)` for the slice loop AND `net.entropy_bottleneck` =
    dcae_b200.EntropyBottleneck; `compress()` / `decompress()` as written with the native coder for both strings;
    x_hat equals the forward pass's reconstruction."""
    import os
    import tempfile
    from oracle.reference_loader import build_reference_net, load_reference_dcae_module, reference_available
    if not reference_available():
        pytest.skip("reference models/dcae.py not staged")
    from dcae_b200 import accelerate, ans
    from dcae_b200.entropy_bottleneck import EntropyBottleneck
    # the reference's evaluation flags (eval.py:3182-3187, 3904).  They are not decoration: compress() and decompress() each
    # run h_z_s1 / h_z_s2 through torch, and with cuDNN's heuristics on the two calls are not bit-reproducible -- one CDF
    # index that flips desynchronises the range decoder (observed here: x_hat garbage in one run out of three without them).
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    saved_cudnn = torch.backends.cudnn.enabled
    torch.backends.cudnn.enabled = False
    net = build_reference_net(lively_params).cuda().eval()
    eb = EntropyBottleneck(192)
    eb.load_state_dict(oeb.init_params(192, seed=4, trained_like=True), strict=False)
    eb.update()
    net.entropy_bottleneck = eb.cuda().eval()
    accelerate(net, device="cuda:0")
    ref = load_reference_dcae_module()
    saved = ref.BufferedRansEncoder, ref.RansDecoder
    ref.BufferedRansEncoder, ref.RansDecoder = ans.BufferedRansEncoder, ans.RansDecoder
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(8)).cuda()
    cwd = os.getcwd()
    try:
        with tempfile.TemporaryDirectory() as td, torch.no_grad():
            os.makedirs(os.path.join(td, "output", "debug"))
            os.chdir(td)
            enc = net.compress(x)
            dec = net.decompress(enc["strings"], enc["shape"])
            fwd = net(x)
    finally:
        os.chdir(cwd)
        ref.BufferedRansEncoder, ref.RansDecoder = saved
        torch.backends.cudnn.enabled = saved_cudnn
    assert isinstance(enc["strings"][1][0], bytes) and len(enc["strings"][1][0]) > 8
    diff = float((dec["x_hat"] - fwd["x_hat"].clamp(0, 1)).abs().max())
    print(f"\nfull model: |x_hat(decompress) - x_hat(forward)| max {diff:.2e}")
    assert diff < 1e-4              # pixels in [0, 1]; g_s (torch deconvolutions) is not bit-reproducible between calls
    assert bool(((fwd["likelihoods"]["z"] >= 1e-9) & (fwd["likelihoods"]["z"] <= 1)).all())

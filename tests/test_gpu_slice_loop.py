"""GPU parity of the whole channel-slice loop (C ABI: dcae_slice_loop_*) against
  (1) the golden outputs of the reference itself (tests/golden/*.npz, from /root/reference), and
  (2) the CPU oracle restatement on the same seeded inputs.
fp32-parity bar (north star): means / scales / y_hat / likelihoods within 1e-5 per-tensor relative
error; symbols / indexes are bit-exact GIVEN identical (y, mu, scale) -- checked by feeding the
device's own mu/scale to the oracle's quantiser -- and their end-to-end mismatch RATE against the
reference is reported and bounded (it cannot be exactly zero when accumulation order differs;
even the torch-CPU oracle shows 3e-5 against the torch-CPU reference)."""
import numpy as np
import pytest
import torch

from _util import GOLDEN_CASES, load_golden, rel_err, mismatch_rate
from oracle import gaussian_conditional as ogc
from oracle.entropy_model import SliceLoopOracle

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
# (per-slice teacher-forced rel tol, max teacher-forced sym/idx mismatch rate, max free-running mismatch rate)
# the rates are <= 5 x the worst value observed on B200 over the three golden cases (printed by the tests;
# profiles/r02/parity_rates.txt): a regression 10 x worse than today fails
MODES = {"fp32": (FP32_TOL, 5e-4, 3e-4), "tf32x3": (FP32_TOL, 7.5e-4, 5e-4), "f16x3": (FP32_TOL, 5e-4, 5e-4), "tf32": (1e-2, 5e-2, 0.3)}

_engines = {}


def engine(params, math):
    from dcae_b200.entropy_model import EntropySliceLoop
    if math not in _engines:
        _engines[math] = EntropySliceLoop(params, device="cuda:0", math=math)
    return _engines[math]


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3", "tf32"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_per_slice_parity_vs_reference_golden(case, math, lively_params):
    """Stage-wise parity, the only well-defined one: slice i is computed from the REFERENCE's
    previous slices (its recorded symbols are replayed through the decode path, exactly what
    DCAE.decompress does, dcae.py:893-896), so one flipped symbol cannot cascade into later slices.
    Per slice: mu, scale within the fp32 bar; indexes / symbols bit-exact except where the reference
    value sits within that tolerance of a rounding / table boundary (rate reported and bounded)."""
    from dcae_b200 import _lib
    g = load_golden(case)
    eng = engine(lively_params, math)
    tol, mm, _ = MODES[math]
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    B, _, h, w = y.shape
    lib, plan, s = eng.lib, eng._plan(B, h, w), _lib.current_stream(eng.device)
    _lib.check(lib.dcae_slice_loop_load(plan.handle, y.data_ptr(), ls.data_ptr(), lm.data_ptr(), s))
    idx = torch.empty(B, 64, h, w, dtype=torch.int32, device="cuda")
    worst = {"mu": 0.0, "scale": 0.0, "idx": 0.0, "sym": 0.0}
    tok2img = lambda t: t.reshape(B, h, w, -1).permute(0, 3, 1, 2)
    for i in range(5):
        sl = slice(64 * i, 64 * i + 64)
        _lib.check(lib.dcae_slice_loop_params(plan.handle, i, s))
        _lib.check(lib.dcae_slice_loop_indexes(plan.handle, i, idx.data_ptr(), s))
        mu = tok2img(eng.tap("means", B, h, w)[:, sl]).cpu()
        sc = tok2img(eng.tap("scales", B, h, w)[:, sl]).cpu()
        worst["mu"] = max(worst["mu"], rel_err(mu, g["means"][:, sl]))
        worst["scale"] = max(worst["scale"], rel_err(sc, g["scales"][:, sl]))
        worst["idx"] = max(worst["idx"], mismatch_rate(idx.cpu(), g["indexes"][i]))
        worst["sym"] = max(worst["sym"], mismatch_rate(ogc.quantize(g["y"][:, sl], "symbols", mu), g["symbols"][i]))
        # kernel 3 is bit-exact given identical inputs: the device's own scale through the oracle
        assert torch.equal(idx.cpu(), ogc.build_indexes(sc, ogc.get_scale_table()))
        sym_ref = g["symbols"][i].cuda().contiguous()
        _lib.check(lib.dcae_slice_loop_decode(plan.handle, i, sym_ref.data_ptr(), s))
    y_hat = torch.empty_like(y)
    _lib.check(lib.dcae_slice_loop_store(plan.handle, y_hat.data_ptr(), None, None, None, None, None, None, s))
    torch.cuda.synchronize()
    worst["y_hat"] = rel_err(y_hat.cpu(), g["y_hat"])
    print(f"\n[{case} {math}] per-slice (teacher-forced) worst: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    assert worst["mu"] < tol and worst["scale"] < tol and worst["y_hat"] < tol
    assert worst["idx"] <= mm and worst["sym"] <= mm


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3", "tf32"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_free_running_forward_and_compress(case, math, lively_params):
    """DCAE.forward / compress slice loop end to end (no teacher forcing).  A symbol that flips in slice i
    changes y_hat by 1.0 and every later slice with it, so only the mismatch RATE is meaningful here; it is
    reported and bounded.  Kernel 3 itself stays bit-exact on the device's own (y, mu, scale)."""
    g = load_golden(case)
    eng = engine(lively_params, math)
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    out = eng.compress(y, ls, lm, with_likelihoods=True)
    torch.cuda.synchronize()
    rates = {k: mismatch_rate(out[k].cpu(), g[k]) for k in ("symbols", "indexes")}
    errs = {k: rel_err(out[k].cpu(), g[k]) for k in ("means", "scales")}
    print(f"\n[{case} {math}] free-running: mismatch vs reference {rates}, rel_err {errs}, launches {eng.last_launches}")
    assert max(rates.values()) <= MODES[math][2]
    if math == "fp32":
        # observed: symbols identical to the reference; at most one index in 40 320 differs (a scale within
        # 2e-6 of a table entry) -- the torch-CPU oracle itself shows 3e-5 against the torch-CPU reference
        assert rates["symbols"] == 0.0 and rates["indexes"] <= 2e-4
        assert errs["means"] < FP32_TOL and errs["scales"] < FP32_TOL
        assert rel_err(out["y_hat"].cpu(), g["y_hat"]) < FP32_TOL
        # Likelihoods: mu/scale differ from the reference by ~3e-6 and d ln(lik)/d mu = v/sigma^2 is large in
        # the tails, so element-wise agreement is loose there by construction (kernel 3 itself is checked on
        # identical inputs in test_gpu_gc.py); the rate (sum of log-likelihoods = bpp numerator) must agree to 1e-5.
        lik, want = out["likelihoods"].cpu().double(), g["lik"].double()
        assert bool(((lik - want).abs() <= 5e-3 * want + 3e-7).all())
        assert float(((lik - want).abs() <= 1e-5 * want + 3e-7).double().mean()) > 0.97
        assert abs(float(torch.log2(lik).sum() - torch.log2(want).sum())) <= 1e-5 * abs(float(torch.log2(want).sum()))
    mu, sc = out["means"].cpu(), out["scales"].cpu()
    for i in range(5):
        sl = slice(64 * i, 64 * i + 64)
        assert torch.equal(out["symbols"][i].cpu(), ogc.quantize(g["y"][:, sl], "symbols", mu[:, sl]))
        assert torch.equal(out["indexes"][i].cpu(), ogc.build_indexes(sc[:, sl], ogc.get_scale_table()))


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3"])
def test_stagewise_taps_vs_oracle(math, lively_params):
    """Stage dumps in the style of the reference's debug_save (dcae_5_fixed.py:29-34): slice 0."""
    from oracle.entropy_model import dictionary_cross_attention, _sub
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, math)
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    B, _, h, w = y.shape
    lib, plan = eng.lib, eng._plan(B, h, w)
    from dcae_b200 import _lib
    s = _lib.current_stream(eng.device)
    _lib.check(lib.dcae_slice_loop_load(plan.handle, y.data_ptr(), ls.data_ptr(), lm.data_ptr(), s))
    _lib.check(lib.dcae_slice_loop_params(plan.handle, 0, s))
    torch.cuda.synchronize()
    taps = {}
    query = torch.cat([g["latent_scales"], g["latent_means"]], 1)
    dict_info = dictionary_cross_attention(query, lively_params["dt"], _sub(lively_params, "dt_cross_attention.0."), taps)
    tokm = lambda t: t.reshape(-1, t.shape[-1])
    for name in ("x0", "x1", "x2", "x3"):
        got = eng.tap16(name, B, h, w) if (math == "f16x3" and name == "x3") else eng.tap(name, B, h, w)
        e = rel_err(got.cpu(), tokm(taps[name]))
        print(f"[{math}] tap {name}: {e:.2e}")
        assert e < FP32_TOL
    # in f16x3 mode tensors that only feed GEMMs exist as fp16 hi/lo planes only
    tap = (lambda n: eng.tap16(n, B, h, w)) if math == "f16x3" else (lambda n: eng.tap(n, B, h, w))
    e = rel_err(tap("attn").cpu(), tokm(taps["attn"]))
    assert e < FP32_TOL
    got = tap("support")[:, :320].cpu()
    assert rel_err(got, dict_info.permute(0, 2, 3, 1).reshape(-1, 320)) < FP32_TOL


@pytest.mark.parametrize("case", ["slice_loop_b1_8x12", "slice_loop_b1_16x16"])
@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3"])
def test_decompress_reproduces_compress_bit_exactly(math, case, lively_params):
    """The codec property the reference fights for: the decoder regenerates the SAME indexes
    from its own scales and the same y_hat, bit for bit, on this device."""
    g = load_golden(case)
    eng = engine(lively_params, math)
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    enc = eng.compress(y, ls, lm)
    dec = eng.decompress(ls, lm, lambda i, idx: enc["symbols"][i])
    assert torch.equal(dec["indexes"], enc["indexes"])
    assert torch.equal(dec["y_hat"], enc["y_hat"])
    # and against the reference's own decompress() output (x_hat = clamp(y_hat, 0, 1) in the golden run)
    assert mismatch_rate(dec["indexes"].cpu(), g["dec_indexes"]) <= MODES[math][2]
    # teacher-forced decode of the REFERENCE's symbols: y_hat of the reference's decompress() (x_hat = clamp(y_hat, 0, 1))
    ref_sym = g["symbols"].cuda()
    forced = eng.decompress(ls, lm, lambda i, idx: ref_sym[i])
    err = float((forced["y_hat"].clamp(0, 1).cpu() - g["dec_y_hat"]).abs().max())
    assert err < FP32_TOL * float(g["y_hat"].abs().max())          # same absolute bar as the unclamped y_hat comparison


def test_batch_invariance_and_determinism(lively_params):
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, "f16x3")
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    a = eng.compress(y, ls, lm, with_likelihoods=True)
    a = {k: v.clone() for k, v in a.items()}
    b = eng.compress(y, ls, lm, with_likelihoods=True)
    one = eng.compress(y[1:], ls[1:], lm[1:], with_likelihoods=True)
    for k in ("means", "scales", "y_hat", "likelihoods"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a[k][1:], one[k]), k
    assert torch.equal(a["symbols"][:, 1:], one["symbols"]) and torch.equal(a["indexes"][:, 1:], one["indexes"])


@pytest.mark.parametrize("math", ["f16x3", "fp32"])
def test_lanes_do_not_change_any_output_bit(math, lively_params):
    """`forward` splits the batch over two plans / streams (lanes=2); every per-element output must be identical to
    the single-plan run, for an odd batch too, and the bpp numerator must agree to float rounding."""
    from dcae_b200 import EntropySliceLoop
    gen = torch.Generator().manual_seed(77)
    B, h, w = 3, 8, 12
    y = (4 * torch.randn(B, 320, h, w, generator=gen)).cuda()
    ls, lm = torch.randn(B, 320, h, w, generator=gen).cuda(), torch.randn(B, 320, h, w, generator=gen).cuda()
    one = EntropySliceLoop(lively_params, device="cuda:0", math=math, lanes=1)
    two = EntropySliceLoop(lively_params, device="cuda:0", math=math, lanes=2)
    two.LANE_MIN_TOKENS = 0           # lanes normally engage from 8 192 tokens per step; force them for this small case
    a = one.compress(y, ls, lm, with_likelihoods=True)
    b = two.compress(y, ls, lm, with_likelihoods=True)
    for k in ("means", "scales", "y_hat", "likelihoods", "symbols", "indexes"):
        assert torch.equal(a[k], b[k]), k
    assert abs(float(a["log2_lik_sum"]) - float(b["log2_lik_sum"])) <= 1e-5 * abs(float(a["log2_lik_sum"]))
    nz = (torch.rand(B, 320, h, w, generator=gen) - 0.5).cuda()
    a, b = one.forward(y, ls, lm, noise=nz), two.forward(y, ls, lm, noise=nz)
    for k in ("means", "scales", "y_hat", "likelihoods"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(one.tap("x2", B, h, w), two.tap("x2", B, h, w))


def test_packed_coder_handoff_equals_the_int32_tensors(lively_params):
    """one packed D2H (int16 symbols, uint8 indexes, coder order) carries exactly what the reference's
    per-slice `.tolist()` lists carry (dcae.py:742-743); a symbol outside int16 is detected and the int32 path is used."""
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, "f16x3")
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    enc = eng.compress(y, ls, lm)
    host = eng.compress_to_host(y, ls, lm)
    assert host["overflow"] == 0 and host["symbols"].dtype == torch.int16 and host["symbols"].is_pinned()
    # the reference's list order: slice-major, then reshape(-1) of [B, 64, h, w]
    want_sym = torch.cat([enc["symbols"][i].reshape(-1) for i in range(5)]).cpu()
    want_idx = torch.cat([enc["indexes"][i].reshape(-1) for i in range(5)]).cpu()
    assert torch.equal(host["symbols"].int(), want_sym) and torch.equal(host["indexes"].int(), want_idx)
    big = y.clone()
    big[0, 3, 2, 2] = 1.0e5                       # a symbol that does not fit int16: flagged, int32 returned
    host = eng.compress_to_host(big, ls, lm)
    enc = eng.compress(big, ls, lm)
    assert host["overflow"] >= 1 and host["symbols"].dtype == torch.int32
    assert torch.equal(host["symbols"], enc["symbols"].flatten().cpu())


def test_training_noise_forward_vs_oracle(lively_params):
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, "fp32")
    noise = torch.empty_like(g["y"]).uniform_(-0.5, 0.5, generator=torch.Generator().manual_seed(3))
    out = eng.forward(g["y"].cuda(), g["latent_scales"].cuda(), g["latent_means"].cuda(), noise=noise.cuda())
    y_hat, means, scales, lik = SliceLoopOracle(lively_params).forward(g["y"], g["latent_scales"], g["latent_means"], noise)
    assert rel_err(out["means"].cpu(), means) < FP32_TOL and rel_err(out["scales"].cpu(), scales) < FP32_TOL
    close = (out["likelihoods"].cpu() - lik).abs() <= 1e-3 * lik + 1e-9
    assert float(close.double().mean()) > 0.995   # a few elements sit next to a rounding boundary of an earlier slice


def test_bpp_reduction_matches_likelihood_tensor(lively_params):
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, "f16x3")
    out = eng.forward(g["y"].cuda(), g["latent_scales"].cuda(), g["latent_means"].cuda())
    want = torch.log2(out["likelihoods"].double()).sum()
    assert abs(float(out["log2_lik_sum"]) - float(want)) <= 1e-5 * abs(float(want))


def test_kodak_shape_properties(lively_params):
    """BASELINE config #2 shape at B=2 (768x512 -> 32x48 tokens): size-independent properties only."""
    eng = engine(lively_params, "f16x3")
    gen = torch.Generator().manual_seed(1234)
    y = (4 * torch.randn(2, 320, 32, 48, generator=gen)).cuda()
    ls = torch.randn(2, 320, 32, 48, generator=gen).cuda()
    lm = torch.randn(2, 320, 32, 48, generator=gen).cuda()
    enc = eng.compress(y, ls, lm, with_likelihoods=True)
    assert bool(torch.isfinite(enc["means"]).all()) and bool(torch.isfinite(enc["scales"]).all())
    assert bool(((enc["likelihoods"] >= 1e-9) & (enc["likelihoods"] <= 1)).all())
    assert int(enc["indexes"].min()) >= 0 and int(enc["indexes"].max()) <= 63
    # encode -> decode round trip
    dec = eng.decompress(ls, lm, lambda i, idx: enc["symbols"][i])
    assert torch.equal(dec["indexes"], enc["indexes"]) and torch.equal(dec["y_hat"], enc["y_hat"])
    # symbols are exactly round(y - mu) for the device's own means
    sym = torch.stack([torch.round(y[:, 64 * i:64 * i + 64] - enc["means"][:, 64 * i:64 * i + 64]).int() for i in range(5)])
    assert torch.equal(sym, enc["symbols"])
    # fp32 SIMT and tf32x3 agree to fp32 accuracy on a full-size problem (slice 0: nothing has cascaded yet)
    ref = engine(lively_params, "fp32").compress(y, ls, lm)
    assert rel_err(enc["means"][:, :64], ref["means"][:, :64]) < FP32_TOL
    assert rel_err(enc["scales"][:, :64], ref["scales"][:, :64]) < FP32_TOL


def test_input_validation(lively_params):
    eng = engine(lively_params, "fp32")
    y = torch.zeros(1, 320, 4, 4).cuda()
    with pytest.raises(ValueError):
        eng.forward(y.cpu(), y, y)
    with pytest.raises(ValueError):
        eng.forward(y[:, :64], y, y)
    with pytest.raises(ValueError):
        eng.forward(y, y[:, :, :2], y)


def test_host_pipeline_matches_direct_calls(lively_params):
    """dcae_b200.HostPipeline (pinned host in/out, overlapped copies) returns exactly what forward() returns."""
    from dcae_b200.pipeline import HostPipeline
    eng = engine(lively_params, "f16x3")
    B, h, w = 2, 7, 9
    gen = torch.Generator().manual_seed(5)
    batches = [[(4 * torch.randn(B, 320, h, w, generator=gen)).pin_memory(), torch.randn(B, 320, h, w, generator=gen).pin_memory(),
                torch.randn(B, 320, h, w, generator=gen).pin_memory()] for _ in range(5)]
    pipe = HostPipeline(eng, B, h, w, depth=2, want_symbols=True)
    got = [{k: v.clone() for k, v in res.items()} for res in pipe.run(batches)]
    assert len(got) == 5
    for b, res in zip(batches, got):
        want = eng.forward(*[t.cuda() for t in b], want_symbols=True)
        for k in ("y_hat", "means", "scales", "likelihoods", "symbols", "indexes"):
            assert torch.equal(res[k], want[k].cpu()), k


@pytest.mark.parametrize("name,B,h,w", [("config1_256x256", 1, 16, 16), ("config3_clic_2048x1408", 1, 88, 128),
                                        ("config5_4k_3840x2176", 1, 136, 240)])
def test_baseline_config_shapes(name, B, h, w, lively_params):
    """BASELINE.json configs #1 / #3 / #5 at their full latent sizes (one image): slice 0 against the CPU oracle
    (nothing has cascaded yet), then the size-independent properties of the whole loop."""
    eng = engine(lively_params, "f16x3")
    gen = torch.Generator().manual_seed(1234)
    y = 4 * torch.randn(B, 320, h, w, generator=gen)
    ls = torch.randn(B, 320, h, w, generator=gen)
    lm = torch.randn(B, 320, h, w, generator=gen)
    yc, lsc, lmc = y.cuda(), ls.cuda(), lm.cuda()
    enc = eng.compress(yc, lsc, lmc, with_likelihoods=True)
    _, mu0, sc0 = SliceLoopOracle(lively_params).slice_params(0, ls, lm, [])
    e_mu, e_sc = rel_err(enc["means"][:, :64].cpu(), mu0), rel_err(enc["scales"][:, :64].cpu(), sc0)
    print(f"\n[{name}] T={B * h * w}: slice-0 rel_err mu {e_mu:.2e} scale {e_sc:.2e}, launches {eng.last_launches}")
    assert e_mu < FP32_TOL and e_sc < FP32_TOL
    assert mismatch_rate(enc["symbols"][0].cpu(), ogc.quantize(y[:, :64], "symbols", mu0)) <= 1e-3
    assert bool(torch.isfinite(enc["y_hat"]).all())
    assert bool(((enc["likelihoods"] >= 1e-9) & (enc["likelihoods"] <= 1)).all())
    assert int(enc["indexes"].min()) >= 0 and int(enc["indexes"].max()) <= 63
    sym = torch.stack([torch.round(yc[:, 64 * i:64 * i + 64] - enc["means"][:, 64 * i:64 * i + 64]).int() for i in range(5)])
    assert torch.equal(sym, enc["symbols"])
    dec = eng.decompress(lsc, lmc, lambda i, idx: enc["symbols"][i])
    assert torch.equal(dec["indexes"], enc["indexes"]) and torch.equal(dec["y_hat"], enc["y_hat"])


@pytest.mark.parametrize("B,h,w", [(1, 5, 17), (3, 13, 6), (2, 1, 1), (1, 24, 40), (5, 9, 9)])
@pytest.mark.parametrize("math", ["f16x3", "tf32x3"])
def test_ragged_shapes_against_the_oracle(B, h, w, math, lively_params):
    """Token grids that are not multiples of the 8x16 / 128-token tiles, odd batches (lanes of unequal size), a single
    token: slice 0 against the CPU oracle (nothing has cascaded yet) + round trip of the whole loop."""
    eng = engine(lively_params, math)
    gen = torch.Generator().manual_seed(100 * B + 10 * h + w)
    y = 4 * torch.randn(B, 320, h, w, generator=gen)
    ls, lm = torch.randn(B, 320, h, w, generator=gen), torch.randn(B, 320, h, w, generator=gen)
    enc = eng.compress(y.cuda(), ls.cuda(), lm.cuda(), with_likelihoods=True)
    _, mu0, sc0 = SliceLoopOracle(lively_params).slice_params(0, ls, lm, [])
    assert rel_err(enc["means"][:, :64].cpu(), mu0) < FP32_TOL and rel_err(enc["scales"][:, :64].cpu(), sc0) < FP32_TOL
    assert bool(torch.isfinite(enc["y_hat"]).all()) and bool(((enc["likelihoods"] >= 1e-9) & (enc["likelihoods"] <= 1)).all())
    dec = eng.decompress(ls.cuda(), lm.cuda(), lambda i, idx: enc["symbols"][i])
    assert torch.equal(dec["indexes"], enc["indexes"]) and torch.equal(dec["y_hat"], enc["y_hat"])


def test_cuda_graph_replay_equals_stream_launches(lively_params):
    """`capture()`: the whole step (both lanes, the side stream, 173 launches) as one CUDA graph; inputs by address."""
    eng = engine(lively_params, "f16x3")
    gen = torch.Generator().manual_seed(9)
    y = (4 * torch.randn(2, 320, 16, 16, generator=gen)).cuda()
    ls, lm = torch.randn(2, 320, 16, 16, generator=gen).cuda(), torch.randn(2, 320, 16, 16, generator=gen).cuda()
    replay, out = eng.capture(y, ls, lm, want_symbols=True)
    want = {k: v.clone() for k, v in eng.forward(y, ls, lm, want_symbols=True).items()}
    for v in out.values():
        v.zero_()
    replay()
    torch.cuda.synchronize()
    for k in ("y_hat", "means", "scales", "likelihoods", "symbols", "indexes"):
        assert torch.equal(out[k], want[k]), k
    y.mul_(0.5)                                   # new data at the same addresses
    replay()
    torch.cuda.synchronize()
    again = eng.forward(y, ls, lm, want_symbols=True)
    assert torch.equal(out["symbols"], again["symbols"]) and torch.equal(out["y_hat"], again["y_hat"])


@pytest.mark.parametrize("name,B,h,w", [("config2_kodak_768x512", 2, 32, 48), ("config3_clic_2048x1408", 1, 88, 128),
                                        ("config5_4k_3840x2176", 1, 136, 240)])
@pytest.mark.parametrize("math", ["f16x3"])
def test_full_depth_parity_at_baseline_sizes(name, B, h, w, math, lively_params):
    """All FIVE slices at the BASELINE.json latent sizes against the CPU oracle, teacher-forced: slice i is computed
    from the oracle's own symbols of slices < i (replayed through the decode path, what DCAE.decompress does), so
    mu / scale of every slice -- the long-K cc1 of slices 1-4, the LRP chain, the y_hat planes feeding later slices,
    multi-wave persistent schedules -- are compared element for element at <= 1e-5, and symbols / indexes by rate."""
    from dcae_b200 import _lib
    eng = engine(lively_params, math)
    gen = torch.Generator().manual_seed(4321)
    y = 4 * torch.randn(B, 320, h, w, generator=gen)
    ls, lm = torch.randn(B, 320, h, w, generator=gen), torch.randn(B, 320, h, w, generator=gen)
    o_sym, o_idx, o_yhat, o_mu, o_sc = SliceLoopOracle(lively_params).compress(y, ls, lm)
    yc, lsc, lmc = y.cuda(), ls.cuda(), lm.cuda()
    lib, plan, s = eng.lib, eng._plan(B, h, w), _lib.current_stream(eng.device)
    _lib.check(lib.dcae_slice_loop_load(plan.handle, yc.data_ptr(), lsc.data_ptr(), lmc.data_ptr(), s))
    idx = torch.empty(B, 64, h, w, dtype=torch.int32, device="cuda")
    tok2img = lambda t: t.reshape(B, h, w, -1).permute(0, 3, 1, 2)
    worst = {"mu": 0.0, "scale": 0.0, "idx": 0.0, "sym": 0.0}
    for i in range(5):
        sl = slice(64 * i, 64 * i + 64)
        _lib.check(lib.dcae_slice_loop_params(plan.handle, i, s))
        _lib.check(lib.dcae_slice_loop_indexes(plan.handle, i, idx.data_ptr(), s))
        mu = tok2img(eng.tap("means", B, h, w)[:, sl]).cpu()
        sc = tok2img(eng.tap("scales", B, h, w)[:, sl]).cpu()
        e_mu, e_sc = rel_err(mu, o_mu[:, sl]), rel_err(sc, o_sc[:, sl])
        assert e_mu < FP32_TOL and e_sc < FP32_TOL, (i, e_mu, e_sc)
        worst["mu"], worst["scale"] = max(worst["mu"], e_mu), max(worst["scale"], e_sc)
        worst["idx"] = max(worst["idx"], mismatch_rate(idx.cpu(), o_idx[i]))
        worst["sym"] = max(worst["sym"], mismatch_rate(ogc.quantize(y[:, sl], "symbols", mu), o_sym[i]))
        assert torch.equal(idx.cpu(), ogc.build_indexes(sc, ogc.get_scale_table()))      # kernel 3 on the device's own scale
        sym = o_sym[i].cuda().contiguous()
        _lib.check(lib.dcae_slice_loop_decode(plan.handle, i, sym.data_ptr(), s))
    y_hat = torch.empty_like(yc)
    _lib.check(lib.dcae_slice_loop_store(plan.handle, y_hat.data_ptr(), None, None, None, None, None, None, s))
    torch.cuda.synchronize()
    worst["y_hat"] = rel_err(y_hat.cpu(), o_yhat)
    print(f"\n[{name} {math}] T={B * h * w} teacher-forced, all 5 slices: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    assert worst["y_hat"] < FP32_TOL
    assert worst["idx"] <= MODES[math][1] and worst["sym"] <= MODES[math][1]


def test_range_coder_round_trip_through_the_slice_loop(lively_params):
    """SURVEY 8f N1 + N2: compress_to_string -> decompress_from_string on the device reproduces y_hat bit for bit, the
    stream decodes (pure-Python restatement of the coder) to exactly the symbols the loop emitted, and its length is the
    likelihoods' bit count to a fraction of a percent (the tables quantise the same Gaussians)."""
    from dcae_b200.gaussian_conditional import GaussianConditional
    from oracle import rans as orans
    g = load_golden("slice_loop_b2_7x9")
    eng = engine(lively_params, "f16x3")
    gcm = GaussianConditional(None).cuda()
    gcm.update_scale_table(ogc.get_scale_table())
    y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
    enc = eng.compress(y, ls, lm, with_likelihoods=True)
    out = eng.compress_to_string(y, ls, lm, gcm)
    assert isinstance(out["y_string"], bytes) and out["overflow"] == 0
    dec = eng.decompress_from_string(out["y_string"], ls, lm, gcm)
    assert torch.equal(dec["y_hat"], enc["y_hat"]) and torch.equal(dec["indexes"], enc["indexes"])
    q, ln, off = gcm.quantized_cdf.cpu().tolist(), gcm.cdf_length.cpu().tolist(), gcm.offset.cpu().tolist()
    sym = orans.Decoder(out["y_string"]).decode(enc["indexes"].flatten().cpu().tolist(), q, ln, off)
    assert sym == enc["symbols"].flatten().cpu().tolist()
    # stream length = the tables' code length of exactly these symbols (in-range: -log2 freq / 2^16; out of range: the
    # sentinel + 4 bits per bypass digit incl. the unary digit count), to the coder's ~0.01 % overhead + the 8-byte flush
    qn, lnn, offn = gcm.quantized_cdf.cpu().numpy(), gcm.cdf_length.cpu().numpy(), gcm.offset.cpu().numpy()
    sy, ix = enc["symbols"].flatten().cpu().numpy().astype("int64"), enc["indexes"].flatten().cpu().numpy()
    v, sent = sy - offn[ix], lnn[ix] - 2
    inside = (v >= 0) & (v < sent)
    vv = np.where(inside, v, sent)
    bits = -np.log2((qn[ix, vv + 1] - qn[ix, vv]) / 65536.0)
    raw = np.where(v < 0, -2 * v - 1, 2 * (v - sent))[~inside]
    digits = np.where(raw > 0, np.floor(np.log2(np.maximum(raw, 1)) / 4).astype("int64") + 1, 0)
    bits_total = float(bits.sum() + 4.0 * (digits + digits // 15 + 1).sum())
    got = len(out["y_string"]) * 8
    print(f"\nstream {got} bits, table code length {bits_total:.0f} bits, {int((~inside).sum())} bypassed symbols of {sy.size}")
    assert abs(got - bits_total) <= 1e-3 * bits_total + 64


def test_fast_math_mode_f16_measured_tolerance(lively_params):
    """`math="f16"`: the reduced-precision fast mode BASELINE.md promises next to the parity mode -- the dense layers
    multiply the fp16 hi planes only (11-bit operands, fp32 accumulate).  NOT a parity mode: what is asserted is the
    tolerance it is documented with: per-slice teacher-forced mu / scale within 3e-3 of the reference golden,
    symbol / index mismatch rates within the stated bounds, and the codec property (decompress reproduces compress bit
    for bit on this device) that any mode must keep."""
    from dcae_b200 import _lib
    eng = engine(lively_params, "f16")
    for case in GOLDEN_CASES:
        g = load_golden(case)
        y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
        B, _, h, w = y.shape
        lib, plan, s = eng.lib, eng._plan(B, h, w), _lib.current_stream(eng.device)
        _lib.check(lib.dcae_slice_loop_load(plan.handle, y.data_ptr(), ls.data_ptr(), lm.data_ptr(), s))
        idx = torch.empty(B, 64, h, w, dtype=torch.int32, device="cuda")
        tok2img = lambda t: t.reshape(B, h, w, -1).permute(0, 3, 1, 2)
        worst = {"mu": 0.0, "scale": 0.0, "idx": 0.0, "sym": 0.0}
        for i in range(5):
            sl = slice(64 * i, 64 * i + 64)
            _lib.check(lib.dcae_slice_loop_params(plan.handle, i, s))
            _lib.check(lib.dcae_slice_loop_indexes(plan.handle, i, idx.data_ptr(), s))
            mu = tok2img(eng.tap("means", B, h, w)[:, sl]).cpu()
            sc = tok2img(eng.tap("scales", B, h, w)[:, sl]).cpu()
            worst["mu"] = max(worst["mu"], rel_err(mu, g["means"][:, sl]))
            worst["scale"] = max(worst["scale"], rel_err(sc, g["scales"][:, sl]))
            worst["idx"] = max(worst["idx"], mismatch_rate(idx.cpu(), g["indexes"][i]))
            worst["sym"] = max(worst["sym"], mismatch_rate(ogc.quantize(g["y"][:, sl], "symbols", mu), g["symbols"][i]))
            sym_ref = g["symbols"][i].cuda().contiguous()
            _lib.check(lib.dcae_slice_loop_decode(plan.handle, i, sym_ref.data_ptr(), s))
        torch.cuda.synchronize()
        print(f"\n[{case} f16 single pass] teacher-forced worst: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
        assert worst["mu"] < 3e-3 and worst["scale"] < 3e-3
        assert worst["sym"] < 2e-2 and worst["idx"] < 8e-2
        enc = eng.compress(y, ls, lm)
        dec = eng.decompress(ls, lm, lambda i, idx: enc["symbols"][i])
        assert torch.equal(dec["indexes"], enc["indexes"]) and torch.equal(dec["y_hat"], enc["y_hat"])


def test_f16x3_parity_holds_for_large_activations(lively_params):
    """Advisor r1: the hi/lo fp16 planes were only exercised with O(1) activations.  Latents of magnitude ~50 and an
    8 x wider x_trans push the pre-LayerNorm activations, the dense concat and the fc1 outputs into the hundreds to
    thousands: slice 0 must still meet the fp32 bar against the CPU oracle, and the range check must report no clamp."""
    from dcae_b200.entropy_model import EntropySliceLoop
    p = {k: v.clone() for k, v in lively_params.items()}
    p["dt_cross_attention.0.x_trans.weight"] *= 32.0
    p["dt_cross_attention.0.mlp.fc1.weight"] *= 4.0
    eng = EntropySliceLoop(p, device="cuda:0", math="f16x3")
    gen = torch.Generator().manual_seed(314)
    B, h, w = 1, 16, 24
    y = 4 * torch.randn(B, 320, h, w, generator=gen)
    ls, lm = 50 * torch.randn(B, 320, h, w, generator=gen), 50 * torch.randn(B, 320, h, w, generator=gen)
    assert eng.check_f16_range(y.cuda(), ls.cuda(), lm.cuda()) == 0
    enc = eng.compress(y.cuda(), ls.cuda(), lm.cuda())
    _, mu0, sc0 = SliceLoopOracle(p).slice_params(0, ls, lm, [])
    q0 = torch.cat([ls, lm], 1).permute(0, 2, 3, 1)            # slice 0's x_trans output, the widest pre-LayerNorm activation
    big = float(torch.nn.functional.linear(q0, p["dt_cross_attention.0.x_trans.weight"], p["dt_cross_attention.0.x_trans.bias"]).abs().max())
    e_mu, e_sc = rel_err(enc["means"][:, :64].cpu(), mu0), rel_err(enc["scales"][:, :64].cpu(), sc0)
    print(f"\nlarge activations: max |x_trans output| {big:.0f}, slice-0 rel_err mu {e_mu:.2e} scale {e_sc:.2e}")
    assert big > 300, big
    assert e_mu < FP32_TOL and e_sc < FP32_TOL, (e_mu, e_sc)
    with pytest.raises(Exception):          # and beyond the fp16 range the first call raises instead of clamping
        EntropySliceLoop(p, device="cuda:0", math="f16x3").forward(y.cuda(), (ls * 1e5).cuda(), lm.cuda())

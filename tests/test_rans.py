"""CPU: the native range coder (libdcae_rans.so, include/dcae_rans.h) against the pure-Python
restatement of the published algorithm (oracle/rans.py), the coder contract of dcae.py:755-756 / :893, and the CDF
tables it consumes."""
import math

import numpy as np
import pytest
import torch

from dcae_b200 import ans
from dcae_b200.gaussian_conditional import GaussianConditional
from oracle import gaussian_conditional as ogc
from oracle import rans as orans


@pytest.fixture(scope="module")
def tables():
    import __graft_entry__ as ge
    ge.build()
    gc = GaussianConditional(None)
    gc.update_scale_table(ogc.get_scale_table())
    return gc.quantized_cdf, gc.cdf_length, gc.offset


def _lists(tables):
    q, ln, off = tables
    return q.tolist(), ln.reshape(-1).int().tolist(), off.reshape(-1).int().tolist()       # dcae.py:718-720


def _draw(n, seed, tables, outliers=0.02):
    """Symbols the way the slice loop produces them: round(N(0, s)) for the table entry of the index; a few far
    outliers on both sides so that the bypass path (sentinel + 4-bit digits) is exercised."""
    g = np.random.default_rng(seed)
    idx = g.integers(0, 64, n).astype(np.int32)
    s = ogc.get_scale_table().numpy()[idx]
    sym = np.rint(g.standard_normal(n) * s).astype(np.int32)
    far = g.random(n) < outliers
    sym[far] = (g.integers(-40000, 40000, far.sum())).astype(np.int32)
    return sym, idx


def test_native_stream_is_byte_identical_to_the_restatement(tables):
    cdf, ln, off = _lists(tables)
    for seed, n in ((0, 1), (1, 7), (2, 300), (3, 2500)):
        sym, idx = _draw(n, seed, tables)
        enc = ans.BufferedRansEncoder()
        enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdf, ln, off)            # Python lists, like the reference
        got = enc.flush()
        want = orans.encode(sym.tolist(), idx.tolist(), cdf, ln, off)
        assert got == want, (seed, n, len(got), len(want))
        dec = ans.RansDecoder()
        dec.set_stream(got)
        assert dec.decode_stream(idx.tolist(), cdf, ln, off) == sym.tolist()
        assert orans.Decoder(got).decode(idx.tolist(), cdf, ln, off) == sym.tolist()


def test_empty_and_split_calls(tables):
    cdf, ln, off = _lists(tables)
    enc = ans.BufferedRansEncoder()
    assert len(enc.flush()) == 8                                                     # just the final state
    sym, idx = _draw(1000, 5, tables)
    a, b = ans.BufferedRansEncoder(), ans.BufferedRansEncoder()
    a.encode_with_indexes(sym, idx, cdf, ln, off)
    for lo in range(0, 1000, 64):                                                    # per-slice extend() of dcae.py:742-743
        b.encode_with_indexes(sym[lo:lo + 64], idx[lo:lo + 64], cdf, ln, off)
    stream = a.flush()
    assert stream == b.flush()
    dec = ans.RansDecoder()
    dec.set_stream(stream)
    out = [dec.decode_array(idx[lo:lo + 100], cdf, ln, off) for lo in range(0, 1000, 100)]   # decode_stream per slice, :893
    assert np.array_equal(np.concatenate(out), sym)


def test_packed_device_formats_code_to_the_same_bytes(tables):
    """int16 symbols / uint8 indexes (what dcae_pack_symbols sends to the host) and the tensors themselves as tables."""
    q, ln, off = tables
    sym, idx = _draw(50_000, 7, tables, outliers=0.0)
    e32, e16 = ans.BufferedRansEncoder(), ans.BufferedRansEncoder()
    e32.encode_with_indexes(sym, idx, q, ln, off)
    e16.encode_with_indexes(torch.from_numpy(sym.astype(np.int16)), torch.from_numpy(idx.astype(np.uint8)), q, ln, off)
    s32 = e32.flush()
    assert s32 == e16.flush()
    dec = ans.RansDecoder()
    dec.set_stream(s32)
    assert np.array_equal(dec.decode_array(idx.astype(np.uint8), q, ln, off), sym)


def test_round_trip_at_clic_size_and_rate_matches_the_entropy(tables):
    """2048x1408 image = 3.6 M symbols: round trip, and the stream length is the table's cross-entropy to ~0.1 %."""
    q, ln, off = tables
    n = 320 * 88 * 128
    sym, idx = _draw(n, 11, tables, outliers=0.0)
    enc = ans.BufferedRansEncoder()
    enc.encode_with_indexes(sym, idx, q, ln, off)
    stream = enc.flush()
    dec = ans.RansDecoder()
    dec.set_stream(stream)
    assert np.array_equal(dec.decode_array(idx, q, ln, off), sym)
    qn, offn, lnn = q.numpy(), off.numpy(), ln.numpy()
    v = np.clip(sym - offn[idx], 0, lnn[idx] - 2)
    freq = qn[idx, v + 1] - qn[idx, v]
    bits = float(-np.log2(freq / 65536.0).sum())
    inside = (sym - offn[idx] >= 0) & (sym - offn[idx] < lnn[idx] - 2)
    assert inside.mean() > 0.9999
    assert abs(len(stream) * 8 - bits) / bits < 2e-3


def test_errors_are_reported_not_crashes(tables):
    cdf, ln, off = _lists(tables)
    enc = ans.BufferedRansEncoder()
    with pytest.raises(ans.RansError):
        enc.encode_with_indexes([0], [64], cdf, ln, off)                             # index outside the table
    with pytest.raises(ans.RansError):
        enc.encode_with_indexes([0, 1], [0], cdf, ln, off)
    dec = ans.RansDecoder()
    with pytest.raises(ans.RansError):
        dec.decode_stream([0], cdf, ln, off)                                         # no stream
    with pytest.raises(ans.RansError):
        dec.set_stream(b"123")
    dec.set_stream(b"\x00" * 8)                                                      # garbage decodes to something or errors, never crashes
    try:
        dec.decode_stream([5] * 100, cdf, ln, off)
    except ans.RansError:
        pass


# ---- G6: the tables ------------------------------------------------------------------------------------------------
def test_pmf_to_quantized_cdf_known_answers():
    """Hand-computed cases of the published algorithm (scale to 2^16, renormalise, partial sums, steal for zero bins)."""
    assert ans.pmf_to_quantized_cdf([0.5, 0.25, 0.25]) == [0, 32768, 49152, 65536]
    assert ans.pmf_to_quantized_cdf([1.0]) == [0, 65536]
    # 0.75 / 0.25 / 0: the zero bin takes one count from the cheapest donor with freq > 1 (bin 1), entries in between shift
    assert ans.pmf_to_quantized_cdf([0.75, 0.25, 0.0]) == [0, 49152, 65535, 65536]
    # zero bin in front of its donor
    assert ans.pmf_to_quantized_cdf([0.0, 0.25, 0.75]) == [0, 1, 16384, 65536]
    # un-normalised input is renormalised with floor division; the last entry is forced to 2^16
    assert ans.pmf_to_quantized_cdf([1.0, 1.0, 1.0]) == [0, 21845, 43690, 65536]
    # std::round is half away from zero: 0.5 + 2^-17 -> 32768.5 -> 32769, total 65538, floor(65536 * 32769 / 65538) = 32768
    assert ans.pmf_to_quantized_cdf([0.5 + 2 ** -17, 0.5 + 2 ** -17]) == [0, 32768, 65536]
    with pytest.raises(ans.RansError):
        ans.pmf_to_quantized_cdf([0.5, float("nan")])
    with pytest.raises(ans.RansError):
        ans.pmf_to_quantized_cdf([0.0, 0.0])


def test_native_pmf_to_quantized_cdf_equals_the_restatement_on_random_pmfs():
    g = np.random.default_rng(0)
    for n in (2, 3, 17, 200, 3131):
        p = g.random(n).astype(np.float32) ** 8                    # many near-zero bins -> many steals
        p /= p.sum()
        assert ans.pmf_to_quantized_cdf(p.tolist()) == ogc.pmf_to_quantized_cdf(p.tolist())


def test_tables_follow_the_gaussian_they_quantise(tables):
    """Implementation-independent check of update(): row i is N(0, table_i) integrated over unit bins, 16-bit
    quantised -- bins of 16 counts or more are within 2.7 counts + 0.2 % of 65536 * pmf, every symbol within
    +-ceil(6.1094 s) is codable, and offset / length follow the tail mass 1e-9 (dcae.py:616-621)."""
    q, ln, off = (t.numpy() for t in tables)
    table = ogc.get_scale_table().double().numpy()
    mult = 6.109410204869      # -Phi^-1(1e-9 / 2)
    assert q.shape == (64, 3133)
    for i in (0, 1, 7, 20, 33, 47, 63):
        c = math.ceil(float(np.float32(table[i])) * np.float32(mult))
        assert off[i] == -c and ln[i] == 2 * c + 1 + 2
        row = q[i, : ln[i]]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) > 0).all()
        k = np.arange(-c, c + 1, dtype=np.float64)
        s = table[i]
        pmf = np.array([0.5 * (math.erf((x + 0.5) / (s * math.sqrt(2))) - math.erf((x - 0.5) / (s * math.sqrt(2)))) for x in k])
        freq = np.diff(row)[:-1].astype(np.float64)              # last entry is the tail / sentinel bin
        ideal = 65536 * pmf
        # every symbol codable; nothing gains more than rounding; bins that round to zero are paid for by the CHEAPEST
        # donors (the published steal rule), which drives bins below ~11 counts down to 1 -- larger bins keep their mass
        assert (freq >= 1).all() and (freq <= ideal + 1.5).all()
        big = ideal >= 16
        assert np.abs(freq - ideal)[big].max() <= 2.7 + 2e-3 * ideal[big].max()
        assert (freq[~big] >= np.minimum(ideal[~big], 1.0) - 1e-9).all()

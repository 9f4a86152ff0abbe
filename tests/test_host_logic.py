"""CPU: host-side logic -- parameter inventory, weight repacking rules, CDF tables."""
import torch
import torch.nn.functional as F

from dcae_b200 import params as P
from dcae_b200.weights import PackedWeights
from dcae_b200.gaussian_conditional import GaussianConditional, _pmf_to_quantized_cdf
from oracle import gaussian_conditional as ogc


def test_param_inventory():
    shapes = P.entropy_param_shapes()
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert 83_000_000 < n < 83_500_000           # hot-path 83.2 M parameters
    assert shapes["dt"] == (128, 640)
    assert shapes["dt_cross_attention.3.x_trans.weight"] == (640, 832)
    assert shapes["cc_mean_transforms.4.0.weight"] == (224, 1216, 3, 3)
    assert shapes["lrp_transforms.0.0.weight"] == (224, 1024, 3, 3)
    a, b = P.init_entropy_params(3), P.init_entropy_params(3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["dt"], P.init_entropy_params(4)["dt"])


def test_conv3x3_repack_is_an_implicit_gemm_with_permuted_support():
    """[N, C, 3, 3] -> [N, 9*C'] tap-major with the library's support order must reproduce F.conv2d."""
    i = 1
    C = P.cs(i)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, C, 3, 3, generator=g)
    x_ref = torch.randn(1, C, 5, 6, generator=g)                      # reference order [ls, lm, y_hat0, dict_info]
    perm = PackedWeights._support_perm(i)
    x_lib = x_ref[:, perm]                                            # library order [dict_info, ls, lm, y_hat0]
    w2d = PackedWeights._conv3x3_to_gemm(w, perm)
    cols = F.unfold(x_lib, 3, padding=1)                              # [1, C*9, L] channel-major, tap-minor
    cols = cols.reshape(1, C, 9, -1).permute(0, 2, 1, 3).reshape(1, 9 * C, -1)   # -> tap-major
    got = (w2d @ cols[0]).reshape(1, 8, 5, 6)
    assert torch.allclose(got, F.conv2d(x_ref, w, padding=1), atol=1e-4)
    assert perm[:320].tolist() == list(range(P.cq(i), P.cq(i) + 320))


def test_cdf_tables_match_oracle_restatement():
    gc = GaussianConditional(None)
    assert gc.update_scale_table(ogc.get_scale_table()) is True
    assert gc.update_scale_table(ogc.get_scale_table()) is False          # already initialised, not forced
    q, off, ln = ogc.build_cdf_tables(ogc.get_scale_table())
    assert torch.equal(gc.quantized_cdf, q) and torch.equal(gc.offset, off) and torch.equal(gc.cdf_length, ln)
    assert tuple(q.shape) == (64, 3133) and int(ln.min()) == 5
    for i in range(64):
        row = q[i, : int(ln[i])]
        assert int(row[0]) == 0 and int(row[-1]) == 65536 and bool((row.diff() > 0).all())
    sd = gc.state_dict()
    assert set(sd) == {"_offset", "_quantized_cdf", "_cdf_length", "scale_table"}      # dcae.py:680-685


def test_pmf_to_quantized_cdf_steals_for_zero_bins():
    cdf = _pmf_to_quantized_cdf([0.5, 1e-12, 0.5, 1e-12])
    assert cdf[0] == 0 and cdf[-1] == 65536
    assert all(b > a for a, b in zip(cdf, cdf[1:]))

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
    # the four the reference resizes before loading (dcae.py:680-685) + the three scalar buffers of compressai's class
    assert set(sd) == {"_offset", "_quantized_cdf", "_cdf_length", "scale_table", "scale_bound",
                       "lower_bound_scale.bound", "likelihood_lower_bound.bound"}


def test_state_dict_round_trip_resizes_the_table_buffers():
    """`strict=False` does not excuse size mismatches; a fresh module (buffers of shape [0]) must load an
    updated module's tables, a baked checkpoint in the reference layout (export_checkpoint.py:34-42), and shrink again."""
    src = GaussianConditional(None)
    src.update_scale_table(ogc.get_scale_table())
    dst = GaussianConditional(None)
    assert dst._offset.numel() == 0
    dst.load_state_dict(src.state_dict(), strict=False)
    for k in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table"):
        assert torch.equal(getattr(dst, k), getattr(src, k)), k
    assert dst.update_scale_table(ogc.get_scale_table()) is False            # already initialised by the load
    # reference checkpoint layout: keys prefixed inside a parent module, only the four table buffers present
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gaussian_conditional = GaussianConditional(None)
    ckpt = {"gaussian_conditional." + k: v.clone() for k, v in src.state_dict().items()
            if k in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table")}
    net = Net()
    missing, unexpected = net.load_state_dict(ckpt, strict=False)
    assert not unexpected and all("bound" in m for m in missing)
    assert torch.equal(net.gaussian_conditional.quantized_cdf, src.quantized_cdf)
    # a different table size (custom levels) replaces the buffers again, and the scalar bounds travel
    small = GaussianConditional(None, scale_bound=0.2)
    small.update_scale_table(ogc.get_scale_table(0.2, 64, 16))
    net.gaussian_conditional.load_state_dict(small.state_dict())
    assert tuple(net.gaussian_conditional._offset.shape) == (16,)
    assert abs(net.gaussian_conditional._bounds[0] - 0.2) < 1e-7


def test_pmf_to_quantized_cdf_steals_for_zero_bins():
    cdf = _pmf_to_quantized_cdf([0.5, 1e-12, 0.5, 1e-12])
    assert cdf[0] == 0 and cdf[-1] == 65536
    assert all(b > a for a, b in zip(cdf, cdf[1:]))

"""GPU parity of each elementary token-major operator (C ABI) against plain torch-CPU fp32, which is
what the reference dispatches to for the same call (dcae.py:300-509, 584-611).
Tolerances (per-tensor relative error, max|a-b|/max|b|):
  fp32 SIMT and element-wise kernels: 2e-6 (4e-6 for K > 4096: summation-order noise grows ~sqrt(K));
  TF32x3 tcgen05: 1e-5 (north-star fp32 bar);
  single-pass TF32: 3e-3 (stated reduced-precision mode, like torch allow_tf32=True)."""
import pytest
import torch
import torch.nn.functional as F

import _cabi as K
from _util import rel_err

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-6, "tf32x3": 1e-5, "f16x3": 1e-5, "tf32": 3e-3}


def tok(x):   # NCHW -> [T, C] (b, y, x) order
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


def untok(t, B, h, w):
    return t.reshape(B, h, w, -1).permute(0, 3, 1, 2).contiguous()


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3", "tf32"])
@pytest.mark.parametrize("B,h,w,K_,N", [(2, 7, 9, 640, 640), (1, 16, 16, 704, 320), (3, 8, 16, 2560, 640), (1, 5, 3, 64, 224)])
def test_linear_gemm(math, B, h, w, K_, N):
    x = torch.randn(B * h * w, K_ + 64, generator=g(1))
    wt = torch.randn(N, K_, generator=g(2)) / K_ ** 0.5
    b = torch.randn(N, generator=g(3))
    want = F.linear(x[:, 32:32 + K_], wt, b)
    got = K.gemm(x.cuda(), B, h, w, 32, K_, wt.cuda(), math=math, bias=b.cuda())
    assert rel_err(got.cpu(), want) < TOL[math]


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3", "tf32"])
@pytest.mark.parametrize("B,h,w,C,N", [(2, 7, 9, 64, 224), (1, 16, 16, 224, 128), (2, 8, 12, 960, 672), (1, 20, 40, 128, 64)])
def test_conv3x3_implicit_gemm(math, B, h, w, C, N):
    x = torch.randn(B, C, h, w, generator=g(4))
    wt = torch.randn(N, C, 3, 3, generator=g(5)) / (9 * C) ** 0.5
    b = torch.randn(N, generator=g(6))
    want = tok(F.conv2d(x, wt, b, padding=1))
    w2d = wt.permute(0, 2, 3, 1).reshape(N, 9 * C)
    got = K.gemm(tok(x).cuda(), B, h, w, 0, C, w2d.cuda(), math=math, taps=9, bias=b.cuda())
    assert rel_err(got.cpu(), want) < TOL[math] * (2 if 9 * C > 4096 else 1)


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3"])
def test_gemm_two_segments_and_epilogues(math):
    B, h, w = 2, 6, 10
    T = B * h * w
    buf = torch.randn(T, 512, generator=g(7))
    wt = torch.randn(128, 9 * (64 + 96), generator=g(8)) * 0.03
    bias, rs = torch.randn(128, generator=g(9)), torch.rand(128, generator=g(10)) + 0.5
    addend, resid = torch.randn(T, 128, generator=g(11)), torch.randn(T, 128, generator=g(12))
    # reference: gather the two column segments, conv as unfold
    a = torch.cat([buf[:, 64:128], buf[:, 320:416]], 1)
    a_img = untok(a, B, h, w)
    w4 = wt.reshape(128, 3, 3, 160).permute(0, 3, 1, 2).contiguous()
    acc = tok(F.conv2d(a_img, w4, None, padding=1))
    v = acc + bias + addend
    want = torch.cat([F.gelu(v[:, :64]), v[:, 64:]], 1) + resid * rs
    got = K.gemm(buf.cuda(), B, h, w, 64, 64, wt.cuda(), math=math, taps=9, col1=320, k1=96, bias=bias.cuda(),
                 addend=addend.cuda(), residual=resid.cuda(), res_scale=rs.cuda(), act=1, act_cols=64)
    assert rel_err(got.cpu(), want) < TOL[math]
    # half-tanh + unscaled residual, written into a column window of a wider buffer (the LRP epilogue)
    out = torch.zeros(T, 256).cuda()
    K.gemm(buf.cuda(), B, h, w, 64, 64, wt.cuda(), math=math, taps=9, col1=320, k1=96, bias=bias.cuda(),
           residual=resid.cuda(), act=2, out=out, out_col=128)
    want2 = resid + 0.5 * torch.tanh(acc + bias)
    assert rel_err(out[:, 128:].cpu(), want2) < TOL[math]
    assert float(out[:, :128].abs().max()) == 0.0


def test_gemm_deterministic_and_batch_invariant():
    x = torch.randn(2, 128, 9, 11, generator=g(13))
    wt = (torch.randn(64, 128, 3, 3, generator=g(14)) * 0.05).permute(0, 2, 3, 1).reshape(64, -1).cuda()
    for math in ("fp32", "tf32x3", "f16x3"):
        full = K.gemm(tok(x).cuda(), 2, 9, 11, 0, 128, wt, math=math, taps=9)
        again = K.gemm(tok(x).cuda(), 2, 9, 11, 0, 128, wt, math=math, taps=9)
        one = K.gemm(tok(x[1:]).cuda(), 1, 9, 11, 0, 128, wt, math=math, taps=9)
        assert torch.equal(full, again)
        assert torch.equal(full[99:], one)


def test_gemm_argument_errors():
    from dcae_b200 import _lib
    x = torch.randn(64, 96).cuda()
    with pytest.raises(_lib.DcaeError):
        K.gemm(x, 1, 8, 8, 0, 48, torch.randn(32, 48).cuda())       # k0 not a multiple of 32
    with pytest.raises(_lib.DcaeError):
        K.gemm(x, 1, 8, 8, 0, 64, torch.randn(32, 96).cuda())       # K mismatch
    with pytest.raises(_lib.DcaeError):
        K.gemm(x, 1, 8, 8, 64, 64, torch.randn(32, 64).cuda())      # columns exceed ld


def test_layernorm_gelu():
    x = torch.randn(333, 640, generator=g(20)) * 3 + 1
    gm, bt = torch.rand(640, generator=g(21)) + 0.5, torch.randn(640, generator=g(22))
    assert rel_err(K.layernorm(x.cuda(), gm.cuda(), bt.cuda()).cpu(), F.layer_norm(x, (640,), gm, bt, 1e-5)) < 2e-6
    assert rel_err(K.gelu(x.cuda()).cpu(), F.gelu(x)) < 2e-6


@pytest.mark.parametrize("C", [640, 1280])
def test_dwconv3x3(C):
    B, h, w = 2, 7, 9
    x = torch.randn(B, C, h, w, generator=g(23))
    wt = torch.randn(C, 1, 3, 3, generator=g(24)) / 3
    b = torch.randn(C, generator=g(25))
    gate = torch.randn(B, C, h, w, generator=g(26))
    want = F.gelu(F.conv2d(x, wt, b, padding=1, groups=C)) * gate
    got = K.dwconv3x3(tok(x).cuda(), wt.reshape(C, 9).t().contiguous().cuda(), b.cuda(), B, h, w, act=1, gate=tok(gate).cuda())
    assert rel_err(got.cpu(), tok(want)) < 2e-6


def test_spatial_gate():
    B, h, w, C = 2, 9, 7, 640
    s = torch.randn(B, C, h, w, generator=g(27))
    x0 = torch.randn(B, C, h, w, generator=g(28))
    rs = torch.rand(C, generator=g(29)) + 0.5
    w7 = torch.randn(1, 2, 7, 7, generator=g(30)) * 0.1
    att = torch.sigmoid(F.conv2d(torch.cat([s.mean(1, keepdim=True), s.max(1, keepdim=True)[0]], 1), w7, padding=3))
    want = s * att + x0 * rs[None, :, None, None]
    got = K.spatial_gate(tok(s).cuda(), tok(x0).cuda(), rs.cuda(), w7.reshape(-1).cuda(), B, h, w)
    assert rel_err(got.cpu(), tok(want)) < 2e-6


@pytest.mark.parametrize("math", ["fp32", "tf32x3", "f16x3", "tf32"])
@pytest.mark.parametrize("T", [300, 128, 4133, 24576])
def test_dict_attention_core(math, T):
    """dcae.py:489-501.  Logits of a few units so that the softmax is neither flat nor one-hot."""
    q = torch.randn(T, 640, generator=g(31))
    Kh = torch.randn(20, 128, 32, generator=g(32)) * 0.5
    Vh = torch.randn(20, 128, 32, generator=g(33))
    sc = torch.rand(20, generator=g(34)) + 0.5
    qh = q.reshape(T, 20, 32).permute(1, 0, 2)
    sim = torch.einsum("enc,edc->end", qh, Kh) * sc[:, None, None]
    want = torch.einsum("end,edc->enc", torch.softmax(sim, -1), Vh).permute(1, 0, 2).reshape(T, 640)
    got = K.dict_attention(q.cuda(), Kh.cuda(), Vh.cuda(), sc.cuda(), math=math)
    assert rel_err(got.cpu(), want) < {"fp32": 2e-6, "tf32x3": 1e-5, "f16x3": 1e-5, "tf32": 2e-2}[math]
    again = K.dict_attention(q.cuda(), Kh.cuda(), Vh.cuda(), sc.cuda(), math=math)
    assert torch.equal(got, again)


def test_f16x3_attention_needs_query_planes():
    """DCAE_MATH_F16X3 reads the query from fp16 planes; without them the call fails loudly (no silent split / fallback)."""
    from dcae_b200 import _lib
    lib = _lib.load()
    q = torch.randn(128, 640, generator=g(60)).cuda()
    Kh, Vh = torch.randn(20, 128, 32, generator=g(61)).cuda(), torch.randn(20, 128, 32, generator=g(62)).cuda()
    kv = _lib.DictKV(Kh.data_ptr(), Vh.data_ptr(), None, None, None, None, torch.ones(20).cuda().data_ptr())
    out = torch.empty_like(q)
    rc = lib.dcae_op_dict_attention(q.data_ptr(), 640, None, kv, 128, out.data_ptr(), 640, None, _lib.MATH["f16x3"], None)
    assert rc != 0 and b"q16" in lib.dcae_last_error()


def test_transposes_roundtrip():
    x = torch.randn(3, 70, 5, 9, generator=g(35))
    t = K.nchw_to_tokens(x.cuda())
    assert torch.equal(t.cpu(), tok(x))
    assert torch.equal(K.tokens_to_nchw(t, 3, 5, 9).cpu(), x)
    xi = torch.randint(-100, 100, (2, 64, 4, 6), dtype=torch.int32)
    ti = K.nchw_to_tokens(xi.cuda())
    assert torch.equal(K.tokens_to_nchw(ti, 2, 4, 6).cpu(), xi)


def _planes_close(hi, lo, want):
    """fp16 hi + lo carries 22 bits: |hi + lo - v| <= 2^-21 |v| (+ the fp16 subnormal quantum)."""
    got = hi.float() + lo.float()
    return bool(((got - want).abs() <= 2.0 ** -21 * want.abs() + 1.2e-7).all())


def test_producers_write_fp16_planes():
    """Every producer of a GEMM operand can emit the fp16 hi/lo planes the f16x3 GEMM reads (no split pass)."""
    x = torch.randn(257, 640, generator=g(40)) * 3
    gm, bt = torch.rand(640, generator=g(41)) + 0.5, torch.randn(640, generator=g(42))
    out, hi, lo = K.layernorm(x.cuda(), gm.cuda(), bt.cuda(), planes=True)
    assert _planes_close(hi, lo, out)
    out, hi, lo = K.gelu(x.cuda(), planes=True)
    assert _planes_close(hi, lo, out)
    B, h, w, C = 2, 7, 9, 640
    xi = torch.randn(B, C, h, w, generator=g(43))
    out, hi, lo = K.dwconv3x3(tok(xi).cuda(), (torch.randn(C, 9, generator=g(44)) / 3).t().contiguous().cuda(),
                              torch.randn(C, generator=g(45)).cuda(), B, h, w, act=1, planes=True)
    assert _planes_close(hi, lo, out)
    out, hi, lo = K.nchw_to_tokens(xi.cuda(), planes=True)
    assert _planes_close(hi, lo, out) and torch.equal(out.cpu(), tok(xi))
    q = torch.randn(300, 640, generator=g(46))
    Kh, Vh = torch.randn(20, 128, 32, generator=g(47)) * 0.5, torch.randn(20, 128, 32, generator=g(48))
    for math in ("tf32x3", "f16x3"):
        out, hi, lo = K.dict_attention(q.cuda(), Kh.cuda(), Vh.cuda(), (torch.rand(20, generator=g(49)) + 0.5).cuda(), math=math, planes=True)
        assert _planes_close(hi, lo, out)


def test_f16x3_gemm_reads_and_writes_planes_directly():
    B, h, w, C, N = 2, 7, 9, 224, 128
    T = B * h * w
    x = torch.randn(B, 704, h, w, generator=g(50))
    wt = torch.randn(N, C, 3, 3, generator=g(51)) / (9 * C) ** 0.5
    b = torch.randn(N, generator=g(52))
    xt = tok(x).cuda()                                    # [T, 704]; the operand is the window [224, 448)
    hi = xt.half()
    lo = (xt - hi.float()).half()
    want = tok(torch.nn.functional.gelu(torch.nn.functional.conv2d(x[:, 224:448], wt, b, padding=1)))
    w2d = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).cuda()
    o_hi = torch.zeros(T, 256, dtype=torch.float16, device="cuda")
    o_lo = torch.zeros_like(o_hi)
    got = K.gemm(xt, B, h, w, 224, C, w2d, math="f16x3", taps=9, bias=b.cuda(), act=1, src16=(hi, lo), out16=(o_hi, o_lo))
    assert rel_err(got.cpu(), want) < 1e-5
    assert _planes_close(o_hi[:, :N], o_lo[:, :N], got)
    assert float(o_hi[:, N:].abs().max()) == 0.0


def test_f16x3_gemm_second_planes_output_carries_the_next_gelu():
    """dcae.py:421-423: GELU is the first op of every dense layer; the producing GEMM writes planes of GELU(out)."""
    B, h, w, C, N = 2, 7, 9, 640, 160
    T = B * h * w
    x = torch.randn(T, C, generator=g(53))
    wt = torch.randn(N, C, generator=g(54)) / C ** 0.5
    b = torch.randn(N, generator=g(55))
    want = x @ wt.t() + b
    o_hi, o_lo, a_hi, a_lo = (torch.zeros(T, 192, dtype=torch.float16, device="cuda") for _ in range(4))
    got = K.gemm(x.cuda(), B, h, w, 0, C, wt.cuda(), math="f16x3", bias=b.cuda(), out16=(o_hi, o_lo),
                 out16_act=(a_hi, a_lo), act2=1)
    assert rel_err(got.cpu(), want) < 1e-5
    assert _planes_close(o_hi[:, :N], o_lo[:, :N], got)
    assert _planes_close(a_hi[:, :N], a_lo[:, :N], torch.nn.functional.gelu(got))
    assert float(a_hi[:, N:].abs().max()) == 0.0

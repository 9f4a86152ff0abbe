"""Thin test-side wrappers that call the C ABI ops with torch CUDA tensors."""
import torch

from dcae_b200 import _lib


def _s(dev):
    return _lib.current_stream(dev)


def split_weight(w2d: torch.Tensor):
    lib = _lib.load()
    w = w2d.contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.check(lib.dcae_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), _s(w.device)))
    return hi, lo


def gemm(buf, B, h, w, col0, k0, weight2d, math="fp32", taps=1, col1=0, k1=0, bias=None, addend=None,
         residual=None, res_scale=None, act=0, act_cols=0, out=None, out_col=0, src16=None, out16=None, out16_act=None, act2=0):
    """src16 = (hi, lo) fp16 planes [T, ld] to read the operand window from (f16x3); out16 = (hi, lo) planes to
    also write the result to."""
    """buf: token-major [T, ld] CUDA fp32.  weight2d: [N, K].  Returns out [T, N] (or writes into `out`)."""
    lib = _lib.load()
    T, ld = buf.shape
    N, K = weight2d.shape
    wt = weight2d.contiguous()
    hi, lo = split_weight(wt)
    from dcae_b200.weights import f16_weight_planes
    nbytes = lib.dcae_planes_bytes(T, k0 + k1)
    planes = torch.empty(nbytes + 128, dtype=torch.uint8, device=buf.device)
    pbase = (planes.data_ptr() + 127) // 128 * 128
    a = _lib.Operand(buf.data_ptr(), ld, col0, k0, col1, k1, taps, B, h, w, pbase, nbytes)
    if src16 is not None:
        a.src16 = _lib.Planes(src16[0].data_ptr(), src16[1].data_ptr(), src16[0].stride(0))
        a.base = None
    h16, l16, K16, descale = f16_weight_planes(lib, wt, taps, _s(buf.device))
    W = _lib.Weight(wt.data_ptr(), hi.data_ptr(), lo.data_ptr(), N, K, h16.data_ptr(), l16.data_ptr(), K16, descale)
    if out is None:
        out = torch.zeros(T, N, device=buf.device)
    e = _lib.Epilogue()
    e.bias = _lib.ptr(bias)
    if addend is not None:
        e.addend, e.addend_ld = addend.data_ptr(), addend.stride(0)
    if residual is not None:
        e.residual, e.residual_ld = residual.data_ptr(), residual.stride(0)
    e.res_scale = _lib.ptr(res_scale)
    e.act, e.act_cols = act, act_cols
    e.out, e.out_ld = out.data_ptr() + 4 * out_col, out.stride(0)
    if out16 is not None:
        e.out16 = _lib.Planes(out16[0].data_ptr(), out16[1].data_ptr(), out16[0].stride(0))
    if out16_act is not None:
        e.out16_act = _lib.Planes(out16_act[0].data_ptr(), out16_act[1].data_ptr(), out16_act[0].stride(0))
        e.act2 = act2
    _lib.check(lib.dcae_op_gemm(a, W, e, _lib.MATH[math], _s(buf.device)), "dcae_op_gemm")
    return out


def _planes(T, C, dev):
    hi = torch.zeros(T, C, dtype=torch.float16, device=dev)
    lo = torch.zeros(T, C, dtype=torch.float16, device=dev)
    return _lib.Planes(hi.data_ptr(), lo.data_ptr(), C), hi, lo


def layernorm(x, g, b, planes=False):
    lib = _lib.load()
    out = torch.empty_like(x)
    p16, hi, lo = _planes(x.shape[0], x.shape[1], x.device) if planes else (None, None, None)
    _lib.check(lib.dcae_op_layernorm(x.data_ptr(), x.stride(0), g.data_ptr(), b.data_ptr(), x.shape[1], x.shape[0],
                                     out.data_ptr(), out.stride(0), p16, _s(x.device)))
    return (out, hi, lo) if planes else out


def gelu(x, planes=False):
    lib = _lib.load()
    out = torch.empty_like(x)
    p16, hi, lo = _planes(x.shape[0], x.shape[1], x.device) if planes else (None, None, None)
    _lib.check(lib.dcae_op_gelu(x.data_ptr(), x.stride(0), x.shape[1], x.shape[0], out.data_ptr(), out.stride(0), p16, _s(x.device)))
    return (out, hi, lo) if planes else out


def dwconv3x3(x, wt9c, bias, B, h, w, act=0, gate=None, planes=False):
    lib = _lib.load()
    C = wt9c.shape[1]
    out = torch.empty(x.shape[0], C, device=x.device)
    p16, hi, lo = _planes(x.shape[0], C, x.device) if planes else (None, None, None)
    _lib.check(lib.dcae_op_dwconv3x3(x.data_ptr(), x.stride(0), wt9c.data_ptr(), bias.data_ptr(), C, B, h, w, act,
                                     _lib.ptr(gate), gate.stride(0) if gate is not None else 0,
                                     out.data_ptr(), out.stride(0), p16, _s(x.device)))
    return (out, hi, lo) if planes else out


def spatial_gate(s_out, x0, res_scale, w7, B, h, w):
    lib = _lib.load()
    out = torch.empty_like(s_out)
    stats = torch.empty(s_out.shape[0], 2, device=s_out.device)
    _lib.check(lib.dcae_op_spatial_gate(s_out.data_ptr(), s_out.stride(0), x0.data_ptr(), x0.stride(0), res_scale.data_ptr(),
                                        w7.data_ptr(), s_out.shape[1], B, h, w, stats.data_ptr(), out.data_ptr(),
                                        out.stride(0), _s(s_out.device)))
    return out


def dict_attention(q, Kh, Vh, head_scale, math="fp32", planes=False):
    lib = _lib.load()
    out = torch.empty_like(q)
    Kh, Vh = Kh.contiguous(), Vh.contiguous()
    Vt = Vh.transpose(1, 2).contiguous()
    khi, klo = split_weight(Kh)
    vhi, vlo = split_weight(Vt)
    kv = _lib.DictKV(Kh.data_ptr(), Vh.data_ptr(), khi.data_ptr(), klo.data_ptr(), vhi.data_ptr(), vlo.data_ptr(),
                     head_scale.data_ptr())
    q16 = keep = None
    if math == "f16x3":          # the f16x3 kernel reads fp16 planes: q, K as [128, 640], V^T as [640, 128]
        from dcae_b200.weights import f16_weight_planes
        s = _s(q.device)
        k2d = Kh.permute(1, 0, 2).reshape(Kh.shape[1], -1).contiguous()                 # 'e n c -> n (e c)'
        vt2d = Vh.permute(0, 2, 1).reshape(-1, Vh.shape[1]).contiguous()                # 'e n c -> (e c) n'
        k_hi, k_lo, _, kv.k_descale = f16_weight_planes(lib, k2d, 1, s)
        v_hi, v_lo, _, kv.v_descale = f16_weight_planes(lib, vt2d, 1, s)
        kv.K16_hi, kv.K16_lo, kv.Vt16_hi, kv.Vt16_lo = k_hi.data_ptr(), k_lo.data_ptr(), v_hi.data_ptr(), v_lo.data_ptr()
        q_hi = q.half()
        q_lo = (q - q_hi.float()).half()
        q16 = _lib.Planes(q_hi.data_ptr(), q_lo.data_ptr(), q_hi.stride(0))
        keep = (k_hi, k_lo, v_hi, v_lo, q_hi, q_lo)
    p16, hi, lo = _planes(q.shape[0], q.shape[1], q.device) if planes else (None, None, None)
    _lib.check(lib.dcae_op_dict_attention(q.data_ptr(), q.stride(0), q16, kv, q.shape[0], out.data_ptr(), out.stride(0), p16,
                                          _lib.MATH[math], _s(q.device)), "dcae_op_dict_attention")
    torch.cuda.synchronize()
    del keep
    return (out, hi, lo) if planes else out


def nchw_to_tokens(x, planes=False):
    lib = _lib.load()
    B, C, H, W = x.shape
    out = torch.empty(B * H * W, C, device=x.device, dtype=x.dtype)
    fn = lib.dcae_op_nchw_to_tokens if x.dtype == torch.float32 else lib.dcae_op_nchw_to_tokens_i32
    if x.dtype == torch.float32:
        p16, hi, lo = _planes(B * H * W, C, x.device) if planes else (None, None, None)
        _lib.check(fn(x.contiguous().data_ptr(), B, C, H * W, out.data_ptr(), C, p16, _s(x.device)))
        return (out, hi, lo) if planes else out
    _lib.check(fn(x.contiguous().data_ptr(), B, C, H * W, out.data_ptr(), C, _s(x.device)))
    return out


def tokens_to_nchw(t, B, H, W):
    lib = _lib.load()
    C = t.shape[1]
    out = torch.empty(B, C, H, W, device=t.device, dtype=t.dtype)
    fn = lib.dcae_op_tokens_to_nchw if t.dtype == torch.float32 else lib.dcae_op_tokens_to_nchw_i32
    _lib.check(fn(t.data_ptr(), t.stride(0), B, C, H * W, out.data_ptr(), _s(t.device)))
    return out

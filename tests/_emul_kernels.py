"""TEST INFRASTRUCTURE ONLY (every result is kept in fp32; the `want` form flags of the CUDA backend are accepted and
ignored): torch-CPU restatement of the C-ABI operator contracts (include/dcae_b200.h) that
`dcae_b200.transforms` composes.  It lets the CPU suite check the host logic -- weight re-indexing (stride-2 conv as
space-to-depth + 3x3, transposed conv as 3x3 + depth-to-space, channel padding, qkv layout) and the block composition
-- against the reference's real modules without a GPU.  Each method follows the header comment of the operator it
stands for, NOT the reference module (the reference module is what the test compares the composition with)."""
import torch
import torch.nn.functional as F

from dcae_b200 import _lib
from dcae_b200.transforms import Act, PackedGemm, _pad_to


class TorchKernels:
    math = "emulated"

    def tensor(self, t):
        return t.detach().to(torch.float32).contiguous()

    def vector(self, t, n_pad=None, fill=0.0):
        t = t.detach().reshape(-1).to(torch.float32)
        if n_pad is not None and t.numel() < n_pad:
            t = torch.cat([t, t.new_full((n_pad - t.numel(),), fill)])
        return t.contiguous()

    def pack_gemm(self, w2d, bias, taps=1):
        pg = PackedGemm(w2d.shape[0], w2d.shape[1], taps)
        pg.w, pg.bias = self.tensor(w2d), self.tensor(_pad_to(bias.reshape(-1), 0, w2d.shape[0]))
        return pg

    def to_tokens(self, x, ld, want=1):
        B, C, H, W = x.shape
        out = torch.zeros(B * H * W, ld)
        out[:, :C] = x.permute(0, 2, 3, 1).reshape(-1, C)
        return Act(out, B, H, W)

    def to_nchw(self, a, C):
        return a.buf[:, :C].reshape(a.B, a.h, a.w, C).permute(0, 3, 1, 2).contiguous()

    def gemm(self, a, pg, act=_lib.ACT_NONE, residual=None, res_scale=None, want=1):
        """dcae_op_gemm: A = the first K / taps columns, gathered over the 9 taps of a 3x3 / stride 1 / pad 1 window when
        taps = 9, K ordered tap-major (tap = 3 (dy + 1) + (dx + 1))."""
        k = pg.K // pg.taps
        x = a.buf[:, :k]
        if pg.taps == 9:
            img = F.pad(x.reshape(a.B, a.h, a.w, k), (0, 0, 1, 1, 1, 1))
            x = torch.cat([img[:, dy:dy + a.h, dx:dx + a.w] for dy in range(3) for dx in range(3)], dim=-1).reshape(a.T, 9 * k)
        acc = x.double() @ pg.w.double().t() + pg.bias.double()
        if act == _lib.ACT_RELU:
            acc = acc.clamp_min(0)
        elif act == _lib.ACT_GELU:
            acc = F.gelu(acc)
        elif act != _lib.ACT_NONE:
            raise NotImplementedError
        if residual is not None:
            r = residual.buf[:, :pg.N].double()
            acc = acc + (r * res_scale.double() if res_scale is not None else r)
        return Act(acc.float(), a.B, a.h, a.w)

    def layernorm(self, a, gamma, beta, C, want=1):
        out = torch.zeros_like(a.buf)
        out[:, :C] = F.layer_norm(a.buf[:, :C], (C,), gamma, beta, 1e-5)
        return Act(out, a.B, a.h, a.w)

    def window_attention(self, qkv, C, c_pad, head_dim, window, shift, rel, want=1):
        """dcae_op_window_attention, written per token from the header comment (slow, small cases only)."""
        B, h, w, P = qkv.B, qkv.h, qkv.w, window
        q = qkv.buf[:, 0:C].reshape(B, h, w, C)
        k = qkv.buf[:, c_pad:c_pad + C].reshape(B, h, w, C)
        v = qkv.buf[:, 2 * c_pad:2 * c_pad + C].reshape(B, h, w, C)
        out = torch.zeros(B, h, w, c_pad)
        nh = C // head_dim
        sp = P - shift
        for wy in range(h // P):
            for wx in range(w // P):
                ys = [(wy * P + py + shift) % h for py in range(P)]
                xs = [(wx * P + px + shift) % w for px in range(P)]
                qi = q[:, ys][:, :, xs].reshape(B, P * P, nh, head_dim)
                ki = k[:, ys][:, :, xs].reshape(B, P * P, nh, head_dim)
                vi = v[:, ys][:, :, xs].reshape(B, P * P, nh, head_dim)
                sim = torch.einsum("bpec,bqec->bepq", qi, ki) / head_dim ** 0.5
                pos = torch.tensor([[py, px] for py in range(P) for px in range(P)])
                d = pos[:, None, :] - pos[None, :, :] + P - 1
                sim = sim + rel[:, d[..., 0], d[..., 1]][None]
                if shift:
                    mask = torch.zeros(P * P, P * P, dtype=torch.bool)
                    if wy == h // P - 1:
                        mask |= (pos[:, None, 0] < sp) != (pos[None, :, 0] < sp)
                    if wx == w // P - 1:
                        mask |= (pos[:, None, 1] < sp) != (pos[None, :, 1] < sp)
                    sim = sim.masked_fill(mask[None, None], float("-inf"))
                o = torch.einsum("bepq,bqec->bpec", sim.softmax(-1), vi).reshape(B, P, P, C)
                for a_, yy in enumerate(ys):
                    for b_, xx in enumerate(xs):
                        out[:, yy, xx, :C] = o[:, a_, b_]
        return Act(out.reshape(B * h * w, c_pad), B, h, w)

    def dwconv_glu(self, f, wt9c, bias, hid, want=1):
        x = f.buf[:, :hid].reshape(f.B, f.h, f.w, hid).permute(0, 3, 1, 2)
        y = F.conv2d(x, wt9c.t().reshape(hid, 1, 3, 3), bias, padding=1, groups=hid)
        y = F.gelu(y).permute(0, 2, 3, 1).reshape(f.T, hid) * f.buf[:, hid:2 * hid]
        return Act(y.contiguous(), f.B, f.h, f.w)

    def space_to_depth(self, a, C, cs, want=1):
        h2, w2 = (a.h + 1) // 2, (a.w + 1) // 2
        img = torch.zeros(a.B, 2 * h2, 2 * w2, C)
        img[:, :a.h, :a.w] = a.buf[:, :C].reshape(a.B, a.h, a.w, C)
        out = torch.zeros(a.B, h2, w2, 4, cs)
        for sy in (0, 1):
            for sx in (0, 1):
                out[:, :, :, sy * 2 + sx, :C] = img[:, sy::2, sx::2]
        return Act(out.reshape(a.B * h2 * w2, 4 * cs), a.B, h2, w2)

    def depth_to_space(self, a, cs, C, c_pad, want=1):
        x = a.buf[:, :4 * cs].reshape(a.B, a.h, a.w, 4, cs)
        out = torch.zeros(a.B, 2 * a.h, 2 * a.w, c_pad)
        for py in (0, 1):
            for px in (0, 1):
                out[:, py::2, px::2, :C] = x[:, :, :, py * 2 + px, :C]
        return Act(out.reshape(-1, c_pad), a.B, 2 * a.h, 2 * a.w)

// Kernel 3: GaussianConditional quantise + likelihood + CDF index in one HBM pass.
// Replaces compressai GaussianConditional.forward/quantize/build_indexes/dequantize at
// /root/reference/models/dcae.py:657-659, :738-740, :891-896 (math: dcae.py:839-857, :57-58).
//
// HBM-bound elementwise kernel: every thread owns float4 groups (16-byte coalesced loads and
// stores), the 64-entry scale table lives in shared memory, and the table search is a log-domain
// guess corrected against the real table entries (exact for ANY sorted table).  Arithmetic follows
// the reference's op order with explicit round-to-nearest intrinsics so that no FMA contraction
// can change a symbol, an index or a likelihood bit.
#include "common.cuh"

namespace dcae {

constexpr int GC_THREADS = 256;
constexpr int GC_MAX_BLOCKS = 148 * 8;
constexpr int GC_MAX_TABLE = 256;

struct GcParams {
  dcae_gc_args a;
  int64_t groups;   // rows * inner/4
  uint32_t inner4;
  uint32_t shift;   // log2(inner4) when inner4 is a power of two
  int32_t n_partials;   // dcae_gc_num_partials(rows, inner): entries past the grid size are written as zero
};

__device__ __forceinline__ float nan_max(float x, float bound) {
  // torch.max(x, bound): NaN propagates
  return (x != x) ? x : fmaxf(x, bound);
}

// DCAE_GC_LIK_REFERENCE: the reference's op order (dcae.py:839-857) with explicit round-to-nearest intrinsics and
// libdevice erfcf -- bit-identical to torch evaluating the reference formula on the GPU (the test mode).
__device__ __forceinline__ float gaussian_likelihood(float out, float mu, float s, float lik_bound) {
  const float c = -0.70710678118654752440f;  // float(-(2 ** -0.5)), dcae.py:855
  float v = fabsf(__fsub_rn(out, mu));
  float up = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(0.5f, v), s))));
  float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(-0.5f, v), s))));
  return nan_max(__fsub_rn(up, lo), lik_bound);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// DCAE_GC_LIK_FAST (default): the same quantity, lik = 1/2 erfc(a) - 1/2 erfc(b) with a = (v - 1/2) k, b = (v + 1/2) k,
// k = 1 / (s sqrt 2), evaluated WITHOUT the subtraction of two rounded erfc values.  For x >= 0, erfc(x) = 2^P(x - C)
// with one degree-11 polynomial on [0, 5.5] (weighted minimax fit; beyond 5.5 erfc < 1e-14, below any likelihood above
// the 1e-9 floor).  Horner at a keeps its intermediates q_j, which are the coefficients of the divided difference
// Q(x) = (P(x) - P(a)) / (x - a); a second Horner gives Q(b), so P(b) - P(a) = (b - a) Q(b) with b - a = k known exactly:
//   a >= 0:           lik = 1/2 2^P(a) (1 - 2^(k Q(b)))                                 [1 - 2^u by series for u > -1]
//   a <  0, b >= 1/2: lik = 1/2 [(1 - 2^P(|a|)) + (1 - 2^P(b))]  = 1/2 [erf|a| + erf b]
//   a <  0, b <  1/2: lik = 1/2 [|a| T(a^2) + b T(b^2)]           (odd polynomial: relative accuracy near 0)
// Branch-free, ~75 instructions, 5 MUFU.  Measured against mpmath on log-uniform scales 0.11 .. 300 (tools/fit_erfc.py,
// tests/test_gpu_gc.py): <= 6e-6 relative everywhere above the floor -- the reference's own fp32 formula is at 1.5e-4
// there (cancellation at large scales), so this variant is closer to the exact value than the reference evaluation.
__device__ __forceinline__ float gaussian_likelihood_fast(float out, float mu, float s, float lik_bound) {
  constexpr float XMAX = 5.5f, CEN = 2.75f;
  const float v = fabsf(out - mu);
  const float sc = fminf(s, 1.0e30f);                    // s = inf: k -> 0 like the reference's x / inf
  float r = rcp_approx(sc);
  r = fmaf(fmaf(-sc, r, 1.0f), r, r);                    // one Newton step
  const float k = r * 0.70710678118654752440f;
  const float a = (v - 0.5f) * k, b = (v + 0.5f) * k;
  const bool neg = a < 0.0f;
  const float ap = fminf(fabsf(a), XMAX), bp = fminf(b, XMAX);
  float d = neg ? (v + v) * k : k;                       // b - |a| without cancellation
  d = b > XMAX ? bp - ap : d;
  const float za = ap - CEN, zb = bp - CEN;
  float q0, q1, q2, q3, q4, q5, q6, q7, q8, q9, q10;
  const float q11 = 6.966848786760238e-08f;
  q10 = fmaf(q11, za, 9.56267314222714e-08f);
  q9 = fmaf(q10, za, -7.519030873481825e-07f);
  q8 = fmaf(q9, za, 1.12355610326631e-06f);
  q7 = fmaf(q8, za, -1.224499737872975e-05f);
  q6 = fmaf(q7, za, 8.600354340160266e-05f);
  q5 = fmaf(q6, za, -0.00046474975533783436f);
  q4 = fmaf(q5, za, 0.002456091344356537f);
  q3 = fmaf(q4, za, -0.012886710464954376f);
  q2 = fmaf(q3, za, -1.3724150657653809f);
  q1 = fmaf(q2, za, -8.405914306640625f);
  q0 = fmaf(q1, za, -13.278767585754395f);               // P(a)
  float Q = fmaf(q11, zb, q10);
  Q = fmaf(Q, zb, q9);
  Q = fmaf(Q, zb, q8);
  Q = fmaf(Q, zb, q7);
  Q = fmaf(Q, zb, q6);
  Q = fmaf(Q, zb, q5);
  Q = fmaf(Q, zb, q4);
  Q = fmaf(Q, zb, q3);
  Q = fmaf(Q, zb, q2);
  Q = fmaf(Q, zb, q1);                                   // (P(b) - P(a)) / (b - a)
  const float delta = fminf(d * Q, 0.0f);
  const float pa = fminf(q0, 0.0f);
  const float pb = fminf(pa + delta, 0.0f);
  const float ea = ex2_approx(pa);
  float t = 1.5252733804059838e-05f;                     // (2^u - 1) / u = ln2 + u ln2^2 / 2 + ...
  t = fmaf(t, delta, 0.00015403530393381606f);
  t = fmaf(t, delta, 0.0013333558146428441f);
  t = fmaf(t, delta, 0.009618129107628477f);
  t = fmaf(t, delta, 0.055504108664821576f);
  t = fmaf(t, delta, 0.2402265069591007f);
  t = fmaf(t, delta, 0.6931471805599453f);
  const float one_m = delta > -1.0f ? -(t * delta) : 1.0f - ex2_approx(delta);
  const float case1 = ea * one_m;
  const float case2 = (1.0f - ea) + (1.0f - ex2_approx(pb));
  const float as = fminf(ap, 0.5f), bs = fminf(bp, 0.5f);
  const float a2 = as * as, b2 = bs * bs;
  float ta = fmaf(-0.02438964508473873f, a2, 0.11245644092559814f), tb = fmaf(-0.02438964508473873f, b2, 0.11245644092559814f);
  ta = fmaf(ta, a2, -0.3761073648929596f); tb = fmaf(tb, b2, -0.3761073648929596f);
  ta = fmaf(ta, a2, 1.128378987312317f); tb = fmaf(tb, b2, 1.128378987312317f);
  const float small = fmaf(as, ta, bs * tb);
  float lik = 0.5f * (neg ? (b < 0.5f ? small : case2) : case1);
  lik = (s != s || a != a) ? __int_as_float(0x7fc00000) : lik;     // NaN in, NaN out (torch.max propagates it)
  return nan_max(lik, lik_bound);
}

// idx = (n-1) - sum_{j<n-1} [s <= tbl[j]]  ==  #{ j < n-1 : tbl[j] < s }   (NaN -> n-1)
// Log-domain guess j0, one branch-free step up or down, then a check of the far neighbour: `bad` says the guess was more
// than one entry off (only for tables that are not log-uniform) and the caller must run table_index_search.
__device__ __forceinline__ int table_index_try(float s, const float* tbl, int n, float log_t0, float inv_step, bool& bad) {
  const float g = (__log2f(s) - log_t0) * inv_step;
  const int j0 = (int)fminf(fmaxf(g, 0.0f), (float)(n - 1));
  const bool up = (j0 < n - 1) && (tbl[j0] < s);
  const bool down = (j0 > 0) && !(tbl[max(j0 - 1, 0)] < s);     // never together with `up` (sorted table)
  const int j = j0 + (int)up - (int)down;
  const float far = tbl[up ? min(j0 + 1, n - 1) : max(j0 - 2, 0)];
  bad = up ? (j < n - 1 && far < s) : (down && j > 0 && !(far < s));
  return (s != s) ? n - 1 : j;
}
__device__ __noinline__ int table_index_search(float s, const float* tbl, int n, int j) {   // exact for ANY sorted table
  if (s != s) return n - 1;
  while (j < n - 1 && tbl[j] < s) ++j;
  while (j > 0 && !(tbl[j - 1] < s)) --j;
  return j;
}
__device__ __forceinline__ int table_index(float s, const float* tbl, int n, float log_t0, float inv_step) {
  bool bad;
  int j = table_index_try(s, tbl, n, log_t0, inv_step, bad);
  if (bad) j = table_index_search(s, tbl, n, j);
  return j;
}

// One float4 group of every tensor lives at row * ld + col; `row, col` come from a shift when inner/4 is a
// power of two (token-major slices: inner = 64), from a 32-bit divide otherwise.  Every thread handles GC_UNROLL
// groups per iteration, all loads issued before any arithmetic (6 x 16 B in flight per thread).
// MODE, LIK (0 = no likelihood, DCAE_GC_LIK_REFERENCE + 1, DCAE_GC_LIK_FAST + 1) and IDX (indexes wanted) are
// compile-time so the element loop is straight-line code.
constexpr int GC_UNROLL = 2;

template <int MODE, int LIK, bool IDX, bool POW2>
__global__ void __launch_bounds__(GC_THREADS) gc_fused_kernel(const GcParams p) {
  __shared__ float tbl[GC_MAX_TABLE];
  __shared__ float red[GC_THREADS / 32];
  const dcae_gc_args& a = p.a;
  const int n = a.n_table;
#pragma unroll 1
  for (int i = threadIdx.x; i < n; i += GC_THREADS) tbl[i] = a.scale_table ? a.scale_table[i] : 0.0f;
  __syncthreads();
  float log_t0 = 0.f, inv_step = 0.f;
  if (IDX && n > 1) {
    log_t0 = __log2f(tbl[0]);
    inv_step = __fdividef((float)(n - 1), __log2f(tbl[n - 1]) - log_t0);   // only seeds the search: the table decides
  }
  constexpr int mode = MODE;
  constexpr bool want_lik = LIK != 0, want_idx = IDX;
  const bool want_log2 = a.log2_partials != nullptr;
  const float scale_bound = a.scale_bound, lik_bound = a.lik_bound;
  float log2_acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * GC_THREADS;

  for (int64_t g0 = (int64_t)blockIdx.x * GC_THREADS + threadIdx.x; g0 < p.groups; g0 += stride * GC_UNROLL) {
    int64_t row[GC_UNROLL], col[GC_UNROLL];
    bool live[GC_UNROLL];
    float4 mu4[GC_UNROLL], y4[GC_UNROLL], sc4[GC_UNROLL], nz4[GC_UNROLL];
    int4 si4[GC_UNROLL];
#pragma unroll
    for (int u = 0; u < GC_UNROLL; ++u) {
      const int64_t g = g0 + u * stride;
      live[u] = g < p.groups;
      if (POW2) {
        row[u] = g >> p.shift;
        col[u] = (g & (int64_t)(p.inner4 - 1)) << 2;
      } else if (p.groups <= 0xffffffffll) {   // 32-bit divide on the common path
        const uint32_t r32 = (uint32_t)g / p.inner4;
        row[u] = r32;
        col[u] = (int64_t)((uint32_t)g - r32 * p.inner4) * 4;
      } else {
        row[u] = g / p.inner4;
        col[u] = (g - row[u] * p.inner4) * 4;
      }
      y4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      mu4[u] = sc4[u] = nz4[u] = y4[u];
      si4[u] = make_int4(0, 0, 0, 0);
      if (live[u]) {
        mu4[u] = __ldg(reinterpret_cast<const float4*>(a.mu + row[u] * a.mu_ld + col[u]));
        if (mode != DCAE_GC_DECODE && a.y != nullptr) y4[u] = __ldg(reinterpret_cast<const float4*>(a.y + row[u] * a.y_ld + col[u]));
        if (a.scale != nullptr) sc4[u] = __ldg(reinterpret_cast<const float4*>(a.scale + row[u] * a.scale_ld + col[u]));
        if (mode == DCAE_GC_NOISE) nz4[u] = __ldg(reinterpret_cast<const float4*>(a.noise + row[u] * a.noise_ld + col[u]));
        if (mode == DCAE_GC_DECODE) si4[u] = __ldg(reinterpret_cast<const int4*>(a.sym_in + row[u] * a.sym_in_ld + col[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < GC_UNROLL; ++u) {
      if (!live[u]) continue;
      const float mu[4] = {mu4[u].x, mu4[u].y, mu4[u].z, mu4[u].w};
      const float y[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
      const float sc[4] = {sc4[u].x, sc4[u].y, sc4[u].z, sc4[u].w};
      const float nz[4] = {nz4[u].x, nz4[u].y, nz4[u].z, nz4[u].w};
      const int si[4] = {si4[u].x, si4[u].y, si4[u].z, si4[u].w};
      float yh[4], lk[4];
      int sy[4], ix[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float r, out;
        if (mode == DCAE_GC_DECODE) {
          r = (float)si[k];                       // dequantize: inputs.type_as(means) + means
          out = __fadd_rn(r, mu[k]);
          yh[k] = out;
        } else {
          r = rintf(__fsub_rn(y[k], mu[k]));      // torch.round: half to even
          yh[k] = __fadd_rn(r, mu[k]);            // ste_round(y - mu) + mu == r + mu (exactly)
          out = (mode == DCAE_GC_NOISE) ? __fadd_rn(y[k], nz[k]) : yh[k];
        }
        sy[k] = (int)r;
        const float s = nan_max(sc[k], scale_bound);
        lk[k] = LIK == DCAE_GC_LIK_REFERENCE + 1 ? gaussian_likelihood(out, mu[k], s, lik_bound)
              : LIK == DCAE_GC_LIK_FAST + 1    ? gaussian_likelihood_fast(out, mu[k], s, lik_bound) : 1.0f;
        ix[k] = want_idx ? table_index(s, tbl, n, log_t0, inv_step) : 0;
        if (want_log2) log2_acc += __log2f(lk[k]);   // bpp numerator: lik in [1e-9, 1], MUFU.LG2 is ample
      }
      const int64_t rw = row[u], cl = col[u];
      if (a.y_hat) *reinterpret_cast<float4*>(a.y_hat + rw * a.y_hat_ld + cl) = make_float4(yh[0], yh[1], yh[2], yh[3]);
      if (a.y_hat16.hi) store_planes4(a.y_hat16, rw, (int)cl, make_float4(yh[0], yh[1], yh[2], yh[3]));
      if (a.lik && want_lik) *reinterpret_cast<float4*>(a.lik + rw * a.lik_ld + cl) = make_float4(lk[0], lk[1], lk[2], lk[3]);
      if (a.sym) *reinterpret_cast<int4*>(a.sym + rw * a.sym_ld + cl) = make_int4(sy[0], sy[1], sy[2], sy[3]);
      if (a.idx) *reinterpret_cast<int4*>(a.idx + rw * a.idx_ld + cl) = make_int4(ix[0], ix[1], ix[2], ix[3]);
    }
  }

  if (a.log2_partials != nullptr) {
    // fixed-order block reduction: lanes by xor-shuffle, warps in index order
    log2_acc = warp_sum(log2_acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = log2_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < GC_THREADS / 32; ++w) t += red[w];
      a.log2_partials[blockIdx.x] = t;
    }
    if (blockIdx.x == 0)        // the caller reduces dcae_gc_num_partials entries whatever this launch's grid was
      for (int i = (int)gridDim.x + (int)threadIdx.x; i < p.n_partials; i += GC_THREADS) a.log2_partials[i] = 0.f;
  }
}

// ---- the hot configuration -----------------------------------------------------------------------------------------
// EVAL mode on row-strided tensors whose row length / 4 is a power of two (token-major slices: inner = 64; the slice loop
// and the roofline microbenchmark), every tensor present, fewer than 2^31 float4 groups and 2^32 elements per tensor:
// 32-bit offsets (one IMAD + one IMAD.WIDE per access instead of 64-bit multiply chains), no per-store pointer tests,
// the rare table-search fallback hoisted out of the element loop.  Same element functions as the generic kernel, so the
// results are the same bits.
struct GcHotParams {
  const float *y, *mu, *scale;
  float *y_hat, *lik;
  int32_t *sym, *idx;
  __half *hi, *lo;
  float* log2_partials;
  const float* scale_table;
  uint32_t y_ld, mu_ld, scale_ld, y_hat_ld, lik_ld, sym_ld, idx_ld, p16_ld;
  uint32_t groups, shift, mask;
  int32_t n_table, n_partials;
  float scale_bound, lik_bound;
};

template <int LIK, bool SYMIDX, bool PLANES>
__global__ void __launch_bounds__(GC_THREADS) gc_eval_hot_kernel(const GcHotParams p) {
  __shared__ float tbl[GC_MAX_TABLE];
  __shared__ float red[GC_THREADS / 32];
  const int n = p.n_table;
  if (SYMIDX) {
#pragma unroll 1
    for (int i = threadIdx.x; i < n; i += GC_THREADS) tbl[i] = p.scale_table[i];
    __syncthreads();
  }
  float log_t0 = 0.f, inv_step = 0.f;
  if (SYMIDX) {
    log_t0 = __log2f(tbl[0]);
    inv_step = __fdividef((float)(n - 1), __log2f(tbl[n - 1]) - log_t0);
  }
  const bool want_log2 = p.log2_partials != nullptr;
  const float scale_bound = p.scale_bound, lik_bound = p.lik_bound;
  float log2_acc = 0.f;
  const uint32_t stride = gridDim.x * GC_THREADS;

  for (uint32_t g0 = blockIdx.x * GC_THREADS + threadIdx.x; g0 < p.groups; g0 += stride * GC_UNROLL) {
    uint32_t row[GC_UNROLL], col[GC_UNROLL];
    bool live[GC_UNROLL];
    float4 mu4[GC_UNROLL], y4[GC_UNROLL], sc4[GC_UNROLL];
#pragma unroll
    for (int u = 0; u < GC_UNROLL; ++u) {
      live[u] = g0 + u * stride < p.groups;
      const uint32_t g = min(g0 + u * stride, p.groups - 1);       // a dead tail group loads the last one and is skipped below
      row[u] = g >> p.shift;
      col[u] = (g & p.mask) << 2;
      mu4[u] = __ldg(reinterpret_cast<const float4*>(p.mu + (row[u] * p.mu_ld + col[u])));
      y4[u] = __ldg(reinterpret_cast<const float4*>(p.y + (row[u] * p.y_ld + col[u])));
      sc4[u] = __ldg(reinterpret_cast<const float4*>(p.scale + (row[u] * p.scale_ld + col[u])));
    }
#pragma unroll
    for (int u = 0; u < GC_UNROLL; ++u) {
      if (!live[u]) continue;
      const float mu[4] = {mu4[u].x, mu4[u].y, mu4[u].z, mu4[u].w};
      const float y[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
      const float sc[4] = {sc4[u].x, sc4[u].y, sc4[u].z, sc4[u].w};
      float yh[4], lk[4], sv[4];
      int sy[4], ix[4];
      bool any_bad = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float r = rintf(__fsub_rn(y[k], mu[k]));      // torch.round: half to even
        yh[k] = __fadd_rn(r, mu[k]);                        // ste_round(y - mu) + mu == r + mu (exactly)
        sy[k] = (int)r;
        const float s = nan_max(sc[k], scale_bound);
        sv[k] = s;
        lk[k] = LIK == DCAE_GC_LIK_REFERENCE + 1 ? gaussian_likelihood(yh[k], mu[k], s, lik_bound)
              : LIK == DCAE_GC_LIK_FAST + 1    ? gaussian_likelihood_fast(yh[k], mu[k], s, lik_bound) : 1.0f;
        if (SYMIDX) {
          bool bad;
          ix[k] = table_index_try(s, tbl, n, log_t0, inv_step, bad);
          any_bad |= bad;
        }
        if (LIK != 0 && want_log2) log2_acc += __log2f(lk[k]);
      }
      if (SYMIDX && any_bad) {                              // not a log-uniform table: the exact search, out of line
#pragma unroll 1
        for (int k = 0; k < 4; ++k) ix[k] = table_index_search(sv[k], tbl, n, ix[k]);
      }
      const uint32_t rw = row[u], cl = col[u];
      *reinterpret_cast<float4*>(p.y_hat + (rw * p.y_hat_ld + cl)) = make_float4(yh[0], yh[1], yh[2], yh[3]);
      if (PLANES) {
        uint2 hv, lv;
        f16_split2(yh[0], yh[1], hv.x, lv.x);
        f16_split2(yh[2], yh[3], hv.y, lv.y);
        *reinterpret_cast<uint2*>(p.hi + (rw * p.p16_ld + cl)) = hv;
        *reinterpret_cast<uint2*>(p.lo + (rw * p.p16_ld + cl)) = lv;
      }
      if (LIK != 0) *reinterpret_cast<float4*>(p.lik + (rw * p.lik_ld + cl)) = make_float4(lk[0], lk[1], lk[2], lk[3]);
      if (SYMIDX) {
        *reinterpret_cast<int4*>(p.sym + (rw * p.sym_ld + cl)) = make_int4(sy[0], sy[1], sy[2], sy[3]);
        *reinterpret_cast<int4*>(p.idx + (rw * p.idx_ld + cl)) = make_int4(ix[0], ix[1], ix[2], ix[3]);
      }
    }
  }
  if (want_log2) {
    log2_acc = warp_sum(log2_acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = log2_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < GC_THREADS / 32; ++w) t += red[w];
      p.log2_partials[blockIdx.x] = t;
    }
    if (blockIdx.x == 0)
      for (int i = (int)gridDim.x + (int)threadIdx.x; i < p.n_partials; i += GC_THREADS) p.log2_partials[i] = 0.f;
  }
}

// ---- kernel 3 backward (training config #4; train.py:165-179 differentiates the rate term through this) ------------
// lik = max(Phi(u) - Phi(l), lik_bound), u = (1/2 - v) / s, l = (-1/2 - v) / s, v = |out - mu|, s = max(scale, scale_bound)
// (dcae.py:839-857).  With g = dL/dlik and phi the standard normal density:
//   g' = g [lik_raw >= lik_bound or g < 0]                       (compressai LowerBound: gradient passes towards the bound)
//   dL/dv = -g' (phi(u) - phi(l)) / s,   dL/ds = -g' (u phi(u) - l phi(l)) / s,   dL/dscale = dL/ds [scale >= scale_bound or dL/ds < 0]
//   NOISE mode (out = y + noise):  dL/dy = dL/dv sign(out - mu),  dL/dmu = -dL/dy
//   EVAL mode  (out = round(y - mu) + mu: no gradient through round, and d out / d mu = 1 cancels d|out - mu| / d mu): 0, 0
__global__ void __launch_bounds__(GC_THREADS) gc_backward_kernel(const dcae_gc_bwd_args a, int64_t groups, uint32_t inner4) {
  for (int64_t g = (int64_t)blockIdx.x * GC_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * GC_THREADS) {
    const int64_t row = g / inner4, col = (g - row * inner4) * 4;
    const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.y + row * a.y_ld + col));
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mu + row * a.mu_ld + col));
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.scale + row * a.scale_ld + col));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.grad_lik + row * a.grad_lik_ld + col));
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.mode == DCAE_GC_NOISE) n4 = __ldg(reinterpret_cast<const float4*>(a.noise + row * a.noise_ld + col));
    const float y[4] = {y4.x, y4.y, y4.z, y4.w}, mu[4] = {m4.x, m4.y, m4.z, m4.w}, sc[4] = {s4.x, s4.y, s4.z, s4.w};
    const float gl[4] = {g4.x, g4.y, g4.z, g4.w}, nz[4] = {n4.x, n4.y, n4.z, n4.w};
    float gy[4], gm[4], gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float out = a.mode == DCAE_GC_NOISE ? y[k] + nz[k] : rintf(y[k] - mu[k]) + mu[k];
      const float d = out - mu[k];
      const float v = fabsf(d);
      const float s = nan_max(sc[k], a.scale_bound);
      const float u = (0.5f - v) / s, l = (-0.5f - v) / s;
      const float raw = 0.5f * erfcf(-0.70710678118654752440f * u) - 0.5f * erfcf(-0.70710678118654752440f * l);
      const float gp = (raw >= a.lik_bound || gl[k] < 0.f) ? gl[k] : 0.f;
      const float pu = 0.3989422804014327f * expf(-0.5f * u * u), pl = 0.3989422804014327f * expf(-0.5f * l * l);
      const float dv = -gp * (pu - pl) / s;
      const float ds = -gp * (u * pu - l * pl) / s;
      gs[k] = (sc[k] >= a.scale_bound || ds < 0.f) ? ds : 0.f;
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      gy[k] = a.mode == DCAE_GC_NOISE ? dv * sgn : 0.f;
      gm[k] = -gy[k];
    }
    if (a.grad_y) *reinterpret_cast<float4*>(a.grad_y + row * a.grad_y_ld + col) = make_float4(gy[0], gy[1], gy[2], gy[3]);
    if (a.grad_mu) *reinterpret_cast<float4*>(a.grad_mu + row * a.grad_mu_ld + col) = make_float4(gm[0], gm[1], gm[2], gm[3]);
    if (a.grad_scale) *reinterpret_cast<float4*>(a.grad_scale + row * a.grad_scale_ld + col) = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ p, int64_t n, float* __restrict__ out) {
  // one block, 256 threads; each thread sums a strided subsequence in index order, then a fixed tree
  __shared__ float sh[256];
  float t = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) t += p[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// The element loop is grid-stride, so the grid is ONE resident wave: SMs x blocks that fit per SM for this variant
// (a fixed 148 x 8 was 1.6 waves at 5 resident blocks/SM, i.e. a 40%-occupied second wave).  The partial-sum buffer is
// sized for the upper bound GC_MAX_BLOCKS, which dcae_gc_num_partials reports.
static int64_t gc_blocks(int64_t rows, int64_t inner, int resident_per_sm = 8) {
  int64_t groups = rows * (inner / 4);
  int64_t b = (groups + GC_THREADS - 1) / GC_THREADS;
  const int64_t cap = (int64_t)num_sms() * resident_per_sm;
  if (b > cap) b = cap;
  if (b > GC_MAX_BLOCKS) b = GC_MAX_BLOCKS;
  if (b < 1) b = 1;
  return b;
}

template <int MODE, int LIK, bool IDX, bool POW2>
static int gc_resident() {
  static const int n = [] {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, gc_fused_kernel<MODE, LIK, IDX, POW2>, GC_THREADS, 0) != cudaSuccess || v < 1) v = 4;
    return v;
  }();
  return n;
}

}  // namespace dcae

extern "C" int64_t dcae_gc_num_partials(int64_t rows, int64_t inner) {
  int64_t b = (rows * (inner / 4) + dcae::GC_THREADS - 1) / dcae::GC_THREADS;     // upper bound of any launch's grid
  return b < 1 ? 1 : (b > dcae::GC_MAX_BLOCKS ? dcae::GC_MAX_BLOCKS : b);
}

extern "C" int dcae_gc_fused(const dcae_gc_args* a, void* stream) {
  using namespace dcae;
  DCAE_REQUIRE(a != nullptr, "dcae_gc_fused: null args");
  DCAE_REQUIRE(a->mode >= DCAE_GC_EVAL && a->mode <= DCAE_GC_DECODE, "dcae_gc_fused: bad mode %d", a->mode);
  DCAE_REQUIRE(a->rows >= 0 && a->inner >= 0 && a->inner % 4 == 0, "dcae_gc_fused: inner (%lld) must be a multiple of 4",
               (long long)a->inner);
  DCAE_REQUIRE(a->mu != nullptr, "dcae_gc_fused: mu is required");
  const bool need_y = a->mode != DCAE_GC_DECODE && (a->y_hat || a->y_hat16.hi || a->lik || a->sym || a->log2_partials);
  DCAE_REQUIRE(planes_ok(&a->y_hat16), "dcae_gc_fused: y_hat16 planes must be 8-byte aligned with ld %% 4 == 0");
  DCAE_REQUIRE(a->mode == DCAE_GC_DECODE ? a->sym_in != nullptr : (!need_y || a->y != nullptr), "dcae_gc_fused: missing input for mode %d", a->mode);
  DCAE_REQUIRE(a->mode != DCAE_GC_NOISE || a->noise != nullptr, "dcae_gc_fused: NOISE mode needs a noise tensor");
  const bool need_scale = a->idx != nullptr || ((a->lik != nullptr || a->log2_partials != nullptr) && a->mode != DCAE_GC_DECODE);
  DCAE_REQUIRE(!need_scale || a->scale != nullptr, "dcae_gc_fused: scale is required for lik/idx");
  DCAE_REQUIRE(a->idx == nullptr || (a->scale_table != nullptr && a->n_table >= 2 && a->n_table <= GC_MAX_TABLE),
               "dcae_gc_fused: idx needs a scale_table of 2..%d entries (update_scale_table not called?)", GC_MAX_TABLE);
#define GC_CHECK_PTR(ptr, ld)                                                                        \
  DCAE_REQUIRE((ptr) == nullptr || (aligned16(ptr) && (ld) % 4 == 0), "dcae_gc_fused: " #ptr " must be 16-byte aligned with ld %% 4 == 0")
  GC_CHECK_PTR(a->y, a->y_ld); GC_CHECK_PTR(a->mu, a->mu_ld); GC_CHECK_PTR(a->scale, a->scale_ld);
  GC_CHECK_PTR(a->noise, a->noise_ld); GC_CHECK_PTR(a->sym_in, a->sym_in_ld); GC_CHECK_PTR(a->y_hat, a->y_hat_ld);
  GC_CHECK_PTR(a->lik, a->lik_ld); GC_CHECK_PTR(a->sym, a->sym_ld); GC_CHECK_PTR(a->idx, a->idx_ld);
#undef GC_CHECK_PTR
  if (a->rows == 0 || a->inner == 0) return DCAE_OK;   // empty input: nothing to do
  GcParams p;
  p.a = *a;
  p.inner4 = (uint32_t)(a->inner / 4);
  p.groups = a->rows * (a->inner / 4);
  p.n_partials = (int32_t)dcae_gc_num_partials(a->rows, a->inner);
  const int n_tensors = (a->y && a->mode != DCAE_GC_DECODE) + 1 + (a->scale != nullptr) + (a->mode == DCAE_GC_NOISE) +
                        (a->mode == DCAE_GC_DECODE) + (a->y_hat != nullptr) + (a->lik != nullptr) + (a->sym != nullptr) + (a->idx != nullptr);
  ProfileScope prof(DCAE_PROF_GC, 4.0 * n_tensors * (double)a->rows * (double)a->inner, stream);
  const bool pow2 = (p.inner4 & (p.inner4 - 1)) == 0;
  p.shift = 0;
  while (pow2 && (1u << p.shift) < p.inner4) ++p.shift;
  DCAE_REQUIRE(a->lik_math == DCAE_GC_LIK_FAST || a->lik_math == DCAE_GC_LIK_REFERENCE, "dcae_gc_fused: bad lik_math %d", a->lik_math);
  const bool want_lik = (a->lik != nullptr || a->log2_partials != nullptr) && a->mode != DCAE_GC_DECODE && a->y != nullptr;
  const int lik = want_lik ? a->lik_math + 1 : 0;
  const bool idx = a->idx != nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  // hot configuration -> gc_eval_hot_kernel (same element functions, 32-bit offsets, no per-store pointer tests)
  {
    const bool both = a->sym && a->idx, neither = !a->sym && !a->idx;
    auto fits = [&](const void* ptr, int64_t ld) { return ptr == nullptr || (ld >= 0 && ld < (1ll << 32) && (a->rows - 1) * ld + a->inner < (1ll << 32)); };
    const bool hot = a->mode == DCAE_GC_EVAL && pow2 && a->y && a->scale && a->y_hat && (lik == 0 || a->lik) && (both || neither) &&
                     p.groups < (1ll << 31) && fits(a->y, a->y_ld) && fits(a->mu, a->mu_ld) && fits(a->scale, a->scale_ld) &&
                     fits(a->y_hat, a->y_hat_ld) && fits(a->lik, a->lik_ld) && fits(a->sym, a->sym_ld) && fits(a->idx, a->idx_ld) &&
                     fits(a->y_hat16.hi, a->y_hat16.ld) && !(lik == 0 && a->log2_partials);
    if (hot) {
      GcHotParams h;
      h.y = a->y; h.mu = a->mu; h.scale = a->scale; h.y_hat = a->y_hat; h.lik = a->lik; h.sym = a->sym; h.idx = a->idx;
      h.hi = static_cast<__half*>(a->y_hat16.hi); h.lo = static_cast<__half*>(a->y_hat16.lo);
      h.log2_partials = a->log2_partials; h.scale_table = a->scale_table;
      h.y_ld = (uint32_t)a->y_ld; h.mu_ld = (uint32_t)a->mu_ld; h.scale_ld = (uint32_t)a->scale_ld; h.y_hat_ld = (uint32_t)a->y_hat_ld;
      h.lik_ld = (uint32_t)a->lik_ld; h.sym_ld = (uint32_t)a->sym_ld; h.idx_ld = (uint32_t)a->idx_ld; h.p16_ld = (uint32_t)a->y_hat16.ld;
      h.groups = (uint32_t)p.groups; h.shift = p.shift; h.mask = p.inner4 - 1;
      h.n_table = a->n_table; h.n_partials = p.n_partials; h.scale_bound = a->scale_bound; h.lik_bound = a->lik_bound;
      const bool planes = a->y_hat16.hi != nullptr;
#define GC_HOT(L, S, P)                                                                                                   \
  do {                                                                                                                    \
    static const int res = [] {                                                                                           \
      int v = 0;                                                                                                          \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, gc_eval_hot_kernel<L, S, P>, GC_THREADS, 0) != cudaSuccess || v < 1) v = 4; \
      return v;                                                                                                           \
    }();                                                                                                                  \
    gc_eval_hot_kernel<L, S, P><<<(unsigned)gc_blocks(a->rows, a->inner, res), GC_THREADS, 0, st>>>(h);                  \
  } while (0)
#define GC_HOT_L(L)                                              \
  do {                                                           \
    if (both && planes) GC_HOT(L, true, true);                   \
    else if (both) GC_HOT(L, true, false);                       \
    else if (planes) GC_HOT(L, false, true);                     \
    else GC_HOT(L, false, false);                                \
  } while (0)
      if (lik == 2) GC_HOT_L(2);
      else if (lik == 1) GC_HOT_L(1);
      else GC_HOT_L(0);
#undef GC_HOT_L
#undef GC_HOT
      DCAE_LAUNCH_CHECK();
      return DCAE_OK;
    }
  }
#define GC_LAUNCH(M, L, I)                                                                                            \
  do {                                                                                                                \
    if (pow2) gc_fused_kernel<M, L, I, true><<<(unsigned)gc_blocks(a->rows, a->inner, gc_resident<M, L, I, true>()), GC_THREADS, 0, st>>>(p);    \
    else gc_fused_kernel<M, L, I, false><<<(unsigned)gc_blocks(a->rows, a->inner, gc_resident<M, L, I, false>()), GC_THREADS, 0, st>>>(p);      \
  } while (0)
#define GC_LAUNCH_MODE(M)                                                         \
  do {                                                                            \
    if (lik == 2 && idx) GC_LAUNCH(M, 2, true);                                   \
    else if (lik == 2) GC_LAUNCH(M, 2, false);                                    \
    else if (lik == 1 && idx) GC_LAUNCH(M, 1, true);                              \
    else if (lik == 1) GC_LAUNCH(M, 1, false);                                    \
    else if (idx) GC_LAUNCH(M, 0, true);                                          \
    else GC_LAUNCH(M, 0, false);                                                  \
  } while (0)
  if (a->mode == DCAE_GC_EVAL) GC_LAUNCH_MODE(DCAE_GC_EVAL);
  else if (a->mode == DCAE_GC_NOISE) GC_LAUNCH_MODE(DCAE_GC_NOISE);
  else {                                                   // DECODE never produces likelihoods
    if (idx) GC_LAUNCH(DCAE_GC_DECODE, 0, true);
    else GC_LAUNCH(DCAE_GC_DECODE, 0, false);
  }
#undef GC_LAUNCH_MODE
#undef GC_LAUNCH
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

extern "C" int dcae_reduce_partials(const float* partials, int64_t n, float* out, void* stream) {
  using namespace dcae;
  DCAE_REQUIRE(partials != nullptr && out != nullptr && n >= 0, "dcae_reduce_partials: bad arguments");
  reduce_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, n, out);
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

extern "C" int dcae_gc_backward(const dcae_gc_bwd_args* a, void* stream) {
  using namespace dcae;
  DCAE_REQUIRE(a && a->y && a->mu && a->scale && a->grad_lik, "dcae_gc_backward: null input");
  DCAE_REQUIRE(a->mode == DCAE_GC_EVAL || (a->mode == DCAE_GC_NOISE && a->noise), "dcae_gc_backward: mode must be EVAL or NOISE (with noise)");
  DCAE_REQUIRE(a->rows >= 0 && a->inner >= 0 && a->inner % 4 == 0, "dcae_gc_backward: inner must be a multiple of 4");
#define GCB_CHECK(ptr, ld) DCAE_REQUIRE((ptr) == nullptr || (aligned16(ptr) && (ld) % 4 == 0), "dcae_gc_backward: " #ptr " must be 16-byte aligned with ld %% 4 == 0")
  GCB_CHECK(a->y, a->y_ld); GCB_CHECK(a->mu, a->mu_ld); GCB_CHECK(a->scale, a->scale_ld); GCB_CHECK(a->noise, a->noise_ld);
  GCB_CHECK(a->grad_lik, a->grad_lik_ld); GCB_CHECK(a->grad_y, a->grad_y_ld); GCB_CHECK(a->grad_mu, a->grad_mu_ld); GCB_CHECK(a->grad_scale, a->grad_scale_ld);
#undef GCB_CHECK
  if (a->rows == 0 || a->inner == 0) return DCAE_OK;
  const int64_t groups = a->rows * (a->inner / 4);
  DCAE_REQUIRE(a->inner / 4 < (1ll << 32), "dcae_gc_backward: row too long");
  ProfileScope prof(DCAE_PROF_GC, 28.0 * (double)a->rows * (double)a->inner, stream);
  gc_backward_kernel<<<(unsigned)gc_blocks(a->rows, a->inner, 8), GC_THREADS, 0, (cudaStream_t)stream>>>(*a, groups, (uint32_t)(a->inner / 4));
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

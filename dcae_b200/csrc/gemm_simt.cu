// FP32 FFMA GEMM with the library's operand / epilogue contract (DCAE_MATH_FP32_SIMT).
// acc[T, N] = A[T, K] * W[N, K]^T where A is gathered from a token-major buffer (two column
// segments, optional 3x3 window = implicit GEMM for the conv stacks of dcae.py:584-611).
// It is the strict-fp32 arithmetic mode and the on-device cross-check of the tcgen05 path
// (gemm_tcgen05.cu); also used for the SIMT dictionary-attention core.
#include "common.cuh"

namespace dcae {

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

struct SimtGemmParams {
  dcae_operand a;
  dcae_epilogue e;
  const float* w;
  int M, N, K, Kc;  // Kc = k0 + k1 = channels per tap
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DCAE_ACT_GELU) return gelu_erf(v);
  if (act == DCAE_ACT_HALF_TANH) return 0.5f * tanhf(v);
  if (act == DCAE_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(const SimtGemmParams p) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread tile 8 (m) x 4 (n)

  // A loader: row = tid/2, 8 consecutive k starting at (tid%2)*8
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int am = m0 + a_row;
  int ab = 0, ay = 0, ax = 0;
  const bool a_row_ok = am < p.M;
  if (a_row_ok) {
    ax = am % p.a.w;
    ay = (am / p.a.w) % p.a.h;
    ab = am / (p.a.w * p.a.h);
  }
  // B loader: row = tid/4, 4 consecutive k at (tid%4)*4
  const int b_row = tid >> 2, b_k = (tid & 3) * 4;
  const bool b_row_ok = (n0 + b_row) < p.N;
  const float* wrow = p.w + (int64_t)(n0 + b_row) * p.K + b_k;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra0, ra1, rb;
  auto fetch = [&](int kt) {
    const int kk = kt * SG_BK;
    const int tap = kk / p.Kc;
    const int c = kk - tap * p.Kc + a_k;
    const int col = (c < p.a.k0) ? (p.a.col0 + c) : (p.a.col1 + (c - p.a.k0));
    int dy = 0, dx = 0;
    if (p.a.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    const int yy = ay + dy, xx = ax + dx;
    if (a_row_ok && (unsigned)yy < (unsigned)p.a.h && (unsigned)xx < (unsigned)p.a.w) {
      const float* src = p.a.base + ((int64_t)(ab * p.a.h + yy) * p.a.w + xx) * p.a.ld + col;
      ra0 = __ldg(reinterpret_cast<const float4*>(src));
      ra1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    } else {
      ra0 = make_float4(0.f, 0.f, 0.f, 0.f);
      ra1 = ra0;
    }
    rb = b_row_ok ? __ldg(reinterpret_cast<const float4*>(wrow + kk)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra0.x; As[buf][a_k + 1][a_row] = ra0.y;
    As[buf][a_k + 2][a_row] = ra0.z; As[buf][a_k + 3][a_row] = ra0.w;
    As[buf][a_k + 4][a_row] = ra1.x; As[buf][a_k + 5][a_row] = ra1.y;
    As[buf][a_k + 6][a_row] = ra1.z; As[buf][a_k + 7][a_row] = ra1.w;
    Bs[buf][b_k + 0][b_row] = rb.x; Bs[buf][b_k + 1][b_row] = rb.y;
    Bs[buf][b_k + 2][b_row] = rb.z; Bs[buf][b_k + 3][b_row] = rb.w;
  };

  const int KT = p.K / SG_BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < KT) stash(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  const dcae_epilogue& e = p.e;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f);
  if (e.bias) bias = __ldg(reinterpret_cast<const float4*>(e.bias + n));
  if (e.res_scale) rs = __ldg(reinterpret_cast<const float4*>(e.res_scale + n));
  const int act_cols = (e.act_cols <= 0 || e.act_cols > p.N) ? p.N : e.act_cols;
  const int act = (n < act_cols) ? e.act : DCAE_ACT_NONE;   // act_cols is a multiple of 4
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) break;
    float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
    if (e.addend) {
      const float4 ad = __ldg(reinterpret_cast<const float4*>(e.addend + (int64_t)m * e.addend_ld + n));
      v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
    }
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    if (e.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(e.residual + (int64_t)m * e.residual_ld + n));
      v.x = fmaf(r.x, rs.x, v.x); v.y = fmaf(r.y, rs.y, v.y); v.z = fmaf(r.z, rs.z, v.z); v.w = fmaf(r.w, rs.w, v.w);
    }
    *reinterpret_cast<float4*>(e.out + (int64_t)m * e.out_ld + n) = v;
  }
}

int gemm_simt(const dcae_operand* a, const dcae_weight* w, const dcae_epilogue* e, cudaStream_t s) {
  SimtGemmParams p;
  p.a = *a;
  p.e = *e;
  p.w = w->w;
  p.M = a->B * a->h * a->w;
  p.N = w->N;
  p.K = w->K;
  p.Kc = a->k0 + a->k1;
  DCAE_REQUIRE(w->w != nullptr, "gemm(simt): fp32 weight pointer is null");
  if (p.M == 0) return DCAE_OK;
  dim3 grid((p.M + SG_BM - 1) / SG_BM, (p.N + SG_BN - 1) / SG_BN);
  gemm_simt_kernel<<<grid, SG_THREADS, 0, s>>>(p);
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

// ---------------------------------------------------------------------------------------------
// SIMT dictionary attention core (dcae.py:489-501): one thread per (token, head); K_h and V_h of the
// head in shared memory (read as warp broadcasts); exact two-pass softmax in fp32.
// ---------------------------------------------------------------------------------------------
constexpr int AT_TOK = 128, AT_DICT = 128, AT_HD = 32;

__global__ void __launch_bounds__(AT_TOK) dict_attention_simt_kernel(const float* __restrict__ q, int64_t q_ld,
                                                                     const float* __restrict__ Kh,
                                                                     const float* __restrict__ Vh,
                                                                     const float* __restrict__ head_scale, int64_t T,
                                                                     float* __restrict__ out, int64_t out_ld) {
  __shared__ __align__(16) float Ks[AT_DICT][AT_HD];
  __shared__ __align__(16) float Vs[AT_DICT][AT_HD];
  const int head = blockIdx.y;
  const float4* kg = reinterpret_cast<const float4*>(Kh + (int64_t)head * AT_DICT * AT_HD);
  const float4* vg = reinterpret_cast<const float4*>(Vh + (int64_t)head * AT_DICT * AT_HD);
  for (int i = threadIdx.x; i < AT_DICT * AT_HD / 4; i += AT_TOK) {
    reinterpret_cast<float4*>(&Ks[0][0])[i] = __ldg(kg + i);
    reinterpret_cast<float4*>(&Vs[0][0])[i] = __ldg(vg + i);
  }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * AT_TOK + threadIdx.x;
  if (t >= T) return;
  const float sc = __ldg(head_scale + head);
  float qv[AT_HD];
  const float4* qr = reinterpret_cast<const float4*>(q + t * q_ld + head * AT_HD);
#pragma unroll
  for (int i = 0; i < AT_HD / 4; ++i) {
    const float4 v = __ldg(qr + i);
    qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
  }
  auto score = [&](int j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < AT_HD; c += 4) {
      const float4 k = *reinterpret_cast<const float4*>(&Ks[j][c]);
      s = fmaf(qv[c], k.x, s); s = fmaf(qv[c + 1], k.y, s); s = fmaf(qv[c + 2], k.z, s); s = fmaf(qv[c + 3], k.w, s);
    }
    return s * sc;   // sim = (q . k) * scale_h   (dcae.py:497-498)
  };
  float mx = -INFINITY;
  for (int j = 0; j < AT_DICT; ++j) mx = fmaxf(mx, score(j));
  float acc[AT_HD];
#pragma unroll
  for (int c = 0; c < AT_HD; ++c) acc[c] = 0.f;
  float l = 0.f;
  for (int j = 0; j < AT_DICT; ++j) {
    const float pj = expf(score(j) - mx);
    l += pj;
#pragma unroll
    for (int c = 0; c < AT_HD; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(&Vs[j][c]);
      acc[c] = fmaf(pj, v.x, acc[c]); acc[c + 1] = fmaf(pj, v.y, acc[c + 1]);
      acc[c + 2] = fmaf(pj, v.z, acc[c + 2]); acc[c + 3] = fmaf(pj, v.w, acc[c + 3]);
    }
  }
  const float inv = 1.0f / l;
  float4* orow = reinterpret_cast<float4*>(out + t * out_ld + head * AT_HD);
#pragma unroll
  for (int i = 0; i < AT_HD / 4; ++i)
    orow[i] = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
}

int dict_attention_simt(const float* q, int64_t q_ld, const float* Kh, const float* Vh, const float* head_scale,
                        int64_t T, float* out, int64_t out_ld, cudaStream_t s) {
  if (T == 0) return DCAE_OK;
  dim3 grid((unsigned)((T + AT_TOK - 1) / AT_TOK), 20);
  dict_attention_simt_kernel<<<grid, AT_TOK, 0, s>>>(q, q_ld, Kh, Vh, head_scale, T, out, out_ld);
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

}  // namespace dcae

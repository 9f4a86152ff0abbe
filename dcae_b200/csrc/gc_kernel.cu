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

__device__ __forceinline__ float gaussian_likelihood(float out, float mu, float s, float lik_bound) {
  const float c = -0.70710678118654752440f;  // float(-(2 ** -0.5)), dcae.py:855
  float v = fabsf(__fsub_rn(out, mu));
  float up = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(0.5f, v), s))));
  float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(-0.5f, v), s))));
  return nan_max(__fsub_rn(up, lo), lik_bound);
}

__device__ __forceinline__ int table_index(float s, const float* tbl, int n, float log_t0, float inv_step) {
  // idx = (n-1) - sum_{j<n-1} [s <= tbl[j]]  ==  #{ j < n-1 : tbl[j] < s }   (NaN -> n-1)
  // Log-domain guess j0, one branch-free step up or down, then a check of the far neighbour; the search
  // loop only runs for tables that are not log-uniform (exact for ANY sorted table).
  const float g = (__log2f(s) - log_t0) * inv_step;
  const int j0 = (int)fminf(fmaxf(g, 0.0f), (float)(n - 1));
  const bool up = (j0 < n - 1) && (tbl[j0] < s);
  const bool down = (j0 > 0) && !(tbl[max(j0 - 1, 0)] < s);     // never together with `up` (sorted table)
  int j = j0 + (int)up - (int)down;
  const float far = tbl[up ? min(j0 + 1, n - 1) : max(j0 - 2, 0)];
  const bool bad = up ? (j < n - 1 && far < s) : (down && j > 0 && !(far < s));
  if (bad) {
    while (j < n - 1 && tbl[j] < s) ++j;
    while (j > 0 && !(tbl[j - 1] < s)) --j;
  }
  return (s != s) ? n - 1 : j;
}

// One float4 group of every tensor lives at row * ld + col; `row, col` come from a shift when inner/4 is a
// power of two (token-major slices: inner = 64), from a 32-bit divide otherwise.
// MODE, LIK (likelihood wanted) and IDX (indexes wanted) are compile-time so the element loop is straight-line code.
template <int MODE, bool LIK, bool IDX, bool POW2>
__global__ void __launch_bounds__(GC_THREADS) gc_fused_kernel(const GcParams p) {
  __shared__ float tbl[GC_MAX_TABLE];
  __shared__ float red[GC_THREADS / 32];
  const dcae_gc_args& a = p.a;
  const int n = a.n_table;
  for (int i = threadIdx.x; i < n; i += GC_THREADS) tbl[i] = a.scale_table ? a.scale_table[i] : 0.0f;
  __syncthreads();
  float log_t0 = 0.f, inv_step = 0.f;
  if (IDX && n > 1) {
    log_t0 = __log2f(tbl[0]);
    inv_step = (float)(n - 1) / (__log2f(tbl[n - 1]) - log_t0);
  }
  constexpr int mode = MODE;
  constexpr bool want_lik = LIK, want_idx = IDX;
  const bool want_log2 = a.log2_partials != nullptr;
  const float scale_bound = a.scale_bound, lik_bound = a.lik_bound;
  float log2_acc = 0.f;

  for (int64_t g = (int64_t)blockIdx.x * GC_THREADS + threadIdx.x; g < p.groups;
       g += (int64_t)gridDim.x * GC_THREADS) {
    int64_t row, col;
    if (POW2) {
      row = g >> p.shift;
      col = (g & (int64_t)(p.inner4 - 1)) << 2;
    } else if (p.groups <= 0xffffffffll) {   // 32-bit divide on the common path
      const uint32_t r32 = (uint32_t)g / p.inner4;
      row = r32;
      col = (int64_t)((uint32_t)g - r32 * p.inner4) * 4;
    } else {
      row = g / p.inner4;
      col = (g - row * p.inner4) * 4;
    }
    const float4 mu4 = __ldg(reinterpret_cast<const float4*>(a.mu + row * a.mu_ld + col));
    float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = y4, nz4 = y4;
    int4 si4 = make_int4(0, 0, 0, 0);
    if (mode != DCAE_GC_DECODE && a.y != nullptr) y4 = __ldg(reinterpret_cast<const float4*>(a.y + row * a.y_ld + col));
    if (a.scale != nullptr) sc4 = __ldg(reinterpret_cast<const float4*>(a.scale + row * a.scale_ld + col));
    if (mode == DCAE_GC_NOISE) nz4 = __ldg(reinterpret_cast<const float4*>(a.noise + row * a.noise_ld + col));
    if (mode == DCAE_GC_DECODE) si4 = __ldg(reinterpret_cast<const int4*>(a.sym_in + row * a.sym_in_ld + col));

    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
    const float y[4] = {y4.x, y4.y, y4.z, y4.w};
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
    const float nz[4] = {nz4.x, nz4.y, nz4.z, nz4.w};
    const int si[4] = {si4.x, si4.y, si4.z, si4.w};
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
      lk[k] = want_lik ? gaussian_likelihood(out, mu[k], s, lik_bound) : 1.0f;
      ix[k] = want_idx ? table_index(s, tbl, n, log_t0, inv_step) : 0;
      if (want_log2) log2_acc += __log2f(lk[k]);   // bpp numerator: lik in [1e-9, 1], MUFU.LG2 is ample
    }
    if (a.y_hat) *reinterpret_cast<float4*>(a.y_hat + row * a.y_hat_ld + col) = make_float4(yh[0], yh[1], yh[2], yh[3]);
    if (a.y_hat16.hi) store_planes4(a.y_hat16, row, (int)col, make_float4(yh[0], yh[1], yh[2], yh[3]));
    if (a.lik && want_lik) *reinterpret_cast<float4*>(a.lik + row * a.lik_ld + col) = make_float4(lk[0], lk[1], lk[2], lk[3]);
    if (a.sym) *reinterpret_cast<int4*>(a.sym + row * a.sym_ld + col) = make_int4(sy[0], sy[1], sy[2], sy[3]);
    if (a.idx) *reinterpret_cast<int4*>(a.idx + row * a.idx_ld + col) = make_int4(ix[0], ix[1], ix[2], ix[3]);
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

template <int MODE, bool LIK, bool IDX, bool POW2>
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
  const bool lik = (a->lik != nullptr || a->log2_partials != nullptr) && a->mode != DCAE_GC_DECODE && a->y != nullptr;
  const bool idx = a->idx != nullptr;
  cudaStream_t st = (cudaStream_t)stream;
#define GC_LAUNCH(M, L, I)                                                                                            \
  do {                                                                                                                \
    if (pow2) gc_fused_kernel<M, L, I, true><<<(unsigned)gc_blocks(a->rows, a->inner, gc_resident<M, L, I, true>()), GC_THREADS, 0, st>>>(p);    \
    else gc_fused_kernel<M, L, I, false><<<(unsigned)gc_blocks(a->rows, a->inner, gc_resident<M, L, I, false>()), GC_THREADS, 0, st>>>(p);      \
  } while (0)
#define GC_LAUNCH_MODE(M)                                                         \
  do {                                                                            \
    if (lik && idx) GC_LAUNCH(M, true, true);                                     \
    else if (lik) GC_LAUNCH(M, true, false);                                      \
    else if (idx) GC_LAUNCH(M, false, true);                                      \
    else GC_LAUNCH(M, false, false);                                              \
  } while (0)
  if (a->mode == DCAE_GC_EVAL) GC_LAUNCH_MODE(DCAE_GC_EVAL);
  else if (a->mode == DCAE_GC_NOISE) GC_LAUNCH_MODE(DCAE_GC_NOISE);
  else GC_LAUNCH_MODE(DCAE_GC_DECODE);
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

// Hyper-latent entropy model (first part): compressai's EntropyBottleneck as the reference uses it for z
// (/root/reference/models/dcae.py:629-633 forward, :705-706 compress, :861 decompress) in one pass:
//   out    = round(z - median) + median   (EVAL)  |  z + noise (NOISE)  |  float(symbols) + median (DECODE)
//   lik    = max(sigmoid(L(out + 1/2)) - sigmoid(L(out - 1/2)), lik_bound)
//   L(v)   = the per-channel monotone network 1 -> 3 -> 3 -> 3 -> 3 -> 1 of the factorised prior (Balle et al. 2018):
//            h = softplus(M_i) h + b_i;  h += tanh(f_i) * tanh(h)  for i < 4
//   sym    = int32(round(z - median)),  z_hat = out
// softplus(M_i) and tanh(f_i) are constant per weight load: the host packs them once, 58 floats per channel.
// One block works on one (image, channel) row of h*w elements with the channel's 58 parameters in shared memory.
#include "common.cuh"

namespace dcae {

constexpr int EB_THREADS = 128;
constexpr int EB_PARAMS = 58;   // matrices 3 + 9 + 9 + 9 + 3, biases 3 + 3 + 3 + 3 + 1, factors 3 + 3 + 3 + 3
// offsets inside the packed row
constexpr int EB_M0 = 0, EB_M1 = 3, EB_M2 = 12, EB_M3 = 21, EB_M4 = 30, EB_B0 = 33, EB_B1 = 36, EB_B2 = 39, EB_B3 = 42, EB_B4 = 45,
              EB_F0 = 46, EB_F1 = 49, EB_F2 = 52, EB_F3 = 55;

__device__ __forceinline__ float eb_logits(float v, const float* p) {
  float h[3], g[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    h[k] = p[EB_M0 + k] * v + p[EB_B0 + k];
    h[k] += p[EB_F0 + k] * tanhf(h[k]);
  }
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
    const float* m = p + EB_M1 + 9 * layer;
    const float* b = p + EB_B1 + 3 * layer;
    const float* f = p + EB_F1 + 3 * layer;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      g[k] = m[3 * k] * h[0] + m[3 * k + 1] * h[1] + m[3 * k + 2] * h[2] + b[k];
      g[k] += f[k] * tanhf(g[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) h[k] = g[k];
  }
  return p[EB_M4] * h[0] + p[EB_M4 + 1] * h[1] + p[EB_M4 + 2] * h[2] + p[EB_B4];
}

__device__ __forceinline__ float eb_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(EB_THREADS) eb_fused_kernel(const dcae_eb_args a) {
  __shared__ float p[EB_PARAMS];
  const int row = blockIdx.x;                  // image * C + channel
  const int c = row % a.C;
  for (int i = threadIdx.x; i < EB_PARAMS; i += EB_THREADS) p[i] = a.params[(int64_t)c * EB_PARAMS + i];
  __syncthreads();
  const float med = a.medians[c];
  const int64_t base = (int64_t)row * a.HW;
  for (int64_t i = (int64_t)blockIdx.y * EB_THREADS + threadIdx.x; i < a.HW; i += (int64_t)gridDim.y * EB_THREADS) {
    float r, out;
    if (a.mode == DCAE_GC_DECODE) {
      r = (float)a.sym_in[base + i];
      out = r + med;
    } else {
      const float z = a.z[base + i];
      r = rintf(z - med);                      // torch.round: half to even
      out = a.mode == DCAE_GC_NOISE ? z + a.noise[base + i] : r + med;
    }
    if (a.sym) a.sym[base + i] = (int)r;
    if (a.z_hat) a.z_hat[base + i] = r + med;  // ste_round(z - median) + median (dcae.py:631-633) == the dequantised value
    if (a.lik) {
      const float lower = eb_logits(out - 0.5f, p), upper = eb_logits(out + 0.5f, p);
      const float lk = eb_sigmoid(upper) - eb_sigmoid(lower);
      a.lik[base + i] = (lk != lk) ? lk : fmaxf(lk, a.lik_bound);
    }
  }
}

}  // namespace dcae

extern "C" int dcae_eb_fused(const dcae_eb_args* a, void* stream) {
  using namespace dcae;
  DCAE_REQUIRE(a && a->params && a->medians, "dcae_eb_fused: null parameters");
  DCAE_REQUIRE(a->mode >= DCAE_GC_EVAL && a->mode <= DCAE_GC_DECODE, "dcae_eb_fused: bad mode %d", a->mode);
  DCAE_REQUIRE(a->B >= 0 && a->C > 0 && a->HW >= 0 && (int64_t)a->B * a->C < (1ll << 31), "dcae_eb_fused: bad shape");
  DCAE_REQUIRE(a->mode == DCAE_GC_DECODE ? a->sym_in != nullptr : a->z != nullptr, "dcae_eb_fused: missing input for mode %d", a->mode);
  DCAE_REQUIRE(a->mode != DCAE_GC_NOISE || a->noise != nullptr, "dcae_eb_fused: NOISE mode needs a noise tensor");
  DCAE_REQUIRE(a->mode != DCAE_GC_DECODE || a->lik == nullptr, "dcae_eb_fused: DECODE produces no likelihood");
  if (a->B == 0 || a->HW == 0) return DCAE_OK;
  ProfileScope prof(DCAE_PROF_OTHER, 0.0, stream);
  int64_t chunks = (a->HW + EB_THREADS - 1) / EB_THREADS;
  if (chunks > 64) chunks = 64;
  dim3 grid((unsigned)(a->B * a->C), (unsigned)chunks);
  eb_fused_kernel<<<grid, EB_THREADS, 0, (cudaStream_t)stream>>>(*a);
  DCAE_LAUNCH_CHECK();
  return DCAE_OK;
}

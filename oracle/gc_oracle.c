/* TEST INFRASTRUCTURE ONLY -- a checker, never linked, imported or shipped by the product (dcae_b200/).
 *
 * Plain-C restatement of the GaussianConditional element math of the reference's slice loop, independent of
 * torch (glibc erfcf instead of Sleef / libdevice), used by tests/test_oracle_c.py to cross-check
 * oracle/gaussian_conditional.py and, through it, kernel 3.  One function per reference call site:
 *
 *   gc_forward_eval   compressai GaussianConditional.forward in eval mode, /root/reference/models/dcae.py:657-659
 *                     (quantize "dequantize" = rint(y - mu) + mu; likelihood = in-tree copy dcae.py:839-857;
 *                     lower bounds 0.11 on the scale, 1e-9 on the likelihood, dcae.py:614, 846)
 *   gc_symbols        quantize(y, "symbols", mu), dcae.py:739      (round half to even, int32)
 *   gc_build_indexes  build_indexes(scale), dcae.py:738, 891:  idx = (n-1) - sum_{j<n-1} [max(scale, bound) <= table[j]]
 *   gc_dequantize     dequantize(symbols, mu), dcae.py:896
 *
 * Arithmetic is single precision in the reference's op order; compile with -ffp-contract=off (no FMA contraction)
 * so that every intermediate rounds as torch's eager fp32 ops do.  Parity pinned by tests/test_oracle_c.py against
 * the torch oracle, which is itself pinned bit for bit against the reference's in-tree likelihood
 * (tests/test_oracle_vs_reference.py).
 */
#include <math.h>
#include <stdint.h>

static float nan_max(float x, float bound) { return (x != x) ? x : (x > bound ? x : bound); } /* torch.max(x, bound) */

/* dcae.py:854-857  _standardized_cumulative(x) = 0.5 * erfc(-(2 ** -0.5) * x), evaluated here as the likelihood uses it */
static float likelihood(float out, float mu, float scale, float scale_bound, float lik_bound) {
  const float c = -0.70710678118654752440f;          /* float(-(2 ** -0.5)), dcae.py:855 */
  const float s = nan_max(scale, scale_bound);       /* dcae.py:846 lower_bound_scale */
  const float v = fabsf(out - mu);                   /* dcae.py:845, 847 */
  const float upper = 0.5f * erfcf(c * ((0.5f - v) / s));   /* dcae.py:848 */
  const float lower = 0.5f * erfcf(c * ((-0.5f - v) / s));  /* dcae.py:849 */
  return nan_max(upper - lower, lik_bound);          /* dcae.py:850 + likelihood_lower_bound */
}

void gc_forward_eval(const float* y, const float* mu, const float* scale, int64_t n, float scale_bound, float lik_bound,
                     float* y_hat, float* lik) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) {
    const float out = rintf(y[i] - mu[i]) + mu[i];   /* quantize "dequantize": round half to even, then + means */
    if (y_hat) y_hat[i] = out;
    if (lik) lik[i] = likelihood(out, mu[i], scale[i], scale_bound, lik_bound);
  }
}

void gc_symbols(const float* y, const float* mu, int64_t n, int32_t* sym) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) sym[i] = (int32_t)rintf(y[i] - mu[i]);
}

void gc_build_indexes(const float* scale, int64_t n, const float* table, int32_t n_table, float scale_bound, int32_t* idx) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) {
    const float s = nan_max(scale[i], scale_bound);
    int32_t k = n_table - 1;
    for (int32_t j = 0; j < n_table - 1; ++j) k -= (s <= table[j]);   /* NaN compares false: stays n_table - 1 */
    idx[i] = k;
  }
}

void gc_dequantize(const int32_t* sym, const float* mu, int64_t n, float* y_hat) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) y_hat[i] = (float)sym[i] + mu[i];
}

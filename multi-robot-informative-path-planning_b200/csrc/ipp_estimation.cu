// the Poisson log-likelihood of every particle of the importance sampler with progressive
// correction (src/estimation/estimation.py:15-55, evaluated for each of n_samples particles at :97, once per stage).
// After the GP and the scoring moved to the device this is the largest cost of a whole strategy run (85 % of an
// 8-agent run, tools/run_profile.py), all of it an N x n x M array expression.
//
//   rate_ij = lambda_b + sum_m A_im / max((ox_j - x_im)^2 + (oy_j - y_im)^2, 1e-6), clamped below at 1e-10
//   ll_i    = sum_j [ k_j log(rate_ij) - lgamma(k_j + 1) - rate_ij ]          (scipy.stats.poisson.logpmf)
//
// The weights exp(gamma * ll) feed np.random.choice(p = w), so the consumer needs to know how far this ll can be
// from the reference's: the rate is formed with the reference's IEEE operations in the reference's order (no FMA
// contraction: bit-identical to numpy), lgamma(k + 1) is supplied by the caller (scipy's own values, one per
// observation), and next to ll the kernel returns the two sums that bound the remaining differences --
//   S1_i = sum_j (|k_j log rate_ij| + lgamma_j + rate_ij)   (log ulp + the three roundings of each term)
//   S2_i = sum_j |term_ij|                                  (summation order)
// The host turns them into an error band on the resampling CDF and re-decides any draw that falls inside it
// (src/estimation/estimation.py in this package).
#include "ipp_common.cuh"
#include "../../include/ipp_b200.h"

namespace ipp {

constexpr int LL_THREADS = 256;
constexpr int LL_MAX_SOURCES = 32;

__global__ void __launch_bounds__(LL_THREADS) poisson_loglik_kernel(const double* __restrict__ particles, int M,
                                                                    const double* __restrict__ obs_xy,
                                                                    const double* __restrict__ counts,
                                                                    const double* __restrict__ lgam, int n,
                                                                    double lambda_b, double* __restrict__ out) {
  __shared__ double src[3 * LL_MAX_SOURCES];
  __shared__ double red[3][LL_THREADS / 32];
  const int i = blockIdx.x, t = threadIdx.x;
  if (t < 3 * M) src[t] = particles[(size_t)i * 3 * M + t];
  __syncthreads();
  double ll = 0.0, s1 = 0.0, s2 = 0.0;
  for (int j = t; j < n; j += LL_THREADS) {
    const double2 o = reinterpret_cast<const double2*>(obs_xy)[j];
    // np.sum over the (contiguous, short) source axis in numpy's order: left to right below 8 sources, else
    // 8 interleaved accumulators combined as ((0+1)+(2+3))+((4+5)+(6+7)) and the remainder added one by one
    auto share = [&](int m) {
      const double dx = __dsub_rn(o.x, src[3 * m]), dy = __dsub_rn(o.y, src[3 * m + 1]);
      const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      return __ddiv_rn(src[3 * m + 2], fmax(d2, 1e-6));
    };
    double s = 0.0;
    if (M < 8) {
      for (int m = 0; m < M; ++m) s = __dadd_rn(s, share(m));
    } else {
      double r[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = share(q);
      int m = 8;
      for (; m < M - (M % 8); m += 8)
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], share(m + q));
      s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
      for (; m < M; ++m) s = __dadd_rn(s, share(m));
    }
    const double rate = fmax(__dadd_rn(lambda_b, s), 1e-10);
    const double k = counts[j], g = lgam[j];
    // scipy _discrete_distns.py poisson._logpmf: xlogy(k, mu) - gamln(k + 1) - mu, -inf outside the support
    const double kl = k == 0.0 ? 0.0 : __dmul_rn(k, log(rate));
    const double term = k < 0.0 ? -INFINITY : __dsub_rn(__dsub_rn(kl, g), rate);
    ll += term;
    s1 += fabs(kl) + fabs(g) + rate;
    s2 += fabs(term);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ll += __shfl_xor_sync(0xffffffffu, ll, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((t & 31) == 0) {
    red[0][t >> 5] = ll;
    red[1][t >> 5] = s1;
    red[2][t >> 5] = s2;
  }
  __syncthreads();
  if (t < 3) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < LL_THREADS / 32; ++w) a += red[t][w];
    out[3 * (size_t)i + t] = a;
  }
}

}  // namespace ipp

using namespace ipp;

extern "C" int ipp_poisson_loglik(const double* particles, int N, int M, const double* obs_xy, const double* counts,
                                  const double* lgam, int n, double lambda_b, double* out, void* stream) {
  if (N < 0 || M <= 0 || M > LL_MAX_SOURCES || n < 0 || !out || (N > 0 && !particles) ||
      (n > 0 && (!obs_xy || !counts || !lgam)))
    return IPP_ERR_ARG;
  if (N == 0) return IPP_OK;
  poisson_loglik_kernel<<<N, LL_THREADS, 0, (cudaStream_t)stream>>>(particles, M, obs_xy, counts, lgam, n, lambda_b, out);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

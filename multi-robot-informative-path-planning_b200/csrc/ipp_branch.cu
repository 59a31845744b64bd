// Information-gain scoring of RRT candidate branches and leaves (SURVEY.md §8a rows I1-I8, T, and
// the selector half of G6).  Replaces the per-node Python gain functions of src/rrt/rrt.py:32-259,
// the penalty loop of bias_beta_path_selection (rrt.py:460-482), mutual_information_gain
// (rrt.py:512-526) and the visited-point penalty of rrt_thread.py:383-387.
//
// Design: the reference re-walks node -> root for every inserted node and recomputes every per-node
// term (including an O(#observations) scan) at every step of every walk.  Here the work is split:
//   phase 1  per-NODE static terms, once per node      (ipp_node_terms, ipp_count_close)
//   phase 2  per-BRANCH walk over versioned parent pointers, summing term(node, parent-at-time)
// "Versioned" because rewire (rrt_utils.py:85-88) edits parent pointers while the tree grows and
// node.information is computed at insertion time and never refreshed (rrt.py:321-323): branch b is
// evaluated against the parents as they were when node eval_time[b] was inserted.
//
// Floating point: np.linalg.norm of a 2-vector is sqrt(fma(y, y, x*x)) with OpenBLAS' ddot kernel
// (verified bit-for-bit in the build container); norm2() mirrors that so the `< d_wp` tests agree.
// Pairs whose squared distance is within a few ulp of the threshold are reported as "ambiguous"
// and re-decided on the host with numpy itself.
#include "ipp_common.cuh"
#include "../../include/ipp_b200.h"

namespace ipp {

__device__ __forceinline__ double norm2(double dx, double dy) { return sqrt(__fma_rn(dy, dy, __dmul_rn(dx, dx))); }

// ---- close-observation counts (exploitation penalty input: rrt.py:246-252, 471-477) ---------------
// One warp per point; observations streamed through shared memory in coalesced tiles.
__global__ void __launch_bounds__(256) count_close_kernel(const double* __restrict__ P, int np,
                                                          const double* __restrict__ obs, int no, double d,
                                                          int* __restrict__ sure, int* __restrict__ amb,
                                                          int* __restrict__ amb_total) {
  __shared__ double so[2][512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = blockIdx.x * 8 + warp;
  double px = 0.0, py = 0.0;
  if (pi < np) {
    px = P[2 * pi];
    py = P[2 * pi + 1];
  }
  int ns = 0, na = 0;
  // r < d is decided on squared distances (no square root per pair): outside a band of ~18 ulp around d^2 the
  // comparison of the squares equals the comparison numpy makes on the correctly rounded roots
  const double d2 = d * d, band = 4e-15 * d2;
  for (int o0 = 0; o0 < no; o0 += 512) {
    __syncthreads();
    for (int e = threadIdx.x; e < 512; e += 256) {
      int oi = o0 + e;
      so[0][e] = oi < no ? obs[2 * oi] : 0.0;
      so[1][e] = oi < no ? obs[2 * oi + 1] : 0.0;
    }
    __syncthreads();
    const int lim = min(512, no - o0);
    for (int e = lane; e < lim; e += 32) {
      const double dx = px - so[0][e], dy = py - so[1][e];
      const double r2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
      if (fabs(r2 - d2) <= band)
        ++na;
      else if (r2 < d2)
        ++ns;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ns += __shfl_xor_sync(0xffffffffu, ns, o);
    na += __shfl_xor_sync(0xffffffffu, na, o);
  }
  if (pi < np && lane == 0) {
    sure[pi] = ns;
    amb[pi] = na;
    if (na > 0) atomicAdd(amb_total, na);
  }
}

// Same counts with the observations spatially grouped: obs is ordered so that every CC_TILE consecutive rows are a
// compact cluster (the host sorts them along a Z-order curve; the count does not depend on the order) and
// tile_bbox[t] = {xmin, xmax, ymin, ymax} of tile t.  A point only visits the tiles whose box is within d of it:
// with d_wp of a few units in a 100 x 100 workspace that is 1-2 % of the observations.  One warp per point:
// lanes test 32 boxes at a time (ballot), then walk the set bits; inside a tile each lane takes CC_TILE / 32 rows.
constexpr int CC_TILE = 64;
__global__ void __launch_bounds__(256) count_close_tiled_kernel(const double* __restrict__ P, int np,
                                                                const double* __restrict__ obs, int no,
                                                                const double* __restrict__ tile_bbox, double d,
                                                                int* __restrict__ sure, int* __restrict__ amb,
                                                                int* __restrict__ amb_total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = blockIdx.x * 8 + warp;
  if (pi >= np) return;
  const double px = P[2 * pi], py = P[2 * pi + 1];
  const double d2 = d * d, band = 4e-15 * d2, reach2 = d2 * (1.0 + 1e-9);
  const int ntile = (no + CC_TILE - 1) / CC_TILE;
  int ns = 0, na = 0;
  for (int t0 = 0; t0 < ntile; t0 += 32) {
    const int tl = t0 + lane;
    bool hit = false;
    if (tl < ntile) {
      const double4 b = *reinterpret_cast<const double4*>(tile_bbox + 4 * tl);
      const double gx = fmax(0.0, fmax(b.x - px, px - b.y)), gy = fmax(0.0, fmax(b.z - py, py - b.w));
      hit = gx * gx + gy * gy <= reach2;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {
      const int tb = (t0 + __ffs(m) - 1) * CC_TILE;
      m &= m - 1;
#pragma unroll
      for (int q = 0; q < CC_TILE / 32; ++q) {
        const int oi = tb + lane + 32 * q;
        if (oi < no) {
          const double2 o = *reinterpret_cast<const double2*>(obs + 2 * oi);
          const double dx = px - o.x, dy = py - o.y;
          const double r2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
          if (fabs(r2 - d2) <= band)
            ++na;
          else if (r2 < d2)
            ++ns;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ns += __shfl_xor_sync(0xffffffffu, ns, o);
    na += __shfl_xor_sync(0xffffffffu, na, o);
  }
  if (lane == 0) {
    sure[pi] = ns;
    amb[pi] = na;
    if (na > 0) atomicAdd(amb_total, na);
  }
}

// ---- suppression factor (rrt.py:32-47) --------------------------------------------------------
__device__ __forceinline__ double suppression_factor(double xt, double yt, const double* __restrict__ src, int k, int S) {
  const double xk = src[3 * k], yk = src[3 * k + 1], ak = src[3 * k + 2];
  double F = 0.0;
  for (int j = 0; j < S; ++j) {
    const double xj = src[3 * j], yj = src[3 * j + 1], aj = src[3 * j + 2];
    if (xj == xk && yj == yk && aj == ak) continue;  // np.array_equal(source, other_source)
    const double mx = __ddiv_rn(__dadd_rn(xk, xj), 2.0), my = __ddiv_rn(__dadd_rn(yk, yj), 2.0);
    const double dd = norm2(xt - mx, yt - my);
    const double C = 2.0 - 1.0 / (1.0 + exp((dd - 2.0) / 16.0));
    F += (1.0 - C) + C / (1.0 + exp((2.0 - dd) / 16.0));
  }
  return F;
}

// ---- phase 1: per-node static terms ---------------------------------------------------------
// variant 0: no_penalties (closest source only, rrt.py:51-70)     1: distance (rrt.py:72-96)
//         2: distance+rotation (rrt.py:98-128)                    3: all (rrt.py:130-259)
__global__ void __launch_bounds__(128) node_terms_kernel(int variant, const double* __restrict__ node_xy, int N,
                                                         const double* __restrict__ src, int S, int closest,
                                                         const double* __restrict__ ws4, const double* __restrict__ obst,
                                                         int n_obst, double* __restrict__ src_gain,
                                                         double* __restrict__ wpen) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N) return;
  const double xt = node_xy[2 * q], yt = node_xy[2 * q + 1];
  double gsum = 0.0;
  if (S > 0) {
    if (variant == 0) {
      const double dsrc = norm2(xt - src[3 * closest], yt - src[3 * closest + 1]);
      const double F = suppression_factor(xt, yt, src, closest, S);
      const double u = dsrc - 2.0;
      gsum += (1.0 + exp(-(u * u) / 2.0 * 16.0)) * F;
    } else {
      for (int k = 0; k < S; ++k) {
        const double dsrc = norm2(xt - src[3 * k], yt - src[3 * k + 1]);
        const double F = suppression_factor(xt, yt, src, k, S);
        const double u = dsrc - 2.0;
        if (variant == 3)
          gsum += exp(-(u * u) / 32.0) * F;
        else
          gsum += (1.0 + exp(-(u * u) / 2.0 * 16.0)) * F;
      }
    }
  }
  src_gain[q] = gsum;
  if (wpen) {
    double w = 0.0;
    if (xt < ws4[0] || xt > ws4[1] || yt < ws4[2] || yt > ws4[3]) w = INFINITY;  // rrt.py:179-182
    for (int o = 0; o < n_obst; ++o) {                                               // rrt.py:185-192
      const double x0 = obst[4 * o], y0 = obst[4 * o + 1];
      const double x1 = x0 + obst[4 * o + 2], y1 = y0 + obst[4 * o + 3];
      if (x0 <= xt && xt <= x1 && y0 <= yt && yt <= y1) w = INFINITY;
    }
    wpen[q] = w;
  }
}

// segment/segment test of rrt.py:221-237
__device__ __forceinline__ bool seg_hit(double x1, double y1, double x2, double y2, double x3, double y3, double x4,
                                        double y4) {
  const double d1x = x2 - x1, d1y = y2 - y1, d2x = x4 - x3, d2y = y4 - y3;
  const double den = __dsub_rn(__dmul_rn(d1x, d2y), __dmul_rn(d1y, d2x));
  if (den == 0.0) return false;
  const double t = __ddiv_rn(__dsub_rn(__dmul_rn(x3 - x1, d2y), __dmul_rn(y3 - y1, d2x)), den);
  const double u = __ddiv_rn(__dsub_rn(__dmul_rn(x3 - x1, d1y), __dmul_rn(y3 - y1, d1x)), den);
  return 0.0 <= t && t <= 1.0 && 0.0 <= u && u <= 1.0;
}

__device__ __forceinline__ int parent_at(int q, int T, const int* __restrict__ parent0, const int* __restrict__ hoff,
                                         const int* __restrict__ htime, const int* __restrict__ hpar) {
  int p = parent0[q];
  if (hoff) {
    for (int e = hoff[q]; e < hoff[q + 1]; ++e) {
      if (htime[e] < T) p = hpar[e];
      else break;
    }
  }
  return p;
}

// ---- per-step term of a branch walk: node q reached from parent p (rrt.py:65-70, 91-96, 123-128, 254-259) -------
struct TermArgs {
  int variant;
  const double* node_xy;
  const double* src_gain;
  const double* wpen;
  const int* n_close;
  int agent_has_obs, have_estimates;
  const double* obst;
  int n_obst;
};
__device__ __forceinline__ double branch_term(const TermArgs& a, int q, int p) {
  const double qx = a.node_xy[2 * q], qy = a.node_xy[2 * q + 1];
  const double px = a.node_xy[2 * p], py = a.node_xy[2 * p + 1];
  if (a.variant == 3) {
    double w = a.wpen[q];
    for (int o = 0; o < a.n_obst && w == 0.0; ++o) {  // edge parent -> node against the 4 sides, rrt.py:203-217
      const double x0 = a.obst[4 * o], y0 = a.obst[4 * o + 1];
      const double x1 = x0 + a.obst[4 * o + 2], y1 = y0 + a.obst[4 * o + 3];
      if (seg_hit(px, py, qx, qy, x0, y0, x0, y1) || seg_hit(px, py, qx, qy, x1, y0, x1, y1) ||
          seg_hit(px, py, qx, qy, x0, y0, x1, y0) || seg_hit(px, py, qx, qy, x0, y1, x1, y1))
        w = INFINITY;
    }
    if (!a.have_estimates) return -w;
    const double r = norm2(qx - px, qy - py);
    const double dp = exp(0.5 * (r * r));
    double ep = 0.0;
    if (a.agent_has_obs) {
      const int nc = a.n_close[q];
      ep = exp(0.05 * (double)((long long)nc * nc));
    }
    return ((a.src_gain[q] - dp) - ep) - w;
  }
  if (!a.have_estimates) return 0.0;
  if (a.variant == 0) return a.src_gain[q];
  const double dx = qx - px, dy = qy - py;
  double term = a.src_gain[q] * exp(0.5 * norm2(dx, dy));
  if (a.variant == 2) {
    const double th = atan2(dy, dx);
    term = term * exp(0.05 * (th * th) / 0.1);
  }
  return term;
}

// Phase 1b: the term of every node against its INSERTION-TIME parent, once per node.  Almost every step of every
// walk uses exactly this value (a step differs only where a rewire changed the node's parent before the branch
// was scored), so phase 2 becomes a gather-sum of 8-byte terms along parent indices instead of two exp() per step.
__global__ void __launch_bounds__(128) node_term0_kernel(TermArgs a, int N, const int* __restrict__ parent0,
                                                         double* __restrict__ term0) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N) return;
  const int p = parent0[q];
  term0[q] = p >= 0 ? branch_term(a, q, p) : 0.0;
}

// ---- phase 2: branch walks -------------------------------------------------------------------
__global__ void __launch_bounds__(128) branch_gain_kernel(TermArgs a, int N, const int* __restrict__ parent0,
                                                          const int* __restrict__ hoff, const int* __restrict__ htime,
                                                          const int* __restrict__ hpar, const double* __restrict__ term0,
                                                          const int* __restrict__ eval_node,
                                                          const int* __restrict__ eval_time, int B,
                                                          double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int q = eval_node[b];
  const int T = eval_time ? eval_time[b] : 0x7fffffff;
  double sum = 0.0;
  for (int guard = 0; guard < N; ++guard) {
    const int p0 = parent0[q];
    const int p = hoff ? parent_at(q, T, parent0, hoff, htime, hpar) : p0;
    if (p < 0) break;
    sum += (p == p0) ? term0[q] : branch_term(a, q, p);  // same value, same summation order as the inline form
    q = p;
  }
  if (a.variant != 3) sum = fmax(sum, 0.0);  // max(final_gain, 0); rrt.py:70,96,128
  out[b] = sum;
}

// ---- stage-1 selector scores (rrt.py:456-482): ((MI - dist) - workspace) - exploitation --------
__global__ void __launch_bounds__(128) leaf_penalty_kernel(const double* __restrict__ leaf_xy,
                                                           const double* __restrict__ parent_xy,
                                                           const int* __restrict__ has_parent, int L,
                                                           const int* __restrict__ n_close, int agent_has_obs,
                                                           const double* __restrict__ ws4, double mi,
                                                           double* __restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double x = leaf_xy[2 * i], y = leaf_xy[2 * i + 1];
  double dp = 0.0;
  if (has_parent[i]) {
    const double r = norm2(x - parent_xy[2 * i], y - parent_xy[2 * i + 1]);
    dp = exp(0.5 * (r * r));
  }
  const double w = (x < ws4[0] || x > ws4[1] || y < ws4[2] || y > ws4[3]) ? INFINITY : 0.0;
  double ep = 0.0;
  if (agent_has_obs) {
    const int nc = n_close[i];
    ep = exp(0.05 * (double)((long long)nc * nc));
  }
  scores[i] = ((mi - dp) - w) - ep;
}

// ---- UCB visited-point penalty (rrt_thread.py:383-387), subtraction order = path order ----------
__global__ void __launch_bounds__(128) ucb_visit_penalty_kernel(const double* __restrict__ leaf_xy, int L,
                                                                const double* __restrict__ path_xy, int P, double d,
                                                                double* __restrict__ acq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const double x = leaf_xy[2 * i], y = leaf_xy[2 * i + 1];
  double a = acq[i];
  for (int k = 0; k < P; ++k) {
    const double r = norm2(x - path_xy[2 * k], y - path_xy[2 * k + 1]);
    if (r < d) a -= 0.1 * (d - r) / d;
  }
  acq[i] = a;
}

// ---- np.argmax semantics: first maximum, a NaN anywhere wins at its first position --------------
__global__ void __launch_bounds__(1024) argmax_first_kernel(const double* __restrict__ v, int n, double* __restrict__ best_val,
                                                            int* __restrict__ best_idx) {
  __shared__ double sv[32];
  __shared__ int si[32];
  const int t = threadIdx.x;
  double bv = -INFINITY;
  int bi = 0x7fffffff;
  bool bnan = false;
  for (int i = t; i < n; i += 1024) {
    const double x = v[i];
    const bool xn = isnan(x);
    if (bnan) continue;  // this thread already holds its earliest NaN (indices ascend)
    if (xn) {
      bnan = true;
      bv = x;
      bi = i;
    } else if (bi == 0x7fffffff || x > bv) {
      bv = x;
      bi = i;
    }
  }
  // combine (value, index, nan) : NaN beats everything, then larger value, then smaller index
  auto better = [](double av, int ai, double cv, int ci) {
    if (ai == 0x7fffffff) return false;
    if (ci == 0x7fffffff) return true;
    const bool an = isnan(av), cn = isnan(cv);
    if (an != cn) return an;
    if (an && cn) return ai < ci;
    if (av != cv) return av > cv;
    return ai < ci;
  };
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if ((t & 31) == 0) {
    sv[t >> 5] = bv;
    si[t >> 5] = bi;
  }
  __syncthreads();
  if (t < 32) {
    bv = sv[t];
    bi = si[t];
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (t == 0) {
      best_val[0] = bv;
      best_idx[0] = (bi == 0x7fffffff) ? -1 : bi;
    }
  }
}

}  // namespace ipp

using namespace ipp;

extern "C" {

int ipp_count_close(const double* P, int np, const double* obs, int no, double d, int* sure, int* amb, int* amb_total,
                    const double* tile_bbox, void* stream) {
  if (np < 0 || no < 0 || !sure || !amb || !amb_total) return IPP_ERR_ARG;
  if (np == 0) return IPP_OK;
  if (!P || (no > 0 && !obs)) return IPP_ERR_ARG;
  if (tile_bbox && no > 0)
    count_close_tiled_kernel<<<(np + 7) / 8, 256, 0, (cudaStream_t)stream>>>(P, np, obs, no, tile_bbox, d, sure, amb,
                                                                             amb_total);
  else
    count_close_kernel<<<(np + 7) / 8, 256, 0, (cudaStream_t)stream>>>(P, np, obs, no, d, sure, amb, amb_total);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_node_terms(int variant, const double* node_xy, int N, const double* sources, int S, int closest_idx,
                   const double* ws4, const double* obstacles, int n_obst, double* src_gain, double* wpen,
                   void* stream) {
  if (variant < 0 || variant > 3 || N < 0 || S < 0 || !src_gain) return IPP_ERR_ARG;
  if (N == 0) return IPP_OK;
  if (!node_xy || (S > 0 && !sources) || (variant == 3 && (!wpen || !ws4)) || (n_obst > 0 && !obstacles))
    return IPP_ERR_ARG;
  if (variant == 0 && S > 0 && (closest_idx < 0 || closest_idx >= S)) return IPP_ERR_ARG;
  node_terms_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(variant, node_xy, N, sources, S, closest_idx, ws4,
                                                                       obstacles, n_obst, src_gain,
                                                                       variant == 3 ? wpen : nullptr);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_branch_gain(int variant, const double* node_xy, int N, const int* parent0, const int* hist_off,
                    const int* hist_time, const int* hist_parent, const double* src_gain, const double* wpen,
                    const int* n_close, int agent_has_obs, int have_estimates, const double* obstacles, int n_obst,
                    const int* eval_node, const int* eval_time, int B, double* term0, double* gain_out, void* stream) {
  if (variant < 0 || variant > 3 || N <= 0 || B < 0 || !node_xy || !parent0 || !src_gain || !gain_out || !term0)
    return IPP_ERR_ARG;
  if (B == 0) return IPP_OK;
  if (!eval_node || (variant == 3 && (!wpen || (agent_has_obs && have_estimates && !n_close)))) return IPP_ERR_ARG;
  if (n_obst > 0 && !obstacles) return IPP_ERR_ARG;
  TermArgs a{variant, node_xy, src_gain, wpen, n_close, agent_has_obs, have_estimates, obstacles, n_obst};
  cudaStream_t st = (cudaStream_t)stream;
  node_term0_kernel<<<(N + 127) / 128, 128, 0, st>>>(a, N, parent0, term0);
  branch_gain_kernel<<<(B + 127) / 128, 128, 0, st>>>(a, N, parent0, hist_off, hist_time, hist_parent, term0, eval_node,
                                                      eval_time, B, gain_out);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_leaf_penalty(const double* leaf_xy, const double* parent_xy, const int* has_parent, int L, const int* n_close,
                     int agent_has_obs, const double* ws4, double mi, double* scores, void* stream) {
  if (L < 0 || !scores) return IPP_ERR_ARG;
  if (L == 0) return IPP_OK;
  if (!leaf_xy || !parent_xy || !has_parent || !ws4 || (agent_has_obs && !n_close)) return IPP_ERR_ARG;
  leaf_penalty_kernel<<<(L + 127) / 128, 128, 0, (cudaStream_t)stream>>>(leaf_xy, parent_xy, has_parent, L, n_close,
                                                                         agent_has_obs, ws4, mi, scores);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_ucb_visit_penalty(const double* leaf_xy, int L, const double* path_xy, int P, double d, double* acq,
                          void* stream) {
  if (L < 0 || P < 0 || !acq) return IPP_ERR_ARG;
  if (L == 0 || P == 0) return IPP_OK;
  if (!leaf_xy || !path_xy) return IPP_ERR_ARG;
  ucb_visit_penalty_kernel<<<(L + 127) / 128, 128, 0, (cudaStream_t)stream>>>(leaf_xy, L, path_xy, P, d, acq);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_argmax_first(const double* v, int n, double* best_val, int* best_idx, void* stream) {
  if (n <= 0 || !v || !best_val || !best_idx) return IPP_ERR_ARG;
  argmax_first_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(v, n, best_val, best_idx);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

}  // extern "C"

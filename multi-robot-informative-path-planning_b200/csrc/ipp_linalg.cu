// GP fit side of the hot path: RBF Gram build, blocked fp64 Cholesky,
// explicit triangular inverse W = L^-1, fused log-marginal-likelihood + gradient.
//
// Replaces, for the reference's GaussianProcessRegressor.fit / log_marginal_likelihood
// (sklearn/gaussian_process/_gpr.py:233-368, 541-656, reached from
// src/point_source/point_source.py:337,349): kernel(X, eval_gradient=True), scipy cholesky,
// cho_solve(L, y), cho_solve(L, I) and the einsum trace.
//
// Data layout (all row-major fp64, device resident, owned by the caller):
//   NP   = round_up(n, 64)         padded order; rows/cols >= n carry an identity block
//   Kbuf = (NP + 64) x NP          Gram -> Cholesky factor in place; row NP holds y, and because
//                                  the factorisation treats it as one more row block it becomes
//                                  z = L^-1 y for free (augmented-row trick)
//   Wbuf = NP x NP                 W = L^-1 (lower), upper part zero
//   Tbuf = NP x NP                 scratch for the recursive-doubling inverse
//   vbuf = vectors + partials      see ipp_gp_layout()
//   hyp  = {c, l, jitter}          in DEVICE memory so a captured CUDA graph replays for any theta
#include "ipp_gemm.cuh"
#include "ipp_chol_plan.h"
#include "../../include/ipp_b200.h"
#include <cooperative_groups.h>
#include <cstdlib>
namespace cg = cooperative_groups;

namespace ipp {

struct GpDims {
  int n, np, rows;
  long ld;
  long npv;       // length of one vector slot
  long max_tiles; // partial-sum slots
};
__host__ __device__ inline GpDims gp_dims(int n) {
  GpDims d;
  d.n = n;
  d.np = round_up(n > 0 ? n : 1, BLK);
  d.rows = d.np + BLK;
  d.ld = d.np;
  d.npv = d.np + BLK;
  long t = d.np / BLK;
  d.max_tiles = t * (t + 1) / 2;
  return d;
}
// vbuf slots
enum { V_XS = 0, V_YS = 1, V_Y = 2, V_W = 3, V_ALPHA = 4, V_NSLOTS = 5 };
__host__ __device__ inline long v_partials_off(const GpDims& d) { return (long)V_NSLOTS * d.npv; }
__host__ __device__ inline long v_result_off(const GpDims& d) { return v_partials_off(d) + 2 * d.max_tiles; }
__host__ __device__ inline long v_total(const GpDims& d) { return v_result_off(d) + 8; }

// A batch of R independent problems of the same size in slot-strided buffers (the optimiser restarts of one fit run
// in lockstep: one launch sequence evaluates the objective of every restart that is still iterating).  Problem r uses
// slot slots[r] (r itself when slots is null); hyp is [slot][4].
struct Bat {
  const int* slots;
  long sK, sW, sT, sV;  // slot strides in doubles
  __device__ __forceinline__ long slot(int r) const { return slots ? (long)slots[r] : (long)r; }
};

// ---------------------------------------------------------------------------------------------
// Gram build into the padded factor buffer.  K_ij = c * exp(-0.5 * |x_i/l - x_j/l|^2), exact c on
// the diagonal (sklearn kernels.py:1561-1565) plus jitter (_gpr.py:589); identity padding; y row.
// Writes 8 * (NP+64) * NP bytes, reads 24 n: HBM-write bound.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_fit_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                       const double* __restrict__ hyp, double* __restrict__ K,
                                                       double* __restrict__ vbuf, GpDims d, Bat bat) {
  __shared__ double sx[2][BLK], sy[2][BLK];
  const int bj = blockIdx.x, bi = blockIdx.y;
  const int t = threadIdx.x;
  {
    const long slot = bat.slot(blockIdx.z);
    hyp += 4 * slot;
    K += slot * bat.sK;
    vbuf += slot * bat.sV;
  }
  const double c = hyp[0], l = hyp[1], jit = hyp[2];
  const int i0 = bi * BLK, j0 = bj * BLK;
  if (t < 2 * BLK) {
    int which = t / BLK, e = t % BLK;
    int idx = (which ? j0 : i0) + e;
    double vx = 0.0, vy = 0.0;
    if (idx < d.n) {
      vx = __ddiv_rn(X[2 * idx], l);
      vy = __ddiv_rn(X[2 * idx + 1], l);
    }
    sx[which][e] = vx;
    sy[which][e] = vy;
    if (which == 1 && bi == bj && idx < d.npv) {  // diagonal blocks publish the scaled coordinates
      vbuf[V_XS * d.npv + idx] = vx;
      vbuf[V_YS * d.npv + idx] = vy;
    }
  }
  if (bi == 0 && bj == 0 && t == 0) vbuf[v_result_off(d) + 3] = 0.0;  // info
  __syncthreads();
  const bool yblock = (i0 >= d.np);
#pragma unroll 4
  for (int e = t; e < BLK * BLK; e += 256) {
    int r = e / BLK, cc = e % BLK;
    int gi = i0 + r, gj = j0 + cc;
    double v = 0.0;
    if (yblock) {
      if (r == 0 && gj < d.n) v = y[gj];
    } else if (gj <= gi) {
      if (gi < d.n) {
        v = (gi == gj) ? __dadd_rn(c, jit) : rbf_scaled(sx[0][r], sy[0][r], sx[1][cc], sy[1][cc], c);
      } else {
        v = (gi == gj) ? 1.0 : 0.0;
      }
    }
    K[(long)gi * d.ld + gj] = v;
  }
}

// Public dense Gram (gp.kernel(X) / gp.kernel(X, Y): rrt.py:447,452).  out is n x m, ld >= m.
__global__ void __launch_bounds__(256) gram_rbf_kernel(const double* __restrict__ X, int n, const double* __restrict__ Y,
                                                       int m, double c, double l, double jitter, int symmetric,
                                                       double* __restrict__ out, long ld, int kind) {
  __shared__ double sx[2][BLK], sy[2][BLK];
  const int i0 = blockIdx.y * BLK, j0 = blockIdx.x * BLK, t = threadIdx.x;
  if (t < 2 * BLK) {
    int which = t / BLK, e = t % BLK;
    const double* P = which ? Y : X;
    int lim = which ? m : n, idx = (which ? j0 : i0) + e;
    double vx = 0.0, vy = 0.0;
    if (idx < lim) {
      vx = __ddiv_rn(P[2 * idx], l);
      vy = __ddiv_rn(P[2 * idx + 1], l);
    }
    sx[which][e] = vx;
    sy[which][e] = vy;
  }
  __syncthreads();
  for (int e = t; e < BLK * BLK; e += 256) {
    int r = e / BLK, cc = e % BLK;
    int gi = i0 + r, gj = j0 + cc;
    if (gi < n && gj < m) {
      double v = (symmetric && gi == gj) ? __dadd_rn(c, jitter) : stationary_scaled(kind, sx[0][r], sy[0][r], sx[1][cc], sy[1][cc], c);
      out[(long)gi * ld + gj] = v;
    }
  }
}

// Symmetric case of the public Gram: one CTA per LOWER 64x64 tile computes it once (half the exp() calls, which
// bound this kernel together with the 8 n^2 bytes it must write) and stores both the tile and, through a
// padded shared-memory transpose, its mirror image -- every global store a full 512-byte row segment,
// streaming (evict-first) because the matrix is written once and is larger than L2 at n = 5000.
__global__ void __launch_bounds__(256) gram_sym_kernel(const double* __restrict__ X, int n, double c, double l,
                                                       double jitter, double* __restrict__ out, long ld, int kind) {
  __shared__ double sx[2][BLK], sy[2][BLK];
  __shared__ double tile[BLK][BLK + 1];
  const int b = blockIdx.x, t = threadIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((long)ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int i0 = ti * BLK, j0 = tj * BLK;
  if (t < 2 * BLK) {
    const int which = t / BLK, e = t % BLK, idx = (which ? j0 : i0) + e;
    double vx = 0.0, vy = 0.0;
    if (idx < n) {
      vx = __ddiv_rn(X[2 * idx], l);
      vy = __ddiv_rn(X[2 * idx + 1], l);
    }
    sx[which][e] = vx;
    sy[which][e] = vy;
  }
  __syncthreads();
  const int cc = t & 63, rq = t >> 6;
#pragma unroll 4
  for (int it = 0; it < BLK / 4; ++it) {
    const int r = rq + 4 * it, gi = i0 + r, gj = j0 + cc;
    const double v = (gi == gj) ? __dadd_rn(c, jitter) : stationary_scaled(kind, sx[0][r], sy[0][r], sx[1][cc], sy[1][cc], c);
    tile[r][cc] = v;
    if (gi < n && gj < n) __stcs(out + (long)gi * ld + gj, v);
  }
  if (ti == tj) return;
  __syncthreads();
#pragma unroll 4
  for (int it = 0; it < BLK / 4; ++it) {
    const int r = rq + 4 * it, gi = j0 + r, gj = i0 + cc;  // mirror tile: row = a column index of the computed tile
    if (gi < n && gj < n) __stcs(out + (long)gi * ld + gj, tile[cc][r]);
  }
}

// ---------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky as ONE persistent cooperative kernel (one CTA per SM, grid-wide
// barriers between dependent phases) instead of three launches per 64-column panel:
//   for panel k:   [grid barrier]  panel solve  A_ik <- A_ik L_kk^-T   (one CTA per 64-row block)
//                  [grid barrier]  trailing update on the DMMA pipe    (lower tiles over all CTAs);
//                                  the CTA that owns the tile holding block (k+1, k+1) updates it
//                                  first and factors it at once, so the next diagonal factor
//                                  overlaps the rest of the update instead of serialising the grid.
// At n = 5000 that is 158 grid barriers instead of 237 dependent launches.
// ---------------------------------------------------------------------------------------------
constexpr int PSTR = BLK + 1;
constexpr int QSTR = BLK + 4;  // pitch of a block read as DMMA fragments (conflict free for the m8n8k4 pattern)
constexpr size_t PANEL_SMEM = sizeof(double) * (BLK * PSTR + BLK * QSTR + BLK);  // two 64x64 blocks + 64 reciprocals

// In-place Cholesky of a 64x64 block held in shared memory (stride PSTR, lower part used), blocked by 8
// columns: warp 0 factors each 64 x 8 panel in registers (lane r owns rows jb + r and jb + r + 32; pivots and
// multipliers travel by shuffle, no barrier inside a panel), then the whole CTA applies the rank-8 update.
// 16 CTA barriers per block instead of 192; every entry sees the same operations in the same order as the
// textbook right-looking column sweep.
template <int NT, class Sync = CtaSync>
__device__ __forceinline__ void potf2_block(double* D, int t, int first_index, double* info, Sync sync = Sync()) {
  const int lane = t & 31, warp = t >> 5;
  for (int jb = 0; jb < BLK; jb += 8) {
    if (warp == 0) {
      const int r0 = jb + lane, r1 = jb + lane + 32;
      const bool has0 = r0 < BLK, has1 = r1 < BLK;
      double a0[8], a1[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        a0[c] = has0 ? D[r0 * PSTR + jb + c] : 1.0;
        a1[c] = has1 ? D[r1 * PSTR + jb + c] : 1.0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double piv = __shfl_sync(0xffffffffu, a0[c], c);  // row jb + c lives in lane c
        // one reciprocal square root per column (no divide on the chain); a non-positive pivot gives NaN
        const double rs = rsqrt(piv);
        if (lane == 0 && !(piv > 0.0) && info[0] == 0.0) info[0] = (double)(first_index + jb + c + 1);
        if (lane == c)
          a0[c] = piv * rs;  // sqrt(piv); NaN for piv <= 0 (rsqrt gives NaN or inf), as sqrt would
        else if (lane > c)
          a0[c] = a0[c] * rs;
        a1[c] = a1[c] * rs;
#pragma unroll
        for (int q = c + 1; q < 8; ++q) {
          const double lqc = __shfl_sync(0xffffffffu, a0[c], q);  // L[jb + q][jb + c]
          if (lane > c) a0[q] -= a0[c] * lqc;
          a1[q] -= a1[c] * lqc;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (has0) D[r0 * PSTR + jb + c] = a0[c];
        if (has1) D[r1 * PSTR + jb + c] = a1[c];
      }
    }
    sync();
    // rank-8 update of the trailing lower triangle in 4x4 register tiles (16 independent FMA chains per thread,
    // 5 shared loads per 8 FMAs instead of 17); every entry still sees its 8 subtractions in column order
    for (int blk = t; blk < (BLK / 4) * (BLK / 4); blk += NT) {
      const int br = blk >> 4, bc = blk & 15, r0 = 4 * br, c0 = 4 * bc;
      if (bc <= br && c0 >= jb + 8) {
        double d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) d[i][j] = D[(r0 + i) * PSTR + c0 + j];
#pragma unroll
        for (int ph = 0; ph < 8; ph += 4) {
          double pr[4][4], pc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              pr[i][pp] = D[(r0 + i) * PSTR + jb + ph + pp];
              pc[i][pp] = D[(c0 + i) * PSTR + jb + ph + pp];
            }
#pragma unroll
          for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) d[i][j] -= pr[i][pp] * pc[j][pp];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) D[(r0 + i) * PSTR + c0 + j] = d[i][j];
      }
    }
    sync();
  }
}

template <int NT>
__device__ __forceinline__ void chol_diag_step(double* __restrict__ A, long ld, int k, double* psm, double* info) {
  double* D = psm;
  const int t = threadIdx.x, j0 = BLK * k;
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    D[r * PSTR + cc] = A[(long)(j0 + r) * ld + j0 + cc];
  }
  __syncthreads();
  potf2_block<NT>(D, t, j0, info);
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    A[(long)(j0 + r) * ld + j0 + cc] = (cc <= r) ? D[r * PSTR + cc] : 0.0;
  }
  __syncthreads();
}

// Diagonal block k after panel kp = k - 1: apply that panel's update to the 64x64 block with plain FMAs
// (it is 1/4 of a 128x128 tensor tile and sits on the critical path), then factor it.
template <int NT>
__device__ __forceinline__ void chol_diag_update_step(double* __restrict__ A, long ld, int k, int kp, double* psm,
                                                      double* info, long long* clk = nullptr) {
  const long long q0 = clk ? clock64() : 0;
  double* D = psm;
  double* P = psm + BLK * PSTR;
  const int t = threadIdx.x, j0 = BLK * k, p0 = BLK * kp;
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    D[r * PSTR + cc] = A[(long)(j0 + r) * ld + j0 + cc];
    P[r * QSTR + cc] = A[(long)(j0 + r) * ld + p0 + cc];
  }
  __syncthreads();
  const long long q1 = clk ? clock64() : 0;
  {  // D -= P P^T on the DMMA pipe: each warp owns 64 / (NT / 32) rows, all 64 columns
    constexpr int RPW = BLK / (NT / 32), MT = RPW / 8;
    const int warp = t >> 5, lane = t & 31, g = lane >> 2, tg = lane & 3;
    double acc[MT][8][2];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < BLK; kk += 4) {
      double af[MT], bf[8];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) af[mi] = P[(warp * RPW + mi * 8 + g) * QSTR + kk + tg];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) bf[ni] = P[(ni * 8 + g) * QSTR + kk + tg];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = warp * RPW + mi * 8 + g, cc = ni * 8 + tg * 2 + e;
          if (cc <= r) D[r * PSTR + cc] -= acc[mi][ni][e];
        }
  }
  __syncthreads();
  const long long q2 = clk ? clock64() : 0;
  potf2_block<NT>(D, t, j0, info);
  const long long q3 = clk ? clock64() : 0;
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    A[(long)(j0 + r) * ld + j0 + cc] = (cc <= r) ? D[r * PSTR + cc] : 0.0;
  }
  __syncthreads();
  if (clk) {
    const long long q4 = clock64();
    clk[0] += q1 - q0;  // block loads
    clk[1] += q2 - q1;  // DMMA update
    clk[2] += q3 - q2;  // potf2
    clk[3] += q4 - q3;  // store
  }
}

// Panel solve of one 64-row block: A_ik <- A_ik * L_kk^-T by forward substitution (backward stable,
// unlike multiplying by the explicit inverse: at cond(K) ~ 1e11 the diagonal blocks reach cond ~ 1e5 and
// the inverse route cost 5 digits of L).  NT / 64 adjacent lanes share a row (columns interleaved), the
// row lives in their registers, the solved entry is broadcast by shuffle, L_kk comes from shared memory.
// ROWS (64, 32 or 16) rows of block i per call, part `h` of the block: when a panel has fewer row blocks than half
// (a quarter of) the CTAs, the blocks are split so that every SM takes part -- the 64 substitution steps stay, but
// each step moves 1/2 (1/4) of the shared-memory traffic that bounds it.  Every row sees the same operations in the
// same order whatever the split, so the factor does not depend on it.
template <int NT, int ROWS>
__device__ __forceinline__ void chol_trsm_block(double* __restrict__ A, long ld, int k, int i, int h, double* psm) {
  constexpr int LPR = NT / ROWS;  // lanes per row
  constexpr int CPL = BLK / LPR;  // columns per lane
  static_assert(LPR >= 1 && LPR <= 32 && CPL * LPR == BLK, "row split must keep a row inside one warp");
  double* Ls = psm;
  double* Xs = psm + BLK * PSTR;
  double* rinv = psm + BLK * PSTR + BLK * QSTR;
  const int t = threadIdx.x, j0 = BLK * k, i0 = BLK * i + ROWS * h;
  // (L2 loads: in the dataflow kernel these blocks were written by other CTAs of the same launch)
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    Ls[r * PSTR + cc] = __ldcg(A + (long)(j0 + r) * ld + j0 + cc);
  }
  for (int e = t; e < ROWS * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    Xs[r * PSTR + cc] = __ldcg(A + (long)(i0 + r) * ld + j0 + cc);
  }
  // One reciprocal per column, then multiplies: 64 dependent fp64 divides per row were most of this step,
  // and rows of exact zeros (identity padding, the y block) sent every one of them down the divide's slow path.
  if (t < BLK) rinv[t] = 1.0 / __ldcg(A + (long)(j0 + t) * ld + j0 + t);
  __syncthreads();
  {
    const int r = t / LPR, q0 = t % LPR, lane = t & 31;
    double x[CPL];
#pragma unroll
    for (int m = 0; m < CPL; ++m) x[m] = Xs[r * PSTR + q0 + m * LPR];
#pragma unroll
    for (int c = 0; c < BLK; ++c) {
      const int owner = c % LPR, mc = c / LPR;
      double xc = x[mc] * rinv[c];  // meaningful in the owner lane only
      xc = __shfl_sync(0xffffffffu, xc, (lane & ~(LPR - 1)) + owner);
      if (q0 == owner) x[mc] = xc;
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        const int q = q0 + m * LPR;
        if (q > c) x[m] -= xc * Ls[q * PSTR + c];
      }
    }
#pragma unroll
    for (int m = 0; m < CPL; ++m) Xs[r * PSTR + q0 + m * LPR] = x[m];
  }
  __syncthreads();
  for (int e = t; e < ROWS * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    A[(long)(i0 + r) * ld + j0 + cc] = Xs[r * PSTR + cc];
  }
  __syncthreads();
}

// Trailing update of one lower tile after panel k: C[i][j] -= sum_p P[i][p] P[j][p], i >= j >= 64(k+1).
template <class C>
__device__ __forceinline__ void chol_syrk_tile(double* __restrict__ A, long ld, int rows_total, int np, int k, int b,
                                               double* smem, long long* tile_clk = nullptr) {
  const long long c0 = tile_clk ? clock64() : 0;
  const int skip = k + 1;  // diagonal block (k+1, k+1) is updated and factored by chol_diag_update_step
  const int j0 = BLK * k, r0 = j0 + BLK;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((long)ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int row0 = r0 + ti * C::BM, col0 = r0 + tj * C::BN;
  if (row0 >= rows_total || col0 >= np) return;
  // The C tile is read-modify-written once per panel and, for n >~ 4000, the trailing matrix no longer fits in
  // L2: start pulling the tile's 128-byte lines into L2 now, so that the epilogue's loads (measured 10 us per tile
  // from HBM against 1.8 us from L2) overlap the DMMA main loop instead of following it.
  {
    constexpr int LINES_PER_ROW = C::BN * 8 / 128;
    for (int e = threadIdx.x; e < C::BM * LINES_PER_ROW; e += C::THREADS) {
      const int gr = row0 + e / LINES_PER_ROW, gc = col0 + (e % LINES_PER_ROW) * 16;
      if (gr < rows_total && gc < np && (gc >> 6) <= (gr >> 6))
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A + (long)gr * ld + gc));
    }
  }
  double acc[C::MI][C::NI][2];
  zero_acc<C>(acc);
  auto fillA = [&](double* s, int k0) {
    load_tile_async<C::BM, Lay::KCONT, C::THREADS>(s, A, ld, row0, k0, rows_total, j0 + BLK);
  };
  auto fillB = [&](double* s, int k0) {
    load_tile_async<C::BN, Lay::KCONT, C::THREADS>(s, A, ld, col0, k0, np, j0 + BLK);
  };
  gemm_mainloop<C, Lay::KCONT, Lay::KCONT>(acc, smem, j0, j0 + BLK, fillA, fillB);
  const long long c1 = tile_clk ? clock64() : 0;
  // C -= acc as 16-byte read-modify-writes, all loads of half a warp tile in flight before the first store
  // (the compiler's own schedule kept ~8 scalar loads in flight: 8 dependent trips to HBM per tile).
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N, g = lane >> 2, tg = lane & 3;
    constexpr int HALF = C::MI >= 8 ? C::MI / 4 : (C::MI / 2 > 0 ? C::MI / 2 : 1);  // rows of 8x8 blocks per batch
#pragma unroll
    for (int h0 = 0; h0 < C::MI; h0 += HALF) {
      double2 cv[HALF][C::NI];
      bool ok[HALF][C::NI];
#pragma unroll
      for (int mi = 0; mi < HALF; ++mi)
#pragma unroll
        for (int ni = 0; ni < C::NI; ++ni) {
          const int gr = row0 + wm * C::WTM + (h0 + mi) * 8 + g, gc = col0 + wn * C::WTN + ni * 8 + tg * 2;
          const int br = gr >> 6, bc = gc >> 6;
          ok[mi][ni] = gr < rows_total && gc < np && bc <= br && !(br == skip && bc == skip);
          if (ok[mi][ni]) cv[mi][ni] = __ldcg(reinterpret_cast<const double2*>(A + (long)gr * ld + gc));
        }
#pragma unroll
      for (int mi = 0; mi < HALF; ++mi)
#pragma unroll
        for (int ni = 0; ni < C::NI; ++ni) {
          const int gr = row0 + wm * C::WTM + (h0 + mi) * 8 + g, gc = col0 + wn * C::WTN + ni * 8 + tg * 2;
          if (ok[mi][ni])
            *reinterpret_cast<double2*>(A + (long)gr * ld + gc) =
                make_double2(cv[mi][ni].x - acc[h0 + mi][ni][0], cv[mi][ni].y - acc[h0 + mi][ni][1]);
        }
    }
  }
  if (tile_clk) {
    __syncthreads();
    const long long c2 = clock64();
    tile_clk[0] += c1 - c0;  // operand pipeline + DMMA
    tile_clk[1] += c2 - c1;  // read-modify-write of the C tile
    tile_clk[2] += 1;
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) chol_coop_kernel(double* __restrict__ A, long ld, int rows_total, int np,
                                                                  double* __restrict__ info, double* __restrict__ phase_cycles) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double smem[];
  const int nb = np / BLK, nrb = rows_total / BLK;
  // phase_cycles (nullable): CTA 0's clock64() spent in {barrier 1, panel solve, barrier 2, diagonal block,
  // trailing tiles} summed over the panels -- the evidence for where a latency-bound factorisation goes
  const bool timing = phase_cycles != nullptr && blockIdx.x < 2 && threadIdx.x == 0;  // CTA 0 (panel) and CTA 1 (tiles)
  long long tc[5] = {0, 0, 0, 0, 0}, t0 = 0, tile_clk[3] = {0, 0, 0};
  long long* tclk = (phase_cycles != nullptr && blockIdx.x == 1) ? tile_clk : nullptr;  // uniform over CTA 1
  long long diag_clk[4] = {0, 0, 0, 0};
  long long* dclk = (phase_cycles != nullptr && blockIdx.x == 0) ? diag_clk : nullptr;
  auto tick = [&](int slot) {
    if (timing) {
      const long long now = clock64();
      tc[slot] += now - t0;
      t0 = now;
    }
  };
  if (blockIdx.x == 0) chol_diag_step<C::THREADS>(A, ld, 0, smem, info);
  if (timing) t0 = clock64();
  for (int k = 0; k < nb; ++k) {
    grid.sync();  // L_kk is final
    tick(0);
    {
      const int rem = nrb - (k + 1), G = (int)gridDim.x;
      if (4 * rem <= G && C::THREADS >= 64) {
        for (int w = blockIdx.x; w < 4 * rem; w += G) chol_trsm_block<C::THREADS, 16>(A, ld, k, k + 1 + w / 4, w % 4, smem);
      } else if (2 * rem <= G) {
        for (int w = blockIdx.x; w < 2 * rem; w += G) chol_trsm_block<C::THREADS, 32>(A, ld, k, k + 1 + w / 2, w % 2, smem);
      } else {
        for (int i = k + 1 + blockIdx.x; i < nrb; i += G) chol_trsm_block<C::THREADS, 64>(A, ld, k, i, 0, smem);
      }
    }
    tick(1);
    grid.sync();  // panel k is final
    tick(2);
    if (k + 1 < nb) {
      const int rem = rows_total - BLK * (k + 1);
      const int T = (rem + C::BM - 1) / C::BM;
      const int ntile = T * (T + 1) / 2;
      // the next diagonal block first (CTA 0), behind the other CTAs' tiles
      // CTA 0 owns the critical path (next diagonal block); the other CTAs share the tiles
      if (blockIdx.x == 0) chol_diag_update_step<C::THREADS>(A, ld, k + 1, k, smem, info, dclk);
      tick(3);
      if (gridDim.x == 1) {
        for (int b = 0; b < ntile; ++b) chol_syrk_tile<C>(A, ld, rows_total, np, k, b, smem);
      } else if (blockIdx.x > 0) {
        for (int b = blockIdx.x - 1; b < ntile; b += gridDim.x - 1) chol_syrk_tile<C>(A, ld, rows_total, np, k, b, smem, tclk);
      }
      tick(4);
    }
  }
  if (timing) {
    for (int q = 0; q < 5; ++q) phase_cycles[5 * blockIdx.x + q] = (double)tc[q];
    if (blockIdx.x == 1)  // (timing is true for thread 0 only)
      for (int q = 0; q < 3; ++q) phase_cycles[10 + q] = (double)tile_clk[q];
    if (blockIdx.x == 0)
      for (int q = 0; q < 4; ++q) phase_cycles[13 + q] = (double)diag_clk[q];
  }
}

// W_kk = L_kk^-1 for every 64x64 diagonal block at once (level 0 of the triangular inverse): the panel solve
// applied to the identity, X = I L_kk^-T = W_kk^T -- 64 substitution steps with the rows in registers (4 lanes per
// row, multipliers by shuffle), the same routine as chol_trsm_block; stored transposed.
__global__ void __launch_bounds__(256) diag_inverse_kernel(const double* __restrict__ L, long ld, double* __restrict__ Winv,
                                                           long ldw, Bat bat) {
  constexpr int NT = 256, LPR = NT / BLK, CPL = BLK / LPR;
  extern __shared__ double psm[];
  L += bat.slot(blockIdx.y) * bat.sK;
  Winv += bat.slot(blockIdx.y) * bat.sW;
  double* Ls = psm;
  double* Xs = psm + BLK * PSTR;
  double* rinv = psm + BLK * PSTR + BLK * QSTR;
  const int t = threadIdx.x, j0 = BLK * blockIdx.x;
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    Ls[r * PSTR + cc] = L[(long)(j0 + r) * ld + j0 + cc];
  }
  if (t < BLK) rinv[t] = 1.0 / L[(long)(j0 + t) * ld + j0 + t];
  __syncthreads();
  {
    const int r = t / LPR, q0 = t % LPR, lane = t & 31;
    double x[CPL];
#pragma unroll
    for (int m = 0; m < CPL; ++m) x[m] = (q0 + m * LPR == r) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < BLK; ++c) {
      const int owner = c % LPR, mc = c / LPR;
      double xc = x[mc] * rinv[c];  // meaningful in the owner lane only
      xc = __shfl_sync(0xffffffffu, xc, (lane & ~(LPR - 1)) + owner);
      if (q0 == owner) x[mc] = xc;
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        const int q = q0 + m * LPR;
        if (q > c) x[m] -= xc * Ls[q * PSTR + c];
      }
    }
#pragma unroll
    for (int m = 0; m < CPL; ++m) Xs[r * PSTR + q0 + m * LPR] = x[m];  // X[r][c] = W[c][r]
  }
  __syncthreads();
  for (int e = t; e < BLK * BLK; e += NT) {
    int r = e / BLK, cc = e % BLK;
    Winv[(long)(j0 + r) * ldw + j0 + cc] = (cc <= r) ? Xs[cc * PSTR + r] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// W = L^-1 by recursive doubling over a BALANCED binary partition of the nb = NP/64 block rows: with `parts`
// (a power of two) parts at the finer level, part i is blocks [floor(i nb / parts), floor((i+1) nb / parts)), and
// the level merges parts 2p and 2p+1 -- rows [r0, r1) and [r1, r2):  W21 = -W22 * (L21 * W11).  The partitions of
// successive levels refine each other (floor(2i nb / 2P) = floor(i nb / P)), the finest has at most one block per
// part (diag_inverse_kernel), empty halves are no-ops.  A grid aligned to powers of two instead makes the top level
// lopsided whenever NP is not a power of two (n = 5000: 4096 + 960 rows, its longest tile 128 x 128 x 4096 = 0.53 ms
// of one SM's peak); balanced halves (2496 + 2560) cut the longest k range to 2496 and even out the tile counts.
// Two batched GEMM launches per level:
//   STEP 1: T[r1+i][r0+j]  =  sum_{k >= j} L[r1+i][r0+k] * W[r0+k][r0+j]
//   STEP 2: W[r1+i][r0+j]  = -sum_{k <= i} W[r1+i][r1+k] * T[r1+k][r0+j]
// ---------------------------------------------------------------------------------------------
// MODE: 0 = cp.async ring with a CTA barrier per slab, 1 = mbarrier pipeline (every warp copies and multiplies),
// 2 = warp-specialised (C::THREADS + 128 threads: eight DMMA warps + a copy warpgroup, gemm_ws_*).
// ROLE (MODE 2 only): 1 = copy warpgroup, 2 = DMMA warps; 0 = every thread does everything.
template <class C, int STEP, int MODE, int ROLE>
__device__ __forceinline__ void trtri_level_body(const double* __restrict__ L, long ld, double* __restrict__ W, long ldw,
                                                 double* __restrict__ T, long ldt, int nb, int parts, int s, const Bat& bat,
                                                 double* smem, uint64_t* bars) {
  L += bat.slot(blockIdx.y) * bat.sK;
  W += bat.slot(blockIdx.y) * bat.sW;
  T += bat.slot(blockIdx.y) * bat.sT;
  // 1-D grid, heaviest tiles first (the hardware hands out CTAs in index order): the k range of a tile is
  // triangular -- STEP 1 sums k in [col0, r1), STEP 2 sums k in [r1, row0 + BM) -- so a plain (x, y) order
  // leaves the longest tiles for the last wave.
  // s = 64 * ceil(nb / parts): the largest half at this level
  const int nx = (s + C::BN - 1) / C::BN, ny = (s + C::BM - 1) / C::BM, pairs = parts / 2;
  const int p = blockIdx.x % pairs, rest = blockIdx.x / pairs;
  const int bx = (STEP == 1) ? rest / ny : rest % nx;
  const int by = (STEP == 1) ? rest % ny : ny - 1 - rest / nx;
  const int r0 = BLK * (int)((long)(2 * p) * nb / parts), r1 = BLK * (int)((long)(2 * p + 1) * nb / parts),
            r2 = BLK * (int)((long)(2 * p + 2) * nb / parts);
  if (r1 == r0 || r2 == r1) return;  // an empty half: the other one is already inverted
  const int row0 = r1 + by * C::BM, col0 = r0 + bx * C::BN;
  if (row0 >= r2 || col0 >= r1) return;
  // STEP 1: T = L21 W11 (k in [col0, r1));  STEP 2: W21 = -W22 T (k in [r1, row0 + BM))
  const double* Am = (STEP == 1) ? L : W;
  const long lda = (STEP == 1) ? ld : ldw;
  const double* Bm = (STEP == 1) ? W : T;
  const long ldb = (STEP == 1) ? ldw : ldt;
  const int klimA = (STEP == 1) ? r1 : r2, klimB = (STEP == 1) ? r1 : r2;
  const int kbeg = (STEP == 1) ? col0 : r1, kend = (STEP == 1) ? r1 : min(row0 + C::BM, r2);
  if constexpr (ROLE == 1) {
    auto wsA = [&](double* sm, int k0, int tid) { load_tile_async<C::BM, Lay::KCONT, 128>(sm, Am, lda, row0, k0, r2, klimA, tid); };
    auto wsB = [&](double* sm, int k0, int tid) { load_tile_async<C::BN, Lay::MCONT, 128>(sm, Bm, ldb, col0, k0, r1, klimB, tid); };
    gemm_ws_produce<C, Lay::KCONT, Lay::MCONT>(smem, bars, kbeg, kend, wsA, wsB);
    return;
  } else {
    double acc[C::MI][C::NI][2];
    zero_acc<C>(acc);
    if constexpr (ROLE == 2) {
      gemm_ws_consume<C, Lay::KCONT, Lay::MCONT>(acc, smem, bars, kbeg, kend);
    } else {
      auto fillA = [&](double* sm, int k0) { load_tile_async<C::BM, Lay::KCONT, C::THREADS>(sm, Am, lda, row0, k0, r2, klimA); };
      auto fillB = [&](double* sm, int k0) { load_tile_async<C::BN, Lay::MCONT, C::THREADS>(sm, Bm, ldb, col0, k0, r1, klimB); };
      if (MODE == 1)
        gemm_mainloop_mb<C, Lay::KCONT, Lay::MCONT>(acc, smem, bars, kbeg, kend, fillA, fillB);
      else
        gemm_mainloop<C, Lay::KCONT, Lay::MCONT>(acc, smem, kbeg, kend, fillA, fillB);
    }
    if (STEP == 1) {
      for_each_acc<C>(acc, [&](int r, int c, double v) {
        int gr = row0 + r, gc = col0 + c;
        if (gr < r2 && gc < r1) T[(long)gr * ldt + gc] = v;
      });
    } else {
      for_each_acc<C>(acc, [&](int r, int c, double v) {
        int gr = row0 + r, gc = col0 + c;
        if (gr < r2 && gc < r1) W[(long)gr * ldw + gc] = -v;
      });
    }
  }
}
template <class C, int STEP, int MODE>
__global__ void __launch_bounds__(C::THREADS + (MODE == 2 ? 128 : 0), 1) trtri_level_kernel(const double* __restrict__ L, long ld,
                                                                 double* __restrict__ W, long ldw, double* __restrict__ T,
                                                                 long ldt, int nb, int parts, int s, Bat bat) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bars[2 * MB_STAGES];
  if constexpr (MODE == 2) {
    if (gemm_ws_split<C>(bars)) {
      trtri_level_body<C, STEP, MODE, 1>(L, ld, W, ldw, T, ldt, nb, parts, s, bat, smem, bars);
      return;
    }
    trtri_level_body<C, STEP, MODE, 2>(L, ld, W, ldw, T, ldt, nb, parts, s, bat, smem, bars);
  } else {
    trtri_level_body<C, STEP, MODE, 0>(L, ld, W, ldw, T, ldt, nb, parts, s, bat, smem, bars);
  }
}

// Zero the strictly-upper blocks of W before the doubling sweep (diagonal 64-blocks come from the
// diagonal kernel, strictly-lower ones from the sweep); the posterior and trace kernels read them.
__global__ void __launch_bounds__(256) w_clear_kernel(double* __restrict__ W, long ldw, int np, Bat bat) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi >= bj) return;  // strictly-lower blocks are fully rewritten by the sweep
  W += bat.slot(blockIdx.z) * bat.sW;
  for (int e = threadIdx.x; e < BLK * BLK; e += 256) {
    int r = e / BLK, c = e % BLK;
    W[(long)(bi * BLK + r) * ldw + bj * BLK + c] = 0.0;
  }
}

// alpha = W^T z  (z = row NP of the factor buffer).  W is 8 NP^2 bytes and read once: HBM bound, so the
// k range is split GEMV_SPLIT ways across CTAs (one 64-column strip alone gives NP/64 CTAs, too few to
// pull HBM bandwidth); partial sums land in `part` and are combined in a fixed order (deterministic).
constexpr int GEMV_SPLIT = 16;
__global__ void __launch_bounds__(256) gemv_wt_kernel(const double* __restrict__ W, long ldw, const double* __restrict__ z,
                                                      double* __restrict__ part, int np, Bat bat) {
  __shared__ double red[4][BLK];
  W += bat.slot(blockIdx.z) * bat.sW;
  z += bat.slot(blockIdx.z) * bat.sK;
  part += bat.slot(blockIdx.z) * bat.sT;
  const int ci = threadIdx.x & 63, kq = threadIdx.x >> 6;
  const int i0 = blockIdx.x * BLK, i = i0 + ci;
  // rows [i0, np) of this strip are non-zero; chunk them evenly over blockIdx.y
  const int rows = np - i0, chunk = (rows + GEMV_SPLIT - 1) / GEMV_SPLIT;
  const int kb = i0 + blockIdx.y * chunk, ke = min(np, kb + chunk);
  double s = 0.0;
  for (int k = kb + kq; k < ke; k += 4) s += W[(long)k * ldw + i] * z[k];
  red[kq][ci] = s;
  __syncthreads();
  if (kq == 0) part[(long)blockIdx.y * np + i] = (red[0][ci] + red[1][ci]) + (red[2][ci] + red[3][ci]);
}
__global__ void __launch_bounds__(256) gemv_reduce_kernel(const double* __restrict__ part, double* __restrict__ alpha, int np,
                                                          Bat bat) {
  part += bat.slot(blockIdx.y) * bat.sT;
  alpha += bat.slot(blockIdx.y) * bat.sV;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= np) return;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < GEMV_SPLIT; ++q) s += part[(long)q * np + i];
  alpha[i] = s;
}

// ---------------------------------------------------------------------------------------------
// Fused K^-1 trace contraction for the LML gradient (_gpr.py:629-651) without ever storing K^-1
// or dK:  g_k = sum_ij (alpha_i alpha_j - (W^T W)_ij) * dK_ij,k  with dK_0 = K0 (no jitter) and
// dK_1 = K0 o D2 (D2 on l-scaled coordinates, kernels.py:1577).  Lower-triangular tiles only,
// off-diagonal tiles weighted 2.  Per-tile partials, reduced in fixed order by lml_finalize.
// ---------------------------------------------------------------------------------------------
template <class C, int MODE, int ROLE>
__device__ __forceinline__ void lml_grad_body(const double* __restrict__ W, long ldw, const double* __restrict__ vbuf,
                                              const double* __restrict__ hyp, double* __restrict__ partials, const GpDims& d,
                                              const Bat& bat, const double* __restrict__ tile_bbox, double* smem,
                                              uint64_t* bars, double (*red)[C::THREADS / 32]) {
  {
    const long slot = bat.slot(blockIdx.y);
    W += slot * bat.sW;
    vbuf += slot * bat.sV;
    partials += slot * bat.sV;
    hyp += 4 * slot;
  }
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((long)ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int row0 = ti * C::BM, col0 = tj * C::BN;
  // Block sparsity of the contraction.  Every term carries dK_ij = K_ij (or K_ij d2_ij), and with the observations in
  // a spatial order a tile pair whose bounding boxes are farther apart than 13.6 l has K_ij d2_ij < 2e-38 c for all
  // of its 128 x 128 entries: against |alpha_i alpha_j - K^-1_ij| <= ~1e14 that is below 1e-23 of anything the sum
  // keeps, so the tile contributes exactly nothing representable and its W^T W product is not formed at all.
  if (tile_bbox != nullptr && ti != tj) {
    const double4 bi = *reinterpret_cast<const double4*>(tile_bbox + 4 * ti), bj = *reinterpret_cast<const double4*>(tile_bbox + 4 * tj);
    const double gx = fmax(0.0, fmax(bi.x - bj.y, bj.x - bi.y)), gy = fmax(0.0, fmax(bi.z - bj.w, bj.z - bi.w));
    const double cut = 13.6 * hyp[1];
    if (gx * gx + gy * gy > cut * cut) {
      if (threadIdx.x == 0 && ROLE != 1) partials[2 * b] = partials[2 * b + 1] = 0.0;
      return;
    }
  }
  if constexpr (ROLE == 1) {
    auto wsA = [&](double* sm, int k0, int tid) { load_tile_async<C::BM, Lay::MCONT, 128>(sm, W, ldw, row0, k0, d.np, d.np, tid); };
    auto wsB = [&](double* sm, int k0, int tid) { load_tile_async<C::BN, Lay::MCONT, 128>(sm, W, ldw, col0, k0, d.np, d.np, tid); };
    gemm_ws_produce<C, Lay::MCONT, Lay::MCONT>(smem, bars, row0, d.np, wsA, wsB);
    return;
  }
  double acc[C::MI][C::NI][2];
  zero_acc<C>(acc);
  if constexpr (ROLE == 2) {
    gemm_ws_consume<C, Lay::MCONT, Lay::MCONT>(acc, smem, bars, row0, d.np);
  } else {
    auto fillA = [&](double* sm, int k0) { load_tile_async<C::BM, Lay::MCONT, C::THREADS>(sm, W, ldw, row0, k0, d.np, d.np); };
    auto fillB = [&](double* sm, int k0) { load_tile_async<C::BN, Lay::MCONT, C::THREADS>(sm, W, ldw, col0, k0, d.np, d.np); };
    if (MODE == 1)
      gemm_mainloop_mb<C, Lay::MCONT, Lay::MCONT>(acc, smem, bars, row0, d.np, fillA, fillB);
    else
      gemm_mainloop<C, Lay::MCONT, Lay::MCONT>(acc, smem, row0, d.np, fillA, fillB);
  }
  const double c = hyp[0];
  const double* xs = vbuf + V_XS * d.npv;
  const double* ys = vbuf + V_YS * d.npv;
  const double* al = vbuf + V_ALPHA * d.npv;
  const double wgt = (ti == tj) ? 1.0 : 2.0;
  double g0 = 0.0, g1 = 0.0;
  for_each_acc<C>(acc, [&](int r, int cc, double v) {
    int gi = row0 + r, gj = col0 + cc;
    if (gi < d.n && gj < d.n) {
      double d2 = sqdist_nofma(xs[gi], ys[gi], xs[gj], ys[gj]);
      double kij = (gi == gj) ? c : __dmul_rn(exp(__dmul_rn(-0.5, d2)), c);
      double tt = al[gi] * al[gj] - v;
      g0 += tt * kij;
      g1 += tt * (kij * d2);
    }
  });
  g0 = warp_sum(g0);
  g1 = warp_sum(g1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = g0;
    red[1][warp] = g1;
  }
  if (ROLE == 2) Group256Sync()();  // the DMMA warps only: the copy warps have returned
  else __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < C::THREADS / 32; ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
    }
    partials[2 * b] = wgt * s0;
    partials[2 * b + 1] = wgt * s1;
  }
}
template <class C, int MODE>
__global__ void __launch_bounds__(C::THREADS + (MODE == 2 ? 128 : 0), 1) lml_grad_kernel(const double* __restrict__ W, long ldw,
                                                              const double* __restrict__ vbuf,
                                                              const double* __restrict__ hyp, double* __restrict__ partials,
                                                              GpDims d, Bat bat, const double* __restrict__ tile_bbox) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[2][C::THREADS / 32];
  __shared__ __align__(8) uint64_t bars[2 * MB_STAGES];
  if constexpr (MODE == 2) {
    if (gemm_ws_split<C>(bars)) {
      lml_grad_body<C, MODE, 1>(W, ldw, vbuf, hyp, partials, d, bat, tile_bbox, smem, bars, red);
      return;
    }
    lml_grad_body<C, MODE, 2>(W, ldw, vbuf, hyp, partials, d, bat, tile_bbox, smem, bars, red);
  } else {
    lml_grad_body<C, MODE, 0>(W, ldw, vbuf, hyp, partials, d, bat, tile_bbox, smem, bars, red);
  }
}

// result = {lml, dlml/dlog c, dlml/dlog l, info, z.z, sum log diag L}
__global__ void __launch_bounds__(256) lml_finalize_kernel(const double* __restrict__ K, double* __restrict__ vbuf,
                                                           int ntiles, int with_grad, GpDims d, Bat bat) {
  __shared__ double red[4][256];
  const int t = threadIdx.x;
  K += bat.slot(blockIdx.x) * bat.sK;
  vbuf += bat.slot(blockIdx.x) * bat.sV;
  const double* z = K + (long)d.np * d.ld;
  double zz = 0.0, sl = 0.0, g0 = 0.0, g1 = 0.0;
  for (int k = t; k < d.n; k += 256) {
    zz += z[k] * z[k];
    sl += log(K[(long)k * d.ld + k]);
  }
  if (with_grad) {
    const double* part = vbuf + v_partials_off(d);
    for (int k = t; k < ntiles; k += 256) {
      g0 += part[2 * k];
      g1 += part[2 * k + 1];
    }
  }
  red[0][t] = zz;
  red[1][t] = sl;
  red[2][t] = g0;
  red[3][t] = g1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o)
      for (int q = 0; q < 4; ++q) red[q][t] += red[q][t + o];
    __syncthreads();
  }
  if (t == 0) {
    double* out = vbuf + v_result_off(d);
    out[0] = -0.5 * red[0][0] - red[1][0] - 0.5 * (double)d.n * 1.8378770664093453;  // log(2 pi)
    out[1] = 0.5 * red[2][0];
    out[2] = 0.5 * red[3][0];
    out[4] = red[0][0];
    out[5] = red[1][0];
  }
}

// Blocked backward substitution alpha = L^-T z for the final fit (_gpr.py:363-367), one launch per
// 64-block from the bottom up.  Every CTA solves the 64x64 diagonal system by substitution (redundantly,
// it is tiny); CTA c < b then updates w_c -= L[b][c]^T alpha_b and CTA b stores alpha_b.
__global__ void __launch_bounds__(BLK) backsub_step_kernel(const double* __restrict__ L, long ld, double* __restrict__ w,
                                                           double* __restrict__ alpha, int b) {
  __shared__ double Ls[BLK * PSTR];
  __shared__ double ab[BLK];
  const int t = threadIdx.x, b0 = b * BLK, c0 = blockIdx.x * BLK;
  for (int e = t; e < BLK * BLK; e += BLK) {
    int r = e / BLK, cc = e % BLK;
    Ls[r * PSTR + cc] = L[(long)(b0 + r) * ld + b0 + cc];
  }
  double wt = w[b0 + t];
  __syncthreads();
  for (int k = BLK - 1; k >= 0; --k) {
    if (t == k) ab[k] = wt / Ls[k * PSTR + k];
    __syncthreads();
    if (t < k) wt -= Ls[k * PSTR + t] * ab[k];
  }
  if ((int)blockIdx.x == b) {
    alpha[b0 + t] = ab[t];
  } else {
    double u = 0.0;
#pragma unroll 8
    for (int k = 0; k < BLK; ++k) u += L[(long)(b0 + k) * ld + c0 + t] * ab[k];
    w[c0 + t] -= u;
  }
}

// The same backward substitution as ONE launch (nb <= BACKSUB_MAX_BLOCKS CTAs, all resident): CTA c owns block c of w.
// It subtracts L[b][c]^T alpha_b for b = nb-1 .. c+1 as soon as alpha_b is published (a flag per block: release store by
// the producer, acquire load by the consumers; the next L block is already in registers while it waits), then solves
// its own 64 x 64 diagonal system in one warp (reciprocal pivots, multipliers by shuffle) and publishes alpha_c.
// 79 dependent launches of 25 us each (1.9 ms at n = 5000) become one chain of ~4 us steps.
constexpr int BACKSUB_MAX_BLOCKS = 256;
__global__ void __launch_bounds__(256) backsub_chain_kernel(const double* __restrict__ L, long ld, const double* __restrict__ z,
                                                            double* __restrict__ alpha, int* __restrict__ flags, int nb) {
  __shared__ double Ls[BLK * PSTR];
  __shared__ double red[4][BLK];
  __shared__ double ab[BLK], wsh[BLK], rinv[BLK];
  const int t = threadIdx.x, c = nb - 1 - (int)blockIdx.x;  // the hardware hands out CTAs in index order: last block first
  const int c0 = c * BLK, ci = t & 63, kq = t >> 6;
  for (int e = t; e < BLK * BLK; e += 256) {
    const int r = e / BLK, cc = e % BLK;
    Ls[r * PSTR + cc] = __ldcg(L + (long)(c0 + r) * ld + c0 + cc);
  }
  if (t < BLK) {
    wsh[t] = __ldcg(z + c0 + t);
    rinv[t] = 1.0 / __ldcg(L + (long)(c0 + t) * ld + c0 + t);
  }
  double w_acc = 0.0;  // thread (kq, ci): partial sum over k = kq (mod 4) of everything subtracted from w[c0 + ci]
  double lreg[16];
  auto load_block = [&](int b) {
#pragma unroll
    for (int j = 0; j < 16; ++j) lreg[j] = __ldcg(L + (long)(b * BLK + kq + 4 * j) * ld + c0 + ci);
  };
  if (nb - 1 > c) load_block(nb - 1);
  for (int b = nb - 1; b > c; --b) {
    if (t == 0) {
      long long spins = 0;
      while (ld_acquire(flags + b) == 0)
        if (++spins > (1LL << 28)) __trap();  // a chain bug must end the kernel, not hang the device
    }
    __syncthreads();
    if (t < BLK) ab[t] = __ldcg(alpha + b * BLK + t);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) w_acc += lreg[j] * ab[kq + 4 * j];
    if (b - 1 > c) load_block(b - 1);
  }
  red[kq][ci] = w_acc;
  __syncthreads();
  if (t < 32) {  // one warp: lane holds w[lane] and w[lane + 32]
    double w0 = wsh[t] - ((red[0][t] + red[1][t]) + (red[2][t] + red[3][t]));
    double w1 = wsh[t + 32] - ((red[0][t + 32] + red[1][t + 32]) + (red[2][t + 32] + red[3][t + 32]));
    for (int k = BLK - 1; k >= 0; --k) {
      const double mine = (k >= 32 ? w1 : w0) * rinv[k];
      const double a = __shfl_sync(0xffffffffu, mine, k & 31);
      if (t == (k & 31)) {
        if (k >= 32) w1 = a; else w0 = a;
      }
      if (t < k) w0 -= Ls[k * PSTR + t] * a;
      if (t + 32 < k) w1 -= Ls[k * PSTR + t + 32] * a;
    }
    alpha[c0 + t] = w0;
    alpha[c0 + t + 32] = w1;
    __syncwarp();
    if (t == 0) st_release(flags + c, 1);
  }
}

__global__ void copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

#include "ipp_cholq.cuh"

// ---------------------------------------------------------------------------------------------
// Host-side launch sequences
// ---------------------------------------------------------------------------------------------
template <class K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static bool use_small_tiles(int np) { return np <= 1536; }

static const Bat kOne = {nullptr, 0, 0, 0, 0};  // a single problem in slot 0

// Main loop of the GEMM-shaped fit kernels (A/B measurements): IPP_GEMM_MB=0 a CTA barrier per k-slab, =1 the mbarrier
// pipeline with every warp copying and multiplying, default (=2) warp-specialised where the tile has eight DMMA warps.
static int gemm_pipeline_mode() {
  static const int mode = [] {
    const char* e = getenv("IPP_GEMM_MB");
    return (e && e[0] == '0') ? 0 : ((e && e[0] == '1') ? 1 : 2);
  }();
  return mode;
}
static bool use_mb_pipeline() { return gemm_pipeline_mode() != 0; }


// How the factorisation is run: the cooperative kernel (grid barriers; one problem) or the dataflow kernel driven by
// a task plan (any number of problems, ipp_cholq.cuh).
struct CholRun {
  const int* plan;  // device copy of ipp_host_chol_plan's table; null = cooperative kernel
  int* sched;       // ipp_chol_sched_ints(n, R) ints of scratch
  int use_hints;
  int max_ctas;
  int stats;
  int eager_scan;
  int red;  // tile updates written back as reductions at the L2
  int ws;  // warp-specialised executor (chol_queue_ws_kernel); 0: every warp claims, copies and multiplies in turn
};

static int device_sms(int* sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IPP_ERR_CUDA;
  return IPP_OK;
}

static long chol_sched_ints(int n, int R) {
  const GpDims d = gp_dims(n);
  const int tb = use_small_tiles(d.np) ? 1 : 2;
  const int nb = d.np / BLK, nrb = nb + 1, tmax = (nrb - 1) / tb;
  const long nflags = nb + nrb + (long)(tmax + 1) * (tmax + 2) / 2;
  // flags | queue state | -, abort | per-CTA counters | per-CTA task trace (measurement aids, flags bits 1 / 2)
  return (long)R * nflags + 8L * R + 8 + (long)QSTATS * QSTATS_CTAS + 4L * QTRACE_EVENTS * QTRACE_CTAS;
}

template <class C>
static int chol_queue_launch(double* Kb, double* vbuf, const GpDims& d, const Bat& bat, int R, const CholRun& run,
                             cudaStream_t st) {
  constexpr size_t SYRK_SM = gemm_smem_bytes<C, Lay::KCONT, Lay::KCONT>();
  constexpr size_t SM1 = SYRK_SM > FACT_SMEM ? SYRK_SM : FACT_SMEM;
  constexpr size_t SM = SM1 > solve_smem<C::BM>() ? SM1 : solve_smem<C::BM>();
  constexpr bool CAN_WS = C::THREADS == 256;
  const bool ws = CAN_WS && run.ws;
  constexpr size_t SMW = cholq_ws_ring_bytes<Cfg128>() > SM ? cholq_ws_ring_bytes<Cfg128>() : SM;
  int sms = 0, per_sm = 0;
  if (device_sms(&sms) != IPP_OK) return IPP_ERR_CUDA;
  if (ws) {
    if (opt_in_smem(chol_queue_ws_kernel<Cfg128>, SMW) != cudaSuccess) return IPP_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_queue_ws_kernel<Cfg128>, 384, SMW) != cudaSuccess || per_sm < 1)
      return IPP_ERR_CUDA;
  } else {
    if (opt_in_smem(chol_queue_kernel<C>, SM) != cudaSuccess) return IPP_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_queue_kernel<C>, C::THREADS, SM) != cudaSuccess || per_sm < 1)
      return IPP_ERR_CUDA;
  }
  // no more CTAs than the widest front has work for: every extra CTA only polls
  const int T0 = (d.rows + C::BM - 1) / C::BM;
  long useful = (long)R * ((long)T0 * (T0 + 1) / 2 + d.rows / BLK);
  long ctas = (long)sms * per_sm;
  if (ctas > useful) ctas = useful;
  if (run.max_ctas > 0 && ctas > run.max_ctas) ctas = run.max_ctas;
  if (ctas < 1) ctas = 1;
  if (cudaMemsetAsync(run.sched, 0, sizeof(int) * chol_sched_ints(d.n, R), st) != cudaSuccess) return IPP_ERR_CUDA;
  CholQ a;
  a.K = Kb;
  a.pstride = bat.sK;
  a.ld = d.ld;
  a.rows_total = d.rows;
  a.np = d.np;
  a.vbuf = vbuf;
  a.vstride = bat.sV;
  a.info_off = v_result_off(d) + 3;
  a.slots = bat.slots;
  a.R = R;
  a.plan = run.plan;
  a.sched = run.sched;
  a.timeout = 4000000000LL;  // ~2 s of polling without work: a scheduling bug, not a slow task
  a.use_hints = run.use_hints;
  a.stats = run.stats;
  a.use_mb = use_mb_pipeline() ? 1 : 0;
  a.eager_scan = run.eager_scan;
  a.red_epilogue = run.red;
  if (ws)
    chol_queue_ws_kernel<Cfg128><<<(unsigned)ctas, 384, SMW, st>>>(a);
  else
    chol_queue_kernel<C><<<(unsigned)ctas, C::THREADS, SM, st>>>(a);
  return IPP_OK;
}

template <class C>
static int chol_inplace(double* Kb, double* Wb, double* vbuf, const GpDims& d, const Bat& bat, int R, const CholRun& run,
                        cudaStream_t st) {
  if (run.plan) {
    if (!run.sched) return IPP_ERR_ARG;
    const int rc = chol_queue_launch<C>(Kb, vbuf, d, bat, R, run, st);
    if (rc != IPP_OK) return rc;
  } else {
    if (R != 1 || bat.slots) return IPP_ERR_ARG;  // the cooperative kernel factors one problem
    constexpr size_t SYRK_SM = gemm_smem_bytes<C, Lay::KCONT, Lay::KCONT>();
    constexpr size_t SM = SYRK_SM > PANEL_SMEM ? SYRK_SM : PANEL_SMEM;
    if (opt_in_smem(chol_coop_kernel<C>, SM) != cudaSuccess) return IPP_ERR_CUDA;
    int sms = 0, per_sm = 0;
    if (device_sms(&sms) != IPP_OK) return IPP_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_coop_kernel<C>, C::THREADS, SM) != cudaSuccess || per_sm < 1)
      return IPP_ERR_CUDA;
    // no more CTAs than the first (largest) trailing update has tiles: grid barriers cost with the grid size
    const int T0 = (d.rows - BLK + C::BM - 1) / C::BM;
    int useful = T0 * (T0 + 1) / 2;
    if (useful < d.rows / BLK) useful = d.rows / BLK;
    int ctas = sms * per_sm;
    if (ctas > useful + 1) ctas = useful + 1;  // + the panel CTA
    if (run.max_ctas > 1 && ctas > run.max_ctas) ctas = run.max_ctas;  // leave SMs to concurrent factorisations
    if (ctas < 1) ctas = 1;
    long ld = d.ld;
    int rows = d.rows, np = d.np;
    double* info = vbuf + v_result_off(d) + 3;
    double* phase = nullptr;  // CTA 0 / 1 cycle counters: a debugging aid, off in the product
    void* args[] = {(void*)&Kb, (void*)&ld, (void*)&rows, (void*)&np, (void*)&info, (void*)&phase};
    if (cudaLaunchCooperativeKernel((void*)chol_coop_kernel<C>, dim3(ctas), dim3(C::THREADS), args, SM, st) != cudaSuccess)
      return IPP_ERR_CUDA;
  }
  if (Wb) {
    if (opt_in_smem(diag_inverse_kernel, PANEL_SMEM) != cudaSuccess) return IPP_ERR_CUDA;
    diag_inverse_kernel<<<dim3(d.np / BLK, R), 256, PANEL_SMEM, st>>>(Kb, d.ld, Wb, d.ld, bat);
  }
  return IPP_OK;
}

template <class C, int MODE>
static int trtri_level_t(const double* Kb, double* Wb, double* Tb, const GpDims& d, int parts, int s, const Bat& bat, int R,
                         cudaStream_t st) {
  constexpr size_t SM = MODE ? gemm_mb_smem_bytes<C, Lay::KCONT, Lay::MCONT>() : gemm_smem_bytes<C, Lay::KCONT, Lay::MCONT>();
  constexpr int NT = C::THREADS + (MODE == 2 ? 128 : 0);
  if (opt_in_smem(trtri_level_kernel<C, 1, MODE>, SM) != cudaSuccess) return IPP_ERR_CUDA;
  if (opt_in_smem(trtri_level_kernel<C, 2, MODE>, SM) != cudaSuccess) return IPP_ERR_CUDA;
  const unsigned grid = (unsigned)(((s + C::BN - 1) / C::BN) * ((s + C::BM - 1) / C::BM) * (parts / 2));
  trtri_level_kernel<C, 1, MODE><<<dim3(grid, R), NT, SM, st>>>(Kb, d.ld, Wb, d.ld, Tb, d.ld, d.np / BLK, parts, s, bat);
  trtri_level_kernel<C, 2, MODE><<<dim3(grid, R), NT, SM, st>>>(Kb, d.ld, Wb, d.ld, Tb, d.ld, d.np / BLK, parts, s, bat);
  return IPP_OK;
}
template <class C>
static int trtri_level(const double* Kb, double* Wb, double* Tb, const GpDims& d, int parts, int s, const Bat& bat, int R,
                       cudaStream_t st) {
  const int mode = gemm_pipeline_mode();
  if constexpr (C::THREADS == 256)
    if (mode == 2) return trtri_level_t<C, 2>(Kb, Wb, Tb, d, parts, s, bat, R, st);
  return mode ? trtri_level_t<C, 1>(Kb, Wb, Tb, d, parts, s, bat, R, st) : trtri_level_t<C, 0>(Kb, Wb, Tb, d, parts, s, bat, R, st);
}

template <class C>
static int trtri_inplace(const double* Kb, double* Wb, double* Tb, const GpDims& d, const Bat& bat, int R, cudaStream_t st) {
  const int nb = d.np / BLK;
  w_clear_kernel<<<dim3(nb, nb, R), 256, 0, st>>>(Wb, d.ld, d.np, bat);
  int parts = 1;
  while (parts < nb) parts *= 2;
  for (; parts >= 2; parts /= 2) {  // finest level first
    const int s = BLK * ((nb + parts - 1) / parts);  // the largest half at this level
    const int pairs = nb < parts ? (nb + 1) / 2 : parts / 2;  // pairs with something to merge (finest level: ~nb / 2)
    // levels whose 128x128 tiling cannot fill the GPU run on 64x64 tiles (4x the CTAs, each a quarter of the work)
    const int tiles128 = ((s + 127) / 128) * ((s + 127) / 128) * pairs * R;
    const int rc = (C::BM > 64 && tiles128 < 2 * 148) ? trtri_level<Cfg64>(Kb, Wb, Tb, d, parts, s, bat, R, st)
                                                  : trtri_level<C>(Kb, Wb, Tb, d, parts, s, bat, R, st);
    if (rc != IPP_OK) return rc;
  }
  return IPP_OK;
}

template <class C, int MODE>
static int lml_grad_launch_t(const double* Wb, double* vbuf, const double* hyp, const GpDims& d, int* ntiles, const Bat& bat,
                             int R, const double* tile_bbox, cudaStream_t st) {
  constexpr size_t SM = MODE ? gemm_mb_smem_bytes<C, Lay::MCONT, Lay::MCONT>() : gemm_smem_bytes<C, Lay::MCONT, Lay::MCONT>();
  if (opt_in_smem(lml_grad_kernel<C, MODE>, SM) != cudaSuccess) return IPP_ERR_CUDA;
  const int T = (d.np + C::BM - 1) / C::BM;
  *ntiles = T * (T + 1) / 2;
  lml_grad_kernel<C, MODE><<<dim3(*ntiles, R), C::THREADS + (MODE == 2 ? 128 : 0), SM, st>>>(
      Wb, d.ld, vbuf, hyp, vbuf + v_partials_off(d), d, bat, tile_bbox);
  return IPP_OK;
}
template <class C>
static int lml_grad_launch(const double* Wb, double* vbuf, const double* hyp, const GpDims& d, int* ntiles, const Bat& bat,
                           int R, const double* tile_bbox, cudaStream_t st) {
  const int mode = gemm_pipeline_mode();
  if constexpr (C::THREADS == 256)
    if (mode == 2) return lml_grad_launch_t<C, 2>(Wb, vbuf, hyp, d, ntiles, bat, R, tile_bbox, st);
  return mode ? lml_grad_launch_t<C, 1>(Wb, vbuf, hyp, d, ntiles, bat, R, tile_bbox, st)
              : lml_grad_launch_t<C, 0>(Wb, vbuf, hyp, d, ntiles, bat, R, tile_bbox, st);
}

static int gp_factor(const double* X, const double* y, int n, const double* hyp, double* Kb, double* Wb, double* Tb,
                     double* vbuf, bool invert, const Bat& bat, int R, const CholRun& run, cudaStream_t st) {
  if (n <= 0 || R <= 0 || !X || !y || !hyp || !Kb || !Wb || !vbuf) return IPP_ERR_ARG;
  const GpDims d = gp_dims(n);
  gram_fit_kernel<<<dim3(d.np / BLK, d.rows / BLK, R), 256, 0, st>>>(X, y, hyp, Kb, vbuf, d, bat);
  int rc = use_small_tiles(d.np) ? chol_inplace<Cfg64>(Kb, Wb, vbuf, d, bat, R, run, st)
                                 : chol_inplace<Cfg128>(Kb, Wb, vbuf, d, bat, R, run, st);
  if (rc != IPP_OK) return rc;
  if (invert) {
    if (!Tb) return IPP_ERR_ARG;
    rc = use_small_tiles(d.np) ? trtri_inplace<Cfg64>(Kb, Wb, Tb, d, bat, R, st) : trtri_inplace<Cfg128>(Kb, Wb, Tb, d, bat, R, st);
    if (rc != IPP_OK) return rc;
  }
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

// One LML (+ gradient) evaluation of each of the R problems of a batch.
static int gp_lml_grad_run(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                           double* Tbuf, double* vbuf, int with_grad, const Bat& bat, int R, const CholRun& run,
                           const double* tile_bbox, cudaStream_t st) {
  int rc = gp_factor(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, with_grad != 0, bat, R, run, st);
  if (rc != IPP_OK) return rc;
  const GpDims d = gp_dims(n);
  int ntiles = 0;
  if (with_grad) {
    // partials go to Tbuf, which the triangular inverse is done with
    gemv_wt_kernel<<<dim3(d.np / BLK, GEMV_SPLIT, R), 256, 0, st>>>(Wbuf, d.ld, Kbuf + (long)d.np * d.ld, Tbuf, d.np, bat);
    gemv_reduce_kernel<<<dim3((d.np + 255) / 256, R), 256, 0, st>>>(Tbuf, vbuf + V_ALPHA * d.npv, d.np, bat);
    rc = use_small_tiles(d.np) ? lml_grad_launch<Cfg64>(Wbuf, vbuf, hyp, d, &ntiles, bat, R, tile_bbox, st)
                               : lml_grad_launch<Cfg128>(Wbuf, vbuf, hyp, d, &ntiles, bat, R, tile_bbox, st);
    if (rc != IPP_OK) return rc;
  }
  lml_finalize_kernel<<<R, 256, 0, st>>>(Kbuf, vbuf, ntiles, with_grad, d, bat);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

// ---------------------------------------------------------------------------------------------
// Mutual-information term of the stage-1 selector (rrt.py:512-526):
//   0.5 * log det(I_L + beta * K_pio K_pio^T),  K_pio = prior kernel(leaves, agent observations).
// By Sylvester's identity the determinant equals det(I_na + beta * K_pio^T K_pio), so the smaller of
// the two orders is factorised.  The reference evaluates it L times with identical arguments and
// builds the unused L x L matrix K_pi (rrt.py:447,457-459); here it is one SYRK + one Cholesky.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_cross_padded_kernel(const double* __restrict__ X, int n,
                                                                const double* __restrict__ Y, int m, double c, double l,
                                                                double* __restrict__ out, long ld) {
  __shared__ double sx[2][BLK], sy[2][BLK];
  const int i0 = blockIdx.y * BLK, j0 = blockIdx.x * BLK, t = threadIdx.x;
  if (t < 2 * BLK) {
    int which = t / BLK, e = t % BLK;
    const double* P = which ? Y : X;
    int lim = which ? m : n, idx = (which ? j0 : i0) + e;
    double vx = 0.0, vy = 0.0;
    if (idx < lim) {
      vx = __ddiv_rn(P[2 * idx], l);
      vy = __ddiv_rn(P[2 * idx + 1], l);
    }
    sx[which][e] = vx;
    sy[which][e] = vy;
  }
  __syncthreads();
  for (int e = t; e < BLK * BLK; e += 256) {
    int r = e / BLK, cc = e % BLK;
    int gi = i0 + r, gj = j0 + cc;
    out[(long)gi * ld + gj] = (gi < n && gj < m) ? rbf_scaled(sx[0][r], sy[0][r], sx[1][cc], sy[1][cc], c) : 0.0;
  }
}

template <class C, bool TRANS>
__global__ void __launch_bounds__(C::THREADS) eye_plus_syrk_kernel(const double* __restrict__ M, long ldm, int rows_m,
                                                                   int cols_m, double beta, double* __restrict__ G,
                                                                   GpDims d) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((long)ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int row0 = ti * C::BM, col0 = tj * C::BN;
  double acc[C::MI][C::NI][2];
  zero_acc<C>(acc);
  if (!TRANS) {
    auto fillA = [&](double* sm, int k0) { load_tile_async<C::BM, Lay::KCONT, C::THREADS>(sm, M, ldm, row0, k0, rows_m, cols_m); };
    auto fillB = [&](double* sm, int k0) { load_tile_async<C::BN, Lay::KCONT, C::THREADS>(sm, M, ldm, col0, k0, rows_m, cols_m); };
    gemm_mainloop<C, Lay::KCONT, Lay::KCONT>(acc, smem, 0, cols_m, fillA, fillB);
  } else {
    auto fillA = [&](double* sm, int k0) { load_tile_async<C::BM, Lay::MCONT, C::THREADS>(sm, M, ldm, row0, k0, cols_m, rows_m); };
    auto fillB = [&](double* sm, int k0) { load_tile_async<C::BN, Lay::MCONT, C::THREADS>(sm, M, ldm, col0, k0, cols_m, rows_m); };
    gemm_mainloop<C, Lay::MCONT, Lay::MCONT>(acc, smem, 0, rows_m, fillA, fillB);
  }
  for_each_acc<C>(acc, [&](int r, int c, double v) {
    int gi = row0 + r, gj = col0 + c;
    if (gi < d.np && gj < d.np && gj <= gi) G[(long)gi * d.ld + gj] = ((gi == gj) ? 1.0 : 0.0) + beta * v;
  });
}

template <class C>
static int mi_build(const double* Gx, long ldg, int Lp, int nap, bool trans, double beta, double* Kb, const GpDims& d,
                    cudaStream_t st) {
  const int T = (d.np + C::BM - 1) / C::BM;
  if (!trans) {
    constexpr size_t SM = gemm_smem_bytes<C, Lay::KCONT, Lay::KCONT>();
    if (opt_in_smem(eye_plus_syrk_kernel<C, false>, SM) != cudaSuccess) return IPP_ERR_CUDA;
    eye_plus_syrk_kernel<C, false><<<T * (T + 1) / 2, C::THREADS, SM, st>>>(Gx, ldg, Lp, nap, beta, Kb, d);
  } else {
    constexpr size_t SM = gemm_smem_bytes<C, Lay::MCONT, Lay::MCONT>();
    if (opt_in_smem(eye_plus_syrk_kernel<C, true>, SM) != cudaSuccess) return IPP_ERR_CUDA;
    eye_plus_syrk_kernel<C, true><<<T * (T + 1) / 2, C::THREADS, SM, st>>>(Gx, ldg, Lp, nap, beta, Kb, d);
  }
  return IPP_OK;
}

}  // namespace ipp

using namespace ipp;

extern "C" {

int ipp_gp_layout(int n, long long* out) {
  if (n <= 0 || !out) return IPP_ERR_ARG;
  const GpDims d = gp_dims(n);
  out[0] = d.np;
  out[1] = d.rows;
  out[2] = d.ld;
  out[3] = (long long)d.rows * d.ld;  // Kbuf doubles
  out[4] = (long long)d.np * d.ld;    // Wbuf doubles
  out[5] = (long long)d.np * d.ld;    // Tbuf doubles
  out[6] = v_total(d);                // vbuf doubles
  out[7] = V_ALPHA * d.npv;           // offset of alpha in vbuf
  out[8] = v_result_off(d);           // offset of the 8-double result block
  out[9] = d.npv;                     // vector slot length (slot order: xs, ys, y, w, alpha)
  return IPP_OK;
}

int ipp_gram_rbf(const double* X, int n, const double* Y, int m, double c, double l, double jitter, double* K,
                 long long ldk, void* stream) {
  return ipp_gram_kernel(0, X, n, Y, m, c, l, jitter, K, ldk, stream);
}

int ipp_gram_kernel(int kind, const double* X, int n, const double* Y, int m, double c, double l, double jitter, double* K,
                    long long ldk, void* stream) {
  if (kind < 0 || kind > 2 || !X || n <= 0 || !K || !(l > 0.0)) return IPP_ERR_ARG;
  const int symmetric = (Y == nullptr);
  if (symmetric) {
    Y = X;
    m = n;
  }
  if (m <= 0 || ldk < m) return IPP_ERR_ARG;
  if (symmetric) {
    const long T = (n + BLK - 1) / BLK;
    gram_sym_kernel<<<(unsigned)(T * (T + 1) / 2), 256, 0, (cudaStream_t)stream>>>(X, n, c, l, jitter, K, (long)ldk, kind);
  } else {
    dim3 grid((m + BLK - 1) / BLK, (n + BLK - 1) / BLK);
    gram_rbf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, n, Y, m, c, l, jitter, 0, K, (long)ldk, kind);
  }
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_gp_lml_grad(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                    double* Tbuf, double* vbuf, int with_grad, int max_ctas, void* stream) {
  const CholRun run = {nullptr, nullptr, 0, max_ctas, 0, 0, 0, 0};
  return gp_lml_grad_run(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, with_grad, kOne, 1, run, nullptr, (cudaStream_t)stream);
}

int ipp_host_chol_plan(int n, int near_slack, int* out_host, long long capacity, long long* need) {
  if (n <= 0 || !need || near_slack < 0) return IPP_ERR_ARG;
  const GpDims d = gp_dims(n);
  std::vector<int> plan;
  chol_plan_build(d.np / BLK, use_small_tiles(d.np) ? 1 : 2, near_slack, plan);
  *need = (long long)plan.size();
  if (out_host) {
    if (capacity < (long long)plan.size()) return IPP_ERR_ARG;
    for (size_t i = 0; i < plan.size(); ++i) out_host[i] = plan[i];
  }
  return IPP_OK;
}

int ipp_chol_sched_ints(int n, int R, long long* out) {
  if (n <= 0 || R <= 0 || !out) return IPP_ERR_ARG;
  *out = chol_sched_ints(n, R);
  return IPP_OK;
}

int ipp_gp_lml_grad_batch(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                          double* Tbuf, double* vbuf, const int* slots, int R, int with_grad, const int* plan, int* sched,
                          int flags, const double* tile_bbox, void* stream) {
  if (R <= 0 || !plan || !sched) return IPP_ERR_ARG;
  const GpDims d = gp_dims(n > 0 ? n : 1);
  const Bat bat = {slots, (long)d.rows * d.ld, (long)d.np * d.ld, (long)d.np * d.ld, v_total(d)};
  const CholRun run = {plan, sched, (flags & 1) ? 0 : 1, flags >> 8, (flags & 4) ? 2 : ((flags >> 1) & 1), (flags >> 3) & 1,
                       (flags & 32) ? 0 : 1, (flags & 16) ? 0 : 1};
  return gp_lml_grad_run(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, with_grad, bat, R, run, tile_bbox, (cudaStream_t)stream);
}

static int gp_fit_final_run(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                            double* Tbuf, double* vbuf, const CholRun& run, cudaStream_t st) {
  int rc = gp_factor(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, true, kOne, 1, run, st);
  if (rc != IPP_OK) return rc;
  const GpDims d = gp_dims(n);
  double* w = vbuf + V_W * d.npv;
  double* alpha = vbuf + V_ALPHA * d.npv;
  const int nb = d.np / BLK;
  int sms = 0;
  if (device_sms(&sms) != IPP_OK) return IPP_ERR_CUDA;
  if (nb <= BACKSUB_MAX_BLOCKS && nb <= 4 * sms) {  // every CTA resident (a few per SM): one launch, flags in Tbuf (free now)
    int* flags = reinterpret_cast<int*>(Tbuf);
    if (cudaMemsetAsync(flags, 0, sizeof(int) * nb, st) != cudaSuccess) return IPP_ERR_CUDA;
    backsub_chain_kernel<<<nb, 256, 0, st>>>(Kbuf, d.ld, Kbuf + (long)d.np * d.ld, alpha, flags, nb);
  } else {
    copy_row_kernel<<<(d.np + 255) / 256, 256, 0, st>>>(Kbuf + (long)d.np * d.ld, w, d.np);
    for (int b = nb - 1; b >= 0; --b) backsub_step_kernel<<<b + 1, BLK, 0, st>>>(Kbuf, d.ld, w, alpha, b);
  }
  lml_finalize_kernel<<<1, 256, 0, st>>>(Kbuf, vbuf, 0, 0, d, kOne);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_gp_fit_final(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                     double* Tbuf, double* vbuf, void* stream) {
  const CholRun run = {nullptr, nullptr, 0, 0, 0, 0, 0, 0};
  return gp_fit_final_run(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, run, (cudaStream_t)stream);
}

int ipp_gp_fit_final_plan(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                          double* Tbuf, double* vbuf, const int* plan, int* sched, int flags, void* stream) {
  const CholRun run = {plan, plan ? sched : nullptr, (flags & 1) ? 0 : 1, 0, 0, (flags >> 3) & 1, (flags & 32) ? 0 : 1, (flags & 16) ? 0 : 1};
  return gp_fit_final_run(X, y, n, hyp, Kbuf, Wbuf, Tbuf, vbuf, run, (cudaStream_t)stream);
}

/* result block (vbuf + out[8] of ipp_gp_layout(min(L, na))): [5] = 0.5 * log det(I + beta K K^T), [3] = info */
int ipp_mi_logdet(const double* leaf_xy, int L, const double* obs_xy, int na, double c, double l, double beta,
                  double* Gtmp, double* Kbuf, double* Wbuf, double* vbuf, void* stream) {
  if (L <= 0 || na <= 0 || !leaf_xy || !obs_xy || !Gtmp || !Kbuf || !Wbuf || !vbuf || !(l > 0.0)) return IPP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int Lp = round_up(L, BLK), nap = round_up(na, BLK);
  const bool trans = na < L;
  const GpDims d = gp_dims(trans ? na : L);
  gram_cross_padded_kernel<<<dim3(nap / BLK, Lp / BLK), 256, 0, st>>>(leaf_xy, L, obs_xy, na, c, l, Gtmp, nap);
  if (cudaMemsetAsync(Kbuf + (long)d.np * d.ld, 0, sizeof(double) * BLK * d.ld, st) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaMemsetAsync(vbuf + v_result_off(d), 0, sizeof(double) * 8, st) != cudaSuccess) return IPP_ERR_CUDA;
  int rc = use_small_tiles(d.np) ? mi_build<Cfg64>(Gtmp, nap, Lp, nap, trans, beta, Kbuf, d, st)
                                 : mi_build<Cfg128>(Gtmp, nap, Lp, nap, trans, beta, Kbuf, d, st);
  if (rc != IPP_OK) return rc;
  const CholRun run = {nullptr, nullptr, 0, 0, 0, 0, 0, 0};
  rc = use_small_tiles(d.np) ? chol_inplace<Cfg64>(Kbuf, Wbuf, vbuf, d, kOne, 1, run, st)
                             : chol_inplace<Cfg128>(Kbuf, Wbuf, vbuf, d, kOne, 1, run, st);
  if (rc != IPP_OK) return rc;
  lml_finalize_kernel<<<1, 256, 0, st>>>(Kbuf, vbuf, 0, 0, d, kOne);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

}  // extern "C"

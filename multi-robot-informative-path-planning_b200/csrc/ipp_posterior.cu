// GP posterior mean / standard deviation over the workspace lattice or explicit query points
// (SURVEY.md §8a rows G5, G6).  Two kernels: posterior_kernel (generic: explicit points, k* by exp()) and
// posterior_lattice_kernel (the workspace lattice: separable coordinate tables, warp-specialised cp.async /
// mbarrier pipeline, block-sparse k*) -- see the comment blocks above each.  Replaces gp.predict(X, return_std=True)
// (sklearn/gaussian_process/_gpr.py:446-500 reached from src/point_source/point_source.py:351 and
// src/rrt/rrt_thread.py:375).
//
// Formulation: sigma^2(q) = c - |W k*(q)|^2 with W = L^-1 explicit (the K^-1 quadratic
// form is numerically unusable at cond 1e11; the per-tile TRSM panel does not fit on chip).  So the
// kernel is a lower-triangular fp64 GEMM  V = W (NP x NP) * K*^T (NP x 128 queries)  on the DMMA
// pipe whose B operand is never read from memory -- each k-slab of K* is generated from coordinates
// straight into shared memory -- and whose C tile is never written: the epilogue squares and
// column-reduces it in registers.  The mean k*^T alpha rides along in the generation pass.
//
// Per 128-query tile: flops = 128 * (NP^2 + ...) fp64, HBM bytes = 16/query out (+16 in for points);
// W is streamed from L2 (all CTAs walk it in the same order).  Bound: fp64 tensor pipe.
#include "ipp_gemm.cuh"
#include "../../include/ipp_b200.h"

namespace ipp {

struct PostArgs {
  const double* W;
  long ldw;
  const double* xs;  // observation coordinates / l, zero padded to NP
  const double* ys;
  const double* alpha;
  int n, np;
  double c, l, y_mean, y_std;
  const double* Q;  // explicit points (m x 2) or nullptr
  double x0, x1, xstep, y0, y1, ystep;
  int nx, ny;
  long g_begin, g_end;
  double* mean;
  double* std;
  double* zpred;
  int* neg_count;
  const double* slab_bbox;  // [np / 16][4] = {xmin, xmax, ymin, ymax} of each 16-observation slab, or nullptr
  double cutoff2;           // skip a slab when its box is farther than sqrt(cutoff2) from the tile's queries; <= 0: never
  const double* tile_bbox;  // tabled explicit points: [tile][4] = {xmin, xmax, ymin, ymax} of each 128-query tile, else nullptr
};

using PC = Cfg128;
constexpr int QT = PC::BN;  // queries per tile

template <bool GRID>
__global__ void __launch_bounds__(PC::THREADS, 1) posterior_kernel(PostArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double sq[2][QT];            // scaled query coordinates
  __shared__ double red[PC::WARPS_M][QT]; // column sums of squares per warp row
  __shared__ double mred[2][QT];          // mean partials per generation half
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int wm = warp / PC::WARPS_N, wn = warp % PC::WARPS_N;
  const int g = lane >> 2, tg = lane & 3;
  const int p = t & (QT - 1), h = t >> 7;  // generation role: query p, k-half h
  const long total = a.g_end - a.g_begin;
  const long ntiles = (total + QT - 1) / QT;

  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (t < QT) {
      long gi = a.g_begin + tile * QT + t;
      double qx = 0.0, qy = 0.0;
      if (gi < a.g_end) {
        if (GRID) {
          int ix = (int)(gi % a.nx), iy = (int)(gi / a.nx);
          // np.linspace: i * step + start with separate roundings, endpoint forced
          qx = (ix == a.nx - 1 && a.nx > 1) ? a.x1 : __dadd_rn(__dmul_rn((double)ix, a.xstep), a.x0);
          qy = (iy == a.ny - 1 && a.ny > 1) ? a.y1 : __dadd_rn(__dmul_rn((double)iy, a.ystep), a.y0);
        } else {
          qx = a.Q[2 * gi];
          qy = a.Q[2 * gi + 1];
        }
      }
      sq[0][t] = __ddiv_rn(qx, a.l);
      sq[1][t] = __ddiv_rn(qy, a.l);
    }
    __syncthreads();
    const double qx = sq[0][p], qy = sq[1][p];
    double csq[PC::NI][2];
#pragma unroll
    for (int ni = 0; ni < PC::NI; ++ni) csq[ni][0] = csq[ni][1] = 0.0;
    double mean_part = 0.0;

    for (int row0 = 0; row0 < a.np; row0 += PC::BM) {
      double acc[PC::MI][PC::NI][2];
      zero_acc<PC>(acc);
      const int kend = min(row0 + PC::BM, a.np);
      auto fillA = [&](double* s, int k0) {
        load_tile_async<PC::BM, Lay::KCONT, PC::THREADS>(s, a.W, a.ldw, row0, k0, a.np, a.np);
      };
      auto fillB = [&](double* s, int k0) {
        const bool first_visit = (k0 >= row0);  // slab k0 is generated for every row block >= k0/BM
        double v[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int j = k0 + h * 8 + kk;
          double val = 0.0;
          if (j < a.n) {
            val = rbf_scaled(qx, qy, __ldg(a.xs + j), __ldg(a.ys + j), a.c);
            if (first_visit) mean_part += __ldg(a.alpha + j) * val;
          }
          v[kk] = val;
        }
        // k-major tile: consecutive threads (queries) hit consecutive 8-byte banks
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) s[(h * 8 + kk) * (QT + 4) + p] = v[kk];
      };
      gemm_mainloop<PC, Lay::KCONT, Lay::MCONT>(acc, smem, 0, kend, fillA, fillB);
#pragma unroll
      for (int mi = 0; mi < PC::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < PC::NI; ++ni) {
          csq[ni][0] += acc[mi][ni][0] * acc[mi][ni][0];
          csq[ni][1] += acc[mi][ni][1] * acc[mi][ni][1];
        }
    }
    // column reduction: over the 8 row-lanes g of the warp, then over the WARPS_M warp rows
#pragma unroll
    for (int ni = 0; ni < PC::NI; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = csq[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) red[wm][wn * PC::WTN + ni * 8 + tg * 2 + e] = v;
      }
    mred[h][p] = mean_part;
    __syncthreads();
    if (t < QT) {
      long gi = a.g_begin + tile * QT + t;
      if (gi < a.g_end) {
        double ss = 0.0;
#pragma unroll
        for (int w = 0; w < PC::WARPS_M; ++w) ss += red[w][t];
        double mu = a.y_std * (mred[0][t] + mred[1][t]) + a.y_mean;
        double var = a.c - ss;
        if (var < 0.0) {
          var = 0.0;
          if (a.neg_count) atomicAdd(a.neg_count, 1);
        }
        long o = gi - a.g_begin;
        a.mean[o] = mu;
        a.std[o] = sqrt(var * (a.y_std * a.y_std));
        if (a.zpred) a.zpred[o] = pow(10.0, mu);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Lattice fast path.  On the workspace lattice the RBF cross-covariance factorises,
//   k*(q, j) = [c exp(-(qy - yj)^2 / 2l^2)] * [exp(-(qx - xj)^2 / 2l^2)] = Ey[iy][j] * Ex[ix][j],
// and a lattice has only nx + ny distinct coordinates: two small tables (nx x NP and ny x NP) replace
// the G x n exponentials (each of which the generic kernel re-evaluates once per 128-row block of W).
// The B operand then arrives by cp.async exactly like W: a tile of 128 Ex rows plus the two Ey rows a
// 128-query tile can touch (nx >= 128), multiplied together when the fragment is read from shared
// memory.  The main loop is a uniform 4-stage async pipeline flattened over (row block, k slab), so
// the fp64 pipe sees DMMAs back to back instead of alternating with latency-bound exp() phases.
// ---------------------------------------------------------------------------------------------
constexpr int SEP_STAGES = 4;
constexpr int SEP_A = PC::BM * KPAD;
constexpr int SEP_B = QT * KPAD;
constexpr int SEP_E = 2 * BK;
constexpr int SEP_STAGE = SEP_A + SEP_B + SEP_E;
constexpr size_t SEP_SMEM = sizeof(double) * SEP_STAGES * SEP_STAGE;

// T[i][k] = scale * exp(-0.5 * (q_i / l - xs_k)^2) for lattice coordinate i in [i_begin, i_begin + gridDim.y),
// k < n; zero for the padded k.  q_i follows np.linspace.
__global__ void __launch_bounds__(256) axis_table_kernel(const double* __restrict__ xs, int n, int np, double l, double q0,
                                                         double q1, double qstep, int nq, double scale, int i_begin,
                                                         double* __restrict__ T, long ldt) {
  const int i = i_begin + blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= np || i >= nq) return;
  const double q = (i == nq - 1 && nq > 1) ? q1 : __dadd_rn(__dmul_rn((double)i, qstep), q0);
  double v = 0.0;
  if (k < n) {
    const double d = __dsub_rn(__ddiv_rn(q, l), xs[k]);
    v = __dmul_rn(exp(__dmul_rn(-0.5, __dmul_rn(d, d))), scale);
  }
  T[(long)i * ldt + k] = v;
}

constexpr int LAT_THREADS = PC::THREADS + 128;  // 8 DMMA (consumer) warps + one copy (producer) warpgroup
constexpr int LAT_MASK_WORDS = 32;              // 1024 slabs = 16384 observations; with more, no slab is skipped (both
                                                // loops then read "all active" instead of the mask: see the ternaries)

// Block sparsity of the cross-covariance.  With the observations in spatial (Morton) order a 16-observation
// slab is a compact cluster; when its bounding box is farther than R = l * sqrt(2 ln(1/eps)) from every query
// of the tile, all 128 x 16 entries of k* are below eps * c (eps = 1e-30 by default: twelve orders of magnitude
// under the rounding of the terms that are kept), so the slab is skipped -- no copy, no DMMA, no stage.  Each warp
// evaluates the same test on the same data (one slab per lane, one ballot per 32 slabs), so the copy warps and
// the DMMA warps agree on the slab sequence without exchanging anything.
__device__ __forceinline__ void lattice_slab_mask(const PostArgs& a, unsigned* mask, int lane, long gq0, int nvalid) {
  const int nslab = a.np / BK;
  const int words = (nslab + 31) / 32;
  const bool on = a.slab_bbox != nullptr && a.cutoff2 > 0.0 && words <= LAT_MASK_WORDS;
  const int iy0 = (int)(gq0 / a.nx), ix0 = (int)(gq0 % a.nx);
  const int ixl = ix0 + nvalid - 1;
  double qx0, qx1, qy0, qy1;
  qy0 = a.y0 + iy0 * a.ystep;
  if (a.tile_bbox) {  // explicit points behind a k* table: the box of the tile's own queries
    const double4 tb = *reinterpret_cast<const double4*>(a.tile_bbox + 4 * ((gq0 - a.g_begin) / QT));
    qx0 = tb.x;
    qx1 = tb.y;
    qy0 = tb.z;
    qy1 = tb.w;
  } else if (ixl < a.nx) {  // one lattice row
    qx0 = a.x0 + ix0 * a.xstep;
    qx1 = a.x0 + ixl * a.xstep;
    qy1 = qy0;
  } else {  // the tile wraps into the next row: take the whole x range
    qx0 = a.x0;
    qx1 = a.x1;
    qy1 = a.y0 + (iy0 + 1) * a.ystep;
  }
  for (int w = 0; w < min(words, LAT_MASK_WORDS); ++w) {
    const int sidx = 32 * w + lane;
    bool act = sidx < nslab;
    if (act && on) {
      const double4 b = *reinterpret_cast<const double4*>(a.slab_bbox + 4 * sidx);
      const double gx = fmax(0.0, fmax(b.x - qx1, qx0 - b.y)), gy = fmax(0.0, fmax(b.z - qy1, qy0 - b.w));
      act = gx * gx + gy * gy <= a.cutoff2;
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (lane == 0) mask[w] = m;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(LAT_THREADS, 1) posterior_lattice_kernel(PostArgs a, const double* __restrict__ Ex,
                                                                           const double* __restrict__ Ey, long ldt) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[PC::WARPS_M][QT];
  __shared__ double mred[2][QT];
  __shared__ double csq_s[2 * PC::NI][PC::THREADS];  // per-thread column norms (kept out of the register file)
  __shared__ __align__(8) uint64_t bar_full[SEP_STAGES], bar_empty[SEP_STAGES];
  __shared__ unsigned smask[LAT_THREADS / 32][LAT_MASK_WORDS];  // per warp: which k slabs this tile needs
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long total = a.g_end - a.g_begin;
  const long ntiles = (total + QT - 1) / QT;
  // Row blocks of W.  Only rows < n matter (the padding rows of W are identity rows times a zero k*), and
  // the short block goes FIRST, where its k range is short too: rows [0, first), then full 128-row blocks.
  // Block j covers rows [row0, rlim) and, W being lower triangular, k in [0, round_up(rlim, 16)).
  const int first = a.n % PC::BM;
  const int nrb = a.n / PC::BM + (first ? 1 : 0);
  auto blk_row0 = [&](int j) { return first ? (j ? first + (j - 1) * PC::BM : 0) : j * PC::BM; };
  auto blk_rlim = [&](int j) { return (first && j == 0) ? first : blk_row0(j) + PC::BM; };
  auto blk_kend = [&](int j) { return (blk_rlim(j) + BK - 1) / BK * BK; };
  if (t == 0) {
    for (int s = 0; s < SEP_STAGES; ++s) {
      mbar_init(&bar_full[s], 128);                // one cp.async-completion arrival per copy thread
      mbar_init(&bar_empty[s], PC::THREADS / 32);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // Slabs are numbered globally across tiles: slab G lives in stage G % STAGES and is that stage's
  // (G / STAGES)-th use, which fixes the phase parity of both of its barriers.
  unsigned gslab = 0;  // next slab of this thread (producer: to copy, consumer: to multiply)

  if (warp >= PC::THREADS / 32) {
    // Register file: the copy warpgroup gives its registers back so that the DMMA warps can hold their
    // 128 accumulator registers plus fragments without spilling (384 threads compile for 168 each).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    const int pt = t - PC::THREADS;  // 0..127 within the copy warpgroup
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long gq0 = a.g_begin + tile * QT;
      const int iy0 = (int)(gq0 / a.nx), ix0 = (int)(gq0 % a.nx);
      const int iy1 = min(iy0 + 1, a.ny - 1);
      const int nvalid = (int)min((long)QT, a.g_end - gq0);
      // ---- producer warpgroup: 16-byte cp.async (LDGSTS) into the padded, bank-conflict-free stage layout; each
      // thread's copies complete on the stage's full barrier (cp.async.mbarrier.arrive.noinc).  Two alternatives
      // were measured and rejected at C4 (769 ms with this version): per-row 128-byte bulk copies (UBLKCP), 3x
      // slower -- ~260 small bulk requests per slab swamp the unit; and multiplying Ex * Ey in this warpgroup
      // (LDG, DMUL, STS) so that the DMMA warps load finished k* values, 880-900 ms -- the copy warps then sit on
      // L2 latency with too few registers to keep a slab of loads in flight.
      lattice_slab_mask(a, smask[warp], lane, gq0, nvalid);
      for (int prb = 0; prb < nrb; ++prb) {
        const int p_row0 = blk_row0(prb), p_rlim = blk_rlim(prb), nsl = blk_kend(prb) / BK;
        for (int w = 0; 32 * w < nsl; ++w) {  // the active slabs below the block's last column, by set bit
          unsigned m = w < LAT_MASK_WORDS ? smask[warp][w] : 0xffffffffu;  // beyond the mask: every slab is active
          if (nsl - 32 * w < 32) m &= (1u << (nsl - 32 * w)) - 1u;
          while (m) {
            const int pk = (32 * w + __ffs(m) - 1) * BK;
            m &= m - 1;
            const unsigned stage = gslab % SEP_STAGES, use = gslab / SEP_STAGES;
            if (use > 0) mbar_wait(&bar_empty[stage], (use - 1) & 1);  // every consumer warp is done with the old contents
            double* s = smem + stage * SEP_STAGE;
            load_tile_async<PC::BM, Lay::KCONT, 128>(s, a.W, a.ldw, p_row0, pk, p_rlim, a.np, pt);
#pragma unroll
            for (int q = 0; q < QT * BK / 2 / 128; ++q) {
              const int c = pt + q * 128;
              const int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
              int ix = ix0 + min(r, nvalid - 1);
              if (ix >= a.nx) ix -= a.nx;
              cp_async16(s + SEP_A + r * KPAD + kc, Ex + (long)ix * ldt + pk + kc, true);
            }
            if (pt < BK) {
              const int row = pt >> 3, kc = (pt & 7) * 2;
              cp_async16(s + SEP_A + SEP_B + row * BK + kc, Ey + (long)(row ? iy1 : iy0) * ldt + pk + kc, true);
            }
            cp_async_mbar_arrive(&bar_full[stage]);
            ++gslab;
          }
        }
      }
    }
    cp_async_wait<0>();
    return;  // the copy warps run ahead across tile boundaries, held back only by the empty barriers
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long gq0 = a.g_begin + tile * QT;
    const int ix0 = (int)(gq0 % a.nx);
    const int nvalid = (int)min((long)QT, a.g_end - gq0);
    {
      // ---- consumer warps: DMMA only ----
      const int wm = warp / PC::WARPS_N, wn = warp % PC::WARPS_N;
      const int g = lane >> 2, tg = lane & 3;
      const int p = t & (QT - 1), h = t >> 7;
      int eoff[PC::NI];  // offset of my B column's lattice row inside the stage's two Ey rows
#pragma unroll
      for (int ni = 0; ni < PC::NI; ++ni) eoff[ni] = ((ix0 + wn * PC::WTN + ni * 8 + g) >= a.nx ? BK : 0) + tg;
      const int prow = (ix0 + p >= a.nx) ? BK : 0;
      double acc[PC::MI][PC::NI][2];
      zero_acc<PC>(acc);
#pragma unroll
      for (int i = 0; i < 2 * PC::NI; ++i) csq_s[i][t] = 0.0;
      double mean_part = 0.0;
      lattice_slab_mask(a, smask[warp], lane, gq0, nvalid);
      for (int crb = 0; crb < nrb; ++crb) {
        const int nsl = blk_kend(crb) / BK;
        for (int w = 0; 32 * w < nsl; ++w) {  // the active slabs below the block's last column, by set bit
          unsigned m = w < LAT_MASK_WORDS ? smask[warp][w] : 0xffffffffu;  // beyond the mask: every slab is active
          if (nsl - 32 * w < 32) m &= (1u << (nsl - 32 * w)) - 1u;
          while (m) {
            const int ck = (32 * w + __ffs(m) - 1) * BK;
            m &= m - 1;
            const unsigned stage = gslab % SEP_STAGES;
            mbar_wait(&bar_full[stage], (gslab / SEP_STAGES) & 1);
            const double* sA = smem + stage * SEP_STAGE + (wm * PC::WTM + g) * KPAD + tg;
            const double* sB = smem + stage * SEP_STAGE + SEP_A + (wn * PC::WTN + g) * KPAD + tg;
            const double* sE = smem + stage * SEP_STAGE + SEP_A + SEP_B;
            // B fragments (Ex * Ey) are prepared one k-step ahead of the DMMAs that use them.  The loads and the
            // multiply are volatile asm, issued between the two halves of the previous k-step's DMMA burst, so that
            // ptxas cannot sink them to the point of use (it does when they are plain C++).
            const uint32_t aB = (uint32_t)__cvta_generic_to_shared(sB);
            const uint32_t aE = (uint32_t)__cvta_generic_to_shared(sE);
            auto lds64 = [](uint32_t addr) {
              double v;
              asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
              return v;
            };
            auto vmul = [](double x, double y) {
              double v;
              asm volatile("mul.rn.f64 %0, %1, %2;\n" : "=d"(v) : "d"(x), "d"(y));
              return v;
            };
            double bf[PC::NI], bn[PC::NI];
#pragma unroll
            for (int ni = 0; ni < PC::NI; ++ni)
              bf[ni] = vmul(lds64(aB + 8 * (ni * 8 * KPAD)), lds64(aE + 8 * eoff[ni]));
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
              double af[PC::MI];
#pragma unroll
              for (int mi = 0; mi < PC::MI; ++mi) af[mi] = sA[mi * 8 * KPAD + kk];
#pragma unroll
              for (int mi = 0; mi < PC::MI / 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < PC::NI; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
              if (kk + 4 < BK) {
#pragma unroll
                for (int ni = 0; ni < PC::NI; ++ni)
                  bn[ni] = vmul(lds64(aB + 8 * (ni * 8 * KPAD + kk + 4)), lds64(aE + 8 * (eoff[ni] + kk + 4)));
              }
#pragma unroll
              for (int mi = PC::MI / 2; mi < PC::MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < PC::NI; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
#pragma unroll
              for (int ni = 0; ni < PC::NI; ++ni) bf[ni] = bn[ni];
            }
            if (crb == nrb - 1) {  // the last row block sees every k slab once: the mean rides along there
              const double* sBq = smem + stage * SEP_STAGE + SEP_A + p * KPAD;
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const int k = h * 8 + kk;
                mean_part += __ldg(a.alpha + ck + k) * __dmul_rn(sBq[k], sE[prow + k]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[stage]);  // this warp no longer reads the stage
            ++gslab;
          }
        }
        // row block finished: fold its rows into the column norms
#pragma unroll
        for (int ni = 0; ni < PC::NI; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = csq_s[2 * ni + e][t];
#pragma unroll
            for (int mi = 0; mi < PC::MI; ++mi) {
              v += acc[mi][ni][e] * acc[mi][ni][e];
              acc[mi][ni][e] = 0.0;
            }
            csq_s[2 * ni + e][t] = v;
          }
      }
#pragma unroll
      for (int ni = 0; ni < PC::NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = csq_s[2 * ni + e][t];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) red[wm][wn * PC::WTN + ni * 8 + tg * 2 + e] = v;
        }
      mred[h][p] = mean_part;
    }
    asm volatile("bar.sync 1, 256;\n" ::: "memory");  // consumer warps only
    if (t < nvalid) {
      double ss = 0.0;
#pragma unroll
      for (int w = 0; w < PC::WARPS_M; ++w) ss += red[w][t];
      const double mu = a.y_std * (mred[0][t] + mred[1][t]) + a.y_mean;
      double var = a.c - ss;
      if (var < 0.0) {
        var = 0.0;
        if (a.neg_count) atomicAdd(a.neg_count, 1);
      }
      const long o = gq0 + t - a.g_begin;
      a.mean[o] = mu;
      a.std[o] = sqrt(var * (a.y_std * a.y_std));
      if (a.zpred) a.zpred[o] = pow(10.0, mu);
    }
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
  }
}

static int launch_lattice(PostArgs& a, double* tables, int max_ctas, cudaStream_t st) {
  const long total = a.g_end - a.g_begin;
  if (total <= 0) return IPP_OK;
  const long ldt = a.np;
  double* Ex = tables;
  double* Ey = tables + (long)a.nx * ldt;
  const int iy_lo = (int)(a.g_begin / a.nx);
  const int iy_hi = (int)min((long)a.ny - 1, (a.g_end - 1) / a.nx + 1);  // +1: a tile may read the next row
  const int kb = (a.np + 255) / 256;
  axis_table_kernel<<<dim3(kb, a.nx), 256, 0, st>>>(a.xs, a.n, a.np, a.l, a.x0, a.x1, a.xstep, a.nx, 1.0, 0, Ex, ldt);
  axis_table_kernel<<<dim3(kb, iy_hi - iy_lo + 1), 256, 0, st>>>(a.ys, a.n, a.np, a.l, a.y0, a.y1, a.ystep, a.ny, a.c, iy_lo,
                                                                  Ey, ldt);
  const long ntiles = (total + QT - 1) / QT;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IPP_ERR_CUDA;
  int ctas = (int)(ntiles < (long)sms ? ntiles : (long)sms);
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  if (cudaFuncSetAttribute(posterior_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEP_SMEM) != cudaSuccess)
    return IPP_ERR_CUDA;
  posterior_lattice_kernel<<<ctas, LAT_THREADS, SEP_SMEM, st>>>(a, Ex, Ey, ldt);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

// ---------------------------------------------------------------------------------------------
// Explicit points through the table kernel.  A chunk of queries is a one-row "lattice" whose Ex table is the
// cross-covariance itself, T[q][k] = exp(-|q / l - x_k / l|^2 / 2) -- ONE exp() per (query, observation) instead of one
// per 128-row block of W that reads it -- and whose single Ey row is the constant c, so the fragment product
// T * c rounds like the generic kernel's c * exp().  The chunk then runs on the warp-specialised lattice pipeline
// (cp.async B operand, DMMA warps that only multiply) with the tile boxes of its own queries for the slab test.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) point_table_kernel(const double* __restrict__ xs, const double* __restrict__ ys, int n,
                                                          int np, double l, const double* __restrict__ Q, long mq,
                                                          double* __restrict__ T, long ldt) {
  const long i = blockIdx.x;
  const int k = blockIdx.y * 256 + threadIdx.x;
  if (k >= np || i >= mq) return;
  double v = 0.0;
  if (k < n) v = rbf_scaled(__ddiv_rn(Q[2 * i], l), __ddiv_rn(Q[2 * i + 1], l), xs[k], ys[k], 1.0);
  T[i * ldt + k] = v;
}

// row[k] = c (the Ey row of the one-row lattice) and, one warp per 128-query tile, the tile's bounding box.
__global__ void __launch_bounds__(256) point_tiles_kernel(const double* __restrict__ Q, long mq, int np, double c,
                                                          double* __restrict__ row, double* __restrict__ tile_bbox) {
  const long gt = (long)blockIdx.x * 256 + threadIdx.x;
  if (gt < np) row[gt] = c;
  const long tile = gt >> 5;
  const int lane = threadIdx.x & 31;
  const long ntiles = (mq + QT - 1) / QT;
  if (tile >= ntiles) return;
  double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
  for (long q = tile * QT + lane; q < min(mq, (tile + 1) * QT); q += 32) {
    const double x = Q[2 * q], y = Q[2 * q + 1];
    x0 = fmin(x0, x);
    x1 = fmax(x1, x);
    y0 = fmin(y0, y);
    y1 = fmax(y1, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) {
    tile_bbox[4 * tile + 0] = x0;
    tile_bbox[4 * tile + 1] = x1;
    tile_bbox[4 * tile + 2] = y0;
    tile_bbox[4 * tile + 3] = y1;
  }
}

// scratch per chunk of `rows` queries: rows * NP (table) + NP (the c row) + 4 * ceil(rows / 128) (tile boxes)
static int launch_points_tabled(PostArgs a, const double* Q, long m, double* tables, long table_rows, int max_ctas,
                                cudaStream_t st) {
  const long ldt = a.np;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaFuncSetAttribute(posterior_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEP_SMEM) != cudaSuccess)
    return IPP_ERR_CUDA;
  double* T = tables;
  double* crow = tables + table_rows * ldt;
  double* tbox = crow + ldt;
  double* mean = a.mean;
  double* std = a.std;
  double* zpred = a.zpred;
  const int kb = (a.np + 255) / 256;
  for (long q0 = 0; q0 < m; q0 += table_rows) {
    const long mq = min(table_rows, m - q0);
    const long ntiles = (mq + QT - 1) / QT;
    point_table_kernel<<<dim3((unsigned)mq, kb), 256, 0, st>>>(a.xs, a.ys, a.n, a.np, a.l, Q + 2 * q0, mq, T, ldt);
    const long threads = max((long)a.np, 32 * ntiles);
    point_tiles_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(Q + 2 * q0, mq, a.np, a.c, crow, tbox);
    a.Q = nullptr;  // the lattice kernel sees a one-row lattice of mq columns
    a.nx = (int)mq;
    a.ny = 1;
    a.x0 = a.x1 = a.y0 = a.y1 = a.xstep = a.ystep = 0.0;
    a.g_begin = 0;
    a.g_end = mq;
    a.mean = mean + q0;
    a.std = std + q0;
    a.zpred = zpred ? zpred + q0 : nullptr;
    a.tile_bbox = tbox;
    int ctas = (int)(ntiles < (long)sms ? ntiles : (long)sms);
    if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
    posterior_lattice_kernel<<<ctas, LAT_THREADS, SEP_SMEM, st>>>(a, T, crow, ldt);
    IPP_LAUNCH_CHECK();
  }
  return IPP_OK;
}

static int launch_posterior(PostArgs& a, bool grid, int max_ctas, cudaStream_t st) {
  constexpr size_t SM = gemm_smem_bytes<PC, Lay::KCONT, Lay::MCONT>();
  const long total = a.g_end - a.g_begin;
  if (total <= 0) return IPP_OK;
  const long ntiles = (total + QT - 1) / QT;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return IPP_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return IPP_ERR_CUDA;
  int ctas = (int)(ntiles < (long)sms ? ntiles : (long)sms);  // persistent: one CTA per SM
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  if (grid) {
    if (cudaFuncSetAttribute(posterior_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM) != cudaSuccess)
      return IPP_ERR_CUDA;
    posterior_kernel<true><<<ctas, PC::THREADS, SM, st>>>(a);
  } else {
    if (cudaFuncSetAttribute(posterior_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM) != cudaSuccess)
      return IPP_ERR_CUDA;
    posterior_kernel<false><<<ctas, PC::THREADS, SM, st>>>(a);
  }
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

static void fill_state(PostArgs& a, const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                       double y_std) {
  const int np = round_up(n, BLK);
  const long npv = np + BLK;
  a.W = Wbuf;
  a.ldw = np;
  a.xs = vbuf;
  a.ys = vbuf + npv;
  a.alpha = vbuf + 4 * npv;
  a.n = n;
  a.np = np;
  a.c = c;
  a.l = l;
  a.y_mean = y_mean;
  a.y_std = y_std;
}

}  // namespace ipp

using namespace ipp;

extern "C" {

int ipp_gp_posterior_grid(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                          double y_std, double x0, double x1, int nx, double y0, double y1, int ny,
                          long long g_begin, long long g_end, double* mean, double* std, double* zpred,
                          int* neg_count, double* tables, const double* slab_bbox, double cutoff, int max_ctas,
                          void* stream) {
  if (!Wbuf || !vbuf || n <= 0 || nx <= 0 || ny <= 0 || !mean || !std || !(l > 0.0)) return IPP_ERR_ARG;
  if (g_begin < 0 || g_end > (long long)nx * ny || g_begin > g_end) return IPP_ERR_ARG;
  PostArgs a{};
  fill_state(a, Wbuf, vbuf, n, c, l, y_mean, y_std);
  a.Q = nullptr;
  a.x0 = x0;
  a.x1 = x1;
  a.y0 = y0;
  a.y1 = y1;
  a.nx = nx;
  a.ny = ny;
  a.xstep = nx > 1 ? (x1 - x0) / (double)(nx - 1) : 0.0;
  a.ystep = ny > 1 ? (y1 - y0) / (double)(ny - 1) : 0.0;
  a.g_begin = g_begin;
  a.g_end = g_end;
  a.mean = mean;
  a.std = std;
  a.zpred = zpred;
  a.neg_count = neg_count;
  a.slab_bbox = slab_bbox;
  // the box test assumes ascending lattice axes (the reference's linspace(w0, w1, .)); otherwise stay dense
  a.cutoff2 = (slab_bbox && cutoff > 0.0 && x1 >= x0 && y1 >= y0) ? cutoff * cutoff : 0.0;
  if (tables && nx >= QT) return launch_lattice(a, tables, max_ctas, (cudaStream_t)stream);
  return launch_posterior(a, true, max_ctas, (cudaStream_t)stream);
}

int ipp_gp_posterior_points(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                            double y_std, const double* Q, long long m, double* mean, double* std, double* zpred,
                            int* neg_count, int max_ctas, void* stream) {
  if (!Wbuf || !vbuf || n <= 0 || !Q || m < 0 || !mean || !std || !(l > 0.0)) return IPP_ERR_ARG;
  PostArgs a{};
  fill_state(a, Wbuf, vbuf, n, c, l, y_mean, y_std);
  a.Q = Q;
  a.nx = a.ny = 1;
  a.g_begin = 0;
  a.g_end = m;
  a.mean = mean;
  a.std = std;
  a.zpred = zpred;
  a.neg_count = neg_count;
  return launch_posterior(a, false, max_ctas, (cudaStream_t)stream);
}

int ipp_gp_posterior_points_tabled(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                                   double y_std, const double* Q, long long m, double* mean, double* std, double* zpred,
                                   int* neg_count, double* tables, long long table_rows, const double* slab_bbox,
                                   double cutoff, int max_ctas, void* stream) {
  if (!Wbuf || !vbuf || n <= 0 || !Q || m < 0 || !mean || !std || !(l > 0.0) || !tables) return IPP_ERR_ARG;
  // a chunk is a whole number of 128-query tiles (the last chunk may be ragged) and fits the one-row lattice's int axis
  if (table_rows < QT || table_rows % QT != 0 || table_rows > (1LL << 30)) return IPP_ERR_ARG;
  PostArgs a{};
  fill_state(a, Wbuf, vbuf, n, c, l, y_mean, y_std);
  a.neg_count = neg_count;
  a.mean = mean;
  a.std = std;
  a.zpred = zpred;
  a.slab_bbox = slab_bbox;
  a.cutoff2 = (slab_bbox && cutoff > 0.0) ? cutoff * cutoff : 0.0;
  return launch_points_tabled(a, Q, m, tables, table_rows, max_ctas, (cudaStream_t)stream);
}

int ipp_abi_version(void) { return 12; }

}  // extern "C"

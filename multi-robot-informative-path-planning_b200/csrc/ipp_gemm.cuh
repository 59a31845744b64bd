// fp64 tensor-core tile engine shared by every GEMM-shaped kernel on the path
// (Cholesky trailing update, triangular inverse, K^-1 trace contraction, posterior variance).
//
// CTA tile BM x BN, k-slab BK = 16, 3-stage cp.async ring, DMMA m8n8k4.
// Operand tiles live in shared memory in one of two layouts, both bank-conflict free for the
// m8n8k4 fragment pattern (lane -> (g = lane/4, tg = lane%4)):
//   KCONT: tile[r][k] at s[r*(BK+4) + k]     (source matrix row-major, k contiguous)
//   MCONT: tile[r][k] at s[k*(R+4) + r]      (source matrix k-major, r contiguous)
// (BK+4) % 16 == 4 and (R+4) % 16 == 4, so the 16 lanes of a half-warp hit 16 distinct 8-byte banks.
#pragma once
#include "ipp_common.cuh"

namespace ipp {

constexpr int BK = 16;
constexpr int KPAD = BK + 4;
constexpr int STAGES = 3;

enum class Lay { KCONT, MCONT };

template <int BM_, int BN_, int WM_, int WN_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WM_, WARPS_N = WN_;
  static constexpr int THREADS = 32 * WM_ * WN_;
  static constexpr int WTM = BM / WM_, WTN = BN / WN_;
  static constexpr int MI = WTM / 8, NI = WTN / 8;
  static_assert(WTM % 8 == 0 && WTN % 8 == 0, "warp tile must be a multiple of the 8x8 mma");
};
using Cfg128 = TileCfg<128, 128, 2, 4>;  // 256 threads, warp tile 64x32, 64 fp64 accumulators / thread
using Cfg64 = TileCfg<64, 64, 2, 2>;     // 128 threads, warp tile 32x32, for small problems

template <int R, Lay L>
__host__ __device__ constexpr int tile_doubles() {
  return (L == Lay::KCONT) ? R * KPAD : BK * (R + 4);
}

// Asynchronously stage an R x BK operand tile.  Element (r, k) of the tile is
//   KCONT: g[(r0 + r) * ld + k0 + k]      MCONT: g[(k0 + k) * ld + r0 + r]
// Elements with row index >= rlim or k index >= klim are zero-filled.  All of r0, k0, rlim, klim,
// ld must be even and g 16-byte aligned.
template <int R, Lay L, int THREADS>
__device__ __forceinline__ void load_tile_async(double* __restrict__ s, const double* __restrict__ g, long ld, int r0,
                                                int k0, int rlim, int klim, int tid = -1) {
  constexpr int CHUNKS = R * BK / 2;
  static_assert(CHUNKS % THREADS == 0, "tile chunks must divide evenly over the copying threads");
  if (tid < 0) tid = threadIdx.x;  // default: the whole CTA copies; a warp-specialised caller passes its own index
#pragma unroll
  for (int it = 0; it < CHUNKS / THREADS; ++it) {
    int c = tid + it * THREADS;
    if (L == Lay::KCONT) {
      int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      int gr = r0 + r, gk = k0 + kc;
      bool ok = (gr < rlim) && (gk < klim);
      cp_async16(s + r * KPAD + kc, ok ? g + (long)gr * ld + gk : g, ok);
    } else {
      constexpr int CPR = R / 2;
      int k = c / CPR, rc = (c % CPR) * 2;
      int gk = k0 + k, gr = r0 + rc;
      bool ok = (gk < klim) && (gr < rlim);
      cp_async16(s + k * (R + 4) + rc, ok ? g + (long)gk * ld + gr : g, ok);
    }
  }
}

template <class C>
__device__ __forceinline__ void zero_acc(double (&acc)[C::MI][C::NI][2]) {
#pragma unroll
  for (int i = 0; i < C::MI; ++i)
#pragma unroll
    for (int j = 0; j < C::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// One BK slab of warp-level DMMAs out of shared memory.
template <class C, Lay LA, Lay LB>
__device__ __forceinline__ void mma_slab(double (&acc)[C::MI][C::NI][2], const double* __restrict__ sA,
                                         const double* __restrict__ sB, int wm, int wn, int lane) {
  const int g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    double a[C::MI], b[C::NI];
#pragma unroll
    for (int mi = 0; mi < C::MI; ++mi) {
      int r = wm * C::WTM + mi * 8 + g;
      a[mi] = (LA == Lay::KCONT) ? sA[r * KPAD + kk + tg] : sA[(kk + tg) * (C::BM + 4) + r];
    }
#pragma unroll
    for (int ni = 0; ni < C::NI; ++ni) {
      int cidx = wn * C::WTN + ni * 8 + g;
      b[ni] = (LB == Lay::KCONT) ? sB[cidx * KPAD + kk + tg] : sB[(kk + tg) * (C::BN + 4) + cidx];
    }
#pragma unroll
    for (int mi = 0; mi < C::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < C::NI; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
  }
}

template <class C, Lay LA, Lay LB>
__host__ __device__ constexpr int stage_doubles() {
  return tile_doubles<C::BM, LA>() + tile_doubles<C::BN, LB>();
}
template <class C, Lay LA, Lay LB>
__host__ __device__ constexpr size_t gemm_smem_bytes() {
  return sizeof(double) * STAGES * stage_doubles<C, LA, LB>();
}

// acc += sum over k in [kbeg, kend) of A(:,k) B(k,:) ; kbeg must be a multiple of 2 (it is always a
// multiple of BK here).  FillA(sA, k0) / FillB(sB, k0) stage one slab each (cp.async or direct
// stores); they are called one slab-pair per iteration, STAGES-1 slabs ahead of the math.
template <class C, Lay LA, Lay LB, class FillA, class FillB>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[C::MI][C::NI][2], double* smem, int kbeg, int kend,
                                              FillA fillA, FillB fillB) {
  constexpr int SD = stage_doubles<C, LA, LB>();
  constexpr int AOFF = tile_doubles<C::BM, LA>();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int nk = (kend - kbeg + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      fillA(smem + s * SD, kbeg + s * BK);
      fillB(smem + s * SD + AOFF, kbeg + s * BK);
    }
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nx = it + STAGES - 1;
    if (nx < nk) {
      double* st = smem + (nx % STAGES) * SD;
      fillA(st, kbeg + nx * BK);
      fillB(st + AOFF, kbeg + nx * BK);
    }
    cp_async_commit();
    const double* cur = smem + (it % STAGES) * SD;
    mma_slab<C, LA, LB>(acc, cur, cur + AOFF, wm, wn, lane);
  }
  cp_async_wait<0>();
  __syncthreads();
}

// The same main loop without a CTA-wide barrier per k-slab: every stage has a "full" mbarrier (one arrival per thread,
// fired by the hardware when that thread's cp.async copies of the stage have landed:
// cp.async.mbarrier.arrive.noinc) and an "empty" mbarrier (one arrival per warp when it has read the stage).  A
// warp waits only for the data it is about to multiply and for the slot it is about to refill -- which the slowest
// warp finished PREF slabs ago -- so the eight warps drift by up to two slabs and cover each other's shared-memory
// latencies, instead of meeting at a barrier sixteen times per rank-256 tile.  (The posterior kernel's pipeline, with
// every warp both copying and multiplying.)  `bars` = 2 * MB_STAGES mbarriers in shared memory, (re)initialised
// here; needs MB_STAGES * stage_doubles of shared memory.  A wait that exceeds ~2^28 polls traps: a pipeline bug
// must end the kernel, not hang the device.
constexpr int MB_STAGES = 4;
constexpr int MB_PREF = 2;
template <class C, Lay LA, Lay LB>
__host__ __device__ constexpr size_t gemm_mb_smem_bytes() {
  return sizeof(double) * MB_STAGES * stage_doubles<C, LA, LB>();
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  for (unsigned spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 28)) __trap();
  }
}
template <class C, Lay LA, Lay LB, class FillA, class FillB>
__device__ __forceinline__ void gemm_mainloop_mb(double (&acc)[C::MI][C::NI][2], double* smem, uint64_t* bars, int kbeg,
                                                 int kend, FillA fillA, FillB fillB) {
  constexpr int SD = stage_doubles<C, LA, LB>();
  constexpr int AOFF = tile_doubles<C::BM, LA>();
  uint64_t* full = bars;
  uint64_t* empty = bars + MB_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MB_STAGES; ++s) {
      mbar_init(&full[s], C::THREADS);
      mbar_init(&empty[s], C::THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto produce = [&](int slab) {
    double* st = smem + (slab % MB_STAGES) * SD;
    fillA(st, kbeg + slab * BK);
    fillB(st + AOFF, kbeg + slab * BK);
    cp_async_mbar_arrive(&full[slab % MB_STAGES]);
  };
#pragma unroll
  for (int s = 0; s < MB_PREF; ++s)
    if (s < nk) produce(s);
  for (int it = 0; it < nk; ++it) {
    const int nx = it + MB_PREF;
    if (nx < nk) {
      if (nx >= MB_STAGES) mbar_wait_bounded(&empty[nx % MB_STAGES], ((nx / MB_STAGES) - 1) & 1);  // every warp has read the old contents
      produce(nx);
    }
    mbar_wait_bounded(&full[it % MB_STAGES], (it / MB_STAGES) & 1);
    const double* cur = smem + (it % MB_STAGES) * SD;
    mma_slab<C, LA, LB>(acc, cur, cur + AOFF, wm, wn, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[it % MB_STAGES]);
  }
  __syncthreads();  // the caller may reuse the stages
}

// Warp-specialised form of the same pipeline, for kernels launched with C::THREADS + 128 threads (C::THREADS = 256):
// warps 0..7 only multiply -- wait on "full", DMMA out of shared memory, release on "empty" -- and a copy warpgroup
// (warps 8..11, registers handed over with setmaxnreg) issues every cp.async, running up to MB_STAGES slabs ahead.
// With the copies off their instruction stream the DMMA warps hold the fp64 pipe at its issue rate (measured in the
// dataflow Cholesky: 2.04 us per 128 x 128 x 16 slab = the 37.2 TFLOP/s peak; 2.5 us with every warp doing both).
// A kernel uses it as
//     if (gemm_ws_split<C>(bars)) { body<copy role>(); return; }  body<DMMA role>();
// with NOTHING but the split before it: ptxas gives the DMMA warps their 224 registers only if every path from the
// kernel entry runs through the setmaxnreg of its role (an early `return` before the split made it keep 168 and
// spill half of the accumulators inside the main loop).  The DMMA role's later CTA-wide synchronisation must be
// Group256Sync (named barrier 1), not __syncthreads().
template <class C>
__device__ __forceinline__ bool gemm_ws_split(uint64_t* bars) {
  static_assert(C::THREADS == 256, "eight DMMA warps + one copy warpgroup");
  if (threadIdx.x == 0) {
    for (int s = 0; s < MB_STAGES; ++s) {
      mbar_init(&bars[s], 128);                              // full: one cp.async-completion arrival per copy thread
      mbar_init(&bars[MB_STAGES + s], C::THREADS / 32);      // empty: one arrival per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x >= C::THREADS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    return true;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
  return false;
}
// Copy role: FillA(sA, k0, tid) / FillB(sB, k0, tid) stage one slab with 128 threads (tid = 0..127).
template <class C, Lay LA, Lay LB, class FillA, class FillB>
__device__ __forceinline__ void gemm_ws_produce(double* smem, uint64_t* bars, int kbeg, int kend, FillA fillA, FillB fillB) {
  constexpr int SD = stage_doubles<C, LA, LB>();
  constexpr int AOFF = tile_doubles<C::BM, LA>();
  uint64_t* full = bars;
  uint64_t* empty = bars + MB_STAGES;
  const int pt = threadIdx.x - C::THREADS;
  const int nk = (kend - kbeg + BK - 1) / BK;
  for (int slab = 0; slab < nk; ++slab) {
    const int stage = slab % MB_STAGES;
    if (slab >= MB_STAGES) mbar_wait_bounded(&empty[stage], ((slab / MB_STAGES) - 1) & 1);
    double* st = smem + stage * SD;
    fillA(st, kbeg + slab * BK, pt);
    fillB(st + AOFF, kbeg + slab * BK, pt);
    cp_async_mbar_arrive(&full[stage]);
  }
  cp_async_wait<0>();
}
// DMMA role: acc += A(:, [kbeg, kend)) B([kbeg, kend), :)
template <class C, Lay LA, Lay LB>
__device__ __forceinline__ void gemm_ws_consume(double (&acc)[C::MI][C::NI][2], const double* smem, uint64_t* bars, int kbeg,
                                                int kend) {
  constexpr int SD = stage_doubles<C, LA, LB>();
  constexpr int AOFF = tile_doubles<C::BM, LA>();
  uint64_t* full = bars;
  uint64_t* empty = bars + MB_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int nk = (kend - kbeg + BK - 1) / BK;
  for (int it = 0; it < nk; ++it) {
    const int stage = it % MB_STAGES;
    mbar_wait_bounded(&full[stage], (it / MB_STAGES) & 1);
    const double* cur = smem + stage * SD;
    mma_slab<C, LA, LB>(acc, cur, cur + AOFF, wm, wn, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
}

// Visit every accumulator element with its tile-local (row, col).
template <class C, class F>
__device__ __forceinline__ void for_each_acc(double (&acc)[C::MI][C::NI][2], F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int mi = 0; mi < C::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < C::NI; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) f(wm * C::WTM + mi * 8 + g, wn * C::WTN + ni * 8 + tg * 2 + e, acc[mi][ni][e]);
}

}  // namespace ipp

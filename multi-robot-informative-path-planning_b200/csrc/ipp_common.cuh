// Shared device helpers for the B200 (sm_100a) informative-path-planning kernels.
// fp64 everywhere: the measurement Gram matrix has cond ~1e11, so
// the tensor path is the fp64 DMMA (mma.sync m8n8k4.f64), not tcgen05.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define IPP_OK 0
#define IPP_ERR_ARG (-1)
#define IPP_ERR_CUDA (-2)

#define IPP_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                 \
    if (e__ != cudaSuccess) return IPP_ERR_CUDA;             \
  } while (0)

namespace ipp {

constexpr int BLK = 64;  // padding / panel granularity of every factor-side matrix

__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }

// D(8x8) += A(8x4,row) * B(4x8,col); lane T holds A[T/4][T%4], B[T%4][T/4], D[T/4][2*(T%4)+{0,1}]
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// 16-byte async global->shared copy; when !ok the destination is zero-filled (src-size 0).
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool ok) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  int bytes = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier helpers (shared::cta).  Used by the lattice posterior pipeline, where each stage has a
// "full" barrier (completed by every thread's cp.async group through cp.async.mbarrier.arrive.noinc)
// and an "empty" barrier (one arrival per consumer warp), instead of a CTA-wide __syncthreads per slab.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// Arrive on `bar` once all cp.async issued so far by this thread have landed (does not add to the
// expected count: the barrier is initialised with one expected arrival per producing thread).
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
// Producer side of a bulk-copy (TMA, non-tensor) stage: announce the bytes that will land on `bar` ...
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
// ... and copy `bytes` (multiple of 16, 16-byte aligned both sides) global -> shared through the async proxy,
// completing on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE_%=;\n"
      "bra MBAR_WAIT_%=;\n"
      "MBAR_DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// ---- barrier functors: block routines written for "all threads of the CTA" take one of these, so that the same
// code runs on the 256 DMMA threads of a warp-specialised kernel (named barrier 1) while its copy / scheduler warps
// keep going.
struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct Group256Sync {  // threads 0..255 of the CTA
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
};
// A non-blocking test of an mbarrier phase (observer or waiter).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// ---- flags in global memory (dataflow Cholesky, backward-substitution chain): relaxed polling, acquire / release
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// RBF on pre-scaled coordinates, arithmetic ordered like scipy cdist('sqeuclidean') followed by
// np.exp(-0.5 * d) * c (sklearn kernels.py:1569-1570, 964): separate mul/add roundings, no FMA.
__device__ __forceinline__ double sqdist_nofma(double ax, double ay, double bx, double by) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}
__device__ __forceinline__ double rbf_scaled(double ax, double ay, double bx, double by, double c) {
  return __dmul_rn(exp(__dmul_rn(-0.5, sqdist_nofma(ax, ay, bx, by))), c);
}
// Stationary kernels of the public Gram builder on pre-scaled coordinates.  kind 0: RBF (above); 1 / 2: Matern with
// nu = 1.5 / 2.5 as sklearn evaluates them (kernels.py:1723-1735: r = cdist(X / l, Y / l), K = r * sqrt(3),
// (1 + K) exp(-K);  K = r * sqrt(5), (1 + K + K^2 / 3) exp(-K)), times c.
__device__ __forceinline__ double stationary_scaled(int kind, double ax, double ay, double bx, double by, double c) {
  if (kind == 0) return rbf_scaled(ax, ay, bx, by, c);
  const double r = sqrt(sqdist_nofma(ax, ay, bx, by));
  if (kind == 1) {
    const double k = __dmul_rn(r, 1.7320508075688772);
    return __dmul_rn(__dmul_rn(__dadd_rn(1.0, k), exp(-k)), c);
  }
  const double k = __dmul_rn(r, 2.23606797749979);
  return __dmul_rn(__dmul_rn(__dadd_rn(__dadd_rn(1.0, k), __ddiv_rn(__dmul_rn(k, k), 3.0)), exp(-k)), c);
}

}  // namespace ipp

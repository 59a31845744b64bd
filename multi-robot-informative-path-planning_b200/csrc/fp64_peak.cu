// Standalone micro-benchmark: measured fp64 peaks of this B200 -- the roofline denominator for the
// Cholesky / posterior-variance kernels (MEASURED_PEAKS.json only carries HBM and bf16 numbers).
// Prints one JSON line.  Register-resident operands, no memory traffic.
//   dfma        : independent DFMA chains on the FP64 pipe
//   dmma884     : mma.sync.m8n8k4.f64     (sm_80+)
//   dmma16816   : mma.sync.m16n8k16.f64   (sm_90+), also m16n8k4 / m16n8k8
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, double a, double b) {
  double d[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i][0] = d[i][1] = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(d[i][0]), "+d"(d[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1684(double* out, double a, double b) {
  double d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
                   : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, double a, double b) {
  double d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
          : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
          : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, double a, double b) {
  double d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
          "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
          : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
          : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fp64 exp throughput (the K* generation cost)
__global__ void k_dexp(double* out, double a) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -a * (i + 1) - 1e-3 * threadIdx.x;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -exp(x[i]) - a;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int threads = 256, blocks = sms * 8;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
  const double nthreads = (double)threads * blocks, nwarps = nthreads / 32;
  double t;
  printf("{\"sms\": %d", sms);
  t = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1e-9, 0.999); }, 10);
  printf(", \"dfma_tflops\": %.2f", 2.0 * 16 * ITERS * nthreads / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, 1e-9, 0.999); }, 10);
  printf(", \"dmma884_tflops\": %.2f", 2.0 * 256 * 16 * ITERS * nwarps / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma1684<<<blocks, threads>>>(out, 1e-9, 0.999); }, 10);
  printf(", \"dmma1684_tflops\": %.2f", 2.0 * 512 * 8 * ITERS * nwarps / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, 1e-9, 0.999); }, 10);
  printf(", \"dmma1688_tflops\": %.2f", 2.0 * 1024 * 8 * ITERS * nwarps / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, 1e-9, 0.999); }, 10);
  printf(", \"dmma16816_tflops\": %.2f", 2.0 * 2048 * 8 * ITERS * nwarps / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_dexp<<<blocks, threads>>>(out, 0.5); }, 10);
  printf(", \"dexp_gops\": %.2f", 8.0 * (ITERS / 8) * nthreads / (t * 1e-3) / 1e9);
  printf("}\n");
  CK(cudaFree(out));
  return 0;
}

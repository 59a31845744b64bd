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

// The posterior kernel's occupancy: ONE 256-thread CTA per SM (2 warps per scheduler), 32 independent
// accumulator tiles per warp (warp tile 64 x 32), operands re-read every k-step.
__global__ void __launch_bounds__(256, 1) k_dmma884_tile(double* out, double a, double b) {
  double d[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j][0] = d[i][j][1] = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
    double af[8], bf[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) af[i] = a + it * 1e-9 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = b + it * 1e-9 + j;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(d[i][j][0]), "+d"(d[i][j][1])
                     : "d"(af[i]), "d"(bf[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j][0] + d[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same tile, operands read from shared memory with the kernel's fragment pattern (12 LDS per 32 DMMA).
__global__ void __launch_bounds__(256, 1) k_dmma884_tile_lds(double* out, double a, double b) {
  __shared__ double sA[128 * 20], sB[128 * 20];
  for (int e = threadIdx.x; e < 128 * 20; e += 256) {
    sA[e] = a + e * 1e-9;
    sB[e] = b + e * 1e-9;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3, g = lane >> 2, tg = lane & 3;
  double d[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j][0] = d[i][j][1] = threadIdx.x;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = sA[(wm * 64 + i * 8 + g) * 20 + kk + tg];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = sB[(wn * 32 + j * 8 + g) * 20 + kk + tg];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(d[i][j][0]), "+d"(d[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j][0] + d[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same 64 x 32 warp tile from m16n8k16 instructions (16 per k16-step), operands from shared memory.
__global__ void __launch_bounds__(256, 1) k_dmma16816_tile_lds(double* out, double a, double b) {
  __shared__ double sA[128 * 20], sB[128 * 20];
  for (int e = threadIdx.x; e < 128 * 20; e += 256) {
    sA[e] = a + e * 1e-9;
    sB[e] = b + e * 1e-9;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3, g = lane >> 2, tg = lane & 3;
  double d[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) d[i][j][e] = threadIdx.x;
  for (int it = 0; it < ITERS / 4; ++it) {
    double af[4][8], bf[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) af[i][e] = sA[(wm * 64 + i * 16 + g + 8 * (e & 1)) * 20 + tg + 4 * (e >> 1)];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) bf[j][e] = sB[(wn * 32 + j * 8 + g) * 20 + tg + 4 * e];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
            "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
            : "+d"(d[i][j][0]), "+d"(d[i][j][1]), "+d"(d[i][j][2]), "+d"(d[i][j][3])
            : "d"(af[i][0]), "d"(af[i][1]), "d"(af[i][2]), "d"(af[i][3]), "d"(af[i][4]), "d"(af[i][5]), "d"(af[i][6]),
              "d"(af[i][7]), "d"(bf[j][0]), "d"(bf[j][1]), "d"(bf[j][2]), "d"(bf[j][3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j][0] + d[i][j][1] + d[i][j][2] + d[i][j][3];
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
  {
    const double w1 = (double)sms * 8;  // warps with one 256-thread CTA per SM
    t = time_ms([&] { k_dmma884_tile<<<sms, 256>>>(out, 1e-9, 0.999); }, 10);
    printf(", \"dmma884_1cta_tile_tflops\": %.2f", 2.0 * 256 * 32 * ITERS * w1 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_dmma884_tile_lds<<<sms, 256>>>(out, 1e-9, 0.999); }, 10);
    printf(", \"dmma884_1cta_tile_lds_tflops\": %.2f", 2.0 * 256 * 32 * ITERS * w1 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_dmma16816_tile_lds<<<sms, 256>>>(out, 1e-9, 0.999); }, 10);
    printf(", \"dmma16816_1cta_tile_lds_tflops\": %.2f", 2.0 * 2048 * 16 * (ITERS / 4) * w1 / (t * 1e-3) / 1e12);
  }
  t = time_ms([&] { k_dexp<<<blocks, threads>>>(out, 0.5); }, 10);
  printf(", \"dexp_gops\": %.2f", 8.0 * (ITERS / 8) * nthreads / (t * 1e-3) / 1e9);
  printf("}\n");
  CK(cudaFree(out));
  return 0;
}

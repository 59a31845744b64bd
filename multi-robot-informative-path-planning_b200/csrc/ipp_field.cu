// Callers on either side of the GP hot path, on the same workspace lattice:
//   * the analytic ground-truth field with rectangle obstacles -- PointSourceField.ground_truth and
//     compute_path_length_in_obstacle (src/point_source/point_source.py:163-260), whose obstacle branch is a Python
//     loop over every lattice point in the reference;
//   * the three whole-grid reductions every run ends with -- RMSE and source-weighted RMSE of log10(Z + 1)
//     (main.py:113-114) and the differential entropy of std (src/visualization/plot_helper.py:383-398) -- fused into
//     one pass over the device-resident posterior, so that a metrics-only caller never reads 16 bytes per point back.
// Neither feeds a planning decision (the measurements come from PointSourceField.intensity at the waypoints, which
// stays on the host and bit-exact); parity is to rounding, tested against the host expressions.
#include "ipp_common.cuh"
#include "../../include/ipp_b200.h"

namespace ipp {

constexpr double AIR_ABSORPTION = 0.001205;   // point_source.py:130,169
constexpr double CONCRETE_ABSORPTION = 2.35;  // point_source.py:131,170

// point_source.py:279-297
__device__ __forceinline__ bool segments_intersect(double x1, double y1, double x2, double y2, double x3, double y3,
                                                   double x4, double y4) {
  const double d1x = x2 - x1, d1y = y2 - y1, d2x = x4 - x3, d2y = y4 - y3;
  const double den = __dsub_rn(__dmul_rn(d1x, d2y), __dmul_rn(d1y, d2x));
  if (den == 0.0) return false;
  const double t = __ddiv_rn(__dsub_rn(__dmul_rn(x3 - x1, d2y), __dmul_rn(y3 - y1, d2x)), den);
  const double u = __ddiv_rn(__dsub_rn(__dmul_rn(x3 - x1, d1y), __dmul_rn(y3 - y1, d1x)), den);
  return 0.0 <= t && t <= 1.0 && 0.0 <= u && u <= 1.0;
}

// Chord of the source -> point line inside the rectangle, or a negative value when the segment misses it
// (point_source.py:214-260: line_intersects_rectangle, then the first two boundary crossings of the LINE, x sides
// before y sides).  Fewer than two crossings makes the reference raise; here the chord is NaN.
__device__ __forceinline__ double chord_in_rectangle(double xn, double yn, double xr, double yr, double x_min,
                                                     double x_max, double y_min, double y_max) {
  const bool inside = (x_min <= xn && xn <= x_max && y_min <= yn && yn <= y_max) ||
                      (x_min <= xr && xr <= x_max && y_min <= yr && yr <= y_max);
  if (!inside && !segments_intersect(xn, yn, xr, yr, x_min, y_min, x_min, y_max) &&
      !segments_intersect(xn, yn, xr, yr, x_max, y_min, x_max, y_max) &&
      !segments_intersect(xn, yn, xr, yr, x_min, y_min, x_max, y_min) &&
      !segments_intersect(xn, yn, xr, yr, x_min, y_max, x_max, y_max))
    return -1.0;
  double hx[2], hy[2];
  int nh = 0;
  if (xn != xr) {
    const double xb[2] = {x_min, x_max};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double yy = __dadd_rn(yn, __dmul_rn(__ddiv_rn(xb[s] - xn, xr - xn), yr - yn));
      if (y_min <= yy && yy <= y_max && nh < 2) {
        hx[nh] = xb[s];
        hy[nh] = yy;
        ++nh;
      }
    }
  }
  if (yn != yr) {
    const double yb[2] = {y_min, y_max};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double xx = __dadd_rn(xn, __dmul_rn(__ddiv_rn(yb[s] - yn, yr - yn), xr - xn));
      if (x_min <= xx && xx <= x_max && nh < 2) {
        hx[nh] = xx;
        hy[nh] = yb[s];
        ++nh;
      }
    }
  }
  if (nh < 2) return nan("");
  const double dx = hx[1] - hx[0], dy = hy[1] - hy[0];
  return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// One thread per lattice point (index = iy * nx + ix, coordinates as np.linspace), sources and obstacles in
// shared memory.  Compute bound (per point S * (exp, asin, cos, hypot) + S * n_obst segment tests), 8 B written.
__global__ void __launch_bounds__(256) ground_truth_kernel(const double* __restrict__ sources, int S,
                                                           const double* __restrict__ obst, int n_obst, double x0,
                                                           double x1, double xstep, int nx, double y0, double y1,
                                                           double ystep, int ny, double r_s, double r_d,
                                                           double* __restrict__ field) {
  extern __shared__ double sh[];
  double* ssrc = sh;           // S x 3
  double* sob = sh + 3 * S;    // n_obst x 4 as {x_min, x_max, y_min, y_max}
  for (int e = threadIdx.x; e < 3 * S; e += blockDim.x) ssrc[e] = sources[e];
  for (int e = threadIdx.x; e < n_obst; e += blockDim.x) {
    sob[4 * e] = obst[4 * e];
    sob[4 * e + 1] = obst[4 * e] + obst[4 * e + 2];
    sob[4 * e + 2] = obst[4 * e + 1];
    sob[4 * e + 3] = obst[4 * e + 1] + obst[4 * e + 3];
  }
  __syncthreads();
  const long gi = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= (long)nx * ny) return;
  const int ix = (int)(gi % nx), iy = (int)(gi / nx);
  const double x = (ix == nx - 1 && nx > 1) ? x1 : __dadd_rn(__dmul_rn((double)ix, xstep), x0);
  const double y = (iy == ny - 1 && ny > 1) ? y1 : __dadd_rn(__dmul_rn((double)iy, ystep), y0);
  const double pi4 = 4.0 * 3.141592653589793;
  double total = 0.0;
  for (int s = 0; s < S; ++s) {
    const double xn = ssrc[3 * s], yn = ssrc[3 * s + 1], A = ssrc[3 * s + 2];
    const double dist = hypot(x - xn, y - yn);  // np.abs of a complex number
    const double theta = asin(fmin(r_d / dist, 1.0));
    double absorb = exp(-AIR_ABSORPTION * dist);
    for (int o = 0; o < n_obst; ++o) {
      const double len = chord_in_rectangle(xn, yn, x, y, sob[4 * o], sob[4 * o + 1], sob[4 * o + 2], sob[4 * o + 3]);
      if (!(len < 0.0)) absorb *= exp(-CONCRETE_ABSORPTION * len);
    }
    const double inten = dist <= r_s ? A / (pi4 * (r_s * r_s)) : A / (pi4 * (dist * dist)) * absorb;
    const double resp = dist <= r_d ? 0.5 * A : 0.5 * A * (1.0 - cos(theta)) * absorb;
    total += inten + resp;
  }
  field[gi] = total;
}

// Fused metrics: partial[b] = {sum err^2, sum z_true err^2, sum z_true, sum log(std sqrt(2 pi e))} of block b's
// grid-stride share, err = log10(z_true + 1) - log10(z_pred + 1); std == 0 counts as 1e-6 (plot_helper.py:393-394).
// Fixed grid and fixed-order tree reductions: the result does not depend on timing.
constexpr int METRIC_BLOCKS = 592;  // 4 x 148
__global__ void __launch_bounds__(256) grid_metrics_kernel(const double* __restrict__ zpred, const double* __restrict__ std,
                                                           const double* __restrict__ ztrue, long G,
                                                           double* __restrict__ partial) {
  __shared__ double red[4][256];
  const int t = threadIdx.x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const double k = sqrt(2.0 * (3.141592653589793 * 2.718281828459045));
  for (long i = (long)blockIdx.x * 256 + t; i < G; i += (long)gridDim.x * 256) {
    const double zt = ztrue[i];
    const double err = log10(zt + 1.0) - log10(zpred[i] + 1.0);
    const double e2 = err * err;
    double sd = std[i];
    if (sd == 0.0) sd = 1e-6;
    a0 += e2;
    a1 += zt * e2;
    a2 += zt;
    a3 += log(sd * k);
  }
  red[0][t] = a0;
  red[1][t] = a1;
  red[2][t] = a2;
  red[3][t] = a3;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o)
      for (int q = 0; q < 4; ++q) red[q][t] += red[q][t + o];
    __syncthreads();
  }
  if (t < 4) partial[4 * blockIdx.x + t] = red[t][0];
}
__global__ void __launch_bounds__(256) grid_metrics_final_kernel(const double* __restrict__ partial, int nblocks, long G,
                                                                 double* __restrict__ out) {
  __shared__ double red[4][256];
  const int t = threadIdx.x;
  for (int q = 0; q < 4; ++q) {
    double s = 0.0;
    for (int b = t; b < nblocks; b += 256) s += partial[4 * b + q];
    red[q][t] = s;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o)
      for (int q = 0; q < 4; ++q) red[q][t] += red[q][t + o];
    __syncthreads();
  }
  if (t == 0) {
    out[0] = sqrt(red[0][0] / (double)G);               // RMSE            main.py:113
    out[1] = sqrt(red[1][0] / red[2][0]);                // weighted RMSE   main.py:114
    out[2] = red[3][0] / 0.6931471805599453;             // entropy in bits plot_helper.py:395-397
    out[3] = red[2][0];
  }
}

}  // namespace ipp

using namespace ipp;

extern "C" {

int ipp_ground_truth_grid(const double* sources, int S, const double* obstacles, int n_obst, double x0, double x1,
                          int nx, double y0, double y1, int ny, double r_s, double r_d, double* field, void* stream) {
  if (S < 0 || n_obst < 0 || nx <= 0 || ny <= 0 || !field || (S > 0 && !sources) || (n_obst > 0 && !obstacles))
    return IPP_ERR_ARG;
  const size_t smem = sizeof(double) * (3 * (size_t)S + 4 * (size_t)n_obst);
  if (smem > 48 * 1024) return IPP_ERR_ARG;
  const long G = (long)nx * ny;
  const double xstep = nx > 1 ? (x1 - x0) / (double)(nx - 1) : 0.0;
  const double ystep = ny > 1 ? (y1 - y0) / (double)(ny - 1) : 0.0;
  ground_truth_kernel<<<(unsigned)((G + 255) / 256), 256, smem, (cudaStream_t)stream>>>(
      sources, S, obstacles, n_obst, x0, x1, xstep, nx, y0, y1, ystep, ny, r_s, r_d, field);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

int ipp_grid_metrics(const double* zpred, const double* std, const double* ztrue, long long G, double* partial,
                     double* out, void* stream) {
  if (!zpred || !std || !ztrue || G <= 0 || !partial || !out) return IPP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  grid_metrics_kernel<<<METRIC_BLOCKS, 256, 0, st>>>(zpred, std, ztrue, (long)G, partial);
  grid_metrics_final_kernel<<<1, 256, 0, st>>>(partial, METRIC_BLOCKS, (long)G, out);
  IPP_LAUNCH_CHECK();
  return IPP_OK;
}

}  // extern "C"

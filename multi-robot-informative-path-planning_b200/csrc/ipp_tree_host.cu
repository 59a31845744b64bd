// the RRT primitives (nearest / steer / near / cost / rewire, src/rrt/rrt_utils.py:59-105) and
// the three tree generators built from them (src/rrt/rrt.py:263-327).  HOST code: the growth of one tree is a chain
// of a few hundred dependent insertions (sample -> nearest -> steer -> near -> add -> rewire), each over a few
// hundred nodes, and the trees of the agents follow each other in one RNG stream -- there is nothing for a GPU to
// batch, and in Python every insertion costs ~50 us of interpreter and small-array overhead (4.4 s of an 8 s
// 8-agent run once everything dense is on the device).  This is the same loop natively, on the structure-of-arrays
// tree the device scorer reads (rrt_utils.TreeArrays): ~0.3 us per insertion.
//
// Decisions and coordinates are the reference's bit for bit.  Every comparison of the reference is on
// np.linalg.norm of a 2-vector = sqrt(x.dot(x)), whose dot comes from the BLAS numpy links (two roundings, or one
// fused multiply-add, depending on the ddot kernel); `norm_variant` selects the matching form and the Python side
// calibrates it against np.linalg.norm itself before the first use (rrt_utils.native_norm_variant).  Everything
// else is IEEE element-wise arithmetic in the reference's order; the file is compiled with -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include "../../include/ipp_b200.h"

#define IPP_FMA_FN __attribute__((target("fma")))

namespace {

// x.dot(x) of a 2-vector in the three rounding forms a BLAS ddot can produce
template <int V>
IPP_FMA_FN inline double dot2(double dx, double dy) {
  if (V == 1) return __builtin_fma(dy, dy, dx * dx);   // dot = 0 + x0*x0, then fused + x1*x1
  if (V == 2) return __builtin_fma(dx, dx, dy * dy);
  return dx * dx + dy * dy;                            // two products, one addition (-ffp-contract=off)
}
template <int V>
IPP_FMA_FN inline double norm2(double dx, double dy) {
  return sqrt(dot2<V>(dx, dy));
}

template <int V>
struct Tree {
  double* xy;
  long long *parent0, *parent, *nchild;
  double* edge;
  int n;
  // rrt_utils.py:71-76: leaf-to-root accumulation, starting from the integer 0
  inline double path_cost(long long i) const {
    double total = 0.0;
    while (parent[i] >= 0) {
      total += edge[i];
      i = parent[i];
    }
    return total;
  }
  IPP_FMA_FN inline double dist(long long i, double x, double y) const { return norm2<V>(xy[2 * i] - x, xy[2 * i + 1] - y); }
  IPP_FMA_FN inline int add(double x, double y, long long p) {
    const int i = n++;
    xy[2 * i] = x;
    xy[2 * i + 1] = y;
    parent0[i] = parent[i] = p;
    nchild[i] = 0;
    edge[i] = norm2<V>(x - xy[2 * p], y - xy[2 * p + 1]);
    nchild[p] += 1;
    return i;
  }
};

template <int V>
IPP_FMA_FN int grow(int mode, double* xy, long long* parent0, long long* parent, double* edge, long long* nchild, int n,
                    int capacity, const double* uniforms, int K, const double* ws4, double step, double radius,
                    double budget_portion, double* travelled_io, long long* rewires, int rewire_cap, int* n_rewires_io,
                    int* scratch, int* consumed, int* n_out, int* finished) {
  Tree<V> t{xy, parent0, parent, nchild, edge, n};
  const double w = ws4[1] - ws4[0], h = ws4[3] - ws4[2];
  // sqrt is monotone: a squared length outside [r2lo, r2hi] decides `norm < radius` without the square root
  const double r2 = radius * radius, r2lo = r2 * (1.0 - 1e-14), r2hi = r2 * (1.0 + 1e-14);
  double travelled = *travelled_io;
  int nrw = *n_rewires_io, k = 0, done = travelled < budget_portion ? 0 : 1;
  while (!done && k < K && t.n < capacity && nrw + t.n <= rewire_cap) {
    // np.random.rand(2) * (w, h) + (x0, y0)
    const double tx = uniforms[2 * k] * w + ws4[0], ty = uniforms[2 * k + 1] * h + ws4[2];
    ++k;
    // nearest: min(..., key=norm) keeps the first minimum (rrt_utils.py:100-101).  A node can only beat the
    // running minimum of sqrt(dot) if its dot is smaller, so the square root is taken for those alone.
    long long nn = 0;
    double best_s = dot2<V>(xy[0] - tx, xy[1] - ty), best = sqrt(best_s);
    for (int i = 1; i < t.n; ++i) {
      const double s = dot2<V>(xy[2 * i] - tx, xy[2 * i + 1] - ty);
      if (s < best_s) {
        const double d = sqrt(s);
        if (d < best) {
          best = d;
          best_s = s;
          nn = i;
        }
      }
    }
    // steer (rrt_utils.py:93-98); rrt* uses the literal step 1.0 (rrt.py:285), passed in by the caller
    double dx = tx - xy[2 * nn], dy = ty - xy[2 * nn + 1];
    double distance = norm2<V>(dx, dy);
    if (distance > 0) {
      dx = dx / distance;
      dy = dy / distance;
    }
    distance = distance < step ? distance : step;   // Python min(distance, d_max_step): the first of equals
    const double px = xy[2 * nn] + dx * distance, py = xy[2 * nn + 1] + dy * distance;
    int nnb = 0;
    if (mode != 0)   // near: ascending ids with norm < radius (rrt_utils.py:90-91), before the node is added
      for (int i = 0; i < t.n; ++i) {
        const double s = dot2<V>(xy[2 * i] - px, xy[2 * i + 1] - py);
        if (s < r2lo || (s <= r2hi && sqrt(s) < radius)) scratch[nnb++] = i;
      }
    long long par = nn;
    if (mode == 1) {   // rrt*: cheapest parent among the near nodes (rrt.py:291-297)
      double c_min = t.path_cost(nn) + t.dist(nn, px, py);
      for (int q = 0; q < nnb; ++q) {
        const double c = t.path_cost(scratch[q]) + t.dist(scratch[q], px, py);
        if (c < c_min) {
          c_min = c;
          par = scratch[q];
        }
      }
    }
    const int nw = t.add(px, py, par);   // rig: add_node overrides choose_parent -- the parent is the nearest node
    travelled += edge[nw];
    if (mode != 0)   // rewire (rrt_utils.py:85-88)
      for (int q = 0; q < nnb; ++q) {
        const long long c = scratch[q];
        if (t.path_cost(nw) + t.dist(c, px, py) < t.path_cost(c)) {
          parent[c] = nw;
          edge[c] = t.dist(c, px, py);
          rewires[3 * nrw] = nw;
          rewires[3 * nrw + 1] = c;
          rewires[3 * nrw + 2] = nw;
          ++nrw;
        }
      }
    if (travelled >= budget_portion) done = 1;
  }
  *travelled_io = travelled;
  *n_rewires_io = nrw;
  *consumed = k;
  *n_out = t.n;
  *finished = done;
  return 0;
}

template <int V>
IPP_FMA_FN void norms(const double* xy, int m, double* out) {
  for (int i = 0; i < m; ++i) out[i] = norm2<V>(xy[2 * i], xy[2 * i + 1]);
}

}  // namespace

extern "C" {

int ipp_host_norm2(int variant, const double* xy, int m, double* out) {
  if (variant < 0 || variant > 2 || m < 0 || (m > 0 && (!xy || !out))) return -1;
  if (variant == 0) norms<0>(xy, m, out);
  else if (variant == 1) norms<1>(xy, m, out);
  else norms<2>(xy, m, out);
  return 0;
}

int ipp_host_tree_grow(int mode, int norm_variant, double* xy, long long* parent0, long long* parent, double* edge,
                       long long* nchild, int n, int capacity, const double* uniforms, int K, const double* ws4,
                       double step, double radius, double budget_portion, double* travelled_io, long long* rewires,
                       int rewire_cap, int* n_rewires_io, int* scratch, int* consumed, int* n_out, int* finished) {
  if (mode < 0 || mode > 2 || norm_variant < 0 || norm_variant > 2 || !xy || !parent0 || !parent || !edge || !nchild ||
      n <= 0 || capacity < n || K < 0 || (K > 0 && !uniforms) || !ws4 || !travelled_io || !rewires || !n_rewires_io ||
      !scratch || !consumed || !n_out || !finished)
    return -1;
#define IPP_GROW(V)                                                                                                    \
  return grow<V>(mode, xy, parent0, parent, edge, nchild, n, capacity, uniforms, K, ws4, step, radius, budget_portion, \
                 travelled_io, rewires, rewire_cap, n_rewires_io, scratch, consumed, n_out, finished)
  if (norm_variant == 0) IPP_GROW(0);
  if (norm_variant == 1) IPP_GROW(1);
  IPP_GROW(2);
#undef IPP_GROW
}

}  // extern "C"

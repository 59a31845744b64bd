// the RRT primitives (nearest / steer / near / cost / rewire, src/rrt/rrt_utils.py:59-105) and
// the three tree generators built from them (src/rrt/rrt.py:263-327).  HOST code: the growth of one tree is a chain
// of a few hundred dependent insertions (sample -> nearest -> steer -> near -> add -> rewire), each over a few
// hundred nodes, and the trees of the agents follow each other in one RNG stream -- there is nothing for a GPU to
// batch, and in Python every insertion costs ~50 us of interpreter and small-array overhead (4.4 s of an 8 s
// 8-agent run once everything dense is on the device).  This is the same loop natively, on the structure-of-arrays
// tree the device scorer reads (rrt_utils.TreeArrays): ~0.3 us per insertion.
//
// The two scans of every insertion (nearest node, nodes within the radius) go through a uniform grid of cells a little
// wider than the radius: the reference grows every tree with the WHOLE budget (rrt.py:659), so at the size of its own
// example config a tree has 5 000 - 7 000 nodes in a 40 x 40 workspace and the all-pairs scans were 0.27 s per tree.
// The grid only selects CANDIDATES; the distances, the comparison and the tie rules (first minimum / ascending ids)
// are applied to them exactly as the full scans apply them to every node.
//
// Decisions and coordinates are the reference's bit for bit.  Every comparison of the reference is on
// np.linalg.norm of a 2-vector = sqrt(x.dot(x)), whose dot comes from the BLAS numpy links (two roundings, or one
// fused multiply-add, depending on the ddot kernel); `norm_variant` selects the matching form and the Python side
// calibrates it against np.linalg.norm itself before the first use (rrt_utils.native_norm_variant).  Everything
// else is IEEE element-wise arithmetic in the reference's order; the file is compiled with -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>
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

// Uniform grid over the workspace; nodes outside it (a start position off the map) are kept in the border cells, which
// keeps every lower bound below valid (such a node is farther away than its cell).
struct Grid {
  double x0, y0, cell;
  int gx, gy;
  std::vector<int> head, next;
  Grid(const double* ws4, double cell_want, int capacity) {
    const double w = ws4[1] - ws4[0], h = ws4[3] - ws4[2];
    const double span = w > h ? w : h;
    cell = cell_want > 0.0 ? cell_want * (1.0 + 1e-6) : span / 64.0;  // wider than the radius: see near()
    if (!(cell >= span / 256.0)) cell = span / 256.0;
    if (!(cell > 0.0)) cell = 1.0;
    x0 = ws4[0];
    y0 = ws4[2];
    gx = (int)floor(w / cell) + 1;
    gy = (int)floor(h / cell) + 1;
    if (gx < 1) gx = 1;
    if (gy < 1) gy = 1;
    head.assign((size_t)gx * gy, -1);
    next.assign((size_t)capacity, -1);
  }
  inline int cx(double x) const {
    const double f = floor((x - x0) / cell);
    return f < 0.0 ? 0 : (f > gx - 1 ? gx - 1 : (int)f);
  }
  inline int cy(double y) const {
    const double f = floor((y - y0) / cell);
    return f < 0.0 ? 0 : (f > gy - 1 ? gy - 1 : (int)f);
  }
  inline void insert(int i, double x, double y) {
    const int c = cy(y) * gx + cx(x);
    next[i] = head[c];
    head[c] = i;
  }
};

// Path costs for the rewire test.  The reference walks leaf-to-root for every near node of every insertion
// (rrt_utils.py:71-76, 85-88): with ~36 near nodes and paths of a few hundred edges that is the whole cost of a dense
// tree.  Here every node carries an APPROXIMATE cost (parent's + own edge, kept current through rewires by updating
// the subtree) with a rigorous bound on its distance from the walked sum -- both are floating-point sums of the same
// m positive terms, each within (m - 1) eps of the true sum -- and a comparison is decided on the approximations
// only when the bound cannot change it; otherwise the two walks are done as the reference does them.
struct Costs {
  std::vector<double> a;
  std::vector<int> depth, first, nxt, prv, stack;
  explicit Costs(int capacity) : a(capacity, 0.0), depth(capacity, 0), first(capacity, -1), nxt(capacity, -1), prv(capacity, -1) {}
  inline void link(int c, int p) {
    prv[c] = -1;
    nxt[c] = first[p];
    if (first[p] >= 0) prv[first[p]] = c;
    first[p] = c;
  }
  inline void unlink(int c, int p) {
    if (prv[c] >= 0) nxt[prv[c]] = nxt[c];
    else first[p] = nxt[c];
    if (nxt[c] >= 0) prv[nxt[c]] = prv[c];
  }
  inline double slack(int i) const { return 4e-16 * (depth[i] + 2) * a[i]; }
  void refresh_subtree(int c, const double* edge) {  // a[c], depth[c] are current: push them down
    stack.clear();
    stack.push_back(c);
    while (!stack.empty()) {
      const int u = stack.back();
      stack.pop_back();
      for (int v = first[u]; v >= 0; v = nxt[v]) {
        a[v] = a[u] + edge[v];
        depth[v] = depth[u] + 1;
        stack.push_back(v);
      }
    }
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
  Grid grid(ws4, radius > step ? radius : step, capacity);
  for (int i = 0; i < n; ++i) grid.insert(i, xy[2 * i], xy[2 * i + 1]);
  Costs cost(mode != 0 ? capacity : 0);
  if (mode != 0) {  // the nodes the tree already has: children lists, then costs from the roots down
    for (int i = 0; i < n; ++i)
      if (parent[i] >= 0) cost.link(i, (int)parent[i]);
    for (int i = 0; i < n; ++i)
      if (parent[i] < 0) cost.refresh_subtree(i, edge);
  }
  double travelled = *travelled_io;
  int nrw = *n_rewires_io, k = 0, done = travelled < budget_portion ? 0 : 1;
  while (!done && k < K && t.n < capacity && nrw + t.n <= rewire_cap) {
    // np.random.rand(2) * (w, h) + (x0, y0)
    const double tx = uniforms[2 * k] * w + ws4[0], ty = uniforms[2 * k + 1] * h + ws4[2];
    ++k;
    // nearest: min(..., key=norm) keeps the first minimum (rrt_utils.py:100-101): the smallest sqrt(dot), and among
    // equal ones the smallest id.  Rings of cells around the target's cell; once every cell within Chebyshev distance
    // rho has been seen, an unseen node is at least rho cells away, so the search stops when the best distance is
    // below that (with a margin far above the rounding of either side).  A node can only match or beat the running
    // minimum of sqrt(dot) if its dot is not larger (up to rounding), so the square root is taken for those alone.
    long long nn = -1;
    double best_s = 0.0, best = 0.0, best_hi = 0.0;
    {
      const int tcx = grid.cx(tx), tcy = grid.cy(ty);
      for (int rho = 0;; ++rho) {
        bool any = false;
        for (int iy = tcy - rho; iy <= tcy + rho; ++iy) {
          if (iy < 0 || iy >= grid.gy) continue;
          const bool edge_row = (iy == tcy - rho || iy == tcy + rho);
          for (int ix = tcx - rho; ix <= tcx + rho; ix += (edge_row || rho == 0) ? 1 : 2 * rho) {
            if (ix < 0 || ix >= grid.gx) continue;
            any = true;
            for (int i = grid.head[(size_t)iy * grid.gx + ix]; i >= 0; i = grid.next[i]) {
              const double s = dot2<V>(xy[2 * i] - tx, xy[2 * i + 1] - ty);
              if (nn < 0 || s <= best_hi) {
                const double d = sqrt(s);
                if (nn < 0 || d < best || (d == best && i < nn)) {
                  best = d;
                  best_s = s;
                  best_hi = s * (1.0 + 1e-14);
                  nn = i;
                }
              }
            }
          }
        }
        if (nn >= 0 && best < rho * grid.cell * (1.0 - 1e-9)) break;
        if (!any) break;  // the ring lies outside the grid: every cell has been seen
      }
    }
    (void)best_s;
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
    if (mode != 0) {   // near: ascending ids with norm < radius (rrt_utils.py:90-91), before the node is added.
      // The cells are wider than the radius by 1e-6, so a node `reach` + 1 or more cells away is farther than the
      // radius by a margin no rounding bridges; the candidates are tested as the full scan tests every node.
      const int pcx = grid.cx(px), pcy = grid.cy(py);
      const int reach = (int)ceil(radius / grid.cell);
      for (int iy = pcy - reach; iy <= pcy + reach; ++iy) {
        if (iy < 0 || iy >= grid.gy) continue;
        for (int ix = pcx - reach; ix <= pcx + reach; ++ix) {
          if (ix < 0 || ix >= grid.gx) continue;
          for (int i = grid.head[(size_t)iy * grid.gx + ix]; i >= 0; i = grid.next[i]) {
            const double s = dot2<V>(xy[2 * i] - px, xy[2 * i + 1] - py);
            if (s < r2lo || (s <= r2hi && sqrt(s) < radius)) scratch[nnb++] = i;
          }
        }
      }
      std::sort(scratch, scratch + nnb);
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
    grid.insert(nw, px, py);
    travelled += edge[nw];
    if (mode != 0) {   // rewire (rrt_utils.py:85-88)
      cost.link(nw, (int)par);
      cost.depth[nw] = cost.depth[par] + 1;
      cost.a[nw] = cost.a[par] + edge[nw];
      for (int q = 0; q < nnb; ++q) {
        const int c = scratch[q];
        const double d = t.dist(c, px, py);
        const double lhs = cost.a[nw] + d, tol = cost.slack(nw) + cost.slack(c) + 4e-16 * lhs;
        bool take;
        if (lhs + tol < cost.a[c]) take = true;
        else if (lhs - tol >= cost.a[c]) take = false;
        else take = t.path_cost(nw) + d < t.path_cost(c);   // too close for the bound: the reference's own walks
        if (take) {
          cost.unlink(c, (int)parent[c]);
          parent[c] = nw;
          edge[c] = d;
          cost.link(c, nw);
          cost.a[c] = cost.a[nw] + d;
          cost.depth[c] = cost.depth[nw] + 1;
          cost.refresh_subtree(c, edge);
          rewires[3 * nrw] = nw;
          rewires[3 * nrw + 1] = c;
          rewires[3 * nrw + 2] = nw;
          ++nrw;
        }
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

int ipp_host_leaves_dfs(const long long* parent0, int n, int* out, int* n_leaves) {
  if (!parent0 || n <= 0 || !out || !n_leaves) return -1;
  // children in insertion order: a stable counting sort of the nodes by insertion-time parent
  std::vector<int> start((size_t)n + 1, 0), kids((size_t)n > 1 ? (size_t)n - 1 : 0), fill, stack;
  for (int i = 1; i < n; ++i) {
    if (parent0[i] < 0 || parent0[i] >= n) return -1;
    ++start[(size_t)parent0[i] + 1];
  }
  for (int i = 0; i < n; ++i) start[i + 1] += start[i];
  fill.assign(start.begin(), start.end() - 1);
  for (int i = 1; i < n; ++i) kids[fill[parent0[i]]++] = i;
  int m = 0;
  stack.push_back(0);
  while (!stack.empty()) {
    const int i = stack.back();
    stack.pop_back();
    if (start[i] == start[i + 1]) out[m++] = i;
    for (int k = start[i + 1] - 1; k >= start[i]; --k) stack.push_back(kids[k]);   // first child on top
  }
  *n_leaves = m;
  return 0;
}

}  // extern "C"

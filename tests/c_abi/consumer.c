/* A plain C consumer of include/ipp_b200.h: proves the header is valid C, that the library links with a C
 * toolchain (no C++ or torch types at the boundary) and exercises the entry points that need no device --
 * the ABI version, the buffer layout, and the host-side tree growth (ipp_host_norm2, ipp_host_tree_grow).
 * Built and run by tests/test_abi.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "ipp_b200.h"

int main(void) {
  long long lay[10];
  if (ipp_abi_version() < 9) return 10;
  if (ipp_gp_layout(5000, lay) != 0 || lay[0] != 5056 || lay[3] < lay[0] * lay[0]) return 11;
  if (ipp_gp_layout(-1, lay) != -1) return 12;

  double v[4] = {3.0, 4.0, 1e-3, 2e-3}, out[2];
  for (int variant = 0; variant < 3; ++variant) {
    if (ipp_host_norm2(variant, v, 2, out) != 0 || out[0] != 5.0 || fabs(out[1] - sqrt(5e-6)) > 1e-18) return 13;
  }
  if (ipp_host_norm2(7, v, 2, out) != -1) return 14;

  /* grow a rig tree (mode 2) from 64 fixed "uniform" pairs in a 10 x 10 workspace, step = radius = 1 */
  enum { CAP = 128, K = 64 };
  double xy[2 * CAP] = {5.0, 5.0}, edge[CAP] = {0}, uni[2 * K], ws4[4] = {0, 10, 0, 10}, travelled = 0.0;
  long long parent0[CAP], parent[CAP], nchild[CAP] = {0}, rewires[3 * 1024];
  int scratch[CAP], nrw = 0, consumed = 0, n_out = 0, finished = 0;
  unsigned s = 12345u;
  for (int i = 0; i < 2 * K; ++i) { s = s * 1664525u + 1013904223u; uni[i] = (s >> 8) / 16777216.0; }
  parent0[0] = parent[0] = -1;
  int rc = ipp_host_tree_grow(2, 0, xy, parent0, parent, edge, nchild, 1, CAP, uni, K, ws4, 1.0, 1.0, 20.0, &travelled, rewires,
                              1024, &nrw, scratch, &consumed, &n_out, &finished);
  if (rc != 0 || !finished || n_out != consumed + 1 || travelled < 20.0 || travelled > 21.0) return 15;
  double total = 0.0;
  for (int i = 1; i < n_out; ++i) {
    if (parent0[i] < 0 || parent0[i] >= i || edge[i] <= 0.0 || edge[i] > 1.0 + 1e-12) return 16;
    if (xy[2 * i] < 0 || xy[2 * i] > 10 || xy[2 * i + 1] < 0 || xy[2 * i + 1] > 10) return 17;
  }
  for (int i = 0; i < n_out; ++i) total += (double)nchild[i];
  if ((int)total != n_out - 1) return 18;
  printf("c_abi ok: version %d, NP(5000) = %lld, tree of %d nodes from %d samples, %d rewires\n", ipp_abi_version(), lay[0], n_out,
         consumed, nrw);
  return 0;
}

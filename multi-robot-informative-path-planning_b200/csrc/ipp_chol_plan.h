// Task plan of the dataflow Cholesky (host side, no CUDA): which task touches which blocks, in which queue, and
// what it waits for.  The device executor (chol_queue_kernel, ipp_cholq.cuh) is generic: it only reads this table.
//
// Blocks are 64 x 64 (BLK); block row nb (the "y" row block, rows NP .. NP+63 of the factor buffer) is one more row
// block, so nrb = nb + 1.  Work is expressed on TILES of TB x TB blocks at absolute positions (TB = 2: 128 x 128 for
// NP > 1536, TB = 1 otherwise).  Columns are taken in SUPER-PANELS of two block columns (2s, 2s + 1):
//   F(k)          factor diagonal block k.  It first applies, itself, the panels the bulk has not applied to that block:
//                 the whole previous super-panel (k even: blocks k-2, k-1) or the panel k-1 (k odd), so that the
//                 critical chain F(k) -> S(k+1, k) -> F(k+1) never waits for a bulk tile.
//   S(i0, k, nblk) panel solve of row blocks i0 .. i0 + nblk - 1 against L_kk: the block right under the diagonal
//                 (i0 = k + 1, the one the chain waits for) alone, the others one task per tile row.  For k odd it
//                 first applies panel k-1 (the first half of its own super-panel) to its blocks itself.
//   U2(tile, s)   rank-128 update of a tile right of super-panel s by panels 2s, 2s+1 (one read-modify-write of the
//                 tile per TWO panels: half the traffic and twice the k depth of a per-panel update).  Tiles at
//                 least two super-panels away from the front take super-panels in PAIRS (rank 256, 2s' and 2s'+1
//                 applied when the second is done): they have the slack, and the tile's read-modify-write and the
//                 pipeline fill are paid once per four panels.
// Flags (per problem; all monotone counters): F[k] in {0,1}; S[i] = number of panels row block i is solved for;
// A[tile] = number of panels applied to the tile.
// Queues: F, S and C (the U2 tiles of the columns right behind the super-panel, i.e. everything the chain waits
// for) are claimed in order and ONLY WHEN READY (a CTA never blocks on them); the bulk queue U (the rest of U2) is
// claimed in order unconditionally and its owner waits for the dependencies -- polling the chain queues meanwhile.
// Every dependency of a task precedes it in this order of queues, so any set of resident CTAs makes progress.
#pragma once
#include <utility>
#include <vector>

namespace ipp {

constexpr int PLAN_MAGIC = 0x43485132;
constexpr int PLAN_HDR = 16;
constexpr int PLAN_TS = 24;     // ints per task
constexpr int PLAN_MAXDEP = 6;
enum { PT_F = 0, PT_S = 1, PT_U = 2 };
// header slots
enum { PH_MAGIC = 0, PH_NB, PH_NRB, PH_TB, PH_NFLAGS, PH_NF, PH_NS, PH_NC, PH_NU, PH_OFFF, PH_OFFS, PH_OFFC, PH_OFFU, PH_TS,
       PH_TOTAL, PH_NTILE };
// task slots
enum { TK_TYPE = 0, TK_A, TK_B, TK_C, TK_D, TK_E, TK_F, TK_G, TK_NDEP, TK_DEP0, TK_OUT = TK_DEP0 + 2 * PLAN_MAXDEP, TK_OUTVAL,
       TK_OUT2 };  // TK_OUT2: second flag receiving TK_OUTVAL (-1: none)
static_assert(TK_OUT2 < PLAN_TS, "task record too small");

struct PlanTask {
  int v[PLAN_TS];
  PlanTask() {
    for (int i = 0; i < PLAN_TS; ++i) v[i] = 0;
    for (int d = 0; d < PLAN_MAXDEP; ++d) v[TK_DEP0 + 2 * d] = -1;
    v[TK_OUT2] = -1;
  }
  void dep(int idx, int val) {
    if (val <= 0) return;
    for (int d = 0; d < v[TK_NDEP]; ++d)
      if (v[TK_DEP0 + 2 * d] == idx) {
        if (v[TK_DEP0 + 2 * d + 1] < val) v[TK_DEP0 + 2 * d + 1] = val;
        return;
      }
    const int d = v[TK_NDEP]++;
    v[TK_DEP0 + 2 * d] = idx;
    v[TK_DEP0 + 2 * d + 1] = val;
  }
};

// Builds the plan for NP = 64 * nb.  Returns the number of ints.
inline long chol_plan_build(int nb, int tb, int near_slack, std::vector<int>& out) {
  const int nrb = nb + 1;
  const int tmax = (nrb - 1) / tb;                    // last tile row
  const int ntile = (tmax + 1) * (tmax + 2) / 2;
  const int fF = 0, fS = nb, fA = nb + nrb, nflags = nb + nrb + ntile;
  auto tile_flag = [&](int ti, int tj) { return fA + ti * (ti + 1) / 2 + tj; };
  std::vector<PlanTask> qF, qS, qC, qU;
  // S tasks of panel k: {k+1} alone, then the rest of its tile row (if any), then whole tile rows
  auto s_groups = [&](int k, std::vector<std::pair<int, int>>& out) {
    out.clear();
    int i = k + 1;
    if (i < nrb) out.push_back({i, 1}), ++i;
    while (i < nrb) {
      int e = (i / tb + 1) * tb;
      if (e > nrb) e = nrb;
      out.push_back({i, e - i});
      i = e;
    }
  };
  std::vector<std::pair<int, int>> grp;
  std::vector<int> s_first(nb, -1);                   // queue index of S(k+1, k)
  {
    int pos = 0;
    for (int k = 0; k < nb; ++k) {
      s_first[k] = pos;
      s_groups(k, grp);
      pos += (int)grp.size();
    }
  }

  for (int k = 0; k < nb; ++k) {                      // ---- F
    PlanTask t;
    t.v[TK_TYPE] = PT_F;
    t.v[TK_A] = k;
    t.v[TK_B] = (k & 1) ? 1 : (k >= 2 ? 2 : 0);       // own update: panels [k - width, k)
    t.dep(fS + k, k);                                 // L(k, p) final for every p < k
    t.dep(tile_flag(k / tb, k / tb), (k & 1) ? k - 1 : k - 2);
    t.v[TK_G] = s_first[k];                           // hint: the solve on the critical chain comes next
    t.v[TK_OUT] = fF + k;
    t.v[TK_OUTVAL] = 1;
    qF.push_back(t);
  }
  for (int k = 0; k < nb; ++k) {                      // ---- S
    s_groups(k, grp);
    for (auto [i0, nblk] : grp) {
      PlanTask t;
      t.v[TK_TYPE] = PT_S;
      t.v[TK_A] = i0;
      t.v[TK_B] = k;
      t.v[TK_C] = nblk;
      t.v[TK_D] = k & 1;                              // own update: panel k - 1 (k odd)
      t.v[TK_G] = (i0 == k + 1 && k + 1 < nb) ? k + 1 : -1;   // hint: F(k+1) comes next
      t.dep(fF + k, 1);
      t.dep(tile_flag(i0 / tb, k / tb), k & ~1);      // every earlier super-panel applied to the blocks (i, k)
      if (k & 1) {                                    // operands of the own update: L(i, k-1), L(k, k-1)
        t.dep(fS + k, k);
        for (int b = 0; b < nblk; ++b) t.dep(fS + i0 + b, k);
      }
      t.v[TK_OUT] = fS + i0;
      t.v[TK_OUTVAL] = k + 1;
      t.v[TK_OUT2] = nblk > 1 ? fS + i0 + 1 : -1;
      qS.push_back(t);
    }
  }
  auto u_task = [&](int ti, int tj, int kb, int ke, int cmin, int skip, int applied_before) {
    PlanTask t;
    t.v[TK_TYPE] = PT_U;
    t.v[TK_A] = ti;
    t.v[TK_B] = tj;
    t.v[TK_C] = kb;
    t.v[TK_D] = ke;
    t.v[TK_E] = cmin;
    t.v[TK_F] = skip;
    t.dep(tile_flag(ti, tj), applied_before);
    for (int i = ti * tb; i < (ti + 1) * tb && i < nrb; ++i) t.dep(fS + i, ke);   // A operand rows
    for (int j = tj * tb; j < (tj + 1) * tb && j < nb; ++j)
      if (j >= cmin) t.dep(fS + j, ke);                                           // B operand rows
    t.v[TK_OUT] = tile_flag(ti, tj);
    t.v[TK_OUTVAL] = ke;
    return t;
  };
  for (int s = 0; 2 * s < nb; ++s) {
    const int p0 = 2 * s, p1 = 2 * s + 1;
    const int ke = (p1 < nb) ? p1 + 1 : p1;
    // U2: tiles whose columns start at or beyond block 2s + 2; block (2s+2, 2s+2) is F(2s+2)'s own
    const int tj0 = (2 * s + 2 + tb - 1) / tb;
    for (int tj = tj0; tj <= (nb - 1) / tb; ++tj)
      for (int ti = tj; ti <= tmax; ++ti) {
        const int skip = 2 * s + 2;
        if (tb == 1 && ti == tj && ti == skip) continue;   // the tile IS the skipped block
        // far tiles: super-panels (s & ~1, s | 1) as one rank-256 task, emitted at the odd one
        const int tjs = (tj * tb) / 2;                     // super-panel the tile's first column belongs to
        const bool paired = ((s | 1) <= tjs - 2) && 2 * (s | 1) + 1 < nb;
        if (paired && !(s & 1)) continue;
        const int kb = paired ? 2 * (s - 1) : p0;
        PlanTask t = u_task(ti, tj, kb, ke, tj * tb, skip, kb);
        // Chain priority (queue C) for every tile the front reaches within near_slack + 1 super-panels: the next
        // super-panel's own columns (slack 0: the chain waits for them at once) and near_slack more.  Left in the
        // bulk queue, a tile of slack 1 sits behind the hundreds of far tiles released one pair earlier, and the
        // chain -- hence the release of the next bulk tiles -- stalls until those are claimed.
        (tjs - (s + 1) <= near_slack ? qC : qU).push_back(t);
      }
  }
  out.assign(PLAN_HDR, 0);
  out[PH_MAGIC] = PLAN_MAGIC;
  out[PH_NB] = nb;
  out[PH_NRB] = nrb;
  out[PH_TB] = tb;
  out[PH_NFLAGS] = nflags;
  out[PH_NF] = (int)qF.size();
  out[PH_NS] = (int)qS.size();
  out[PH_NC] = (int)qC.size();
  out[PH_NU] = (int)qU.size();
  out[PH_TS] = PLAN_TS;
  out[PH_NTILE] = ntile;
  const std::vector<PlanTask>* qs[4] = {&qF, &qS, &qC, &qU};
  for (int q = 0; q < 4; ++q) {
    out[PH_OFFF + q] = (int)out.size();
    for (const PlanTask& t : *qs[q]) out.insert(out.end(), t.v, t.v + PLAN_TS);
  }
  out[PH_TOTAL] = (int)out.size();
  return (long)out.size();
}

}  // namespace ipp

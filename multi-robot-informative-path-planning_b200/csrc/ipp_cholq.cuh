// Dataflow Cholesky: one persistent kernel that executes the task plan of ipp_chol_plan.h for R independent
// factorisations at once (the optimiser restarts of one fit, or R = 1).  No grid-wide barrier: every task waits
// only for the blocks it reads, through monotone flags in global memory (release store by the producer after a CTA
// barrier, relaxed polling + fence by the consumer), so the critical chain
//     F(k) -> S(k+1, k) -> F(k+1)                    (64 x 64 blocks, one CTA hands over to itself by "hints")
// runs ahead of the bulk rank-128 tile updates instead of alternating with them, and the chains of several problems
// fill each other's bubbles.  Included by ipp_linalg.cu inside namespace ipp (uses its block routines).
#pragma once

constexpr int QSTR2 = 2 * BLK + 4;  // pitch of a 64 x 128 operand read as DMMA fragments (132 % 16 == 4: conflict free)
constexpr size_t FACT_SMEM = sizeof(double) * (BLK * PSTR + BLK * QSTR2 + BLK);

struct CholQ {
  double* K;          // problem r lives at K + slot(r) * pstride
  long pstride, ld;
  int rows_total, np;
  double* vbuf;       // info slot of problem r: vbuf + slot(r) * vstride + info_off
  long vstride, info_off;
  const int* slots;   // nullable: slot(r) = r
  int R;
  const int* plan;
  int* sched;         // [R][nflags] flags | [R][QSTATE] queue heads and ready frontiers | -, abort | per-CTA stats
  long long timeout;  // cycles a CTA may poll without finding work before it raises `abort`
  int use_hints;
  int stats;          // != 0: every CTA leaves {tasks F/S/U, cycles F/S/U, cycles acquiring, cycles total} after the abort word
  int use_mb;         // tile updates on the mbarrier pipeline (gemm_mainloop_mb) instead of a CTA barrier per k-slab
  int eager_scan;     // advance the frontiers after every chain task, even if a chain task is claimable (A/B aid)
  int red_epilogue;   // tile updates written back as RED.ADD.F64 instead of load / subtract / store
};
constexpr int QSTATS = 16;       // ints per CTA
constexpr int QSTATS_CTAS = 512;
constexpr int QTRACE_EVENTS = 448;  // stats == 2: per CTA, {queue | problem << 8, task index, start ns, end ns} per task
constexpr int QTRACE_CTAS = 160;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// All dependency flags of a task at or above their thresholds?  The loads are relaxed and independent (in flight
// together); the caller fences before the CTA touches the data they guard.
__device__ __forceinline__ bool task_ready(const int* __restrict__ T, const int* flags) {
  int idx[PLAN_MAXDEP], val[PLAN_MAXDEP], got[PLAN_MAXDEP];
#pragma unroll
  for (int d = 0; d < PLAN_MAXDEP; ++d) {
    idx[d] = __ldg(T + TK_DEP0 + 2 * d);
    val[d] = __ldg(T + TK_DEP0 + 2 * d + 1);
  }
#pragma unroll
  for (int d = 0; d < PLAN_MAXDEP; ++d) got[d] = idx[d] >= 0 ? ld_relaxed(flags + idx[d]) : 0x7fffffff;
  bool ok = true;
#pragma unroll
  for (int d = 0; d < PLAN_MAXDEP; ++d) ok = ok && (got[d] >= val[d]);
  return ok;
}

// F task: D = A(k, k); D -= P P^T with P = A(k, [k - width, k)) (the panels the bulk leaves to this block); factor.
template <int NT, class Sync = CtaSync>
__device__ __forceinline__ void cholq_factor(double* __restrict__ A, long ld, int k, int width, double* psm, double* info,
                                             Sync sync = Sync()) {
  double* D = psm;
  double* P = psm + BLK * PSTR;
  const int t = threadIdx.x, j0 = BLK * k, p0 = BLK * (k - width), pw = BLK * width;
  // every load of the task in flight at once (these blocks come from L2: one round trip, not one per row)
  for (int e = t; e < BLK * pw / 2; e += NT) {
    const int r = e / (pw / 2), cc = (e % (pw / 2)) * 2;
    cp_async16(P + r * QSTR2 + cc, A + (long)(j0 + r) * ld + p0 + cc, true);
  }
  cp_async_commit();
  {
    constexpr int PER = BLK * BLK / 2 / NT;
    double2 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = t + i * NT, r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
      v[i] = __ldcg(reinterpret_cast<const double2*>(A + (long)(j0 + r) * ld + j0 + cc));
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = t + i * NT, r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
      D[r * PSTR + cc] = v[i].x;
      D[r * PSTR + cc + 1] = v[i].y;
    }
  }
  cp_async_wait<0>();
  sync();
  if (width > 0) {  // lower part of D -= P P^T on the DMMA pipe: each warp owns 64 / (NT / 32) rows
    constexpr int RPW = BLK / (NT / 32), MT = RPW / 8;
    const int warp = t >> 5, lane = t & 31, g = lane >> 2, tg = lane & 3;
    double acc[MT][8][2];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int nlim = (warp * RPW + RPW + 7) / 8;  // 8-column groups at or left of this warp's last row
#pragma unroll 2
    for (int kk = 0; kk < pw; kk += 4) {
      double af[MT], bf[8];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) af[mi] = P[(warp * RPW + mi * 8 + g) * QSTR2 + kk + tg];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) bf[ni] = (ni < nlim) ? P[(ni * 8 + g) * QSTR2 + kk + tg] : 0.0;
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
          if (ni < nlim) dmma884(acc[mi][ni], af[mi], bf[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = warp * RPW + mi * 8 + g, cc = ni * 8 + tg * 2 + e;
          if (cc <= r) D[r * PSTR + cc] -= acc[mi][ni][e];
        }
    sync();
  }
  potf2_block<NT, Sync>(D, t, j0, info, sync);
  for (int e = t; e < BLK * BLK / 2; e += NT) {
    const int r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
    *reinterpret_cast<double2*>(A + (long)(j0 + r) * ld + j0 + cc) =
        make_double2(cc <= r ? D[r * PSTR + cc] : 0.0, cc + 1 <= r ? D[r * PSTR + cc + 1] : 0.0);
  }
  sync();
}

// S task: rows [64 i0, 64 (i0 + nblk)) of block column k:  X <- (X - Pa Pb^T) L_kk^-T, where Pa = A(rows, panel k-1),
// Pb = A(k, panel k-1) is the one panel the bulk leaves to this task (width 1: k odd, the first half of its
// super-panel; width 0: nothing).  Blocked forward substitution: 8-column groups; inside a group every row is solved
// by one thread with the 8 x 8 triangle of L_kk read as shared-memory broadcasts (reciprocal pivots: one divide per
// column, as in chol_trsm_block), the groups to its right are then updated on the DMMA pipe.  16 CTA barriers per
// task; 64 dependent steps per row in all, but each is an FMA out of registers instead of a shared-memory sweep
// (chol_trsm_block: 6 us per 64 rows; this: ~2 us per 128 rows).
template <int SOLVE_ROWS>
constexpr size_t solve_smem() {  // X, Pa, Pb, L_kk (+ reciprocal pivots)
  return sizeof(double) * ((SOLVE_ROWS + SOLVE_ROWS + BLK + BLK) * QSTR + BLK);
}
template <int NT, int SOLVE_ROWS, class Sync = CtaSync>
__device__ __forceinline__ void cholq_solve(double* __restrict__ A, long ld, int k, int i0, int nblk, int width, double* psm,
                                            Sync sync = Sync()) {
  const int rows = BLK * nblk, t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tg = lane & 3;
  double* Xs = psm;                          // rows x 64, pitch QSTR
  double* Pa = psm + SOLVE_ROWS * QSTR;      // rows x 64 (own update)
  double* Pb = psm + 2 * SOLVE_ROWS * QSTR;  // 64 x 64 (own update)
  double* Ls = Pb + BLK * QSTR;              // L_kk
  double* rinv = Ls + BLK * QSTR;
  const int j0 = BLK * k, r0 = BLK * i0, p0 = BLK * (k - 1);
  auto load_block = [&](double* dst, int grow0, int gcol0, int nrows) {  // asynchronous: all blocks in flight at once
    for (int e = t; e < nrows * (BLK / 2); e += NT) {
      const int r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
      cp_async16(dst + r * QSTR + cc, A + (long)(grow0 + r) * ld + gcol0 + cc, true);
    }
  };
  load_block(Xs, r0, j0, rows);
  if (width > 0) {
    load_block(Pa, r0, p0, rows);
    load_block(Pb, j0, p0, BLK);
  }
  cp_async_commit();
  load_block(Ls, j0, j0, BLK);
  cp_async_commit();
  if (t < BLK) rinv[t] = 1.0 / __ldcg(A + (long)(j0 + t) * ld + j0 + t);
  cp_async_wait<1>();  // X and the own-update operands have landed; L_kk may still be on its way
  sync();
  constexpr int NW = NT / 32;
  if (width > 0) {  // Xs -= Pa Pb^T: warp w owns rows [w * rows / NW, ...) in 8-row groups, all 64 columns
    const int rpw = rows / NW;
    for (int m0 = 0; m0 < rpw; m0 += 8) {
      const int rr = warp * rpw + m0 + g;
      double acc[8][2];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
#pragma unroll 4
      for (int kk = 0; kk < BLK; kk += 4) {
        const double af = Pa[rr * QSTR + kk + tg];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) dmma884(acc[ni], af, Pb[(ni * 8 + g) * QSTR + kk + tg]);
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        double2* x = reinterpret_cast<double2*>(Xs + rr * QSTR + ni * 8 + tg * 2);
        const double2 v = *x;
        *x = make_double2(v.x - acc[ni][0], v.y - acc[ni][1]);
      }
    }
  }
  cp_async_wait<0>();
  sync();
  for (int c = 0; c < 8; ++c) {
    if (t < rows) {  // one thread per row: the 8 x 8 triangle of group c
      double x[8];
      double* xr = Xs + t * QSTR + 8 * c;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = xr[j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int p = 0; p < j; ++p) x[j] -= x[p] * Ls[(8 * c + j) * QSTR + 8 * c + p];
        x[j] *= rinv[8 * c + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[j] = x[j];
    }
    sync();
    if (c < 7) {  // columns right of the group: Xs[:, n] -= X[:, group c] L[n, group c]^T
      const int rpw = rows / NW, ntile = 7 - c;
      for (int m0 = 0; m0 < rpw; m0 += 8) {
        const int rr = warp * rpw + m0 + g;
        const double a0 = -Xs[rr * QSTR + 8 * c + tg], a1 = -Xs[rr * QSTR + 8 * c + 4 + tg];
        double acc[7][2], b0[7], b1[7];  // all the column groups of this row group in flight together
#pragma unroll
        for (int ni = 0; ni < 7; ++ni)
          if (ni < ntile) {
            const int cn = 8 * (c + 1 + ni);
            const double2 v = *reinterpret_cast<const double2*>(Xs + rr * QSTR + cn + tg * 2);
            acc[ni][0] = v.x;
            acc[ni][1] = v.y;
            b0[ni] = Ls[(cn + g) * QSTR + 8 * c + tg];
            b1[ni] = Ls[(cn + g) * QSTR + 8 * c + 4 + tg];
          }
#pragma unroll
        for (int ni = 0; ni < 7; ++ni)
          if (ni < ntile) {
            dmma884(acc[ni], a0, b0[ni]);
            dmma884(acc[ni], a1, b1[ni]);
          }
#pragma unroll
        for (int ni = 0; ni < 7; ++ni)
          if (ni < ntile)
            *reinterpret_cast<double2*>(Xs + rr * QSTR + 8 * (c + 1 + ni) + tg * 2) = make_double2(acc[ni][0], acc[ni][1]);
      }
      sync();
    }
  }
  for (int e = t; e < rows * (BLK / 2); e += NT) {
    const int r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
    *reinterpret_cast<double2*>(A + (long)(r0 + r) * ld + j0 + cc) = *reinterpret_cast<const double2*>(Xs + r * QSTR + cc);
  }
  sync();
}

// C(tile) -= acc over the lower part, columns >= cmin, except the diagonal block `skip`: the loads of half the tile are
// issued before the first dependent store, so the read-modify-write costs two L2 round trips, not one per element.
template <class C>
__device__ __forceinline__ void cholq_tile_epilogue(double* __restrict__ A, long ld, int rows_total, int np, int row0, int col0,
                                                    int cmin, int skip, double (&acc)[C::MI][C::NI][2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N, g = lane >> 2, tg = lane & 3;
  constexpr int HALF = C::MI >= 8 ? C::MI / 4 : (C::MI / 2 > 0 ? C::MI / 2 : 1);
#pragma unroll
  for (int h0 = 0; h0 < C::MI; h0 += HALF) {
    double2 cv[HALF][C::NI];
    bool ok[HALF][C::NI];
#pragma unroll
    for (int mi = 0; mi < HALF; ++mi)
#pragma unroll
      for (int ni = 0; ni < C::NI; ++ni) {
        const int gr = row0 + wm * C::WTM + (h0 + mi) * 8 + g, gc = col0 + wn * C::WTN + ni * 8 + tg * 2;
        const int br = gr >> 6, bc = gc >> 6;
        ok[mi][ni] = gr < rows_total && gc < np && gc >= cmin && bc <= br && !(br == skip && bc == skip);
        if (ok[mi][ni]) cv[mi][ni] = __ldcg(reinterpret_cast<const double2*>(A + (long)gr * ld + gc));
      }
#pragma unroll
    for (int mi = 0; mi < HALF; ++mi)
#pragma unroll
      for (int ni = 0; ni < C::NI; ++ni) {
        const int gr = row0 + wm * C::WTM + (h0 + mi) * 8 + g, gc = col0 + wn * C::WTN + ni * 8 + tg * 2;
        if (ok[mi][ni])
          *reinterpret_cast<double2*>(A + (long)gr * ld + gc) =
              make_double2(cv[mi][ni].x - acc[h0 + mi][ni][0], cv[mi][ni].y - acc[h0 + mi][ni][1]);
      }
  }
}

// The same update as fire-and-forget reductions (RED.ADD.F64 at the L2): C + (-acc) rounds exactly like C - acc, every
// element is touched by one task at a time (the flags serialise the updates of a tile), so the result is the same bits
// -- without the load -> subtract -> store round trips of the read-modify-write.
template <class C>
__device__ __forceinline__ void cholq_tile_epilogue_red(double* __restrict__ A, long ld, int rows_total, int np, int row0,
                                                        int col0, int cmin, int skip, double (&acc)[C::MI][C::NI][2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N, g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int mi = 0; mi < C::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < C::NI; ++ni) {
      const int gr = row0 + wm * C::WTM + mi * 8 + g, gc = col0 + wn * C::WTN + ni * 8 + tg * 2;
      const int br = gr >> 6, bc = gc >> 6;
      if (gr < rows_total && gc < np && gc >= cmin && bc <= br && !(br == skip && bc == skip)) {
        double* p = A + (long)gr * ld + gc;
        asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;\n" ::"l"(p), "d"(-acc[mi][ni][0]) : "memory");
        asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;\n" ::"l"(p + 1), "d"(-acc[mi][ni][1]) : "memory");
      }
    }
}

// U task: tile (row0, col0) -= A(rows, [kb, ke)) A(cols, [kb, ke))^T over the lower part, columns >= cmin, except the
// diagonal block `skip` (block index; that one is its F task's own).
template <class C>
__device__ __forceinline__ void cholq_update_tile(double* __restrict__ A, long ld, int rows_total, int np, int row0, int col0,
                                                  int kb, int ke, int cmin, int skip, double* smem, uint64_t* bars) {
  {  // start pulling the C tile into L2 (see chol_syrk_tile)
    constexpr int LINES_PER_ROW = C::BN * 8 / 128;
    for (int e = threadIdx.x; e < C::BM * LINES_PER_ROW; e += C::THREADS) {
      const int gr = row0 + e / LINES_PER_ROW, gc = col0 + (e % LINES_PER_ROW) * 16;
      if (gr < rows_total && gc < np && gc >= cmin && (gc >> 6) <= (gr >> 6))
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A + (long)gr * ld + gc));
    }
  }
  double acc[C::MI][C::NI][2];
  zero_acc<C>(acc);
  auto fillA = [&](double* s, int k0) { load_tile_async<C::BM, Lay::KCONT, C::THREADS>(s, A, ld, row0, k0, rows_total, ke); };
  auto fillB = [&](double* s, int k0) { load_tile_async<C::BN, Lay::KCONT, C::THREADS>(s, A, ld, col0, k0, np, ke); };
  if (bars)
    gemm_mainloop_mb<C, Lay::KCONT, Lay::KCONT>(acc, smem, bars, kb, ke, fillA, fillB);
  else
    gemm_mainloop<C, Lay::KCONT, Lay::KCONT>(acc, smem, kb, ke, fillA, fillB);
  cholq_tile_epilogue<C>(A, ld, rows_total, np, row0, col0, cmin, skip, acc);
  __syncthreads();
}

constexpr int NQ = 4;      // queues per problem: F, S, C (chain), U (bulk)
constexpr int QSTATE = 8;  // ints of queue state per problem: {head, ready frontier} x NQ

// Claiming.  Each queue (F, S, C, U of every problem) has a HEAD (next task to hand out) and a READY FRONTIER (every
// task below it has its dependencies met).  The frontier is advanced by whoever has reason to believe it can move --
// a CTA that has just finished a chain task (its flag is what the next tasks wait for) and every CTA that finds
// nothing to claim: the lanes of the polling warp test a window of tasks beyond the frontier in parallel and add the
// leading run of ready ones with one atomicMax.  Tasks are claimed with atomicAdd on the head whenever
// head < frontier, so any number of CTAs claim different ready tasks in the same round trip (compare-and-swap on the
// head hands out ONE task per round trip to all contenders together: measured 4x slower than the cooperative
// kernel).  A racing claim can land just beyond the frontier; its owner HOLDS that task (tests it every poll, runs
// it as soon as it is ready) and keeps serving the queues meanwhile -- with compare-and-swap claims, which cannot
// overshoot, so a CTA never holds more than one task.  (Waiting for a held task without serving deadlocks: every CTA
// can end up holding a task behind an F task nobody is free to run -- found by the host simulation of these rules,
// tests/test_chol_plan.py.)  No task is ever claimed before it is ready or about to be, so a CTA never sits on work
// of one problem while another problem has tiles to update (an in-order, claim-then-wait bulk queue did: 27 % of
// the CTA time at four problems went into waiting for the slowest chain).
// Priority: F, S, C, U; among the problems each CTA starts from its own (blockIdx mod R), which spreads the CTAs.
struct QueueView {
  const int* plan;
  int* sched;
  int* qstate;
  int nflags, R;
};

// Advance the ready frontiers of the NQ queues of problem r (8 lanes per queue, two tasks per lane: a window of 16;
// the bulk queue, where hundreds of tiles become ready at once, is re-scanned while its whole window was ready).
// Returns true (warp-uniform) if any frontier moved.
__device__ __forceinline__ bool scan_frontiers(const QueueView& v, int r, int lane) {
  constexpr int PER = 32 / NQ;
  const int q = lane / PER, o = lane % PER;
  int* st = v.qstate + QSTATE * r + 2 * q;
  const int cnt = __ldg(v.plan + PH_NF + q);
  const int* Tq = v.plan + __ldg(v.plan + PH_OFFF + q);
  const int* fl = v.sched + (long)r * v.nflags;
  bool advanced = false;
  for (int rep = 0; rep < 8; ++rep) {
    const int f = ld_relaxed(st + 1);
    const bool r0 = (f + o < cnt) && task_ready(Tq + (f + o) * PLAN_TS, fl);
    const bool r1 = (f + PER + o < cnt) && task_ready(Tq + (f + PER + o) * PLAN_TS, fl);
    const unsigned m0 = __ballot_sync(0xffffffffu, r0), m1 = __ballot_sync(0xffffffffu, r1);
    int m = 0;
    if (o == 0) {
      const unsigned gmask = (1u << PER) - 1u;
      const unsigned z0 = ~(m0 >> lane) & gmask, z1 = ~(m1 >> lane) & gmask;
      m = z0 ? __ffs(z0) - 1 : PER;  // leading run of ready tasks
      if (m == PER) m += z1 ? __ffs(z1) - 1 : PER;
      if (m > 0) advanced = atomicMax(st + 1, f + m) >= 0;  // (the returned value makes the warp wait for the update)
    }
    const int m_bulk = __shfl_sync(0xffffffffu, m, (NQ - 1) * PER);
    if (m_bulk < 2 * PER) break;  // the bulk window was not all ready: nothing more to release right now
  }
  return __any_sync(0xffffffffu, advanced);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) chol_queue_kernel(CholQ a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task[4];  // {queue (0 F, 1 S, 2 C, 3 U; < 0: stop), task index, problem, -}
  __shared__ __align__(8) uint64_t bars[2 * MB_STAGES];
  const int* __restrict__ plan = a.plan;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nflags = __ldg(plan + PH_NFLAGS);
  const int R = a.R, nq = NQ * R;
  int* qstate = a.sched + (long)R * nflags;
  int* abortf = qstate + QSTATE * R + 1;
  const QueueView qv = {plan, a.sched, qstate, nflags, R};
  const int rot = blockIdx.x % R;           // this CTA's first problem
  // ---- warp 0 scheduling state (uniform over the warp)
  int hint_q = -1, hint_i = 0, hint_r = 0;  // successor of the task just finished (tried first)
  int held_q = -1, held_i = 0, held_r = 0;  // task claimed past its queue's frontier, not started yet
  int scan_r = -1;                          // problem whose frontiers to advance before claiming (-1: none)
  int idle_r = rot;                         // problem an idle poll scans next
  long long idle_since = 0;
  bool idle = false;
  long long st_cyc[4] = {0, 0, 0, 0}, st_t0 = clock64(), st_mark = st_t0;  // thread 0: F, S, U, acquiring
  int st_n[3] = {0, 0, 0}, st_polls = 0, tr_n = 0;
  unsigned long long tr_start = 0;
  int* trace = abortf + 7 + QSTATS * QSTATS_CTAS;  // after the per-CTA counters

  for (;;) {
    if (warp == 0) {
      int got_q = -1, got_i = 0, got_r = 0;
      unsigned backoff = 0;
      for (;;) {
        ++st_polls;
        // ---- one pass of independent loads: lane 6 the hinted successor, lane 7 the held task, lanes 8.. the queues
        int special = 0;  // lane 6: hint can be claimed; lane 7: held task is ready
        if (lane == 6 && hint_q >= 0) {
          const int* T = plan + __ldg(plan + PH_OFFF + hint_q) + hint_i * PLAN_TS;
          const int h = ld_relaxed(qstate + QSTATE * hint_r + 2 * hint_q);
          special = (task_ready(T, a.sched + (long)hint_r * nflags) && h == hint_i) ? 1 : 0;
        }
        if (lane == 7 && held_q >= 0)
          special = task_ready(plan + __ldg(plan + PH_OFFF + held_q) + held_i * PLAN_TS, a.sched + (long)held_r * nflags) ? 1 : 0;
        // candidate key = queue * R + (problem - rot mod R): lowest key wins
        int best = 0x7fffffff, best_rdy = 0, best_h = 0;
        bool exhausted = held_q < 0;
        if (lane >= 8)
          for (int c = lane - 8; c < nq; c += 24) {
            const int q = c / R, r = (c % R + rot) % R;
            const int* st = qstate + QSTATE * r + 2 * q;
            const int h = ld_relaxed(st), f = ld_relaxed(st + 1);
            if (h < __ldg(plan + PH_NF + q)) exhausted = false;
            if (h < f && c < best) {
              best = c;
              best_rdy = f;
              best_h = h;
            }
          }
        const int hint_ok = __shfl_sync(0xffffffffu, special, 6), held_ok = __shfl_sync(0xffffffffu, special, 7);
        int win = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) win = min(win, __shfl_xor_sync(0xffffffffu, win, o));
        const bool all_exhausted = __all_sync(0xffffffffu, exhausted);
        // ---- decide: held > hint > queues in priority order
        if (held_q >= 0 && held_ok) {
          got_q = held_q;
          got_i = held_i;
          got_r = held_r;
          held_q = -1;
          break;
        }
        if (hint_q >= 0) {
          const int hq = hint_q;
          hint_q = -1;
          int ok = 0;
          if (hint_ok) {
            if (lane == 0) {
              int* hp = qstate + QSTATE * hint_r + 2 * hq;
              ok = (atomicCAS(hp, hint_i, hint_i + 1) == hint_i);
              if (ok) atomicMax(hp + 1, hint_i + 1);
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
          }
          if (ok) {
            got_q = hq;
            got_i = hint_i;
            got_r = hint_r;
            break;
          }
        }
        // A chain task has just finished: what it enables is not below any frontier yet.  If a chain-class task is
        // claimable anyway, take it (the frontiers are ahead of the heads: nothing is waiting to be discovered that
        // would not be found by the next CTA that runs dry); otherwise advance this problem's frontiers first.
        if (scan_r >= 0 && (a.eager_scan || !(win != 0x7fffffff && win / R < NQ - 1))) {
          const int sr = scan_r;
          scan_r = -1;
          if (scan_frontiers(qv, sr, lane)) continue;
        }
        if (win != 0x7fffffff) {
          const int q = win / R, r = (win % R + rot) % R;
          const int owner = 8 + win % 24;
          int h = 0, ok = 0;
          if (lane == owner) {
            int* hp = qstate + QSTATE * r + 2 * q;
            if (held_q < 0) {
              h = atomicAdd(hp, 1);
              ok = h < __ldg(plan + PH_NF + q) ? (h < best_rdy ? 1 : 2) : 0;  // 2: beyond the frontier seen -> test it
              if (ok == 2 && task_ready(plan + __ldg(plan + PH_OFFF + q) + h * PLAN_TS, a.sched + (long)r * nflags)) ok = 1;
            } else {  // already holding a task: claim exactly the head seen (below the frontier, hence ready)
              h = best_h;
              ok = (atomicCAS(hp, best_h, best_h + 1) == best_h) ? 1 : 0;
            }
          }
          ok = __shfl_sync(0xffffffffu, ok, owner);
          h = __shfl_sync(0xffffffffu, h, owner);
          if (ok == 1) {
            got_q = q;
            got_i = h;
            got_r = r;
            break;
          }
          if (ok == 2) {
            held_q = q;
            held_i = h;
            held_r = r;
          }
          continue;  // lost a race, claimed past the end, or now holding: look again
        }
        if (all_exhausted) {
          got_q = -1;
          break;
        }
        // ---- nothing to claim: advance the frontiers of one problem after the other; if a full round over the
        //      problems moved nothing, watchdog and back off
        bool moved = false;
        for (int k = 0; k < R && !moved; ++k) {
          moved = scan_frontiers(qv, idle_r, lane);
          idle_r = (idle_r + 1) % R;
        }
        if (moved) continue;
        int stop = 0;
        if (lane == 0) {
          const long long now = clock64();
          if (!idle) {
            idle = true;
            idle_since = now;
          }
          if (now - idle_since > a.timeout) atomicExch(abortf, 1);
          stop = ld_relaxed(abortf);
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) {
          got_q = -2;
          break;
        }
        backoff = backoff < 512 ? backoff + 64 : 512;
        __nanosleep(backoff);
      }
      idle = false;
      if (lane == 0) {
        s_task[0] = got_q;
        s_task[1] = got_i;
        s_task[2] = got_r;
      }
      // acquire side: the flags were read with relaxed loads (possibly by another CTA, via the frontier): fence before
      // this CTA touches the data they guard
      __threadfence();
    }
    __syncthreads();
    const int q = s_task[0], ti = s_task[1], r = s_task[2];
    if (a.stats && t == 0) {
      const long long now = clock64();
      st_cyc[3] += now - st_mark;
      st_mark = now;
    }
    if (q < 0) {
      if (q == -2 && t == 0)  // watchdog: report through every problem's info slot (the host raises)
        for (int rr = 0; rr < R; ++rr) a.vbuf[(a.slots ? (long)a.slots[rr] : (long)rr) * a.vstride + a.info_off] = -7777.0;
      break;
    }
    if (a.stats == 2 && t == 0) tr_start = global_ns();
    const int* T = plan + __ldg(plan + PH_OFFF + q) + ti * PLAN_TS;
    const int slot = a.slots ? __ldg(a.slots + r) : r;
    double* A = a.K + (long)slot * a.pstride;
    int* flags = a.sched + (long)r * nflags;
    const int type = __ldg(T + TK_TYPE), pa = __ldg(T + TK_A), pb = __ldg(T + TK_B), pc = __ldg(T + TK_C), pd = __ldg(T + TK_D);
    if (type == PT_F) {
      cholq_factor<C::THREADS>(A, a.ld, pa, pb, smem, a.vbuf + (long)slot * a.vstride + a.info_off);
    } else if (type == PT_S) {
      cholq_solve<C::THREADS, C::BM>(A, a.ld, pb, pa, pc, pd, smem);
    } else {
      cholq_update_tile<C>(A, a.ld, a.rows_total, a.np, pa * C::BM, pb * C::BN, pc * BLK, pd * BLK, __ldg(T + TK_E) * BLK,
                           __ldg(T + TK_F), smem, a.use_mb ? bars : nullptr);
    }
    if (a.stats && t == 0) {
      const long long now = clock64();
      st_cyc[type] += now - st_mark;
      st_n[type] += 1;
      st_mark = now;
      if (a.stats == 2 && blockIdx.x < QTRACE_CTAS && tr_n < QTRACE_EVENTS) {
        int* e = trace + 4 * (QTRACE_EVENTS * blockIdx.x + tr_n++);
        e[0] = q | (r << 8);
        e[1] = ti;
        e[2] = (int)(tr_start & 0x7fffffffu);
        e[3] = (int)(global_ns() & 0x7fffffffu);
      }
    }
    // every executor ends with a CTA barrier after its last global store: publish
    if (t == 0) {
      const int oi = __ldg(T + TK_OUT), ov = __ldg(T + TK_OUTVAL), o2 = __ldg(T + TK_OUT2);
      st_release(flags + oi, ov);
      if (o2 >= 0) st_release(flags + o2, ov);
    }
    if (warp == 0) {
      // a finished chain task is what the next tasks wait for: advance this problem's frontiers before claiming
      // (and a bulk tile close behind the front: the chain tasks of the next super-panels wait for its flag)
      scan_r = (q != 3 || pb * (C::BM / BLK) - pd <= 6) ? r : -1;
      // successor on the critical chain: F(k) -> S(k+1, k), S(k+1, k) -> F(k+1)
      if (a.use_hints) {
        const int g = __ldg(T + TK_G);
        if (type == PT_F && g >= 0) {
          hint_q = 1;
          hint_i = g;
          hint_r = r;
        } else if (type == PT_S && g >= 0) {
          hint_q = 0;
          hint_i = g;
          hint_r = r;
        }
      }
    }
    __syncthreads();  // s_task is rewritten by warp 0 in the next round
  }
  if (a.stats && t == 0 && blockIdx.x < QSTATS_CTAS) {
    int* o = abortf + 7 + QSTATS * blockIdx.x;
    for (int i = 0; i < 3; ++i) o[i] = st_n[i];
    for (int i = 0; i < 4; ++i) o[3 + i] = (int)(st_cyc[i] >> 6);  // units of 64 cycles
    o[7] = (int)((clock64() - st_t0) >> 6);
    o[8] = st_polls;
  }
}

// =====================================================================================================================
// Warp-specialised executor of the same plan and the same claiming rules.  In chol_queue_kernel every warp does
// everything in turn: claim (two round trips to L2), fill the pipeline, multiply, read-modify-write the tile -- a
// rank-256 tile update takes 45 us of which 33.4 us are DMMA issue.  Here a CTA has twelve warps:
//   warps 0..7   DMMA warps (224 registers): execute the tasks; for a tile update they only wait on "full" barriers,
//                multiply out of shared memory and release stages on "empty" barriers;
//   warps 9..11  copy warps: issue every cp.async of a tile update into a 4-stage ring that is numbered globally
//                ACROSS tasks, so the first slabs of the next tile land while the DMMA warps are still in the last
//                slabs and the read-modify-write of the current one;
//   warp 8       scheduler: claims the next task while the current one runs.  It is released for task n when the copy
//                warps have issued the last slab of task n - 1 (named barrier A: the DMMA warps are then <= 4 slabs +
//                epilogue from the end), and hands the task over through a two-slot mailbox (mbarrier for the DMMA
//                warps, named barrier B for the copy warps).
// Rules that keep the critical chain as fast as before: a chain-class task (F, S, front tile) is claimed ahead only
// while no CTA is idle (an idle one starts it now, this one in up to 13 us), and after a task that hands its successor
// to the same CTA (F, the critical S) the next claim waits for the task to finish -- the successor is not ready
// before.  F and S use the ring's shared memory as scratch: a tile update claimed while one of them runs is loaded
// when it has finished.
// Every wait is bounded (trap) or under the scheduler's watchdog: a protocol bug ends the kernel, it cannot hang it.
constexpr int WS_STAGES = 4;
constexpr int WS_LOADERS = 96;
template <class C>
__host__ __device__ constexpr size_t cholq_ws_ring_bytes() {
  return sizeof(double) * WS_STAGES * stage_doubles<C, Lay::KCONT, Lay::KCONT>();
}
__device__ __forceinline__ void ws_pair_sync() { asm volatile("bar.sync 2, 128;\n" ::: "memory"); }  // warps 8..11
__device__ __forceinline__ void ws_wait_done(volatile int* s_done, int need) {
  for (unsigned spin = 0; *s_done < need; ++spin) {
    __nanosleep(40);
    if (spin > (1u << 27)) __trap();
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS + 128, 1) chol_queue_ws_kernel(CholQ a) {
  static_assert(C::THREADS == 256 && C::BM == 128 && C::BN == 128, "eight DMMA warps on 128 x 128 tiles");
  extern __shared__ __align__(16) double smem[];
  // mailbox: {queue (0 F, 1 S, 2 C, 3 U; < 0: stop), task index, problem, type, A..F of the task record, slot, OUT, OUTVAL, OUT2}
  __shared__ int s_task[2][16];
  __shared__ int s_cnt[2];  // DMMA warps that have finished the tile update in mailbox slot n & 1
  __shared__ __align__(8) uint64_t bar_full[WS_STAGES], bar_empty[WS_STAGES], bar_task[2];
  __shared__ volatile int s_done;  // tasks this CTA has finished and published
  constexpr int SD = stage_doubles<C, Lay::KCONT, Lay::KCONT>();
  constexpr int AOFF = tile_doubles<C::BM, Lay::KCONT>();
  const int* __restrict__ plan = a.plan;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nflags = __ldg(plan + PH_NFLAGS);
  const int R = a.R, nq = NQ * R;
  int* qstate = a.sched + (long)R * nflags;
  int* abortf = qstate + QSTATE * R + 1;
  if (t == 0) {
    for (int s = 0; s < WS_STAGES; ++s) {
      mbar_init(&bar_full[s], WS_LOADERS);        // one cp.async-completion arrival per copy thread
      mbar_init(&bar_empty[s], C::THREADS / 32);  // one arrival per DMMA warp
    }
    mbar_init(&bar_task[0], 1);
    mbar_init(&bar_task[1], 1);
    s_done = 0;
    s_cnt[0] = s_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= C::THREADS / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    if (warp == C::THREADS / 32) {
      // ------------------------------------------------------------------------------------------ scheduler warp
      const QueueView qv = {plan, a.sched, qstate, nflags, R};
      const int rot = blockIdx.x % R;
      int* n_idle = abortf + 1;  // CTAs polling with nothing to run
      int hint_q = -1, hint_i = 0, hint_r = 0;
      int held_q = -1, held_i = 0, held_r = 0;
      int scan_r = -1, idle_r = rot;
      // What the two tasks in flight enable once they are done (A: task n - 1, B: task n - 2): their problem's frontiers
      // (a chain task, or a tile close behind the front) and, after F and the critical S, the successor on the chain.
      int pa_task = -1, pa_scan = -1, pa_hq = -1, pa_hi = 0, pa_hr = 0;
      int pb_task = -1, pb_scan = -1;
      int scan_r2 = -1;
      int n = 0;
      for (;;) {
        ws_pair_sync();  // (A) the copy warps have issued everything of task n - 1
        // mailbox slot n & 1 was task n - 2's; a task that hands its successor to this CTA (F, the critical S) is
        // waited for -- the successor is not ready before.  After any other task the next one is claimed while it runs.
        ws_wait_done(&s_done, (pa_task >= 0 && pa_hq >= 0) ? n : n - 1);
        int got_q = -1, got_i = 0, got_r = 0;
        unsigned backoff = 0;
        long long idle_since = 0;
        bool idle = false, counted_idle = false;
        for (;;) {
          const int dn = __shfl_sync(0xffffffffu, (int)s_done, 0);
          const bool ahead = dn < n;  // the DMMA warps are still in task n - 1
          if (pb_task >= 0 && dn > pb_task) {
            pb_task = -1;
            if (pb_scan >= 0) {
              if (scan_r < 0) scan_r = pb_scan;
              else if (scan_r2 < 0) scan_r2 = pb_scan;
            }
          }
          if (pa_task >= 0 && dn > pa_task) {
            pa_task = -1;
            if (pa_scan >= 0) {
              if (scan_r < 0) scan_r = pa_scan;
              else if (scan_r2 < 0) scan_r2 = pa_scan;
            }
            hint_q = pa_hq;
            hint_i = pa_hi;
            hint_r = pa_hr;
          }
          int special = 0;
          if (lane == 6 && hint_q >= 0) {
            const int* T = plan + __ldg(plan + PH_OFFF + hint_q) + hint_i * PLAN_TS;
            const int h = ld_relaxed(qstate + QSTATE * hint_r + 2 * hint_q);
            special = (task_ready(T, a.sched + (long)hint_r * nflags) && h == hint_i) ? 1 : 0;
          }
          if (lane == 7 && held_q >= 0)
            special = task_ready(plan + __ldg(plan + PH_OFFF + held_q) + held_i * PLAN_TS, a.sched + (long)held_r * nflags) ? 1 : 0;
          int best = 0x7fffffff, best_rdy = 0, best_h = 0;
          int bulk = 0x7fffffff, bulk_rdy = 0, bulk_h = 0;  // the same among the bulk queues only
          bool exhausted = held_q < 0;
          if (lane >= 8)
            for (int c = lane - 8; c < nq; c += 24) {
              const int q = c / R, r = (c % R + rot) % R;
              const int* st = qstate + QSTATE * r + 2 * q;
              const int h = ld_relaxed(st), f = ld_relaxed(st + 1);
              if (h < __ldg(plan + PH_NF + q)) exhausted = false;
              if (h < f && c < best) {
                best = c;
                best_rdy = f;
                best_h = h;
              }
              if (h < f && q == NQ - 1 && c < bulk) {
                bulk = c;
                bulk_rdy = f;
                bulk_h = h;
              }
            }
          const int others_idle = ahead ? __shfl_sync(0xffffffffu, lane == 0 ? ld_relaxed(n_idle) : 0, 0) : 0;
          const int hint_ok = __shfl_sync(0xffffffffu, special, 6), held_ok = __shfl_sync(0xffffffffu, special, 7);
          int win = best;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) win = min(win, __shfl_xor_sync(0xffffffffu, win, o));
          // This CTA is still busy: a chain-class task is claimed ahead only if no CTA is idle (an idle one starts it
          // now, this one in up to 13 us); otherwise look at the bulk queues alone.
          if (ahead && win != 0x7fffffff && win / R < NQ - 1 && others_idle > 0) {
            win = bulk;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) win = min(win, __shfl_xor_sync(0xffffffffu, win, o));
            best_rdy = bulk_rdy;
            best_h = bulk_h;
          }
          const bool all_exhausted = __all_sync(0xffffffffu, exhausted);
          if (held_q >= 0 && held_ok && (!ahead || held_q == NQ - 1 || others_idle == 0)) {
            got_q = held_q;
            got_i = held_i;
            got_r = held_r;
            held_q = -1;
            break;
          }
          if (hint_q >= 0) {  // only set once the task it follows is done
            const int hq = hint_q;
            hint_q = -1;
            int ok = 0;
            if (hint_ok) {
              if (lane == 0) {
                int* hp = qstate + QSTATE * hint_r + 2 * hq;
                ok = (atomicCAS(hp, hint_i, hint_i + 1) == hint_i);
                if (ok) atomicMax(hp + 1, hint_i + 1);
              }
              ok = __shfl_sync(0xffffffffu, ok, 0);
            }
            if (ok) {
              got_q = hq;
              got_i = hint_i;
              got_r = hint_r;
              break;
            }
          }
          if (scan_r >= 0 && (a.eager_scan || !(win != 0x7fffffff && win / R < NQ - 1))) {
            const int sr = scan_r;
            scan_r = scan_r2;
            scan_r2 = -1;
            if (scan_frontiers(qv, sr, lane)) continue;
          }
          if (win != 0x7fffffff) {
            const int q = win / R, r = (win % R + rot) % R;
            const int owner = 8 + win % 24;
            int h = 0, ok = 0;
            if (lane == owner) {
              int* hp = qstate + QSTATE * r + 2 * q;
              if (held_q < 0) {
                h = atomicAdd(hp, 1);
                ok = h < __ldg(plan + PH_NF + q) ? (h < best_rdy ? 1 : 2) : 0;
                if (ok == 2 && task_ready(plan + __ldg(plan + PH_OFFF + q) + h * PLAN_TS, a.sched + (long)r * nflags)) ok = 1;
              } else {
                h = best_h;
                ok = (atomicCAS(hp, best_h, best_h + 1) == best_h) ? 1 : 0;
              }
            }
            ok = __shfl_sync(0xffffffffu, ok, owner);
            h = __shfl_sync(0xffffffffu, h, owner);
            if (ok == 1) {
              got_q = q;
              got_i = h;
              got_r = r;
              break;
            }
            if (ok == 2) {
              held_q = q;
              held_i = h;
              held_r = r;
            }
            continue;
          }
          if (all_exhausted) {
            got_q = -1;
            break;
          }
          bool moved = false;
          for (int k = 0; k < R && !moved; ++k) {
            moved = scan_frontiers(qv, idle_r, lane);
            idle_r = (idle_r + 1) % R;
          }
          if (moved) continue;
          if (!ahead && !counted_idle) {  // nothing to claim and nothing running here: this CTA is idle
            counted_idle = true;
            if (lane == 0) atomicAdd(n_idle, 1);
          }
          int stop = 0;
          if (lane == 0) {
            const long long now = clock64();
            if (!idle || ahead) {  // the watchdog counts polling with nothing running on this CTA
              idle = true;
              idle_since = now;
            }
            if (now - idle_since > a.timeout) atomicExch(abortf, 1);
            stop = ld_relaxed(abortf);
          }
          stop = __shfl_sync(0xffffffffu, stop, 0);
          if (stop) {
            got_q = -2;
            break;
          }
          backoff = backoff < 512 ? backoff + 64 : 512;
          __nanosleep(backoff);
        }
        if (counted_idle && lane == 0) atomicSub(n_idle, 1);
        // what this task enables, applied when it is done: its problem's frontiers (chain task, or a tile close
        // behind the front) and the successor on the critical chain
        if (got_q >= 0) {
          const int* T = plan + __ldg(plan + PH_OFFF + got_q) + got_i * PLAN_TS;
          const int type = __ldg(T + TK_TYPE), g = __ldg(T + TK_G);
          if (pa_task >= 0) {  // task n - 1 is still running: keep its scan (task n - 2 is done: applied above)
            pb_task = pa_task;
            pb_scan = pa_scan;
          }
          pa_task = n;
          pa_scan = (got_q != 3 || __ldg(T + TK_B) * (C::BM / BLK) - __ldg(T + TK_D) <= 6) ? got_r : -1;
          pa_hq = -1;
          if (a.use_hints && g >= 0 && (type == PT_F || type == PT_S)) {
            pa_hq = type == PT_F ? 1 : 0;
            pa_hi = g;
            pa_hr = got_r;
          }
        }
        if (lane == 0) {
          int* m = s_task[n & 1];
          m[0] = got_q;
          m[1] = got_i;
          m[2] = got_r;
          if (got_q >= 0) {  // the task record, so that nobody else has to go to the plan for it
            const int* T = plan + __ldg(plan + PH_OFFF + got_q) + got_i * PLAN_TS;
            int f[10];
#pragma unroll
            for (int k = 0; k < 7; ++k) f[k] = __ldg(T + TK_TYPE + k);
            f[7] = __ldg(T + TK_OUT);
            f[8] = __ldg(T + TK_OUTVAL);
            f[9] = __ldg(T + TK_OUT2);
#pragma unroll
            for (int k = 0; k < 7; ++k) m[3 + k] = f[k];
            m[10] = a.slots ? __ldg(a.slots + got_r) : got_r;
            m[11] = f[7];
            m[12] = f[8];
            m[13] = f[9];
          }
        }
        __threadfence();  // acquire side of the relaxed flag loads, before this CTA touches the data they guard
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_task[n & 1]);
        ws_pair_sync();  // (B) mailbox n & 1 is valid for the copy warps
        if (got_q < 0) break;
        ++n;
      }
    } else {
      // ------------------------------------------------------------------------------------------ copy warps
      const int lt = t - (C::THREADS + 32);  // 0..95
      unsigned gslab = 0;
      int prev_type = PT_U;
      for (int n = 0;; ++n) {
        ws_pair_sync();  // (A)
        ws_pair_sync();  // (B)
        const int* m = s_task[n & 1];
        if (m[0] < 0) break;
        const int after = prev_type;
        prev_type = m[3];
        if (m[3] != PT_U) continue;
        // an F or S task uses the ring as scratch: a tile update claimed while it was still running is loaded only
        // when it has finished
        if (after != PT_U) ws_wait_done(&s_done, n);
        const double* A = a.K + (long)m[10] * a.pstride;
        const int row0 = m[4] * C::BM, col0 = m[5] * C::BN, kb = m[6] * BLK, ke = m[7] * BLK, cmin = m[8] * BLK;
        if (!a.red_epilogue) {  // start pulling the C tile into L2 for the read-modify-write at the end
          constexpr int LINES_PER_ROW = C::BN * 8 / 128;
          for (int e = lt; e < C::BM * LINES_PER_ROW; e += WS_LOADERS) {
            const int gr = row0 + e / LINES_PER_ROW, gc = col0 + (e % LINES_PER_ROW) * 16;
            if (gr < a.rows_total && gc < a.np && gc >= cmin && (gc >> 6) <= (gr >> 6))
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A + (long)gr * a.ld + gc));
          }
        }
        for (int k0 = kb; k0 < ke; k0 += BK) {
          const unsigned stage = gslab % WS_STAGES, use = gslab / WS_STAGES;
          if (use > 0) mbar_wait_bounded(&bar_empty[stage], (use - 1) & 1);  // every DMMA warp has read the old contents
          double* sA = smem + stage * SD;
          double* sB = sA + AOFF;
          for (int c = lt; c < C::BM * BK / 2; c += WS_LOADERS) {
            const int rr = c / (BK / 2), kc = (c % (BK / 2)) * 2;
            const int gr = row0 + rr, gc = col0 + rr;
            const bool oka = gr < a.rows_total, okb = gc < a.np;
            cp_async16(sA + rr * KPAD + kc, oka ? A + (long)gr * a.ld + k0 + kc : A, oka);
            cp_async16(sB + rr * KPAD + kc, okb ? A + (long)gc * a.ld + k0 + kc : A, okb);
          }
          cp_async_mbar_arrive(&bar_full[stage]);
          ++gslab;
        }
      }
      cp_async_wait<0>();
    }
    return;
  }

  // ---------------------------------------------------------------------------------------------- DMMA warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
  const Group256Sync gsync;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  long long st_cyc[4] = {0, 0, 0, 0}, st_t0 = clock64(), st_mark = st_t0;  // thread 0: F, S, U, waiting for a task
  int st_n[3] = {0, 0, 0}, tr_n = 0;
  unsigned long long tr_start = 0, tr_first = 0, tr_loop = 0, tr_epi = 0;
  int* trace = abortf + 7 + QSTATS * QSTATS_CTAS;
  unsigned gslab = 0;
  for (int n = 0;; ++n) {
    mbar_wait_bounded(&bar_task[n & 1], (n >> 1) & 1);
    const int* m = s_task[n & 1];
    const int q = m[0], ti = m[1], r = m[2];
    if (a.stats && t == 0) {
      const long long now = clock64();
      st_cyc[3] += now - st_mark;
      st_mark = now;
    }
    if (q < 0) {
      if (q == -2 && t == 0)
        for (int rr = 0; rr < R; ++rr) a.vbuf[(a.slots ? (long)a.slots[rr] : (long)rr) * a.vstride + a.info_off] = -7777.0;
      break;
    }
    if (a.stats == 2 && t == 0) tr_start = global_ns();
    const int type = m[3], pa = m[4], pb = m[5], pc = m[6], pd = m[7], pe = m[8], pf = m[9], slot = m[10];
    const int oi = m[11], ov = m[12], o2 = m[13];
    double* A = a.K + (long)slot * a.pstride;
    int* flags = a.sched + (long)r * nflags;
    // F and S use the ring as scratch and are written for the eight warps together; after a tile update the warps
    // arrive here one by one
    if (type != PT_U) gsync();
    if (type == PT_F) {
      cholq_factor<C::THREADS, Group256Sync>(A, a.ld, pa, pb, smem, a.vbuf + (long)slot * a.vstride + a.info_off, gsync);
    } else if (type == PT_S) {
      cholq_solve<C::THREADS, C::BM, Group256Sync>(A, a.ld, pb, pa, pc, pd, smem, gsync);
    } else {
      double acc[C::MI][C::NI][2];
      zero_acc<C>(acc);
      for (int k0 = pc * BLK; k0 < pd * BLK; k0 += BK) {
        const unsigned stage = gslab % WS_STAGES, use = gslab / WS_STAGES;
        mbar_wait_bounded(&bar_full[stage], use & 1);
        if (a.stats == 2 && t == 0 && k0 == pc * BLK) tr_first = global_ns();
        const double* cur = smem + stage * SD;
        mma_slab<C, Lay::KCONT, Lay::KCONT>(acc, cur, cur + AOFF, wm, wn, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[stage]);
        ++gslab;
      }
      if (a.stats == 2 && t == 0) tr_loop = global_ns();
      if (a.red_epilogue)
        cholq_tile_epilogue_red<C>(A, a.ld, a.rows_total, a.np, pa * C::BM, pb * C::BN, pe * BLK, pf, acc);
      else
        cholq_tile_epilogue<C>(A, a.ld, a.rows_total, a.np, pa * C::BM, pb * C::BN, pe * BLK, pf, acc);
      if (a.stats == 2 && t == 0) tr_epi = global_ns();
      // No barrier at the end of a tile update: every warp counts itself out and goes on to the next task (its ring
      // accesses are ordered by the stage barriers); the last one out publishes the tile.  Its release store is
      // cumulative over the other warps' updates, which it has observed through the counter (fence - atomic - fence
      // at CTA scope), exactly as thread 0's was over the barrier.
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&s_cnt[n & 1], 1) == C::THREADS / 32 - 1) {
          s_cnt[n & 1] = 0;
          __threadfence_block();
          st_release(flags + oi, ov);
          if (o2 >= 0) st_release(flags + o2, ov);
          s_done = n + 1;
        }
      }
    }
    if (a.stats && t == 0) {
      const long long now = clock64();
      st_cyc[type] += now - st_mark;
      st_n[type] += 1;
      st_mark = now;
      if (a.stats == 2 && blockIdx.x < QTRACE_CTAS && tr_n < QTRACE_EVENTS) {
        int* e = trace + 4 * (QTRACE_EVENTS * blockIdx.x + tr_n++);
        e[0] = q | (r << 8);
        // tile updates: where the time went inside the task, in units of 64 ns -- first slab landed, main loop done,
        // stores issued -- in place of the task index (measurement aid)
        e[1] = type == PT_U ? (int)(((tr_first - tr_start) >> 6 & 0x3ff) | ((tr_loop - tr_start) >> 6 & 0x3ff) << 10 |
                                    ((tr_epi - tr_start) >> 6 & 0x3ff) << 20)
                            : ti;
        e[2] = (int)(tr_start & 0x7fffffffu);
        e[3] = (int)(global_ns() & 0x7fffffffu);
      }
    }
    if (t == 0 && type != PT_U) {  // F and S end with a barrier of the DMMA warps after their last global store: publish
      st_release(flags + oi, ov);
      if (o2 >= 0) st_release(flags + o2, ov);
      s_done = n + 1;
    }
  }
  if (a.stats && t == 0 && blockIdx.x < QSTATS_CTAS) {
    int* o = abortf + 7 + QSTATS * blockIdx.x;
    for (int i = 0; i < 3; ++i) o[i] = st_n[i];
    for (int i = 0; i < 4; ++i) o[3 + i] = (int)(st_cyc[i] >> 6);
    o[7] = (int)((clock64() - st_t0) >> 6);
    o[8] = 0;
  }
}

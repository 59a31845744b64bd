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
  int* sched;         // [R][nflags] flags | [R][4] queue heads (F, S, C, -) | bulk head | abort | stats...
  long long timeout;  // cycles a CTA may poll without finding work before it raises `abort`
  int use_hints;
};

__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
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
template <int NT>
__device__ __forceinline__ void cholq_factor(double* __restrict__ A, long ld, int k, int width, double* psm, double* info) {
  double* D = psm;
  double* P = psm + BLK * PSTR;
  const int t = threadIdx.x, j0 = BLK * k, p0 = BLK * (k - width), pw = BLK * width;
  for (int e = t; e < BLK * BLK / 2; e += NT) {
    const int r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
    const double2 v = __ldcg(reinterpret_cast<const double2*>(A + (long)(j0 + r) * ld + j0 + cc));
    D[r * PSTR + cc] = v.x;
    D[r * PSTR + cc + 1] = v.y;
  }
  for (int e = t; e < BLK * pw / 2; e += NT) {
    const int r = e / (pw / 2), cc = (e % (pw / 2)) * 2;
    const double2 v = __ldcg(reinterpret_cast<const double2*>(A + (long)(j0 + r) * ld + p0 + cc));
    *reinterpret_cast<double2*>(P + r * QSTR2 + cc) = v;
  }
  __syncthreads();
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
    __syncthreads();
  }
  potf2_block<NT>(D, t, j0, info);
  for (int e = t; e < BLK * BLK / 2; e += NT) {
    const int r = e / (BLK / 2), cc = (e % (BLK / 2)) * 2;
    *reinterpret_cast<double2*>(A + (long)(j0 + r) * ld + j0 + cc) =
        make_double2(cc <= r ? D[r * PSTR + cc] : 0.0, cc + 1 <= r ? D[r * PSTR + cc + 1] : 0.0);
  }
  __syncthreads();
}

// U task: tile (row0, col0) -= A(rows, [kb, ke)) A(cols, [kb, ke))^T over the lower part, columns >= cmin, except the
// diagonal block `skip` (block index; that one is its F task's own).
template <class C>
__device__ __forceinline__ void cholq_update_tile(double* __restrict__ A, long ld, int rows_total, int np, int row0, int col0,
                                                  int kb, int ke, int cmin, int skip, double* smem) {
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
  gemm_mainloop<C, Lay::KCONT, Lay::KCONT>(acc, smem, kb, ke, fillA, fillB);
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
  __syncthreads();
}

enum { QS_NONE = -1, QS_EXHAUSTED = -2 };

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) chol_queue_kernel(CholQ a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task[4];  // {queue (0 F, 1 S, 2 C, 3 U; < 0: stop), task index, problem, -}
  const int* __restrict__ plan = a.plan;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nflags = __ldg(plan + PH_NFLAGS);
  const int R = a.R;
  int* heads = a.sched + (long)R * nflags;
  int* bulk_head = heads + 4 * R;
  int* abortf = bulk_head + 1;
  int pend = QS_NONE;                  // this CTA's claimed, not yet started bulk task (warp 0, uniform)
  int hint_q = -1, hint_i = 0, hint_r = 0;  // successor of the task just finished (tried first)
  long long idle_since = 0;
  bool idle = false;

  for (;;) {
    if (warp == 0) {
      int got_q = -1, got_i = 0, got_r = 0;
      for (;;) {
        // ---- 0. the successor of the task this CTA just finished
        if (hint_q >= 0) {
          int ok = 0;
          if (lane == 0) {
            int* hp = heads + 4 * hint_r + hint_q;
            const int* T = plan + __ldg(plan + PH_OFFF + hint_q) + hint_i * PLAN_TS;
            const int h = ld_relaxed(hp);
            const bool rdy = task_ready(T, a.sched + (long)hint_r * nflags);
            if (h == hint_i && rdy) ok = (atomicCAS(hp, hint_i, hint_i + 1) == hint_i);
          }
          ok = __shfl_sync(0xffffffffu, ok, 0);
          const int hq = hint_q;
          hint_q = -1;
          if (ok) {
            got_q = hq;
            got_i = hint_i;
            got_r = hint_r;
            break;
          }
        }
        // ---- 1. lanes 0..30: heads of the chain queues (candidate c = queue * R + problem, lowest c wins);
        //         lane 31: the bulk queue (claim the next task if this CTA holds none, then test it)
        int best = 0x7fffffff, best_h = 0;
        bool exhausted = true;
        if (lane < 31) {
          for (int c = lane; c < 3 * R; c += 31) {
            const int q = c / R, r = c % R;
            const int h = ld_relaxed(heads + 4 * r + q);
            if (h < __ldg(plan + PH_NF + q)) {
              exhausted = false;
              const int* T = plan + __ldg(plan + PH_OFFF + q) + h * PLAN_TS;
              if (c < best && task_ready(T, a.sched + (long)r * nflags)) {
                best = c;
                best_h = h;
              }
            }
          }
        } else {
          if (pend == QS_NONE) {
            const int total = __ldg(plan + PH_NU) * R;
            const int pos = (ld_relaxed(bulk_head) < total) ? atomicAdd(bulk_head, 1) : total;
            pend = pos < total ? pos : QS_EXHAUSTED;
          }
          if (pend >= 0) {
            exhausted = false;
            const int ti = pend / R, r = pend % R;
            const int* T = plan + __ldg(plan + PH_OFFU) + ti * PLAN_TS;
            if (task_ready(T, a.sched + (long)r * nflags)) {
              best = 3 * R;
              best_h = pend;
            }
          }
        }
        pend = __shfl_sync(0xffffffffu, pend, 31);
        int win = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) win = min(win, __shfl_xor_sync(0xffffffffu, win, o));
        const bool all_done = __all_sync(0xffffffffu, exhausted);
        if (win != 0x7fffffff) {
          const unsigned owner_mask = __ballot_sync(0xffffffffu, best == win);
          const int owner = __ffs(owner_mask) - 1;
          int ok = 0;
          if (win == 3 * R) {
            ok = 1;
          } else if (lane == owner) {
            const int q = win / R, r = win % R;
            ok = (atomicCAS(heads + 4 * r + q, best_h, best_h + 1) == best_h);
          }
          ok = __shfl_sync(0xffffffffu, ok, owner);
          const int h = __shfl_sync(0xffffffffu, best_h, owner);
          if (ok) {
            if (win == 3 * R) {
              got_q = 3;
              got_i = h / R;
              got_r = h % R;
              pend = QS_NONE;
            } else {
              got_q = win / R;
              got_i = h;
              got_r = win % R;
            }
            break;
          }
          continue;  // lost the race: look again at once
        }
        if (all_done) {
          got_q = -1;
          break;
        }
        // ---- nothing ready: watchdog, then poll again
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
        __nanosleep(32);
      }
      idle = false;
      if (lane == 0) {
        s_task[0] = got_q;
        s_task[1] = got_i;
        s_task[2] = got_r;
      }
      __threadfence();  // acquire side of the flags just read (relaxed): order them before this CTA's data loads
    }
    __syncthreads();
    const int q = s_task[0], ti = s_task[1], r = s_task[2];
    if (q < 0) {
      if (q == -2 && t == 0)  // watchdog: report through every problem's info slot (the host raises)
        for (int rr = 0; rr < R; ++rr) a.vbuf[(a.slots ? (long)a.slots[rr] : (long)rr) * a.vstride + a.info_off] = -7777.0;
      break;
    }
    const int* T = plan + __ldg(plan + PH_OFFF + q) + ti * PLAN_TS;
    const int slot = a.slots ? __ldg(a.slots + r) : r;
    double* A = a.K + (long)slot * a.pstride;
    int* flags = a.sched + (long)r * nflags;
    const int type = __ldg(T + TK_TYPE), pa = __ldg(T + TK_A), pb = __ldg(T + TK_B), pc = __ldg(T + TK_C), pd = __ldg(T + TK_D);
    if (type == PT_F) {
      cholq_factor<C::THREADS>(A, a.ld, pa, pb, smem, a.vbuf + (long)slot * a.vstride + a.info_off);
    } else if (type == PT_S) {
      if (pd == 64)
        chol_trsm_block<C::THREADS, 64>(A, a.ld, pb, pa, pc, smem);
      else if (pd == 32)
        chol_trsm_block<C::THREADS, 32>(A, a.ld, pb, pa, pc, smem);
      else
        chol_trsm_block<C::THREADS, 16>(A, a.ld, pb, pa, pc, smem);
    } else {
      cholq_update_tile<C>(A, a.ld, a.rows_total, a.np, pa * C::BM, pb * C::BN, pc * BLK, pd * BLK, __ldg(T + TK_E) * BLK,
                           __ldg(T + TK_F), smem);
    }
    // every executor ends with a CTA barrier after its last global store: publish
    if (t == 0) {
      const int oi = __ldg(T + TK_OUT), ov = __ldg(T + TK_OUTVAL);
      if (ov < 0)
        red_add_release(flags + oi, 1);
      else
        st_release(flags + oi, ov);
    }
    // successor on the critical chain: F(k) -> S(k+1, k) (first solve task of panel k), S(k+1, k) -> F(k+1)
    if (a.use_hints && warp == 0) {
      if (type == PT_F) {
        hint_q = 1;
        hint_i = __ldg(T + TK_G);  // index of S(k+1, k, 0) in the S queue (-1: none)
        hint_r = r;
        if (hint_i < 0) hint_q = -1;
      } else if (type == PT_S && pa == pb + 1 && __ldg(T + TK_G) >= 0) {
        hint_q = 0;
        hint_i = __ldg(T + TK_G);  // F(k+1), set on the last part of S(k+1, k)
        hint_r = r;
      }
    }
    __syncthreads();  // s_task is rewritten by warp 0 in the next round
  }
}

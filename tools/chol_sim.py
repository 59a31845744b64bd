"""Event-driven performance model of the dataflow Cholesky's scheduler (csrc/ipp_cholq.cuh) on the host: the real task
plan (ipp_host_chol_plan), measured task durations (profiles/chol_queue_timeline_*), in-order queues per problem with
priority F, S, C, U, and a claim latency per task.  Used to compare scheduling policies before spending GPU time:

    python tools/chol_sim.py n R [policy] [u256_us] [claim_us]

policy: rot   -- every CTA prefers problem (cta mod R), then the next ones (what the kernel does)
        prio  -- every CTA prefers the lowest-numbered problem with a ready task (problems desynchronise)
        stag  -- rot, but problem r may not start before r * stagger_us
"""
import ctypes as C
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mripp_b200 import _lib  # noqa: E402

TS, MAXDEP = 24, 6
PH_NFLAGS, PH_NF, PH_OFFF = 4, 5, 9
TK_TYPE, TK_A, TK_B, TK_C, TK_D, TK_DEP0 = 0, 1, 2, 3, 4, 9
TK_OUT, TK_OUTVAL, TK_OUT2 = TK_DEP0 + 2 * MAXDEP, TK_DEP0 + 2 * MAXDEP + 1, TK_DEP0 + 2 * MAXDEP + 2


def build_plan(n, near=0):
    lib = _lib.load()
    need = C.c_longlong(0)
    assert lib.ipp_host_chol_plan(n, near, None, 0, C.byref(need)) == 0
    buf = (C.c_int * need.value)()
    assert lib.ipp_host_chol_plan(n, near, buf, need.value, C.byref(need)) == 0
    return np.frombuffer(buf, dtype=np.int32).copy()


def simulate(n, R, policy="rot", u256=45.4, claim=3.0, n_ctas=148, stagger=0.0, quiet=False, u_base=None, prio_chain=True):
    plan = build_plan(n)
    cnt = [int(plan[PH_NF + q]) for q in range(4)]
    off = [int(plan[PH_OFFF + q]) for q in range(4)]
    nflags = int(plan[PH_NFLAGS])
    # per queue: deps (idx, val) lists, outputs, duration
    deps, outs, dur = [], [], []
    per128 = (u256 - (u_base if u_base is not None else 12.0)) / 2.0
    base = u256 - 2 * per128
    for q in range(4):
        dq, oq, tq = [], [], []
        for i in range(cnt[q]):
            T = plan[off[q] + i * TS: off[q] + (i + 1) * TS]
            dq.append([(int(T[TK_DEP0 + 2 * d]), int(T[TK_DEP0 + 2 * d + 1])) for d in range(MAXDEP) if T[TK_DEP0 + 2 * d] >= 0])
            oq.append(([int(T[TK_OUT])] + ([int(T[TK_OUT2])] if T[TK_OUT2] >= 0 else []), int(T[TK_OUTVAL])))
            ty = int(T[TK_TYPE])
            if ty == 0:
                t = 19.9
            elif ty == 1:
                t = 12.4 if int(T[TK_C]) <= 1 else 18.0
            else:
                ranks = (int(T[TK_D]) - int(T[TK_C])) / 2.0  # 64-blocks -> units of 128
                t = base + per128 * ranks
            tq.append(t)
        deps.append(dq)
        outs.append(oq)
        dur.append(tq)
    flags = np.zeros((R, nflags), dtype=np.int64)
    heads = np.zeros((R, 4), dtype=np.int64)
    start_at = [r * stagger if policy == "stag" else 0.0 for r in range(R)]

    def ready(r, q, i):
        fl = flags[r]
        for idx, val in deps[q][i]:
            if fl[idx] < val:
                return False
        return True

    def pick(cta, now):
        rot = 0 if policy == "prio" else cta % R
        order = [(q, (j + rot) % R) for q in range(4) for j in range(R)]
        if policy == "prio" and not prio_chain:
            order = [(q, r) for r in range(R) for q in range(4)]
        for q, r in order:
            if now < start_at[r]:
                continue
            h = heads[r, q]
            if h < cnt[q] and ready(r, q, h):
                heads[r, q] += 1
                return q, r, int(h)
        return None

    events = []  # (time, seq, cta, q, r, i)
    idle = list(range(n_ctas))
    busy_time = np.zeros(4)
    now, seq, done, total = 0.0, 0, 0, R * sum(cnt)
    timeline = []
    last_finish = [0.0] * R
    while done < total:
        # hand work to idle CTAs
        still = []
        for c in idle:
            got = pick(c, now)
            if got is None:
                still.append(c)
                continue
            q, r, i = got
            d = dur[q][i] + claim
            busy_time[q] += d
            seq += 1
            heapq.heappush(events, (now + d, seq, c, q, r, i))
            timeline.append((now, now + d, q))
        idle = still
        if not events:
            nxt = min((s for s in start_at if s > now), default=None)
            assert nxt is not None, "deadlock in the model"
            now = nxt
            continue
        t, _, c, q, r, i = heapq.heappop(events)
        now = t
        oi, ov = outs[q][i]
        for o in oi:
            flags[r, o] = max(flags[r, o], ov)
        done += 1
        last_finish[r] = now
        idle.append(c)
        # everything finishing at the same instant
        while events and events[0][0] <= now:
            t, _, c, q, r, i = heapq.heappop(events)
            oi, ov = outs[q][i]
            for o in oi:
                flags[r, o] = max(flags[r, o], ov)
            done += 1
            last_finish[r] = now
            idle.append(c)
    util = busy_time.sum() / (now * n_ctas)
    if not quiet:
        print(f"n={n} R={R} policy={policy} u256={u256} claim={claim}: {now / 1e3:.3f} ms total, {now / 1e3 / R:.3f} ms/problem, "
              f"utilisation {util:.3f}; problems finish at {[round(x / 1e3, 2) for x in last_finish]}")
        tl = np.array(timeline)
        bucket = now / 40
        line = []
        for k in range(40):
            lo, hi = k * bucket, (k + 1) * bucket
            ov = np.clip(np.minimum(tl[:, 1], hi) - np.maximum(tl[:, 0], lo), 0, None).sum()
            line.append(ov / (bucket * n_ctas))
        print("utilisation by 1/40 of the run:", " ".join(f"{x:.2f}" for x in line))
    return now, util


if __name__ == "__main__":
    n, R = int(sys.argv[1]), int(sys.argv[2])
    policy = sys.argv[3] if len(sys.argv) > 3 else "rot"
    u256 = float(sys.argv[4]) if len(sys.argv) > 4 else 45.4
    claim = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
    stagger = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    simulate(n, R, policy, u256, claim, stagger=stagger)

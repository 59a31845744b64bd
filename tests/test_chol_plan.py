"""Host-side check of the dataflow Cholesky's task plan (csrc/ipp_chol_plan.h, executed on the device by
chol_queue_kernel): a discrete-event simulation of the kernel's own claiming rules -- chain queues with a ready
frontier, claimed by atomicAdd on a stale view (so that racing claims overshoot the frontier and are HELD), the bulk
queue claimed in order unconditionally, successor hints -- over random CTA counts, task durations and polling
orders, once with the rules of chol_queue_kernel and once with those of chol_queue_ws_kernel on top (the next task
claimed while a tile update is still running; chain-class tasks only while no CTA is idle).  It checks what the device
cannot tell us cheaply:

* no deadlock for any number of resident CTAs (1 included);
* every block receives every panel to its left exactly once, in ascending order, before it is factored / solved;
* every operand block a task reads is final when the task starts, and no two running tasks touch the same block.
"""
import ctypes as C
import random

import numpy as np
import pytest

from mripp_b200 import _lib

HDR, TS, MAXDEP = 16, 24, 6
(PH_MAGIC, PH_NB, PH_NRB, PH_TB, PH_NFLAGS, PH_NF, PH_NS, PH_NC, PH_NU, PH_OFFF, PH_OFFS, PH_OFFC, PH_OFFU, PH_TS, PH_TOTAL,
 PH_NTILE) = range(16)
TK_TYPE, TK_A, TK_B, TK_C, TK_D, TK_E, TK_F, TK_G, TK_NDEP, TK_DEP0 = range(10)
TK_OUT, TK_OUTVAL, TK_OUT2 = TK_DEP0 + 2 * MAXDEP, TK_DEP0 + 2 * MAXDEP + 1, TK_DEP0 + 2 * MAXDEP + 2


def build_plan(n, near):
    lib = _lib.load()
    need = C.c_longlong(0)
    assert lib.ipp_host_chol_plan(n, near, None, 0, C.byref(need)) == 0
    buf = (C.c_int * need.value)()
    assert lib.ipp_host_chol_plan(n, near, buf, need.value, C.byref(need)) == 0
    return np.frombuffer(buf, dtype=np.int32).copy()


class Sim:
    def __init__(self, plan, R, n_ctas, seed, hints=True, ws=False):
        self.p, self.R, self.rng, self.hints = plan, R, random.Random(seed), hints
        # ws: the rules of chol_queue_ws_kernel on top -- a CTA in a tile update claims its NEXT task before the current
        # one is finished (queue 3 freely; a chain-class task only while no CTA is idle), except during a task that hands
        # its successor to the same CTA
        self.ws, self.idle_now = ws, 0
        self.nb, self.nrb, self.tb = int(plan[PH_NB]), int(plan[PH_NRB]), int(plan[PH_TB])
        self.cnt = [int(plan[PH_NF + q]) for q in range(4)]
        self.off = [int(plan[PH_OFFF + q]) for q in range(4)]
        self.flags = np.zeros((R, int(plan[PH_NFLAGS])), dtype=np.int64)
        self.heads = np.zeros((R, 4), dtype=np.int64)
        self.front = np.zeros((R, 4), dtype=np.int64)  # ready frontiers
        self.window = self.rng.choice((1, 2, 20))
        # per problem and block: panels applied so far, final?, solve parts done; blocks being written right now
        self.applied = [dict() for _ in range(R)]
        self.final = [set() for _ in range(R)]
        self.parts = [dict() for _ in range(R)]
        self.writing = [dict() for _ in range(R)]
        self.ctas = [dict(hint=None, busy_until=0, task=None, blocked=None, scan=None, rot=c % R, idle_r=c % R, next=None)
                     for c in range(n_ctas)]
        self.executed = 0

    def task(self, q, i):
        o = self.off[q] + i * TS
        return self.p[o:o + TS]

    def ready(self, T, r):
        for d in range(MAXDEP):
            idx = int(T[TK_DEP0 + 2 * d])
            if idx >= 0 and self.flags[r, idx] < int(T[TK_DEP0 + 2 * d + 1]):
                return False
        return True

    # ---- what a task reads and writes, at block level -------------------------------------------------
    def blocks_of(self, T):
        ty = int(T[TK_TYPE])
        if ty == 0:
            k, w = int(T[TK_A]), int(T[TK_B])
            return dict(kind="F", writes=[(k, k)], reads=[(k, p) for p in range(k - w, k)], before=k - w, panels=range(k - w, k))
        if ty == 1:
            i, k, nblk, w = int(T[TK_A]), int(T[TK_B]), int(T[TK_C]), int(T[TK_D])
            assert nblk >= 1 and i + nblk <= self.nrb and w == (k & 1)
            reads = [(k, k)] + [(k, p) for p in range(k - w, k)] + [(i + b, p) for b in range(nblk) for p in range(k - w, k)]
            return dict(kind="S", writes=[(i + b, k) for b in range(nblk)], reads=reads, before=k - w, panels=range(k - w, k))
        ti, tj, kb, ke, cmin, skip = (int(T[x]) for x in (TK_A, TK_B, TK_C, TK_D, TK_E, TK_F))
        writes, reads = [], set()
        for i in range(ti * self.tb, min((ti + 1) * self.tb, self.nrb)):
            for j in range(tj * self.tb, min((tj + 1) * self.tb, self.nb)):
                if j < cmin or j > i or (i == j == skip):
                    continue
                assert j >= ke, "a panel may only update columns to its right"
                writes.append((i, j))
                for p in range(kb, ke):
                    reads.add((i, p))
                    reads.add((j, p))
        assert writes, "empty update task"
        return dict(kind="U", writes=writes, reads=sorted(reads), before=kb, panels=range(kb, ke))

    def start(self, cta, q, i, r, now):
        T = self.task(q, i)
        assert self.ready(T, r)
        b = self.blocks_of(T)
        for blk in b["reads"]:
            assert blk in self.final[r], (b["kind"], "reads a block that is not final", blk)
            assert blk not in self.writing[r]
        for blk in b["writes"]:
            assert blk not in self.writing[r], (b["kind"], "two tasks write", blk)
            self.writing[r][blk] = 1
            assert blk not in self.final[r]
            assert self.applied[r].get(blk, 0) == b["before"], (b["kind"], blk, self.applied[r].get(blk, 0), b["before"])
        cta["task"] = (T, r, b)
        cta["busy_until"] = now + self.rng.choice((1, 1, 2, 3, 5, 9))

    def finish(self, cta):
        T, r, b = cta["task"]
        for blk in b["writes"]:
            del self.writing[r][blk]
            self.applied[r][blk] = self.applied[r].get(blk, 0) + len(b["panels"])
            if b["kind"] in ("F", "S"):
                assert self.applied[r][blk] == blk[1]
                self.final[r].add(blk)
        ov = int(T[TK_OUTVAL])
        for oi in (int(T[TK_OUT]), int(T[TK_OUT2])):
            if oi >= 0:
                assert self.flags[r, oi] < ov
                self.flags[r, oi] = ov
        ty = int(T[TK_TYPE])
        cta["hint"] = None
        if self.hints and ty in (0, 1) and int(T[TK_G]) >= 0:
            cta["hint"] = (1 - ty, int(T[TK_G]), r)
        cta["scan"] = r if ty != 2 or self.rng.random() < 0.3 else None  # chain finishers advance their problem's frontiers
        cta["task"] = None
        self.executed += 1

    def scan(self, r, front0=None):
        moved = False
        for q in range(4):
            f = int(self.front[r, q] if front0 is None else front0[r, q])
            m = 0
            while m < self.window and f + m < self.cnt[q] and self.ready(self.task(q, f + m), r):
                m += 1
            if m and f + m > self.front[r, q]:
                self.front[r, q] = f + m
                moved = True
        return moved

    def launch(self, cta, q, i, r, now, ahead):
        if ahead:
            assert self.ready(self.task(q, i), r)   # a task handed over early is complete in its dependencies
            cta["next"] = (q, i, r)
        else:
            self.start(cta, q, i, r, now)

    def acquire(self, cta, now, snap, ahead=False):
        """One polling round of one CTA against the queue state `snap` = (heads, frontiers) as it was at the start of
        the tick (every CTA of a tick decides on the same stale view, like concurrent pollers on the device): frontier
        scan if the last task asked for one, held task, hint, then an atomicAdd claim from the first queue (priority
        F, S, C, U; problems rotated by CTA) whose head is behind its frontier -- landing beyond the frontier the
        task is HELD -- else a scan of one problem after the other.  Returns 'exit', 'idle' or 'started'."""
        heads0, front0 = snap
        chain_ok = (not ahead) or self.idle_now == 0    # busy CTA: chain-class tasks only while nobody is idle
        if cta["scan"] is not None and not ahead:
            self.scan(cta["scan"])
            cta["scan"] = None
            front0 = self.front.copy()
        if cta["blocked"] is not None:  # a task claimed past the frontier: held, tested first, the CTA keeps serving
            q, h, r = cta["blocked"]
            if self.ready(self.task(q, h), r) and (q == 3 or chain_ok):
                cta["blocked"] = None
                self.launch(cta, q, h, r, now, ahead)
                return "started"
        if cta["hint"] is not None and not ahead:
            q, i, r = cta["hint"]
            cta["hint"] = None
            if self.heads[r, q] == i and self.ready(self.task(q, i), r):
                self.heads[r, q] += 1
                self.front[r, q] = max(self.front[r, q], i + 1)
                self.start(cta, q, i, r, now)
                return "started"
        exhausted = cta["blocked"] is None
        rot = cta["rot"]
        for c in range(4 * self.R):
            q, j = divmod(c, self.R)
            r = (j + rot) % self.R
            if heads0[r, q] < self.cnt[q]:
                exhausted = False
        for c in range(4 * self.R):
            q, j = divmod(c, self.R)
            r = (j + rot) % self.R
            if heads0[r, q] < front0[r, q]:
                if q != 3 and not chain_ok:
                    continue
                if cta["blocked"] is None:
                    h = int(self.heads[r, q])  # atomicAdd on the live head
                    self.heads[r, q] += 1
                    if h >= self.cnt[q]:
                        return "idle"
                    if h < front0[r, q] or self.ready(self.task(q, h), r):
                        self.launch(cta, q, h, r, now, ahead)
                        return "started"
                    cta["blocked"] = (q, h, r)
                    return "idle"
                h = int(heads0[r, q])  # holding: compare-and-swap on exactly the head seen
                if self.heads[r, q] == h:
                    self.heads[r, q] += 1
                    self.launch(cta, q, h, r, now, ahead)
                    return "started"
                return "idle"
        if ahead:
            return "idle"
        if exhausted:
            return "exit"
        for _ in range(self.R):
            moved = self.scan(cta["idle_r"], front0)
            cta["idle_r"] = (cta["idle_r"] + 1) % self.R
            if moved:
                break
        return "idle"

    def run(self):
        now, live = 0, list(range(len(self.ctas)))
        stall = 0
        while live:
            progressed = False
            self.rng.shuffle(live)
            snap = (self.heads.copy(), self.front.copy())
            self.idle_now = sum(1 for c in live if self.ctas[c]["task"] is None)
            for c in list(live):
                cta = self.ctas[c]
                if cta["task"] is not None:
                    if cta["busy_until"] <= now:
                        self.finish(cta)
                        progressed = True
                        if cta["next"] is not None:      # the task claimed ahead starts at once
                            q, i, r = cta["next"]
                            cta["next"] = None
                            self.start(cta, q, i, r, now)
                            continue
                    else:
                        T_cur = cta["task"][0]
                        hands_over = self.hints and int(T_cur[TK_TYPE]) in (0, 1) and int(T_cur[TK_G]) >= 0
                        if self.ws and cta["next"] is None and not hands_over and self.rng.random() < 0.5:
                            before = (self.heads.sum(), self.front.sum())
                            if self.acquire(cta, now, snap, ahead=True) == "started" or before != (self.heads.sum(), self.front.sum()):
                                progressed = True
                        continue
                before = (self.heads.sum(), self.front.sum())
                res = self.acquire(cta, now, snap)
                if res == "exit":
                    live.remove(c)
                    progressed = True
                elif res == "started" or before != (self.heads.sum(), self.front.sum()):
                    progressed = True
            running = any(self.ctas[c]["task"] is not None for c in live)
            stall = 0 if (progressed or running) else stall + 1
            assert stall < 3, "deadlock: every resident CTA polls or waits and nothing runs"
            now += 1
        total = self.R * sum(self.cnt)
        assert self.executed == total, (self.executed, total)
        for r in range(self.R):
            for i in range(self.nrb):
                for j in range(min(i + 1, self.nb)):
                    assert (i, j) in self.final[r], ("block never finalised", i, j)
                    assert self.applied[r].get((i, j), 0) == j, ("panels applied", i, j, self.applied[r].get((i, j), 0))


@pytest.mark.parametrize("n,near,R,n_sim", [
    (64, 0, 1, 1), (100, 2, 2, 3), (300, 0, 1, 7), (500, 1, 3, 5), (700, 0, 1, 1), (1000, 2, 2, 40),
    (1600, 0, 1, 1), (1700, 3, 1, 16), (2000, 0, 2, 33), (2100, 1, 3, 148),
])
def test_plan_is_deadlock_free_and_complete(n, near, R, n_sim):
    plan = build_plan(n, near)
    assert plan[PH_MAGIC] == 0x43485132 and plan[PH_TS] == TS and plan[PH_TOTAL] == len(plan)
    for seed in range(2):
        Sim(plan, R, n_sim, seed, hints=bool(seed % 2 == 0)).run()
        Sim(plan, R, n_sim, seed + 7, hints=bool(seed % 2 == 0), ws=True).run()


def test_plan_shape_at_c4():
    """The C4 size (n = 5000): tile mode, queue sizes, scratch size reported by the library."""
    lib = _lib.load()
    plan = build_plan(5000, 0)
    nb = 5056 // 64
    assert plan[PH_NB] == nb and plan[PH_NRB] == nb + 1 and plan[PH_TB] == 2 and plan[PH_NF] == nb
    out = C.c_longlong(0)
    assert lib.ipp_chol_sched_ints(5000, 11, C.byref(out)) == 0
    assert out.value == 11 * int(plan[PH_NFLAGS]) + 8 * 11 + 8 + 16 * 512 + 4 * 448 * 160
    # every dependency index is a valid flag
    for q in range(4):
        for i in range(int(plan[PH_NF + q])):
            T = plan[int(plan[PH_OFFF + q]) + i * TS:][:TS]
            assert 0 <= T[TK_OUT] < plan[PH_NFLAGS]
            for d in range(int(T[TK_NDEP])):
                assert 0 <= T[TK_DEP0 + 2 * d] < plan[PH_NFLAGS] and T[TK_DEP0 + 2 * d + 1] > 0

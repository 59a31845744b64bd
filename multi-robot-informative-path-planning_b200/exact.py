"""Decision-exact tie handling for the selectors.

The device scores every candidate (scoring.py); the host only decides WHICH candidate wins.  Two
properties of the reference make that decision sensitive to the last bit of a score:

  * When the agent has no observations yet, `mutual_information_gain` returns the int 0, so the
    reference's score array is int64 and every `-=` truncates toward zero (rrt.py:457-482, 521-522):
    many leaves collapse onto the same integer and np.argmax takes the first.  `integer_scores`
    reproduces that from the device's float scores.
  * RRT edges have the nominal length d_wp (or 1.0 for RRT*), so the distance penalties of most
    leaves agree to within a few ulp and np.argmax is decided by rounding noise of numpy's own
    exp / det.  When the device's top candidates are closer than `REL_TIE` (but not bit-identical),
    only THOSE candidates are re-scored on the host with the reference's exact numpy expressions
    (`exact_leaf_scores`, `exact_branch_gain`) and the first maximum among them wins.

This is selection logic on a handful of candidates, not a compute fallback: every score the
strategies consume still comes from the device.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
from scipy.spatial.distance import cdist

import threading

_norm = np.linalg.norm
# Two places draw uniforms from the global np.random stream ahead of the decision that consumes them:
# TreeArrays.grow_native pre-draws a block of samples and puts the stream back to where the reference's loop would
# have left it; estimation.resample_indices draws the uniforms of RandomState.choice itself.  They hold this lock for
# that span.  The mirror runs its agents one after the other (a legal interleaving of the reference's threads), so
# nothing else draws from the global stream meanwhile; callers that drive several strategies from their own threads
# must not share the global RNG across them (the reference has the same nondeterminism there).
RNG_SECTION = threading.RLock()
REL_TIE = 1e-9
INT64_MIN = np.iinfo(np.int64).min


def integer_scores(scores: np.ndarray) -> np.ndarray:
    """int64 view of `0 - dist - workspace - 0` with numpy's cast rules: truncation toward zero,
    -inf -> INT64_MIN."""
    s = np.asarray(scores, dtype=np.float64)
    out = np.full(s.shape, INT64_MIN, dtype=np.int64)
    ok = np.isfinite(s)
    out[ok] = np.trunc(s[ok]).astype(np.int64)
    return out


def near_tie_candidates(values: np.ndarray, rel: float = REL_TIE) -> Optional[np.ndarray]:
    """Indices (ascending) whose value is within `rel` of the maximum, or None when np.argmax is
    already unambiguous: a single candidate, candidates that are all bit-identical, or a maximum
    that is NaN / infinite (those patterns are exact on the device)."""
    v = np.asarray(values, dtype=np.float64)
    if v.size == 0 or np.isnan(v).any():
        return None
    best = v.max()
    if not np.isfinite(best):
        return None
    cand = np.nonzero(v >= best - rel * max(1.0, abs(best)))[0]
    if cand.size <= 1 or np.all(v[cand] == v[cand[0]]):
        return None
    return cand


_OBS_CACHE: dict = {}


def _obs_array(obs) -> np.ndarray:
    """The (n, 2) array of one agent's observation list, converted once per length of the list (the lists only grow;
    converting 7 000 small arrays on every call was 1.2 ms, half a run at the reference's example size)."""
    n = len(obs)
    hit = _OBS_CACHE.get(id(obs))
    if hit is not None and hit[0] == n and (n == 0 or hit[2] is obs[-1]):
        return hit[1]
    arr = np.asarray(obs, dtype=np.float64).reshape(-1, 2)
    if len(_OBS_CACHE) > 64:
        _OBS_CACHE.clear()
    _OBS_CACHE[id(obs)] = (n, arr, obs[-1] if n else None)
    return arr


def count_close_exact_many(points: np.ndarray, agents_obs_wp: Sequence[Sequence[np.ndarray]], d: float) -> np.ndarray:
    """rrt.py:246-252 / 471-477 for several points at once: observations strictly closer than d, with numpy's scalar
    norm on the pairs near the threshold (everywhere else the comparison is decided far above rounding)."""
    P = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    cnt = np.zeros(P.shape[0], dtype=np.int64)
    for obs in agents_obs_wp:
        if len(obs) == 0:
            continue
        o = _obs_array(obs)
        for lo in range(0, P.shape[0], 64):     # blocks of points: the (points x observations) temporaries stay small
            Pb = P[lo:lo + 64]
            dx = o[None, :, 0] - Pb[:, None, 0]
            dy = o[None, :, 1] - Pb[:, None, 1]
            r = np.sqrt(dx * dx + dy * dy)
            edge = np.abs(r - d) <= 1e-12 * d
            cnt[lo:lo + 64] += np.count_nonzero((r < d) & ~edge, axis=1)
            for i, j in np.argwhere(edge):
                if _norm(Pb[i] - obs[j]) < d:
                    cnt[lo + i] += 1
    return cnt


def count_close_exact(point: np.ndarray, agents_obs_wp: Sequence[Sequence[np.ndarray]], d: float) -> int:
    """One point of count_close_exact_many."""
    return int(count_close_exact_many(np.asarray(point, dtype=np.float64)[None, :], agents_obs_wp, d)[0])


def prior_kernel_exact(A: np.ndarray, B: np.ndarray, c: float, l: float) -> np.ndarray:
    """(ConstantKernel(c) * RBF(l))(A, B) the way sklearn evaluates it (kernels.py:964, 1569-1570)."""
    return np.full((A.shape[0], B.shape[0]), c, dtype=np.float64) * np.exp(-0.5 * cdist(A / l, B / l, metric="sqeuclidean"))


def exact_leaf_scores(cand: np.ndarray, leaf_points: List[np.ndarray], parent_points: List[Optional[np.ndarray]], ctx,
                      beta: float, c: float, l: float, mi_dev: Optional[float] = None) -> np.ndarray:
    """Reference-exact stage-1 scores of the candidate leaves (agent HAS observations).

    The candidates share the mutual-information term and differ only in their penalties, which are exact on the
    host at no cost (no kernel matrix).  With the device's `mi_dev` (right to ~1e-10 relative) the common term is
    only re-evaluated in the reference's own arithmetic -- an L x n_a kernel block and an L x L determinant -- when
    the penalty sums of the best candidates are so close that the rounding of ((mi - dp) - w) - ep could decide
    between them; otherwise the ordering is that of the exact penalties and `mi_dev` stands in for the value."""
    ws = ctx.workspace
    pens = []
    for i in cand:
        p, par = leaf_points[i], parent_points[i]
        dp = np.exp(0.5 * _norm(p - par) ** 2) if par is not None else 0
        w = np.inf if (p[0] < ws[0] or p[0] > ws[1] or p[1] < ws[2] or p[1] > ws[3]) else 0
        ep = np.exp(0.05 * count_close_exact(p, ctx.agents_obs_wp, ctx.d_wp) ** 2)
        pens.append((dp, w, ep))
    mi = None
    if mi_dev is not None and np.isfinite(mi_dev):
        tot = np.array([dp + w + ep for dp, w, ep in pens], dtype=np.float64)
        order = np.sort(tot)
        scale = max(abs(float(mi_dev)), float(order[0]) if np.isfinite(order[0]) else 0.0, 1.0)
        all_equal = all(pn == pens[0] for pn in pens)
        if all_equal or (len(order) > 1 and order[1] - order[0] > 64 * np.finfo(float).eps * scale):
            mi = float(mi_dev)  # decided by the penalties alone (or a bit-identical tie: the first wins either way)
    if mi is None:
        leaf_xy = np.array([np.asarray(p, dtype=np.float64) for p in leaf_points])
        obs = np.array(ctx.agents_obs_wp[ctx.agent_idx])
        K_pio = prior_kernel_exact(leaf_xy, obs, c, l)
        mi = 0.5 * np.log(np.linalg.det(np.eye(leaf_xy.shape[0]) + beta * np.dot(K_pio, K_pio.T)))
        STATS["exact_mi"] += 1
    STATS["near_ties"] += 1
    out = np.empty(len(cand), dtype=np.float64)
    for k, (dp, w, ep) in enumerate(pens):
        v = np.float64(mi)
        v = v - dp
        v = v - w
        v = v - ep
        out[k] = v
    return out


STATS = {"near_ties": 0, "exact_mi": 0}  # selector near-ties re-decided on the host / of those, with the exact MI term


def _suppression(point, source, sources) -> float:
    xt, yt = point
    xk, yk, _ = source
    total = 0
    for other in sources:
        if np.array_equal(source, other):
            continue
        mx, my = (xk + other[0]) / 2, (yk + other[1]) / 2
        d = _norm([xt - mx, yt - my])
        cd = 2 - 1 / (1 + np.exp((d - 2) / 16))
        total += (1 - cd + cd / (1 + np.exp((2 - d) / 16)))
    return total


def _segments_cross(x1, y1, x2, y2, x3, y3, x4, y4) -> bool:
    ax, ay, bx, by = x2 - x1, y2 - y1, x4 - x3, y4 - y3
    den = ax * by - ay * bx
    if den == 0:
        return False
    t = ((x3 - x1) * by - (y3 - y1) * bx) / den
    u = ((x3 - x1) * ay - (y3 - y1) * ax) / den
    return 0 <= t <= 1 and 0 <= u <= 1


def _workspace_penalty(ctx, p, par) -> float:
    ws = ctx.workspace
    if p[0] < ws[0] or p[0] > ws[1] or p[1] < ws[2] or p[1] > ws[3]:
        return np.inf
    rects = [o for o in ctx.obstacles if o.get("type") == "rectangle"]
    for o in rects:
        if o["x"] <= p[0] <= o["x"] + o["width"] and o["y"] <= p[1] <= o["y"] + o["height"]:
            return np.inf
    for o in rects:
        x0, y0, x1, y1 = o["x"], o["y"], o["x"] + o["width"], o["y"] + o["height"]
        if (_segments_cross(par[0], par[1], p[0], p[1], x0, y0, x0, y1) or _segments_cross(par[0], par[1], p[0], p[1], x1, y0, x1, y1)
                or _segments_cross(par[0], par[1], p[0], p[1], x0, y0, x1, y0)
                or _segments_cross(par[0], par[1], p[0], p[1], x0, y1, x1, y1)):
            return np.inf
    return 0


def exact_branch_gain(variant: int, ctx, chain: List[np.ndarray]) -> float:
    """Reference-exact gain of one branch.  `chain` = points from the evaluated node up to and
    including the root, in walk order (rrt.py:65-70, 91-96, 123-128, 254-259)."""
    est = ctx.estimates()
    have = est.size > 0
    total = 0
    close = None
    if variant == 3 and have and ctx.agent_has_obs() and len(chain) > 1:   # every node of the branch in one pass
        close = count_close_exact_many(np.array([np.asarray(q, dtype=np.float64) for q in chain[:-1]]), ctx.agents_obs_wp, ctx.d_wp)
    for k in range(len(chain) - 1):
        p, par = chain[k], chain[k + 1]
        xt, yt = p
        if variant == 3:
            if not have:
                total += -_workspace_penalty(ctx, p, par)
                continue
            g = 0
            for s in est:
                d = _norm([xt - s[0], yt - s[1]])
                g += (np.exp(-(d - 2) ** 2 / (32))) * _suppression(p, s, est)
            dist_pen = np.exp(0.5 * _norm(p - par) ** 2)
            expl = np.exp(0.05 * int(close[k]) ** 2) if ctx.agent_has_obs() else 0
            total += g - dist_pen - expl - _workspace_penalty(ctx, p, par)
            continue
        if not have:
            total += 0
            continue
        g = 0
        if variant == 0:
            s = est[ctx.closest_estimate_index()]
            d = _norm([xt - s[0], yt - s[1]])
            g += (1 + np.exp(-(d - 2) ** 2 / 2 * 16)) * _suppression(p, s, est)
        else:
            for s in est:
                d = _norm([xt - s[0], yt - s[1]])
                g += (1 + np.exp(-(d - 2) ** 2 / 2 * 16)) * _suppression(p, s, est)
            g = g * np.exp(0.5 * _norm(p - par))
            if variant == 2:
                th = np.arctan2(p[1] - par[1], p[0] - par[0])
                g = g * np.exp(0.05 * (th ** 2) / 0.1)
        total += g
    return total if variant == 3 else max(total, 0)

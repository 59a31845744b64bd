"""Device-side scoring of RRT candidates: the batched replacement of the reference's per-node Python
gain functions and selector penalty loops (src/rrt/rrt.py:32-259, 438-526; rrt_thread.py:371-401).

`GpuScorer` is the product implementation (C ABI -> sm_100a kernels).  The strategy classes only
talk to the small interface below, which is also what the CPU tests implement with the oracle to
check the host logic without a GPU (a test seam, not a fallback: the product constructs GpuScorer
and GpuScorer raises without CUDA).

    tree_information(variant, tree, ctx)                 -> float64[N]   node.information for every node
    leaf_scores(leaf_xy, parent_xy, has_parent, ctx, ..) -> (scores float64[L], mi float)
    ucb_acquisition(gp, leaf_xy, visited_xy, beta, d)    -> float64[L]
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

VARIANT_NO_PENALTIES, VARIANT_DISTANCE, VARIANT_DISTANCE_ROTATION, VARIANT_ALL = 0, 1, 2, 3


@dataclass
class ScoreContext:
    """What the gain functions / selectors read from the strategy object."""
    best_estimates: np.ndarray  # (S, 3) rows [x, y, A], S may be 0
    last_point: Optional[np.ndarray]  # agents_full_path[agent][-1] (variant 0 only)
    agents_obs_wp: Sequence[Sequence[np.ndarray]]  # observations per agent
    agent_idx: int
    d_wp: float
    workspace: Sequence[float]
    obstacles: List[dict] = field(default_factory=list)
    obs_cache: Optional[dict] = None  # strategy-owned: arrays of the (append-only) observation lists, by length
    _all: Optional[np.ndarray] = None
    _tiled: Optional[tuple] = None

    def estimates(self):
        return np.ascontiguousarray(np.asarray(self.best_estimates, dtype=np.float64).reshape(-1, 3))

    def all_obs(self):
        """Observations of all agents as one (O, 2) array, agent by agent (built once per context; the strategy
        passes a cache shared by every context of a planning step: turning 5000 list entries into an array costs
        more than the kernels that read them)."""
        if self.obs_cache is not None:
            return _cached_obs(self.obs_cache, self.agents_obs_wp, None)
        if self._all is None:
            rows = [np.asarray(o, dtype=np.float64).reshape(-1, 2) for o in self.agents_obs_wp if len(o) > 0]
            self._all = np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((0, 2))
        return self._all

    def all_obs_tiled(self):
        """all_obs() sorted spatially plus its tile boxes (cached with the observation arrays)."""
        if self.obs_cache is not None:
            base = self.all_obs()
            hit = self.obs_cache.get("tiled")
            if hit is None or hit[0] is not base:
                hit = (base,) + spatial_tiles(base)
                self.obs_cache["tiled"] = hit
            return hit[1], hit[2]
        if self._tiled is None:
            self._tiled = spatial_tiles(self.all_obs())
        return self._tiled

    def agent_obs(self):
        """This agent's own observations as an (n_a, 2) array."""
        if self.obs_cache is not None:
            return _cached_obs(self.obs_cache, self.agents_obs_wp, self.agent_idx)
        return np.asarray(self.agents_obs_wp[self.agent_idx], dtype=np.float64).reshape(-1, 2)

    def agent_has_obs(self):
        return len(self.agents_obs_wp[self.agent_idx]) > 0

    def obstacle_rows(self):
        rows = [[o["x"], o["y"], o["width"], o["height"]] for o in self.obstacles if o.get("type") == "rectangle"]
        return np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1, 4))

    def closest_estimate_index(self):
        """rrt.py:57: first minimum of the distance from the agent's last path point."""
        est = self.estimates()
        if est.shape[0] == 0 or self.last_point is None:
            return 0
        lp = self.last_point
        d = [np.linalg.norm([lp[0] - s[0], lp[1] - s[1]]) for s in est]
        return int(np.argmin(d))


CC_TILE = 64  # observations per bounding-box tile of ipp_count_close (csrc/ipp_branch.cu)


def spatial_tiles(obs):
    """(observations in Z-order, tile boxes) for the pruned close-count kernel, or (obs, None) when there are too few
    to be worth sorting.  The count of observations within d of a point does not depend on their order."""
    from .gp import morton_order, slab_bounding_boxes

    O = obs.shape[0]
    if O < 4 * CC_TILE:
        return obs, None
    srt = np.ascontiguousarray(obs[morton_order(obs)])
    return srt, slab_bounding_boxes(srt, (O + CC_TILE - 1) // CC_TILE * CC_TILE, slab=CC_TILE)


def _cached_obs(cache, agents_obs_wp, agent):
    """Array view of the append-only observation lists, rebuilt only when a list has grown.
    agent = None: all agents stacked in agent order."""
    lens = tuple((id(o), len(o), id(o[-1]) if len(o) else 0) for o in agents_obs_wp)  # list identity, length, tail
    key = "all" if agent is None else agent
    hit = cache.get(key)
    want = lens if agent is None else lens[agent]
    if hit is not None and hit[0] == want:
        return hit[1]
    if agent is None:
        rows = [np.asarray(o, dtype=np.float64).reshape(-1, 2) for o in agents_obs_wp if len(o) > 0]
        arr = np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((0, 2))
    else:
        arr = np.ascontiguousarray(np.asarray(agents_obs_wp[agent], dtype=np.float64).reshape(-1, 2))
    cache[key] = (want, arr)
    return arr


def resolve_close_counts(points, obs, d, sure, amb):
    """Host re-decision for the points the device flagged as having knife-edge pairs (|dist - d|
    within a few ulp).  The whole count of such a point is redone: pairs clearly inside / outside are
    counted vectorised, pairs near the edge with numpy's own scalar norm, so the `< d` outcome is the
    one the reference computes (rrt.py:250,475)."""
    counts = sure.astype(np.int64).copy()
    for i in np.nonzero(amb)[0]:
        p = points[i]
        dd = obs - p
        r = np.sqrt(dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1])
        edge = np.abs(r - d) <= 1e-12 * d
        n = int(np.count_nonzero((r < d) & ~edge))
        for j in np.nonzero(edge)[0]:
            if np.linalg.norm(p - obs[j]) < d:
                n += 1
        counts[i] = n
    return counts


class GpuScorer:
    """All scoring through libipp_b200.so on the current CUDA device."""

    def __init__(self):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.launches = 0
        # device copies that survive between scoring calls: the tree arrays (a tree is append-only between two
        # initialize_trees: key = identity, node count, rewire count) and the observation arrays (key = identity of
        # the host array the context caches).  Re-uploading them on every call was 80 % of a scoring call.
        self._resident = {}
        self._resident_lock = __import__("threading").Lock()
        # a torch.distributed group (True = default group): shard the branch walks of one tree over its ranks in
        # contiguous batches of nodes, node table / sources / observations replicated; see tree_information
        self.branch_group = None

    # -- helpers ---------------------------------------------------------------------------------
    def _dev(self, a, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(a))
        return t.cuda(non_blocking=True)

    def _resident_get(self, key, build):
        """Device tensors for `key`, built once (at most 64 entries are kept: the agents' trees and observation sets)."""
        with self._resident_lock:
            hit = self._resident.get(key)
        if hit is None:
            hit = build()
            with self._resident_lock:
                if len(self._resident) >= 64:
                    self._resident.pop(next(iter(self._resident)))
                self._resident[key] = hit
        return hit

    def close_counts(self, points, obs, d, bbox=None):
        """#observations closer than d to each point (exploitation penalty input).  bbox: tile boxes of a
        spatially sorted `obs` (spatial_tiles), which lets every point skip the tiles out of reach."""
        torch, lib = self.torch, self.lib
        P = points.shape[0]
        if P == 0 or obs.shape[0] == 0:
            return np.zeros(P, dtype=np.int64), None
        Pd, Od = self._dev(points), self._dev(obs)
        Bd = self._dev(bbox) if bbox is not None else None
        sure = torch.empty(P, dtype=torch.int32, device="cuda")
        amb = torch.empty(P, dtype=torch.int32, device="cuda")
        tot = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.ipp_count_close(_lib.ptr(Pd), P, _lib.ptr(Od), obs.shape[0], float(d), _lib.ptr(sure), _lib.ptr(amb),
                                       _lib.ptr(tot), _lib.ptr(Bd), _lib.stream_ptr()), "ipp_count_close")
        self.launches += 1
        if int(tot.item()) == 0:
            return None, sure  # device-resident, nothing ambiguous
        counts = resolve_close_counts(points, obs, d, sure.cpu().numpy(), amb.cpu().numpy())
        return counts, None

    # -- I1-I5 + T: node.information for the whole tree ----------------------------------------------
    def _upload_tree(self, variant, tree, ctx: ScoreContext, eval_range=None):
        """Host tree arrays / context -> device tensors (one dict per scoring pass)."""
        torch = self.torch
        N = tree.n
        est = ctx.estimates()
        S = est.shape[0]
        d = {"N": N, "S": S, "pts": tree.points()}

        def tree_arrays():
            hoff, htime, hpar = tree.history_csr()
            return (self._dev(d["pts"]), self._dev(tree.parent0[:N].astype(np.int32)), self._dev(hoff) if htime.size else None,
                    self._dev(htime) if htime.size else None, self._dev(hpar) if htime.size else None, tree)

        # (the entry keeps a reference to the tree, so its id cannot be reused while the entry lives)
        d["node_xy"], d["parent0"], d["hoff"], d["htime"], d["hpar"], _ = self._resident_get(
            ("tree", id(tree), N, len(getattr(tree, "rewires", ()))), tree_arrays)
        d["src"] = self._dev(est) if S else None
        d["ws"] = self._dev(np.asarray(ctx.workspace, dtype=np.float64))
        ob = ctx.obstacle_rows()
        d["n_obst"] = ob.shape[0]
        d["obst"] = self._dev(ob) if ob.shape[0] else None
        d["src_gain"] = torch.empty(N, dtype=torch.float64, device="cuda")
        d["wpen"] = torch.empty(N, dtype=torch.float64, device="cuda") if variant == VARIANT_ALL else None
        d["closest"] = ctx.closest_estimate_index() if (variant == VARIANT_NO_PENALTIES and S) else 0
        d["has_obs"] = 1 if ctx.agent_has_obs() else 0
        d["need_close"] = bool(variant == VARIANT_ALL and d["has_obs"] and S)
        if d["need_close"]:
            d["obs_host"], bbox = ctx.all_obs_tiled()
            oh = d["obs_host"]
            d["obs"], d["obs_bbox"], _ = self._resident_get(
                ("obs", id(oh), oh.shape[0]), lambda: (self._dev(oh), self._dev(bbox) if bbox is not None else None, oh))
            d["sure"] = torch.empty(N, dtype=torch.int32, device="cuda")
            d["amb"] = torch.empty(N, dtype=torch.int32, device="cuda")
            d["amb_total"] = torch.zeros(1, dtype=torch.int32, device="cuda")
        d["nclose"] = None
        lo, hi = eval_range if eval_range is not None else (0, N)
        d["B"] = hi - lo
        d["eval_node"] = torch.arange(lo, hi, dtype=torch.int32, device="cuda")
        d["out"] = torch.empty(max(hi - lo, 1), dtype=torch.float64, device="cuda")
        d["term0"] = torch.empty(N, dtype=torch.float64, device="cuda")
        d["d_wp"] = float(ctx.d_wp)
        return d

    def _launch_terms(self, variant, d):
        """Phase 1: per-node static terms (+ close-observation counts for variant 3)."""
        lib, st = self.lib, _lib.stream_ptr()
        _lib.check(lib.ipp_node_terms(variant, _lib.ptr(d["node_xy"]), d["N"], _lib.ptr(d["src"]), d["S"], d["closest"],
                                      _lib.ptr(d["ws"]), _lib.ptr(d["obst"]), d["n_obst"], _lib.ptr(d["src_gain"]),
                                      _lib.ptr(d["wpen"]), st), "ipp_node_terms")
        self.launches += 1
        if d["need_close"]:
            _lib.check(lib.ipp_count_close(_lib.ptr(d["node_xy"]), d["N"], _lib.ptr(d["obs"]), d["obs_host"].shape[0],
                                           d["d_wp"], _lib.ptr(d["sure"]), _lib.ptr(d["amb"]), _lib.ptr(d["amb_total"]),
                                           _lib.ptr(d["obs_bbox"]), st),
                       "ipp_count_close")
            self.launches += 1
            d["nclose"] = d["sure"]

    def _launch_walk(self, variant, d):
        """Phase 2: one versioned branch walk per node."""
        lib, st = self.lib, _lib.stream_ptr()
        _lib.check(lib.ipp_branch_gain(variant, _lib.ptr(d["node_xy"]), d["N"], _lib.ptr(d["parent0"]), _lib.ptr(d["hoff"]),
                                       _lib.ptr(d["htime"]), _lib.ptr(d["hpar"]), _lib.ptr(d["src_gain"]), _lib.ptr(d["wpen"]),
                                       _lib.ptr(d["nclose"]), d["has_obs"], 1 if d["S"] else 0, _lib.ptr(d["obst"]),
                                       d["n_obst"], _lib.ptr(d["eval_node"]), _lib.ptr(d["eval_node"]), d["B"],
                                       _lib.ptr(d["term0"]), _lib.ptr(d["out"]), st), "ipp_branch_gain")
        self.launches += 2

    def tree_information(self, variant, tree, ctx: ScoreContext):
        """node.information of every node (the gain of the branch that ends in it).  With `branch_group` set, every
        rank evaluates the per-node terms of the whole (replicated) tree and walks only its contiguous batch of
        nodes; one all_gather assembles the vector, identical on every rank and to the single-GPU result (each walk
        reads the same tables in the same order)."""
        if tree.n == 0:
            return np.zeros(0)
        group, eval_range, sizes = self.branch_group, None, None
        if group is not None:
            from . import dist as D
            import torch.distributed as tdist
            group = None if group is True else group
            world, rank = tdist.get_world_size(group), tdist.get_rank(group)
            bounds = [D.slab_bounds(tree.n, world, r) for r in range(world)]
            eval_range, sizes = bounds[rank], [b - a for a, b in bounds]
        d = self._upload_tree(variant, tree, ctx, eval_range)
        self._launch_terms(variant, d)
        if d["need_close"] and int(d["amb_total"].item()) > 0:
            # knife-edge distance pairs: the host re-decides them with numpy's own norm
            counts = resolve_close_counts(d["pts"], d["obs_host"], d["d_wp"], d["sure"].cpu().numpy(), d["amb"].cpu().numpy())
            d["nclose"] = self._dev(counts.astype(np.int32))
        self._launch_walk(variant, d)
        if sizes is not None:
            from . import dist as D
            info = D.gather_slabs(d["out"][: d["B"]], sizes, group).cpu().numpy()
        else:
            info = d["out"].cpu().numpy()
        info[0] = 0.0  # the root is never scored (InformativeTreeNode.information = 0, rrt_utils.py:35)
        return info

    def time_device_pass(self, variant, tree, ctx: ScoreContext, reps=20):
        """Milliseconds per scoring pass (phase 1 + phase 2) with every input already resident,
        timed with CUDA events on the launching stream (bench.py)."""
        torch = self.torch
        d = self._upload_tree(variant, tree, ctx)
        for _ in range(3):
            self._launch_terms(variant, d)
            self._launch_walk(variant, d)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(reps):
            self._launch_terms(variant, d)
            self._launch_walk(variant, d)
        e1.record()
        for _ in range(reps):
            self._launch_walk(variant, d)
        e2.record()
        torch.cuda.synchronize()
        self.last_walk_ms = e1.elapsed_time(e2) / reps  # phase 2 alone (phase 1 = the difference)
        return e0.elapsed_time(e1) / reps

    # -- I6 + I7: stage-1 selector ---------------------------------------------------------------------
    def mutual_information(self, leaf_xy, agent_obs, c, l, beta):
        """0.5 * log(det(I + beta K K^T)) with numpy.linalg.det's overflow semantics (rrt.py:525-526:
        det = exp(logdet) overflows to inf above log(DBL_MAX))."""
        torch, lib = self.torch, self.lib
        L, na = leaf_xy.shape[0], agent_obs.shape[0]
        if na == 0:
            return 0
        m = min(L, na)
        lay = (C.c_longlong * 10)()
        _lib.check(lib.ipp_gp_layout(m, lay), "ipp_gp_layout")
        lay = [int(v) for v in lay]
        r64 = lambda v: (v + 63) // 64 * 64
        G = torch.empty(r64(L) * r64(na), dtype=torch.float64, device="cuda")
        Kb = torch.empty(lay[3], dtype=torch.float64, device="cuda")
        Wb = torch.empty(lay[4], dtype=torch.float64, device="cuda")
        vb = torch.zeros(lay[6], dtype=torch.float64, device="cuda")
        A, B = self._dev(leaf_xy), self._dev(agent_obs)
        _lib.check(lib.ipp_mi_logdet(_lib.ptr(A), L, _lib.ptr(B), na, float(c), float(l), float(beta), _lib.ptr(G),
                                     _lib.ptr(Kb), _lib.ptr(Wb), _lib.ptr(vb), _lib.stream_ptr()), "ipp_mi_logdet")
        self.launches += 1
        res = vb[lay[8]:lay[8] + 8].cpu().numpy()
        if res[3] != 0.0:
            return float("nan")
        logdet = 2.0 * float(res[5])
        with np.errstate(over="ignore"):
            return 0.5 * np.log(np.exp(np.float64(logdet)))

    def leaf_scores(self, leaf_xy, parent_xy, has_parent, ctx: ScoreContext, beta, kernel_c, kernel_l):
        torch, lib = self.torch, self.lib
        L = leaf_xy.shape[0]
        agent_obs = ctx.agent_obs()
        mi = self.mutual_information(leaf_xy, agent_obs, kernel_c, kernel_l, beta)
        has_obs = 1 if ctx.agent_has_obs() else 0
        nclose_d = None
        if has_obs:
            obs_sorted, bbox = ctx.all_obs_tiled()
            counts, dev_counts = self.close_counts(leaf_xy, obs_sorted, ctx.d_wp, bbox)
            nclose_d = dev_counts if dev_counts is not None else self._dev(counts.astype(np.int32))
        scores = torch.empty(L, dtype=torch.float64, device="cuda")
        lx, px = self._dev(leaf_xy), self._dev(parent_xy)
        hp = self._dev(np.asarray(has_parent, dtype=np.int32))
        ws_d = self._dev(np.asarray(ctx.workspace, dtype=np.float64))
        _lib.check(lib.ipp_leaf_penalty(_lib.ptr(lx), _lib.ptr(px), _lib.ptr(hp), L, _lib.ptr(nclose_d), has_obs,
                                        _lib.ptr(ws_d), float(mi), _lib.ptr(scores), _lib.stream_ptr()), "ipp_leaf_penalty")
        self.launches += 1
        return scores.cpu().numpy(), mi

    # -- G6 selector half: UCB on the posterior -----------------------------------------------------------
    def ucb_acquisition(self, gp, leaf_xy, visited_xy, beta, d_wp):
        """rrt_thread.py:375-387: normalised mu + beta * normalised sigma, minus the visited penalty."""
        torch, lib = self.torch, self.lib
        L = leaf_xy.shape[0]
        if hasattr(gp, "predict_points_device") and hasattr(gp, "X_train_"):
            mu_d, sd_d, _ = gp.predict_points_device(self._dev(leaf_xy))
            mu, sd = mu_d.cpu().numpy(), sd_d.cpu().numpy()
        else:
            mu, sd = gp.predict(leaf_xy, return_std=True)
        mu_n = (mu - np.mean(mu)) / np.std(mu) if np.std(mu) != 0 else mu
        sd_n = (sd - np.mean(sd)) / np.std(sd) if np.std(sd) != 0 else sd
        acq = mu_n + beta * sd_n
        visited = np.asarray(visited_xy, dtype=np.float64).reshape(-1, 2)
        if visited.shape[0]:
            acq_d = self._dev(acq)
            lx, vx = self._dev(leaf_xy), self._dev(visited)
            _lib.check(lib.ipp_ucb_visit_penalty(_lib.ptr(lx), L, _lib.ptr(vx), visited.shape[0], float(d_wp),
                                                 _lib.ptr(acq_d), _lib.stream_ptr()), "ipp_ucb_visit_penalty")
            self.launches += 1
            acq = acq_d.cpu().numpy()
        return acq, float(np.mean(sd))

    def argmax_first(self, values):
        """np.argmax semantics on the device (used by the sharded selectors)."""
        torch, lib = self.torch, self.lib
        v = self._dev(np.asarray(values, dtype=np.float64))
        bv = torch.empty(1, dtype=torch.float64, device="cuda")
        bi = torch.empty(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.ipp_argmax_first(_lib.ptr(v), v.shape[0], _lib.ptr(bv), _lib.ptr(bi), _lib.stream_ptr()),
                   "ipp_argmax_first")
        self.launches += 1
        return int(bi.item()), float(bv.item())

    # -- particle log-likelihoods of the importance sampler ---------------------------
    def prepare_counts(self, obs_wp, obs_vals):
        """Observation side of the Poisson likelihood (estimation.py:36-37), uploaded once per source estimate and
        shared by every stage and every source count: coordinates, rounded counts, and scipy's own
        gammaln(k + 1) -- one value per observation, the same bits the reference computes per (particle, observation)."""
        from scipy.special import gammaln
        obs = np.ascontiguousarray(np.array(obs_wp, dtype=np.float64).reshape(-1, 2))
        counts = np.round(obs_vals).astype(int)
        return {"n": int(obs.shape[0]), "obs": self._dev(obs), "counts": self._dev(counts.astype(np.float64)),
                "lgam": self._dev(np.asarray(gammaln(counts + 1), dtype=np.float64))}

    def particle_log_likelihood(self, thetas, prepared, lambda_b, M):
        """(ll, S1, S2) per particle: the log-likelihood and the two error-bound sums of ipp_poisson_loglik."""
        torch, lib = self.torch, self.lib
        P = np.ascontiguousarray(thetas, dtype=np.float64)
        N = P.shape[0]
        if P.shape[1] != 3 * M:
            raise ValueError(f"particles have {P.shape[1]} columns, expected 3 * {M}")
        Pd = self._dev(P)
        out = torch.empty((N, 3), dtype=torch.float64, device="cuda")
        _lib.check(lib.ipp_poisson_loglik(_lib.ptr(Pd), N, int(M), _lib.ptr(prepared["obs"]), _lib.ptr(prepared["counts"]),
                                          _lib.ptr(prepared["lgam"]), prepared["n"], float(lambda_b), _lib.ptr(out),
                                          _lib.stream_ptr()), "ipp_poisson_loglik")
        self.launches += 1
        h = out.cpu().numpy()
        return h[:, 0], h[:, 1], h[:, 2]

"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: slab partition + all_gather assembly of the
grid posterior, global argmax with numpy tie/NaN semantics, path broadcast, round sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mripp_b200 import dist as D
from oracle import gp_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _optimum(X, y, i):
    """Restart i of a fit on the oracle's objective: L-BFGS-B from a seeded start (sklearn _gpr.py:312-333)."""
    from scipy.optimize import minimize
    yn, _, _ = O.normalise_y(y)
    b = O.log_bounds()
    t0 = np.zeros(2) if i == 0 else np.random.RandomState(100 + i).uniform(b[:, 0], b[:, 1])

    def f(theta):
        v, g = O.lml(theta, X, yn, True)
        return -v, -g

    res = minimize(f, t0, method="L-BFGS-B", jac=True, bounds=b)
    return res.x, res.fun


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)
        X = rng.uniform(0, 6, (40, 2))
        y = np.sin(X[:, 0]) + 0.1 * X[:, 1]
        st = O.fit(X, y, 1.0, 1.5, optimize=False)
        ws, nx, ny = (0, 6, 0, 5), 60, 50
        _, _, r, _ = O.workspace_grid(ws)

        def slab(a, b):
            m, s, _ = O.predict(st, r[a:b])
            return torch.from_numpy(10 ** m), torch.from_numpy(s)

        z, s = D.sharded_grid_posterior(None, ws, nx, ny, compute_slab=slab)

        class WorkModelGP:
            """The two methods sharded_grid_posterior uses of a fitted estimator, on the oracle: a per-row work
            estimate (uneven on purpose) and the posterior of a flat index range."""
            def lattice_row_work(self, x0, x1, nx_, y0, y1, ny_):
                return 1.0 + 4.0 * np.exp(-0.5 * ((np.arange(ny_) - 0.3 * ny_) / (0.1 * ny_)) ** 2)

            def predict_grid_device(self, x0, x1, nx_, y0, y1, ny_, g_begin=0, g_end=None, want_zpred=True):
                zz, ss = slab(g_begin, g_end)
                return None, ss, zz

        zb, sb = D.sharded_grid_posterior(WorkModelGP(), ws, nx, ny)
        assert np.array_equal(zb, z) and np.array_equal(sb, s)          # the split does not change the field
        bal = D.lattice_slabs(WorkModelGP(), ws, nx, ny, world)
        assert bal[0][1] != D.row_slab(nx, ny, world, 0)[1]              # ... and it is not the equal-rows split
        scores = np.array([1.0, 5.0, 5.0, np.nan, 2.0, np.nan, 9.0])
        lo, hi = D.slab_bounds(len(scores), world, rank)
        k_nan = D.sharded_argmax(torch.from_numpy(scores[lo:hi]), lo)
        clean = np.array([1.0, 5.0, 3.0, 5.0, 5.0, 0.0, 4.0])
        k_tie = D.sharded_argmax(torch.from_numpy(clean[lo:hi]), lo)
        path = D.broadcast_path(np.array([[1.0, 2.0], [3.0, 4.0]]) if rank == 1 else None, src=1)
        optima = D.sharded_restarts(lambda idx: [_optimum(X, y, i) for i in idx], 5, 2)
        q.put((rank, z, s, k_nan, k_tie, path, D.rounds_for_rank(7, world, rank), optima))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_matches_single_process(world):
    """world 2, and world 3 for the ragged cases: 5 restarts, 7 scores and 50 lattice rows over 3 ranks."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    X = rng.uniform(0, 6, (40, 2))
    y = np.sin(X[:, 0]) + 0.1 * X[:, 1]
    st = O.fit(X, y, 1.0, 1.5, optimize=False)
    _, _, r, _ = O.workspace_grid((0, 6, 0, 5))
    m, s, _ = O.predict(st, r)
    serial = [_optimum(X, y, i) for i in range(5)]
    for rank, z, sd, k_nan, k_tie, path, rounds, optima in out:
        # restarts spread over the ranks: the gathered list is the serial list, in start order, on every rank
        assert len(optima) == 5
        assert all(np.array_equal(a[0], b[0]) and a[1] == b[1] for a, b in zip(optima, serial))
        assert int(np.argmin([o[1] for o in optima])) == int(np.argmin([o[1] for o in serial]))
        assert np.array_equal(z, 10 ** m) and np.array_equal(sd, s)  # slabs are row independent: exact
        assert k_nan == 3 and k_tie == 1
        assert np.array_equal(path, [[1.0, 2.0], [3.0, 4.0]])
        assert rounds == [r_ for r_ in range(7) if r_ % world == rank]


def test_partition_helpers():
    for total in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            b = [D.slab_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert D.row_slab(10, 7, 2, 1) == (40, 70)
    assert D.combine_argmax(np.array([np.nan, 1.0]), np.array([5, 0])) == 5
    assert D.combine_argmax(np.array([2.0, 2.0, 3.0]), np.array([4, 1, -1])) == 1


def test_balanced_row_slabs_partition_and_balance():
    """Work-balanced lattice split: contiguous whole rows, covers the lattice once, near-equal work, and never an
    empty slab while rows remain (dist.balanced_row_slabs)."""
    from mripp_b200.dist import balanced_row_slabs
    rng = np.random.RandomState(0)
    for ny, nx, world in [(1000, 1000, 8), (50, 37, 3), (7, 5, 7), (3, 10, 8), (1, 4, 2)]:
        w = rng.rand(ny) + np.concatenate([np.linspace(0, 3, ny // 2), np.linspace(3, 0, ny - ny // 2)])
        b = balanced_row_slabs(w, nx, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == nx * ny
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert all(lo % nx == 0 and hi % nx == 0 and lo <= hi for lo, hi in b)
        if ny >= world:
            assert all(hi > lo for lo, hi in b)
        if ny >= 50 * world:
            loads = np.array([w[lo // nx:hi // nx].sum() for lo, hi in b])
            assert loads.max() <= 1.05 * loads.mean()
    assert balanced_row_slabs(np.zeros(10), 10, 3) == [(0, 40), (40, 70), (70, 100)]  # no information: near-equal rows


# ---- independent rounds sharded over the ranks (BASELINE configs[4]): main.py --shard-rounds ------------------
_ROUNDS_CONFIG = {
    "scenarios": [{"class": "PointSourceField",
                   "params": {"num_sources": 2, "workspace_size": [20, 20], "intensity_range": [100000, 1000000]}}],
    "strategies": ["RRTRIG_PointSourceInformative_All_SourceMetric_PathPlanning", "MultiAgentBoustrophedon"],
    "args": {"rounds": 3, "beta_t": 5.0, "save": False, "show": False, "seed": 21, "budget": 30, "d_waypoint_distance": 2.5,
             "budget_iter": 3, "run_number": 1, "lambda_b": 1, "max_sources": 2, "n_samples": 40, "s_stages": 3, "num_agents": 2,
             "stage_lambda": 0.5},
}


def _run_rounds(shard_rounds):
    import argparse
    import copy
    from mripp_b200 import main as M
    from tests._doubles import OracleGP, OracleScorer

    class FastGP(OracleGP):
        full = False

    a = argparse.Namespace(sweep=False, out=None, debug=False, shard_rounds=shard_rounds)
    res = M.run(a, copy.deepcopy(_ROUNDS_CONFIG), gp_factory=FastGP, scorer_factory=OracleScorer)
    return [(r["scenario"], r["round"], r["strategy"], r["rmse"], r["wrmse"], r["entropy"], r["n_obs"]) for r in res]


def _rounds_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    q.put((rank, _run_rounds(True)))


def test_rounds_sharded_over_two_ranks_equal_one_process(monkeypatch):
    """Every rank ends up with the full, ordered list of per-run metrics, equal to what one process computes for the
    same individually seeded rounds."""
    monkeypatch.setenv("WORLD_SIZE", "1")
    single = _run_rounds(True)
    assert [(s, r) for s, r, *_ in single] == [(1, 1), (1, 1), (1, 2), (1, 2), (1, 3), (1, 3)]
    assert len({t[3] for t in single if t[2].startswith("RRTRIG")}) == 3      # the rounds differ (own seeds)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rounds_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert res == single


def _worker8(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        optima = {k: D.sharded_restarts(lambda idx: [(np.array([i, 2.0 * i]), float((i * 7) % 5)) for i in idx], k, 2)
                  for k in (5, 11, 16)}
        scores = np.arange(8192, dtype=float) % 97
        lo, hi = D.slab_bounds(len(scores), world, rank)
        best = D.sharded_argmax(torch.from_numpy(scores[lo:hi]), lo)

        def slab(a, b):
            return torch.arange(a, b, dtype=torch.float64), -torch.arange(a, b, dtype=torch.float64)

        z, s = D.sharded_grid_posterior(None, (0, 100, 0, 100), 1000, 1000, compute_slab=slab)
        q.put((rank, optima, best, float(z.sum()), float(s.sum()), z.shape, bool(np.array_equal(z, np.arange(10 ** 6)))))
    finally:
        dist.destroy_process_group()


def test_eight_ranks_restarts_argmax_and_lattice_gather():
    """The rank count of the scaling run: 11 restarts (the stock fit), fewer restarts than ranks (5) and more (16) over
    8 ranks; the first-maximum argmax of 8192 scores; the gather of a 10^6-point lattice from 8 row slabs."""
    world, port = 8, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker8, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, optima, best, zs, ss, shape, exact in out:
        for k, opt in optima.items():
            assert len(opt) == k
            assert all(np.array_equal(o[0], [i, 2.0 * i]) and o[1] == float((i * 7) % 5) for i, o in enumerate(opt))
        assert best == 96 and shape == (10 ** 6,) and exact and ss == -zs

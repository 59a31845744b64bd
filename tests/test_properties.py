"""Property tests (hypothesis) of the host-side decision logic: the pieces whose job is to return exactly what
numpy would have returned, for any input."""
import numpy as np
from hypothesis import given, settings, strategies as st

from mripp_b200 import dist as D
from mripp_b200.src.estimation import estimation as E
from tests._doubles import NoisyLikelihoodScorer, OracleScorer

SET = settings(max_examples=60, deadline=None, derandomize=True)


@SET
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), power=st.sampled_from([1, 4, 30]), zeros=st.booleans())
def test_choice_from_draws_matches_numpy(seed, n, power, zeros):
    rng = np.random.RandomState(seed)
    w = rng.rand(n) ** power + 1e-300
    if zeros and n > 1:
        w[rng.rand(n) < 0.6] = 0.0
        w[rng.randint(n)] = 1.0
    w /= w.sum()
    np.random.seed(seed)
    want = np.random.choice(n, size=n, p=w)
    np.random.seed(seed)
    _, got = E.choice_from_draws(w, np.random.random_sample(n))
    assert np.array_equal(want, got)


@SET
@given(seed=st.integers(0, 2**31 - 1), world=st.integers(1, 9), parts=st.integers(1, 40))
def test_combine_argmax_is_numpy_argmax(seed, world, parts):
    rng = np.random.RandomState(seed)
    v = np.round(rng.randn(parts), 1)                      # coarse values: ties are common
    v[rng.rand(parts) < 0.15] = np.nan
    bounds = [D.slab_bounds(parts, world, r) for r in range(world)]
    vals, idx = [], []
    for lo, hi in bounds:
        if hi > lo:
            k = int(np.argmax(v[lo:hi]))
            vals.append(v[lo + k]); idx.append(lo + k)
        else:
            vals.append(0.0); idx.append(-1)
    assert D.combine_argmax(np.array(vals), np.array(idx)) == int(np.argmax(v))


@SET
@given(seed=st.integers(0, 2**31 - 1), ny=st.integers(1, 400), nx=st.integers(1, 50), world=st.integers(1, 12))
def test_balanced_row_slabs_is_a_partition(seed, ny, nx, world):
    w = np.random.RandomState(seed).rand(ny) ** 3
    b = D.balanced_row_slabs(w, nx, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == nx * ny
    assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
    assert all(lo % nx == 0 and lo <= hi for lo, hi in b)
    if ny >= world:
        assert all(hi > lo for lo, hi in b)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31 - 1), n_obs=st.integers(1, 60), M=st.integers(1, 3), gamma=st.sampled_from([0.1, 0.55, 1.0]),
       scale=st.sampled_from([1.0, 1e3, 1e6]), extreme=st.booleans())
def test_resampling_survives_a_maximally_wrong_backend(seed, n_obs, M, gamma, scale, extreme):
    """resample_indices with a likelihood backend that errs by up to its reported bound returns np.random.choice's
    indices on the reference's weights and leaves the RNG where that call leaves it."""
    rng = np.random.RandomState(seed)
    N = 64
    obs = rng.uniform(0, 20, (n_obs, 2))
    meas = rng.uniform(5, 5e3, n_obs)
    P = np.column_stack([c for _ in range(M) for c in (rng.uniform(0, 20, N), rng.uniform(0, 20, N), rng.uniform(1e4, 1e5, N))])
    if seed % 3 == 0:
        P = P[rng.randint(N)] * (1 + 1e-4 * rng.randn(N, 3 * M))        # a tight cloud: weights spread over many particles
    np.random.seed(seed % 1000)
    w = E.stage_weights(E.batched_poisson_log_likelihood(P, list(obs), list(meas), 1, M), gamma, N)
    want = np.random.choice(N, size=N, p=w)
    after = np.random.uniform(size=2)
    for backend in (OracleScorer(), NoisyLikelihoodScorer(scale, seed=seed, extreme=extreme)):
        np.random.seed(seed % 1000)
        got = E.resample_indices(P, gamma, N, list(obs), list(meas), 1, M, backend, backend.prepare_counts(list(obs), list(meas)))
        assert np.array_equal(want, got) and np.array_equal(after, np.random.uniform(size=2))


def _python_growth(tree, mode, ws, step, radius, budget):
    """The loops of mripp_b200/src/rrt/rrt.py (rrt / rrt_star / rig) on TreeArrays' decision-exact primitives."""
    from mripp_b200.src.rrt import rrt_utils as U
    norm = np.linalg.norm
    travelled = 0
    while travelled < budget:
        target = np.random.rand(2) * (ws[1] - ws[0], ws[3] - ws[2]) + (ws[0], ws[2])
        nn = tree.nearest(target)
        point = U.steer_point(tree.point(nn), target, step)
        nb = tree.near(point, radius) if mode else []
        best = nn
        if mode == 1:
            best_cost = tree.path_cost(nn) + norm(tree.point(nn) - point)
            for q in nb:
                c = tree.path_cost(q) + norm(tree.point(q) - point)
                if c < best_cost:
                    best, best_cost = q, c
        new = tree.add(point, best)
        travelled += norm(tree.point(new) - tree.point(best))
        if mode:
            tree.rewire(nb, new)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31 - 1), mode=st.sampled_from([0, 1, 2]), step=st.sampled_from([0.5, 1.0, 2.5]),
       radius=st.sampled_from([1.0, 2.5, 4.0]), budget=st.sampled_from([0.0, 3.0, 25.0, 60.0]),
       ws=st.sampled_from([(0, 40, 0, 40), (0, 10, 0, 25), (-5, 5, 2, 4), (0.5, 12.25, 0, 7)]), calls=st.integers(1, 3))
def test_native_tree_growth_equals_python_primitives(seed, mode, step, radius, budget, ws, calls):
    from mripp_b200.src.rrt import rrt_utils as U
    if U.native_norm_variant() is None:
        return
    out = []
    for native in (True, False):
        np.random.seed(seed % 100000)
        root = np.array([ws[0] + 0.3 * (ws[1] - ws[0]), ws[2] + 0.6 * (ws[3] - ws[2])])
        t = U.TreeArrays(root, capacity=4)
        for _ in range(calls):
            if native:
                assert t.grow_native(mode, ws, step, radius, budget, block=16) is not None     # small blocks: many refills
            else:
                _python_growth(t, mode, ws, step, radius, budget)
        out.append((t.points().copy(), t.parent0[: t.n].copy(), t.parent[: t.n].copy(), t.edge[: t.n].copy(), t.nchild[: t.n].copy(),
                    list(t.rewires), np.random.uniform(size=2)))
    a, b = out
    assert all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and a[5] == b[5] and np.array_equal(a[6], b[6])


@SET
@given(seed=st.integers(0, 2**31 - 1), n_obs=st.integers(1, 80), M=st.integers(1, 3), distinct=st.integers(1, 12),
       N=st.integers(2, 90))
def test_distinct_rows_likelihood_is_the_full_evaluation(seed, n_obs, M, distinct, N):
    """distinct_rows_log_likelihood (what a re-decided sampler stage evaluates) returns the bits of the full
    batched_poisson_log_likelihood for any cloud of repeated particles."""
    rng = np.random.RandomState(seed)
    obs = rng.uniform(0, 40, (n_obs, 2))
    vals = rng.poisson(10.0 ** rng.uniform(0, 4, n_obs)).astype(float)
    rows = np.column_stack([rng.uniform(0, 40, (distinct, 2)), rng.uniform(1e5, 1e6, distinct)] * M)
    parts = rows[rng.randint(0, distinct, N)]
    assert np.array_equal(E.distinct_rows_log_likelihood(parts, obs, vals, 1.0, M),
                          E.batched_poisson_log_likelihood(parts, obs, vals, 1.0, M))


@SET
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(0, 60), cuts=st.integers(0, 6))
def test_rewire_log_is_the_list_it_replaces(seed, n, cuts):
    """Any interleaving of Python appends and native blocks gives the list of tuples, in order."""
    from mripp_b200.src.rrt.rrt_utils import RewireLog

    rng = np.random.RandomState(seed)
    ev = [(int(a), int(b), int(a)) for a, b in zip(np.sort(rng.randint(1, 500, n)), rng.randint(0, 400, n))]
    edges = sorted(set([0, n] + rng.randint(0, n + 1, cuts).tolist()))
    log = RewireLog()
    for k, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
        if k % 2:
            log.extend_array(np.array(ev[a:b], dtype=np.int64).reshape(-1, 3))
        else:
            for e in ev[a:b]:
                log.append(e)
    assert len(log) == n and list(log) == ev and log == ev and bool(log) == (n > 0)
    assert np.array_equal(log.as_array(), np.array(ev, dtype=np.int64).reshape(-1, 3))


@SET
@given(m=st.integers(1, 3 * 10 ** 6), NP=st.sampled_from([64, 512, 2048, 5056, 16384]), free=st.integers(0, 180 << 30),
       have=st.integers(0, 1 << 29), sms=st.sampled_from([108, 132, 148]))
def test_point_table_rows_fit_their_budget(m, NP, free, have, sms):
    from mripp_b200 import gp as G

    rows = G.point_table_rows(m, NP, free, have, sms)
    assert rows >= 128 and rows % 128 == 0 and rows <= max(128, (m + 127) // 128 * 128)
    if rows > 128:
        scratch = rows * NP + NP + 4 * (rows // 128)
        assert 8 * rows * NP <= min(4 << 30, free // 4) or scratch <= have

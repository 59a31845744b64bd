"""Property tests (hypothesis) of the host-side decision logic: the pieces whose job is to return exactly what
numpy would have returned, for any input."""
import numpy as np
from hypothesis import given, settings, strategies as st

from mripp_b200 import dist as D
from mripp_b200.src.estimation import estimation as E
from tests._doubles import NoisyLikelihoodScorer, OracleScorer

SET = settings(max_examples=60, deadline=None)


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


@settings(max_examples=25, deadline=None)
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

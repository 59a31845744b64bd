"""Multi-source parameter estimation (Morelande / Ristic / Gunatilaka importance sampling with progressive
correction): mirror of the reference's src/estimation/estimation.py.

What runs where.  The Poisson log-likelihood of all `n_samples` particles -- an N x n_obs x M array expression per
stage, 85 % of a whole multi-agent run once the GP and the scoring are on the device -- is one launch of
ipp_poisson_loglik.  Everything that consumes the global RNG (prior draws, np.random.choice, multivariate_normal.rvs)
and everything that is returned to the caller (particles, the estimate, its log-likelihood and BIC) stays on the host
in the reference's own numpy / scipy calls, so those values keep the reference's bits.

Decision parity.  The device log-likelihoods only decide WHICH particles are resampled.  They differ from the
reference's by rounding (log ulps, summation order), so a resampling draw u could in principle fall on the other side
of a CDF step.  The kernel returns, next to ll_i, the sums that bound |ll_i - ll_i^reference| (BOUND below); the host
turns them into a band around every CDF step and, when any of the stage's draws lies inside the band, re-decides that
stage from the reference's own expression (batched_poisson_log_likelihood, identical numpy / scipy calls) with the
same draws.  The resampled indices -- and with them every particle, estimate and waypoint downstream -- are
therefore the reference's, bit for bit; STATS counts how often the band was hit.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np
from scipy.stats import multivariate_normal, poisson, uniform

EPS = float(np.finfo(np.float64).eps)
# |ll_gpu - ll_reference| <= EPS * (C_TERM * S1 + c_sum(n) * S2).  Per observation, both sides hold the same rate,
# count and lgamma; they differ by the error of log() (CUDA: <= 1 ulp, documented; glibc, which scipy's xlogy calls:
# < 1 ulp -> 2 EPS |k log rate| at worst), by the rounding of k * log on either side (1 EPS together), and by the two
# subtractions on either side (<= 2 EPS of the partial results): <= 3 EPS |k log rate| + 2 EPS (|k log rate| + lgamma +
# rate) <= 5 EPS S1.  C_TERM = 6 leaves a fifth of headroom (16 in round 1: at the reference's example size, 1000 draws
# per stage, the wider band had a draw inside it in one stage out of seven).
# The summations differ in order: numpy's pairwise sum (blocks of 128 on 8 accumulators: <= 24 additions deep for
# n <= 8192, + 1 per doubling) against the kernel's (ceil(n / 256) sequential + 8 tree levels); each addition
# contributes EPS / 2 of the running sum, itself <= S2.
C_TERM = 6.0
STATS = {"stages": 0, "redecided": 0, "device_evaluations": 0, "why_nonfinite": 0, "why_loose_bound": 0, "why_in_band": 0,
         "rows": 0, "distinct_rows": 0}   # rows / distinct_rows: particles of the re-decided stages, and how many were evaluated


def c_sum(n: int) -> float:
    """Summation part of the bound in units of EPS: (additions in sequence on the device + on the host) * EPS / 2, with
    1 % for the second-order terms (round 1 charged a whole EPS per addition)."""
    depth = np.ceil(n / 256.0) + 8.0 + 24.0 + max(0.0, np.ceil(np.log2(max(n, 1) / 8192.0)))
    return float(0.505 * depth)


def poisson_log_likelihood(theta: np.ndarray, obs_wp: np.ndarray, obs_vals: np.ndarray, lambda_b: float, M: int) -> float:
    """estimation.py:15-55 for one parameter vector (host: its value is returned to the caller through the BIC)."""
    return float(batched_poisson_log_likelihood(np.asarray(theta, dtype=float)[None, :], obs_wp, obs_vals, lambda_b, M)[0])


_LOGPMF_DIRECT = None   # None: untested; True / False after the self-check
_HOST_BLOCK = 16        # particles per block of the host expression: its temporaries stay in the CPU's caches
_HOST_THREADS = max(1, min(8, (os.cpu_count() or 1) - 1))
_POOL = None


def _host_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=_HOST_THREADS, thread_name_prefix="ipp-host-ll")
    return _POOL


def _logpmf_direct_ok() -> bool:
    """scipy.stats.poisson.logpmf(k, mu) validates and masks its arguments (a third of the call at these sizes) and then
    evaluates xlogy(k, mu) - gammaln(k + 1) - mu on the valid entries (scipy/stats/_discrete_distns.py, poisson_gen).
    Checked once on this scipy for integer k >= 0 and mu > 0: the bare expression gives the same bits, including k = 0,
    huge counts and the 1e-10 rate floor."""
    global _LOGPMF_DIRECT
    if _LOGPMF_DIRECT is None:
        try:
            from scipy.special import gammaln, xlogy
            rng = np.random.RandomState(5)
            k = np.concatenate([rng.poisson(10.0 ** rng.uniform(-1, 6, 4000)), [0, 0, 1, 2, 10 ** 7]]).astype(int)
            mu = np.concatenate([10.0 ** rng.uniform(-3, 6.5, 4000), [1e-10, 3.7, 1e-10, 1e6, 1e7 + 0.5]])
            a = poisson.logpmf(k[None, :], np.stack([mu, mu[::-1]]))
            b = xlogy(k[None, :], np.stack([mu, mu[::-1]])) - gammaln(k + 1)[None, :] - np.stack([mu, mu[::-1]])
            _LOGPMF_DIRECT = bool(np.array_equal(a, b))
        except Exception:
            _LOGPMF_DIRECT = False
    return _LOGPMF_DIRECT


def batched_poisson_log_likelihood(thetas: np.ndarray, obs_wp, obs_vals, lambda_b: float, M: int) -> np.ndarray:
    """Log-likelihood of every row of `thetas` (N x 3M) in array expressions instead of the reference's Python loop over
    particles (estimation.py:97) -- element for element the same numpy / scipy operations, hence the same bits.
    rate = lambda_b + sum_i A_i / max(d2, 1e-6).  Used for the values that leave this module and to re-decide a stage
    whose draws fall inside the error band of the device likelihoods.  The particles go through in blocks (every row's
    reductions are its own, so blocking changes nothing but the size of the temporaries), and where the self-check
    above holds the Poisson term is evaluated directly, with gammaln(k + 1) once per observation."""
    obs_wp = np.array(obs_wp)
    counts = np.round(obs_vals).astype(int)
    thetas = np.asarray(thetas, dtype=float)
    src = thetas.reshape(thetas.shape[0], M, 3)
    direct = counts.size > 0 and counts.min() >= 0 and _logpmf_direct_ok()
    if direct:
        from scipy.special import gammaln, xlogy
        lg = gammaln(counts + 1)
    out = np.empty(src.shape[0])
    ox, oy = obs_wp[None, :, None, 0], obs_wp[None, :, None, 1]

    def block(b0: int) -> None:
        sb = src[b0:b0 + _HOST_BLOCK]
        d2 = (ox - sb[:, None, :, 0]) ** 2 + (oy - sb[:, None, :, 1]) ** 2
        rate = lambda_b + np.sum(sb[:, None, :, 2] / np.maximum(d2, 1e-6), axis=2)
        rate = np.maximum(rate, 1e-10)
        if direct:
            out[b0:b0 + _HOST_BLOCK] = np.sum(xlogy(counts[None, :], rate) - lg[None, :] - rate, axis=1)
        else:
            out[b0:b0 + _HOST_BLOCK] = np.sum(poisson.logpmf(counts[None, :], rate), axis=1)

    starts = range(0, src.shape[0], _HOST_BLOCK)
    if direct and src.shape[0] * counts.size >= 1 << 18 and _HOST_THREADS > 1:
        # the blocks are independent rows of the same element-wise expression (numpy releases the GIL inside its loops)
        list(_host_pool().map(block, starts))
    else:
        for b0 in starts:
            block(b0)
    return out


def stage_weights(ll: np.ndarray, gamma: float, n_samples: int) -> np.ndarray:
    """calc_weights, estimation.py:97-109, from the particles' log-likelihoods."""
    log_w = gamma * ll
    w = np.exp(log_w - np.max(log_w))
    w /= np.sum(w)
    if np.any(np.isnan(w)):
        print("Warning: NaN weights encountered. Returning uniform weights.")
        w = np.full(n_samples, 1 / n_samples)
    return w


def choice_from_draws(w: np.ndarray, u: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The index computation of RandomState.choice(n, size, p=w) given its uniform draws: cdf = cumsum(p) / cdf[-1],
    searchsorted(side='right') (numpy/random/mtrand.pyx, legacy choice with replacement)."""
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return cdf, cdf.searchsorted(u, side="right").astype(np.int64)


def draws_inside_band(cdf: np.ndarray, u: np.ndarray, idx: np.ndarray, band: float) -> bool:
    """True when some draw lies within `band` of a CDF step it could cross (the steps are cdf[0 .. N-2]; the last
    entry is exactly 1 on both sides and u < 1)."""
    n = cdf.shape[0]
    below = np.where(idx > 0, cdf[np.maximum(idx - 1, 0)], -np.inf)        # the step at or below u
    above = np.where(idx < n - 1, cdf[np.minimum(idx, n - 1)], np.inf)      # the next step above u
    return bool(np.any(u - below <= band) or np.any(above - u <= band))


def resample_indices(particles, gamma, n_samples, obs_wp, obs_vals, lambda_b, M, scorer, prepared) -> np.ndarray:
    """np.random.choice(n_samples, size=n_samples, p=weights) of estimation.py:117 with the weights' log-likelihoods
    from the device; consumes exactly the n_samples uniforms the reference's call consumes."""
    from mripp_b200.exact import RNG_SECTION
    with RNG_SECTION:
        return _resample_indices(particles, gamma, n_samples, obs_wp, obs_vals, lambda_b, M, scorer, prepared)


def _resample_indices(particles, gamma, n_samples, obs_wp, obs_vals, lambda_b, M, scorer, prepared) -> np.ndarray:
    STATS["stages"] += 1
    ll, s1, s2 = scorer.particle_log_likelihood(particles, prepared, lambda_b, M)
    exact = s1 is None                        # a host backend (tests) hands back the reference's own values
    if not exact:
        STATS["device_evaluations"] += 1
    u = np.random.random_sample(n_samples)        # the draws RandomState.choice makes for this call, nothing else
    if not exact and bool(np.all(particles == particles[0])):
        # A collapsed cloud (one particle repeated -- the usual state after the first stages once the likelihood is
        # peaked: all draws pick the same row and the perturbation covariance is zero).  Equal rows get equal
        # log-likelihoods from the reference's element-wise expression, so its weights are exactly uniform
        # (or NaN -> replaced by uniform, estimation.py:106-108): no band to check.
        return choice_from_draws(stage_weights(np.zeros(n_samples), gamma, n_samples), u)[1]
    ok = exact or bool(np.all(np.isfinite(ll)) and np.all(np.isfinite(s1)) and np.all(np.isfinite(s2)))
    if ok:
        w = stage_weights(ll, gamma, n_samples)
        cdf, idx = choice_from_draws(w, u)
        if exact:
            return idx
        bound = EPS * (C_TERM * s1 + c_sum(len(obs_vals)) * s2)
        top = int(np.argmax(ll))
        shift = gamma * (bound + bound[top])
        # A particle whose weight stays below e^-700 of the largest even at the far end of its bound cannot move a CDF
        # step by more than 1e-304: hopeless particles (huge residuals, hence huge S2 and a loose bound) do not count.
        live = gamma * (ll - ll[top]) + shift >= -700.0
        if np.all(shift[live] < 1.0):
            band = 2.0 * float(np.sum(w[live] * np.expm1(shift[live]))) + 8.0 * n_samples * EPS
            if not draws_inside_band(cdf, u, idx, band):
                return idx
            STATS["why_in_band"] += 1
        else:
            STATS["why_loose_bound"] += 1
    else:
        STATS["why_nonfinite"] += 1
    # inside the band (or a non-finite likelihood): the reference's own expression for the weights, same draws
    STATS["redecided"] += 1
    w = stage_weights(distinct_rows_log_likelihood(particles, obs_wp, obs_vals, lambda_b, M), gamma, n_samples)
    return choice_from_draws(w, u)[1]


def distinct_rows_log_likelihood(particles: np.ndarray, obs_wp, obs_vals, lambda_b: float, M: int) -> np.ndarray:
    """batched_poisson_log_likelihood evaluated once per DISTINCT particle.  The stages that get re-decided are those
    of a nearly collapsed cloud -- resampling has left a few distinct rows, each many times over (measured at the
    reference's example size: 2 to 142 distinct rows among 1000, log-likelihoods within 1e-8 of each other, which is
    exactly why a draw can land inside the band) -- and the reference's expression is element-wise per row with
    per-row reductions, so equal rows get equal values: the result is the full evaluation's, bit for bit, at the cost
    of the distinct rows only."""
    particles = np.asarray(particles, dtype=float)
    if particles.ndim == 2 and particles.shape[0] > 1 and np.all(np.isfinite(particles)):
        uniq, inv = np.unique(particles, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        if uniq.shape[0] < particles.shape[0] and np.array_equal(uniq[inv], particles):
            STATS["distinct_rows"] = STATS.get("distinct_rows", 0) + int(uniq.shape[0])
            STATS["rows"] = STATS.get("rows", 0) + int(particles.shape[0])
            return batched_poisson_log_likelihood(uniq, obs_wp, obs_vals, lambda_b, M)[inv]
    STATS["distinct_rows"] = STATS.get("distinct_rows", 0) + int(particles.shape[0])
    STATS["rows"] = STATS.get("rows", 0) + int(particles.shape[0])
    return batched_poisson_log_likelihood(particles, obs_wp, obs_vals, lambda_b, M)


_MVN_DIRECT = None   # None: untested; True / False after the self-check


def _mvn_direct_ok() -> bool:
    """scipy.stats.multivariate_normal.rvs(mean, cov, size) validates its arguments (an eigendecomposition of cov among
    other things: half of the call at these sizes) and then draws with the global RandomState's own
    multivariate_normal(mean, cov, size).  Checked once on this scipy / numpy: same values, same RNG position, for a
    regular, a rank-deficient and an all-zero covariance.  The global stream is restored afterwards."""
    global _MVN_DIRECT
    if _MVN_DIRECT is None:
        saved = np.random.get_state()
        try:
            rng = np.random.RandomState(7)
            ok = True
            for dim, kind in ((3, "full"), (6, "deficient"), (9, "zero"), (3, "zero")):
                A = rng.randn(dim, dim if kind == "full" else 2)
                cov = np.zeros((dim, dim)) if kind == "zero" else np.cov(A @ rng.randn(A.shape[1], 50))
                mean = rng.uniform(0, 100, dim)
                np.random.seed(11)
                a = multivariate_normal.rvs(mean=mean, cov=cov * 0.001, size=40)
                ra = np.random.uniform(size=2)
                np.random.seed(11)
                b = np.random.multivariate_normal(mean, cov * 0.001, 40)
                rb = np.random.uniform(size=2)
                ok = ok and a.shape == b.shape and np.array_equal(a, b) and np.array_equal(ra, rb)
            _MVN_DIRECT = bool(ok)
        except Exception:
            _MVN_DIRECT = False
        finally:
            np.random.set_state(saved)
    return _MVN_DIRECT


def perturb(mean: np.ndarray, cov: np.ndarray, n_samples: int) -> np.ndarray:
    """multivariate_normal.rvs(mean=mean, cov=cov, size=n_samples) of estimation.py:122: through scipy, or -- where the
    self-check has shown it to be the same draw -- straight from numpy's generator.  Anything scipy's validation would
    reject (non-finite entries, a single sample whose output scipy squeezes) still goes through scipy."""
    if n_samples > 1 and _mvn_direct_ok() and np.all(np.isfinite(cov)) and np.all(np.isfinite(mean)):
        return np.random.multivariate_normal(mean, cov, n_samples)
    return multivariate_normal.rvs(mean=mean, cov=cov, size=n_samples)


def importance_sampling_with_progressive_correction(obs_wp, obs_vals, lambda_b: float, M: int, n_samples: int, s_stages: int,
                                                    prior_dist: List, scenario, alpha: float = 0.001, scorer=None,
                                                    prepared=None) -> Tuple[np.ndarray, np.ndarray]:
    """estimation.py:57-130; RNG order per stage: np.random.choice, then multivariate_normal.rvs."""
    scorer = _default_scorer() if scorer is None else scorer
    prepared = scorer.prepare_counts(obs_wp, obs_vals) if prepared is None else prepared
    gammas = np.linspace(0.1, 1.0, s_stages)
    particles = np.column_stack([d.rvs(n_samples) for d in prior_dist])
    ws, ir = scenario.workspace_size, scenario.intensity_range
    for gamma in gammas:
        picked = particles[resample_indices(particles, gamma, n_samples, obs_wp, obs_vals, lambda_b, M, scorer, prepared)]
        mean = np.mean(picked, axis=0)
        cov = np.cov(picked, rowvar=False)
        moved = perturb(mean, cov * alpha, n_samples)
        moved[:, 0] = np.clip(moved[:, 0], ws[0], ws[1])
        moved[:, 1] = np.clip(moved[:, 1], ws[2], ws[3])
        moved[:, 2] = np.clip(moved[:, 2], ir[0], ir[1])
        particles = moved
    return np.mean(particles, axis=0), particles


def calculate_bic(log_likelihood: float, num_params: int, num_data_points: int) -> float:
    """estimation.py:132-136 (sign convention of the reference: larger is better)."""
    return 2 * log_likelihood + num_params * np.log(num_data_points)


_SCORER = None


def _default_scorer():
    """The device scorer (raises without CUDA or without libipp_b200.so: there is no CPU route through here)."""
    global _SCORER
    if _SCORER is None:
        from mripp_b200.scoring import GpuScorer
        _SCORER = GpuScorer()
    return _SCORER


def estimate_sources_bayesian(obs_wp, obs_vals, lambda_b: float, max_sources: int, n_samples: int, s_stages: int, scenario,
                              scorer=None) -> Tuple[np.ndarray, int, float]:
    """estimation.py:138-189: model selection over M = 1..max_sources by BIC."""
    scorer = _default_scorer() if scorer is None else scorer
    prepared = scorer.prepare_counts(obs_wp, obs_vals)
    ws, ir = scenario.workspace_size, scenario.intensity_range
    best = (None, 0, -np.inf)
    # scipy's uniform(loc, scale): the reference passes the upper bounds as `scale` (estimation.py:171-173).  The
    # frozen distributions hold no state of their own (they draw from the global RNG), so the three the reference
    # rebuilds for every M are built once.
    prior3 = [uniform(loc=ws[0], scale=ws[1]), uniform(loc=ws[2], scale=ws[3]), uniform(loc=ir[0], scale=ir[1])]
    for M in range(1, max_sources + 1):
        priors = prior3 * M
        est, _ = importance_sampling_with_progressive_correction(obs_wp, obs_vals, lambda_b, M, n_samples, s_stages, priors,
                                                                 scenario, scorer=scorer, prepared=prepared)
        bic = calculate_bic(poisson_log_likelihood(est, obs_wp, obs_vals, lambda_b, M), 3 * M, len(obs_vals))
        if bic > best[2]:
            best = (est, M, bic)
    return best

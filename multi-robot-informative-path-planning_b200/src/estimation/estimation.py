"""Multi-source parameter estimation (Morelande / Ristic / Gunatilaka importance sampling with
progressive correction): host-side mirror of the reference's src/estimation/estimation.py.

It is NOT on the accelerated path this round: the resampling and
perturbation steps consume the global RNG in a fixed order that the waypoint sequences depend on,
so it stays on the host.  The only restructuring is that the Poisson log-likelihood of all
`n_samples` particles is evaluated in ONE array expression instead of a Python loop over particles
(estimation.py:97) -- element for element the same numpy / scipy calls, hence the same bits.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
from scipy.stats import multivariate_normal, poisson, uniform


def poisson_log_likelihood(theta: np.ndarray, obs_wp: np.ndarray, obs_vals: np.ndarray, lambda_b: float, M: int) -> float:
    """estimation.py:15-55 for one parameter vector."""
    return float(batched_poisson_log_likelihood(np.asarray(theta, dtype=float)[None, :], obs_wp, obs_vals, lambda_b, M)[0])


def batched_poisson_log_likelihood(thetas: np.ndarray, obs_wp, obs_vals, lambda_b: float, M: int) -> np.ndarray:
    """Log-likelihood of every row of `thetas` (N x 3M): rate = lambda_b + sum_i A_i / max(d2, 1e-6)."""
    obs_wp = np.array(obs_wp)
    counts = np.round(obs_vals).astype(int)
    src = np.asarray(thetas, dtype=float).reshape(thetas.shape[0], M, 3)
    d2 = (obs_wp[None, :, None, 0] - src[:, None, :, 0]) ** 2 + (obs_wp[None, :, None, 1] - src[:, None, :, 1]) ** 2
    rate = lambda_b + np.sum(src[:, None, :, 2] / np.maximum(d2, 1e-6), axis=2)
    rate = np.maximum(rate, 1e-10)
    return np.sum(poisson.logpmf(counts[None, :], rate), axis=1)


def importance_sampling_with_progressive_correction(obs_wp, obs_vals, lambda_b: float, M: int, n_samples: int, s_stages: int,
                                                    prior_dist: List, scenario, alpha: float = 0.001
                                                    ) -> Tuple[np.ndarray, np.ndarray]:
    """estimation.py:57-130; RNG order per stage: np.random.choice, then multivariate_normal.rvs."""
    gammas = np.linspace(0.1, 1.0, s_stages)
    particles = np.column_stack([d.rvs(n_samples) for d in prior_dist])
    ws, ir = scenario.workspace_size, scenario.intensity_range
    for gamma in gammas:
        log_w = gamma * batched_poisson_log_likelihood(particles, obs_wp, obs_vals, lambda_b, M)
        w = np.exp(log_w - np.max(log_w))
        w /= np.sum(w)
        if np.any(np.isnan(w)):
            print("Warning: NaN weights encountered. Returning uniform weights.")
            w = np.full(n_samples, 1 / n_samples)
        picked = particles[np.random.choice(n_samples, size=n_samples, p=w)]
        mean = np.mean(picked, axis=0)
        cov = np.cov(picked, rowvar=False)
        moved = multivariate_normal.rvs(mean=mean, cov=cov * alpha, size=n_samples)
        moved[:, 0] = np.clip(moved[:, 0], ws[0], ws[1])
        moved[:, 1] = np.clip(moved[:, 1], ws[2], ws[3])
        moved[:, 2] = np.clip(moved[:, 2], ir[0], ir[1])
        particles = moved
    return np.mean(particles, axis=0), particles


def calculate_bic(log_likelihood: float, num_params: int, num_data_points: int) -> float:
    """estimation.py:132-136 (sign convention of the reference: larger is better)."""
    return 2 * log_likelihood + num_params * np.log(num_data_points)


def estimate_sources_bayesian(obs_wp, obs_vals, lambda_b: float, max_sources: int, n_samples: int, s_stages: int, scenario
                              ) -> Tuple[np.ndarray, int, float]:
    """estimation.py:138-189: model selection over M = 1..max_sources by BIC."""
    ws, ir = scenario.workspace_size, scenario.intensity_range
    best = (None, 0, -np.inf)
    for M in range(1, max_sources + 1):
        # scipy's uniform(loc, scale): the reference passes the upper bounds as `scale` (estimation.py:171-173)
        priors = [uniform(loc=ws[0], scale=ws[1]), uniform(loc=ws[2], scale=ws[3]), uniform(loc=ir[0], scale=ir[1])] * M
        est, _ = importance_sampling_with_progressive_correction(obs_wp, obs_vals, lambda_b, M, n_samples, s_stages, priors,
                                                                 scenario)
        bic = calculate_bic(poisson_log_likelihood(est, obs_wp, obs_vals, lambda_b, M), 3 * M, len(obs_vals))
        if bic > best[2]:
            best = (est, M, bic)
    return best

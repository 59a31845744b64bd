"""CPU restatement of the reference's source-estimator likelihood (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this).

Follows /root/reference/src/estimation/estimation.py:
  poisson_log_likelihood   :15-55   one parameter vector, numpy + scipy.stats.poisson.logpmf
  particle_log_likelihoods :97      the Python loop over the n_samples particles inside calc_weights
Pinned against tests/golden/field_golden.npz (ll_thetas / ll_values, est_*: produced by the unmodified reference,
oracle/make_golden.py gold_field) in tests/test_oracle_golden.py.
"""
import numpy as np
from scipy.stats import poisson


def poisson_log_likelihood(theta, obs_wp, obs_vals, lambda_b, M):
    """estimation.py:15-55."""
    obs_wp = np.array(obs_wp)
    counts = np.round(obs_vals).astype(int)
    sources = np.asarray(theta, dtype=float).reshape((M, 3))
    d_ji = (obs_wp[:, None, 0] - sources[:, 0]) ** 2 + (obs_wp[:, None, 1] - sources[:, 1]) ** 2
    lam = lambda_b + np.sum(sources[:, 2] / np.maximum(d_ji, 1e-6), axis=1)
    lam = np.maximum(lam, 1e-10)
    return np.sum(poisson.logpmf(counts, lam))


def particle_log_likelihoods(thetas, obs_wp, obs_vals, lambda_b, M):
    """estimation.py:97: one call per particle."""
    return np.array([poisson_log_likelihood(thetas[i], obs_wp, obs_vals, lambda_b, M) for i in range(len(thetas))])


class HostLikelihood:
    """The likelihood backend interface of the product's sampler (prepare_counts / particle_log_likelihood) on the
    restatement above; S1 = S2 = None marks the values as the reference's own (no error band)."""

    launches = 0

    def prepare_counts(self, obs_wp, obs_vals):
        return {"obs_wp": obs_wp, "obs_vals": obs_vals}

    def particle_log_likelihood(self, thetas, prepared, lambda_b, M):
        return particle_log_likelihoods(np.asarray(thetas, float), prepared["obs_wp"], prepared["obs_vals"], lambda_b, M), None, None

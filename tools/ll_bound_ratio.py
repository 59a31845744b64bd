"""How much of the sampler's error bound the device log-likelihoods actually use: max |ll_device - ll_reference| / bound
over particle clouds of the sizes a run sees (src/estimation/estimation.py: C_TERM, c_sum)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mripp_b200.scoring import GpuScorer
from mripp_b200.src.estimation import estimation as E
sc = GpuScorer()
worst = 0.0
for n, N, M, seed in ((500, 200, 1, 1), (2000, 200, 3, 2), (5000, 1000, 2, 3), (15000, 1000, 3, 4), (6000, 1000, 3, 5)):
    rng = np.random.RandomState(seed)
    obs = rng.uniform(0, 40, (n, 2))
    src = np.array([[5.0, 35.0, 1e6], [35.0, 35.0, 1e6]])
    lam = 1.0 + sum(a / np.maximum((obs[:, 0] - x) ** 2 + (obs[:, 1] - y) ** 2, 1e-6) for x, y, a in src)
    meas = rng.poisson(lam).astype(float)
    th = rng.uniform(0, 40, (N, 3 * M))
    th[:, 2::3] = rng.uniform(1e5, 2e6, (N, M))
    th[: N // 2, :3] = src[0] + rng.normal(0, 0.3, (N // 2, 3)) * np.array([1, 1, 1e4])   # half of the cloud near a true source
    prep = sc.prepare_counts(list(obs), list(meas))
    ll, s1, s2 = sc.particle_log_likelihood(th, prep, 1.0, M)
    want = E.batched_poisson_log_likelihood(th, list(obs), list(meas), 1.0, M)
    bound = E.EPS * (E.C_TERM * s1 + E.c_sum(n) * s2)
    r = np.abs(ll - want) / bound
    worst = max(worst, float(r.max()))
    print(f"n_obs {n} particles {N} sources {M}: max |d ll| / bound {r.max():.3f}, median {np.median(r):.3f}, "
          f"term share of the bound {float(np.median(E.C_TERM * s1 / (E.C_TERM * s1 + E.c_sum(n) * s2))):.2f}")
print(f"C_TERM {E.C_TERM}: worst ratio {worst:.3f}")

"""Accuracy of the fitted state in Morton order vs the given order, against the CPU oracle, over regimes where the
block-sparse path switches on.  Prints max |d sigma|, max |d var| / (c y_std^2), max |d mean|."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
from oracle import gp_oracle as O
from bench import synth_observations
for n, W, l in [(2000, 70, 2.0), (5000, 70, 2.0), (5000, 100, 2.8), (3000, 100, 1.0), (5000, 100, 2.0), (1000, 70, 2.0)]:
    X, meas, _ = synth_observations(n, W, 1)
    y = np.log10(meas)
    q = np.random.RandomState(2).uniform(0, W, (1200, 2)); q[:300] = X[:300] + 1e-3
    st = O.fit(X, y, 1.0, l, optimize=False)
    mo, so, _ = O.predict(st, q)
    res = {}
    for mode in ("curve", "shuffled", "maximin", "given"):
        gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, l), optimizer=None, normalize_y=True)
        gp.cluster_order = mode
        if mode == "given":
            gp.skip_below = 0
        gp.fit(X, y)
        m, s = gp.predict(q, return_std=True)
        res[mode] = (bool(gp._bbox_d is not None), float(np.abs(s - so).max()), float(np.abs(s**2 - so**2).max() / st["y_std"]**2),
                     float(np.abs(m - mo).max()))
    print(f"n={n} W={W} l={l}: " + " | ".join(f"{k}: dsig {v[1]:.1e} dvar {v[2]:.1e} dmean {v[3]:.1e}" for k, v in res.items()), flush=True)

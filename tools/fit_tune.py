"""Full 11-restart fit vs number of restart streams.  python tools/fit_tune.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
from bench import synth_observations
for n, W, opts in [(131, 40, (1, 4, 11)), (500, 20, (1, 4, 6, 11)), (2000, 50, (1, 3, 4, 6, 11)), (5000, 100, (1, 2, 3, 4, 6))]:
    X, meas, _ = synth_observations(n, W, 0)
    y = np.log10(meas)
    for streams in opts:
        gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), n_restarts_optimizer=10, normalize_y=True)
        gp.restart_streams = streams
        if n < 5000:
            np.random.seed(0); gp.fit(X, y)
        np.random.seed(0); torch.cuda.synchronize(); t = time.perf_counter()
        gp.fit(X, y); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"n={n} streams={streams}: fit {dt:.3f} s  ({gp.log_marginal_likelihood_value_:.4f})", flush=True)
        del gp
        torch.cuda.empty_cache()

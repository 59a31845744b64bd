"""Full 11-restart fit, serial vs concurrent restart streams.  python tools/fit_timing.py n [n ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
from bench import synth_observations
for n in [int(a) for a in sys.argv[1:]] or [131, 500, 2000]:
    W = {131: 40, 500: 20, 2000: 50, 5000: 100}.get(n, 30)
    X, meas, _ = synth_observations(n, W, 0)
    y = np.log10(meas)
    for streams in (1, None):
        gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), n_restarts_optimizer=10, normalize_y=True)
        gp.restart_streams = streams
        np.random.seed(0); gp.fit(X, y)  # warm-up (allocations, module load)
        np.random.seed(0); torch.cuda.synchronize(); t = time.perf_counter()
        gp.fit(X, y); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"n={n} streams={streams}: fit {dt:.3f} s, {gp.n_lml_evals_ // 2} evals, theta {gp.kernel_.theta}, lml {gp.log_marginal_likelihood_value_:.6f}", flush=True)

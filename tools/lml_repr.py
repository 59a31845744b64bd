"""Prints LML + gradient of fixed hyper-parameters at full precision (hex) for several sizes: run once per setting of
IPP_GEMM_MB / IPP_CHOL_WS / IPP_CHOL_RED (they are read once per process) and diff the outputs -- the pipeline variants
must agree bit for bit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_observations
from mripp_b200 import gp as G
for n in (300, 1700, 2000, 5000):
    W = {500: 20, 2000: 50, 5000: 100}.get(n, max(10, int(np.sqrt(n) * 1.4)))
    X, meas, _ = synth_observations(n, W, 0)
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, np.log10(meas))
    for th in (np.log([1.3, 1.7]), np.log([0.7, 3.1])):
        lml, gr = gp.log_marginal_likelihood(th, eval_gradient=True)
        print(n, float(lml).hex(), float(gr[0]).hex(), float(gr[1]).hex())
    print(n, "alpha", float(np.sum(gp.alpha_)).hex(), "L", float(np.sum(gp.L_)).hex())

"""Small fixed workload for ncu: one fit_final + LML evals + one grid posterior.  Args: n W G n_lml [Gy [m_points]]
(m_points > 0: one dense tabled explicit-point posterior of that many random queries)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
n, W, Gx, n_lml = (int(a) for a in sys.argv[1:5])
Gy = int(sys.argv[5]) if len(sys.argv) > 5 else Gx  # optional: fewer lattice rows (a slab) to keep ncu captures short
rng = np.random.RandomState(0)
X = rng.uniform(0, W, (n, 2)); X[-n // 5:] = X[:n // 5]
y = np.sin(X[:, 0] / 3) + 0.1 * X[:, 1] + 1e-3 * rng.randn(n)
gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, y)
th = np.log([1.0, 2.0])
for _ in range(n_lml):
    gp.log_marginal_likelihood(th, eval_gradient=True)
if n_lml:
    G.gram_device(X, None, 1.0, 1.0)  # the public symmetric Gram (gp.kernel(X)): HBM-write bound
if Gx:
    gp.predict_grid_device(0, W, Gx, 0, W * Gy / Gx, Gy)
if len(sys.argv) > 6 and int(sys.argv[6]) > 0:
    Qd = torch.from_numpy(rng.uniform(0, W, (int(sys.argv[6]), 2))).cuda()
    gp.predict_points_device(Qd, tabled=True, skip_below=0)
torch.cuda.synchronize()
print("ok")

"""Fixed workload for ncu: batched LML + gradient evaluations (the lockstep fit's launch sequence).  Args: n R reps"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_observations
from mripp_b200 import gp as G
n, R, reps = (int(a) for a in sys.argv[1:4])
W = {500: 20, 2000: 50, 5000: 100}.get(n, 30)
X, meas, _ = synth_observations(n, W, 0)
gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, np.log10(meas))
ws = G._Workspace().ensure(n, slots=R)
rng = np.random.RandomState(1)
thetas = [np.array([rng.uniform(-1, 1), rng.uniform(0, 1.5)]) for _ in range(R)]
for _ in range(reps):
    gp._eval_batch(ws, list(range(R)), thetas, True)
torch.cuda.synchronize()
print("ok")

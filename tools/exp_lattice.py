"""Kernel experiment driver: time the lattice posterior (one tile per CTA) under a given libipp variant.
IPP_B200_LIB=<path> python tools/exp_lattice.py n nx ny"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
n, nx, ny = (int(a) for a in sys.argv[1:4])
W = 100
rng = np.random.RandomState(0)
X = rng.uniform(0, W, (n, 2)); X[-n // 5:] = X[:n // 5]
y = np.sin(X[:, 0] / 3) + 0.1 * X[:, 1] + 1e-3 * rng.randn(n)
gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, y)
for _ in range(2):
    gp.predict_grid_device(0, W, nx, 0, W, ny)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    gp.predict_grid_device(0, W, nx, 0, W, ny)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
NP = gp._ws.layout[0]
print(os.environ.get("IPP_B200_LIB", "default").split("/")[-1], f"n={n} G={nx*ny} ms={ms:.3f} TF={nx*ny*(float(n)**2)/ms/1e9:.2f}")

"""End-to-end predict_spatial_field at the reference's everyday sizes (fixed theta): where the milliseconds go."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
from mripp_b200.src.point_source.point_source import PointSourceField
from bench import synth_observations
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
for n, W in [(131, 40), (500, 20), (500, 40), (2000, 50)]:
    X, meas, _ = synth_observations(n, W, 0)
    f = PointSourceField(num_sources=3, workspace_size=(0, W, 0, W), seed=1)
    f.gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True)
    wp = [X[i] for i in range(n)]
    ylog = np.log10(meas)
    nx = W * 10
    print(f"n={n} grid {nx}x{nx}: predict_spatial_field {t(lambda: f.predict_spatial_field(wp, meas)):.3f} ms | "
          f"fit(list) {t(lambda: f.gp.fit(wp, ylog)):.3f} | fit(array) {t(lambda: f.gp.fit(X, ylog)):.3f} | "
          f"grid_host {t(lambda: f.gp.predict_grid_host(0, W, nx, 0, W, nx)):.3f} | grid_device {t(lambda: f.gp.predict_grid_device(0, W, nx, 0, W, nx)):.3f} | "
          f"as_points(list) {t(lambda: G._as_points(wp)):.3f}", flush=True)

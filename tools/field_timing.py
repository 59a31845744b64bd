"""Timing of the section-8f kernels next to their host (reference-expression) counterparts.  Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.visualization.plot_helper import calculate_differential_entropy, field_metrics_device

obst = [{"type": "rectangle", "x": 4.0 + 4 * i, "y": 3.0 + 3 * (i % 3), "width": 2.0, "height": 1.5} for i in range(4)]
# host: the reference's expressions on a 20 x 20 workspace (200 x 200 lattice) -- the obstacle branch is a Python loop
t0 = time.perf_counter()
small = PointSourceField(num_sources=10, workspace_size=(0, 20, 0, 20), obstacles=obst, seed=4)  # ctor computes g_truth
t_host = time.perf_counter() - t0
dev = small.ground_truth_device()
print(f"ground truth 200x200, 10 sources, 4 rectangles: host {t_host:.2f} s ({small.g_truth.size / t_host:.0f} points/s), "
      f"max rel diff device vs host {np.abs(dev / small.g_truth - 1).max():.1e}")
big = PointSourceField(num_sources=10, workspace_size=(0, 200, 0, 200), obstacles=[], seed=4)  # 2000 x 2000, no host loop
big.obstacles = obst
for _ in range(2):
    big.ground_truth_device(as_tensor=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gt = big.ground_truth_device(as_tensor=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
G = gt.numel()
print(f"ground truth 2000x2000 (C5 lattice), 10 sources, 4 rectangles: device {ms:.3f} ms = {G / ms * 1e3:.3e} points/s "
      f"({8 * G / ms / 1e6:.0f} GB/s written)")
# fused metrics on the C5 lattice
rng = np.random.RandomState(0)
zt = gt.view(-1)
zp = zt * torch.exp(0.2 * torch.randn(G, dtype=torch.float64, device="cuda"))
sd = torch.rand(G, dtype=torch.float64, device="cuda") * 0.3
for _ in range(2):
    field_metrics_device(zp, sd, zt)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    m = field_metrics_device(zp, sd, zt)
torch.cuda.synchronize(); t_dev = (time.perf_counter() - t0) / 10
zph, sdh, zth = zp.cpu().numpy(), sd.cpu().numpy(), zt.cpu().numpy()
t0 = time.perf_counter()
err = np.log10(zth + 1) - np.log10(zph + 1)
ref = (np.sqrt(np.mean(err ** 2)), np.sqrt(np.sum(zth * err ** 2) / np.sum(zth)), calculate_differential_entropy(sdh))
t_np = time.perf_counter() - t0
t0 = time.perf_counter(); zp.cpu(); sd.cpu(); torch.cuda.synchronize(); t_d2h = time.perf_counter() - t0
print(f"grid metrics over {G} points: device {t_dev * 1e3:.3f} ms ({24 * G / t_dev / 1e9:.0f} GB/s read incl. launch + 32 B read-back), "
      f"numpy {t_np * 1e3:.1f} ms + D2H of Z and std {t_d2h * 1e3:.1f} ms; rel diff {max(abs(a / b - 1) for a, b in zip(m, ref)):.1e}")

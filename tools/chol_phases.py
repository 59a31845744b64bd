"""Where the cooperative Cholesky spends CTA 0's cycles (clock64 counters the kernel leaves in vbuf).
Usage: python tools/chol_phases.py n [n ...]   (run under gpurun)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G

for n in [int(a) for a in sys.argv[1:]] or [500, 2000, 5000]:
    W = max(10, int(np.sqrt(n) * 1.4))
    rng = np.random.RandomState(0)
    X = rng.uniform(0, W, (n, 2)); X[-n // 5:] = X[:n // 5]
    y = np.sin(X[:, 0] / 3) + 0.1 * X[:, 1] + 1e-3 * rng.randn(n)
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, y)
    th = np.log([1.0, 2.0])
    for _ in range(2):
        gp.log_marginal_likelihood(th, eval_gradient=False)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        gp.log_marginal_likelihood(th, eval_gradient=False)
    torch.cuda.synchronize(); t_nograd = (time.perf_counter() - t) / 5
    lay = gp._ws.layout
    NP = lay[0]; nb = NP // 64
    ntile = (nb * (nb + 1)) // 2
    off = lay[8] - 2 * ntile
    ph10 = gp._ws.v[off:off + 10].cpu().numpy() / 1.965e3  # us at 1965 MHz
    tl = gp._ws.v[off + 10:off + 13].cpu().numpy()
    dg = gp._ws.v[off + 13:off + 17].cpu().numpy() / 1.965e3 / max(nb - 1, 1)
    print(f"   CTA0 diagonal block per panel: loads {dg[0]:.1f} us, DMMA update {dg[1]:.1f} us, potf2 {dg[2]:.1f} us, store {dg[3]:.1f} us")
    if tl[2] > 0:
        print(f"   CTA1 trailing tiles: {int(tl[2])} tiles, mainloop {tl[0] / tl[2] / 1.965e3:.1f} us, C read-modify-write {tl[1] / tl[2] / 1.965e3:.1f} us per tile")
    ph = ph10[:5]
    print("   CTA1 per panel: " + ", ".join(f"{v/nb:.1f}" for v in ph10[5:]))
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        gp.log_marginal_likelihood(th, eval_gradient=True)
    torch.cuda.synchronize(); t_grad = (time.perf_counter() - t) / 5
    names = ["barrier1", "panel_solve", "barrier2", "diag_block", "trailing"]
    print(f"n={n} NP={NP} panels={nb}: lml(no grad) {t_nograd*1e3:.3f} ms, lml+grad {t_grad*1e3:.3f} ms; CTA0 us total "
          + ", ".join(f"{k}={v:.0f}" for k, v in zip(names, ph)) + " | per panel " + ", ".join(f"{v/nb:.1f}" for v in ph), flush=True)

"""GPU first-light check: GP kernels vs the numpy oracle + raw timings.  Run under gpurun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200 import gp as G
from oracle import gp_oracle as O

def make(n, W, seed=0, dup=0.2):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0, W, (n, 2))
    k = int(dup * n)
    if k: X[-k:] = X[:k]
    src = np.array([[0.3*W, 0.6*W, 5e5], [0.7*W, 0.2*W, 8e5], [0.5*W, 0.5*W, 2e5]])
    d2 = ((X[:, None, :] - src[None, :, :2])**2).sum(-1)
    inten = (src[None, :, 2] / (4*np.pi*np.maximum(d2, 0.25))).sum(1)
    meas = np.maximum(inten * (1 + 1e-3 * rng.randn(n)), 1e-6)
    return X, np.log10(meas)

def ev(t0=None):
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

out = {}
for n, W in [(40, 10), (64, 10), (131, 40), (500, 20), (1000, 30), (2000, 50)]:
    X, y = make(n, W)
    kern = G.ConstantRBFKernel(1.0, 1.0)
    gp = G.GaussianProcessRegressor(kernel=kern, n_restarts_optimizer=0, normalize_y=True, optimizer=None)
    res = {}
    for c0, l0 in [(1.0, 1.0), (2.5, 3.0), (0.7, 0.3)]:
        gp.kernel = G.ConstantRBFKernel(c0, l0)
        gp.fit(X, y)
        st = O.fit(X, y, c0, l0, optimize=False)
        th = np.log([c0, l0])
        lml_g, grad_g = gp.log_marginal_likelihood(th, eval_gradient=True)
        lml_o, grad_o = O.lml(th, X, st["yn"])
        L = gp.L_; a = gp.alpha_
        rel_L = np.abs(L - st["L"]).max() / np.abs(st["L"]).max()
        rel_a = np.abs(a - st["alpha"]).max() / np.abs(st["alpha"]).max()
        q = np.random.RandomState(1).uniform(0, W, (3000, 2))
        kq = min(200, n); q[:kq] = X[:kq] + 1e-3
        mg, sg = gp.predict(q, return_std=True)
        mo, so, _ = O.predict(st, q)
        res[f"c{c0}_l{l0}"] = dict(lml_rel=abs(lml_g - lml_o) / abs(lml_o), lml=lml_o, grad_g=list(grad_g), grad_o=list(grad_o),
                                   final_lml_rel=abs(gp.log_marginal_likelihood_value_ - lml_o) / abs(lml_o),
                                   rel_L=rel_L, rel_alpha=rel_a, amax=float(np.abs(st["alpha"]).max()),
                                   dmean=float(np.abs(mg - mo).max()), dstd=float(np.abs(sg - so).max()),
                                   dvar=float(np.abs(sg**2 - so**2).max()))
    out[f"n{n}"] = res
    print(n, json.dumps(res), flush=True)

# grid posterior vs oracle
X, y = make(500, 20)
gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.3, 2.0), optimizer=None, normalize_y=True).fit(X, y)
st = O.fit(X, y, 1.3, 2.0, optimize=False)
m, s, z = gp.predict_grid_device(0, 20, 200, 0, 20, 200)
_, _, r, shape = O.workspace_grid((0, 20, 0, 20))
mo, so, _ = O.predict(st, r)
print("grid", float(np.abs(m.cpu().numpy() - mo).max()), float(np.abs(s.cpu().numpy() - so).max()),
      float(np.abs(z.cpu().numpy() / 10**mo - 1).max()), flush=True)

# timings
tim = {}
for n, W, Gx in [(131, 40, 400), (500, 20, 200), (2000, 50, 500), (5000, 100, 1000)]:
    X, y = make(n, W)
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True)
    gp.fit(X, y)
    th = np.log([1.0, 2.0])
    for _ in range(2): gp.log_marginal_likelihood(th, eval_gradient=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    reps = 5
    for _ in range(reps): gp.log_marginal_likelihood(th, eval_gradient=True)
    torch.cuda.synchronize(); t_lml = (time.perf_counter() - t) / reps
    for _ in range(2): gp.predict_grid_device(0, W, Gx, 0, W, Gx)
    torch.cuda.synchronize()
    e0 = ev(); 
    for _ in range(3): gp.predict_grid_device(0, W, Gx, 0, W, Gx)
    e1 = ev(); torch.cuda.synchronize()
    t_post = e0.elapsed_time(e1) / 3 * 1e-3
    NP = (n + 63) // 64 * 64
    tim[f"n{n}"] = dict(lml_ms=t_lml * 1e3, lml_tflops=NP**3 / t_lml / 1e12, post_ms=t_post * 1e3, pts_per_s=Gx * Gx / t_post,
                        post_tflops=Gx * Gx * float(NP) ** 2 / t_post / 1e12)
    print(n, json.dumps(tim[f"n{n}"]), flush=True)
json.dump(dict(parity=out, timing=tim), open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w"), indent=1)

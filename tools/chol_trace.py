"""Timeline of one dataflow Cholesky launch (IPP_CHOL_STATS=2): how many CTAs are busy with which kind of task in every
time bucket, and when each super-panel's chain tasks run.  Args: n R [bucket_us]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["IPP_CHOL_STATS"] = "2"
import torch
from bench import synth_observations
from mripp_b200 import gp as G
n, R = int(sys.argv[1]), int(sys.argv[2])
bucket = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
W = {500: 20, 2000: 50, 5000: 100}.get(n, 30)
X, meas, _ = synth_observations(n, W, 0)
gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X, np.log10(meas))
ws = G._Workspace().ensure(n, slots=R)
rng = np.random.RandomState(1)
thetas = [np.array([rng.uniform(-1, 1), rng.uniform(0, 1.5)]) for _ in range(R)]
for _ in range(2):
    gp._eval_batch(ws, list(range(R)), thetas, False)
ev = G.chol_trace(ws, R)
print("events", len(ev), "span us", ev[:, 5].max() / 1e3, "ctas", len(np.unique(ev[:, 0])))
names = "FSCU"
T = ev[:, 5].max() / 1e3
nb = int(T / bucket) + 1
busy = np.zeros((4, nb))
for cta, q, r, i, t0, t1 in ev:
    a, b = t0 / 1e3, t1 / 1e3
    k0, k1 = int(a / bucket), int(b / bucket)
    for k in range(k0, k1 + 1):
        lo, hi = max(a, k * bucket), min(b, (k + 1) * bucket)
        if hi > lo:
            busy[q, k] += hi - lo
ncta = len(np.unique(ev[:, 0]))
print(f"bucket {bucket} us; busy fraction of {ncta} CTAs by queue F S C U | total")
for k in range(nb):
    f = busy[:, k] / (bucket * ncta)
    print(f"{k * bucket:8.0f} " + " ".join(f"{x:5.2f}" for x in f) + f" | {f.sum():5.2f}")
for q in range(4):
    d = (ev[ev[:, 1] == q][:, 5] - ev[ev[:, 1] == q][:, 4]) / 1e3
    if len(d):
        print(f"queue {names[q]}: {len(d)} tasks, mean {d.mean():.1f} us, p50 {np.median(d):.1f}, p90 {np.percentile(d, 90):.1f}, max {d.max():.1f}")
# the F chain of problem 0: when each diagonal block was factored
f0 = ev[(ev[:, 1] == 0) & (ev[:, 2] == 0)]
f0 = f0[np.argsort(f0[:, 3])]
print("F(k) end times of problem 0 (us), every 8th:", [(int(k), round(t / 1e3)) for k, t in zip(f0[::8, 3], f0[::8, 5])])
if os.environ.get("IPP_CHOL_WS", "1") != "0":  # the warp-specialised executor packs in-task marks into the index field
    for q in (2, 3):
        u = ev[ev[:, 1] == q]
        if len(u):
            w = u[:, 3]
            first, loop, epi = (w & 0x3ff) * 0.064, ((w >> 10) & 0x3ff) * 0.064, ((w >> 20) & 0x3ff) * 0.064
            tot = (u[:, 5] - u[:, 4]) / 1e3
            print(f"queue {names[q]} inside the task (us, mean): first slab landed {first.mean():.2f}, main loop done {loop.mean():.2f}, "
                  f"stores issued {epi.mean():.2f}, barrier passed {tot.mean():.2f}")
    # gap between consecutive tasks of one CTA
    gaps = []
    for c in np.unique(ev[:, 0]):
        e = ev[ev[:, 0] == c]
        e = e[np.argsort(e[:, 4])]
        gaps.append((e[1:, 4] - e[:-1, 5]) / 1e3)
    g = np.concatenate(gaps)
    print(f"gap between tasks of a CTA (us): median {np.median(g):.2f}, mean of those below 20 us {g[g < 20].mean():.2f}, share below 2 us {(g < 2).mean():.2f}")

"""Source estimator on the device: kernel time of ipp_poisson_loglik, one whole
estimate_sources_bayesian against the same sampler on the host expression, and how often the error band
re-decides a stage.  Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_observations
from mripp_b200.scoring import GpuScorer
from mripp_b200.src.estimation import estimation as E
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200 import _lib


class BatchedHost:
    """The one-expression host form (what the mirror ran before the kernel; faster than the reference's loop)."""
    def prepare_counts(self, obs_wp, obs_vals):
        return {"obs_wp": obs_wp, "obs_vals": obs_vals}
    def particle_log_likelihood(self, thetas, prepared, lambda_b, M):
        return E.batched_poisson_log_likelihood(thetas, prepared["obs_wp"], prepared["obs_vals"], lambda_b, M), None, None


sc = GpuScorer()
for n, N, M in [(500, 200, 3), (2000, 200, 3), (5000, 200, 3), (5000, 1000, 3)]:
    X, meas, src = synth_observations(n, 100, 0)
    prep = sc.prepare_counts(list(X), list(meas))
    rng = np.random.RandomState(1)
    P = np.column_stack([c for _ in range(M) for c in (rng.uniform(0, 100, N), rng.uniform(0, 100, N), rng.uniform(1e5, 1e6, N))])
    Pd = torch.from_numpy(P).cuda(); out = torch.empty((N, 3), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    def launch():
        lib.ipp_poisson_loglik(_lib.ptr(Pd), N, M, _lib.ptr(prep["obs"]), _lib.ptr(prep["counts"]), _lib.ptr(prep["lgam"]), n, 1.0,
                               _lib.ptr(out), _lib.stream_ptr())
    for _ in range(5): launch()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): launch()
    e1.record(); torch.cuda.synchronize(); k_us = e0.elapsed_time(e1) * 1e3 / 200
    for _ in range(3): sc.particle_log_likelihood(P, prep, 1, M)
    t0 = time.perf_counter()
    for _ in range(50): sc.particle_log_likelihood(P, prep, 1, M)
    call_us = (time.perf_counter() - t0) / 50 * 1e6
    t0 = time.perf_counter()
    for _ in range(3): E.batched_poisson_log_likelihood(P, list(X), list(meas), 1, M)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"loglik n={n} N={N} M={M}: kernel {k_us:.1f} us ({N * n * M / k_us * 1e6:.3e} particle-observation-source terms/s), "
          f"call incl. H2D/D2H {call_us:.0f} us, host expression {host_ms:.1f} ms", flush=True)

for n in (500, 2000, 5000):
    X, meas, src = synth_observations(n, 100, 0)
    scen = PointSourceField(num_sources=3, workspace_size=(0, 100, 0, 100), seed=1)
    res = {}
    for name, backend in (("device", sc), ("host", BatchedHost())):
        before = dict(E.STATS)
        for rep in range(2 if name == "device" else 1):
            np.random.seed(3)
            t0 = time.perf_counter()
            est, M, bic = E.estimate_sources_bayesian(list(X), list(meas), 1, 3, 200, 10, scen, scorer=backend)
            dt = time.perf_counter() - t0
        res[name] = (dt, est, M, bic, np.random.uniform(size=2))
        d = {k: E.STATS[k] - before[k] for k in E.STATS}
        print(f"estimate_sources_bayesian n={n} obs, 200 particles, 10 stages, M=1..3 [{name}]: {dt:.3f} s, M*={M}, bic={bic:.6e}, {d}", flush=True)
    same = np.array_equal(res["device"][1], res["host"][1]) and res["device"][2:4] == res["host"][2:4] and np.array_equal(res["device"][4], res["host"][4])
    print(f"  device == host bit for bit (estimate, M, BIC, RNG position): {same}; speed-up {res['host'][0] / res['device'][0]:.1f}x", flush=True)

"""BASELINE configs[0]: the reference's shipped example config at its stated size (budget 5000, d_wp 1.0, 1000
particles, 25 stages) through mripp_b200.main.

  python tools/stock_config.py CONFIG OUT.json [WALL_LIMIT_S] [profile|timers]

`profile` (default) runs the worker under cProfile (a third slower on this call-heavy host loop); `timers` only wraps
the four top-level phases of an iteration in perf_counter pairs, so `seconds` is the run's own wall time.

With a wall limit the run is sampled every 10 s (measurements taken, sampler stages, LML evaluations) and stopped at
the limit: the budget-5000 run spends its time in ~1000 calls of the 75-stage source estimator, far beyond a GPU-box
time slice, so the record is the progress curve plus a profile of where the time goes."""
import cProfile
import io
import json
import os
import pstats
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch  # noqa: E402

from mripp_b200 import exact as EXACT  # noqa: E402
from mripp_b200 import gp as G  # noqa: E402
from mripp_b200 import main as M  # noqa: E402
from mripp_b200.src.estimation import estimation as EST  # noqa: E402
from mripp_b200.src.point_source.point_source import PointSourceField  # noqa: E402

cfg, out = sys.argv[1], sys.argv[2]
limit = float(sys.argv[3]) if len(sys.argv) > 3 else None
mode = sys.argv[4] if len(sys.argv) > 4 else "profile"
count = {"measurements": 0, "calls": 0, "fits": 0}
_sim = PointSourceField.simulate_measurements
_fit = G.GaussianProcessRegressor.fit


def counted_sim(self, waypoints, noise_level=0.001):
    r = _sim(self, waypoints, noise_level)
    count["measurements"] += len(r)
    count["calls"] += 1
    return r


def counted_fit(self, X, y):
    count["fits"] += 1
    return _fit(self, X, y)


def counted_budget(self, path):
    r = _spent(self, path)
    count["budget_spent"] = count.get("budget_spent", 0.0) + float(r)
    return r


from mripp_b200.src.rrt import rrt as _R  # noqa: E402

_spent = _R.InformativeRRTBaseClass.calculate_budget_spent
_R.InformativeRRTBaseClass.calculate_budget_spent = counted_budget
count["budget_spent"] = 0.0
PointSourceField.simulate_measurements = counted_sim
G.GaussianProcessRegressor.fit = counted_fit
state = {"res": None, "err": None, "done": False}
prof = cProfile.Profile()
phase = {}


def timed_phase(owner, name, label):
    """Wall time and call count of one phase (class or module attribute) without a profiler."""
    fn = getattr(owner, name)

    def wrapper(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            rec_ = phase.setdefault(label, [0, 0.0])
            rec_[0] += 1
            rec_[1] += time.perf_counter() - t

    setattr(owner, name, wrapper)


if mode == "timers":
    from mripp_b200.src.rrt import rrt as R  # noqa: E402
    from mripp_b200.src.rrt.rrt_utils import TreeArrays  # noqa: E402

    timed_phase(TreeArrays, "grow_native", "tree growth (ipp_host_tree_grow)")
    timed_phase(R, "_score_tree", "tree scoring (device)")
    timed_phase(R, "informative_source_metric_path_selection", "stage-2 path selection")
    timed_phase(R, "bias_beta_path_selection", "stage-1 path selection")
    timed_phase(R, "estimate_sources_bayesian", "source estimator")  # rrt.py binds the name at import
    timed_phase(EST, "batched_poisson_log_likelihood", "  of which host likelihood (re-decided stages + BIC values)")
    timed_phase(PointSourceField, "predict_spatial_field", "predict_spatial_field (lazy: deferred until read)")


def work():
    try:
        if mode != "timers":
            prof.enable()
        state["res"] = M.main(["-config", cfg, "--out", out + ".metrics"])
    except BaseException as e:  # noqa: BLE001
        state["err"] = repr(e)
    finally:
        prof.disable()
        state["done"] = True


t0 = time.perf_counter()
th = threading.Thread(target=work, daemon=True)
th.start()
curve = []
while not state["done"]:
    th.join(timeout=10.0)
    now = time.perf_counter() - t0
    curve.append({"t": round(now, 1), **count, "sampler_stages": EST.STATS["stages"], "sampler_redecided": EST.STATS["redecided"]})
    print(json.dumps(curve[-1]), flush=True)
    if limit is not None and now > limit:
        break
buf = io.StringIO()
try:
    pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(30)
except Exception as e:  # noqa: BLE001 -- the profiler is still attached to the worker thread when stopped early
    buf.write(f"(no profile: {e!r})")
cfgd = json.load(open(cfg))
rec = {"config": cfg, "args": cfgd["args"], "finished": state["done"], "error": state["err"], "seconds": time.perf_counter() - t0,
       "results": state["res"], "progress": curve, "sampler_stats": dict(EST.STATS), "selector_stats": dict(EXACT.STATS),
       "peak_device_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "mode": mode,
       "phases": {k: {"calls": v[0], "seconds": round(v[1], 3)} for k, v in phase.items()},
       "profile_top": buf.getvalue()[:8000] if mode != "timers" else None}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps({k: rec[k] for k in ("finished", "seconds", "results", "sampler_stats")})[:1500], flush=True)
os._exit(0)

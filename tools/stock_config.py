"""BASELINE configs[0]: the reference's shipped example config at its stated size (budget 5000, d_wp 1.0, 1000
particles, 25 stages) through mripp_b200.main.

  python tools/stock_config.py CONFIG OUT.json [WALL_LIMIT_S]

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


PointSourceField.simulate_measurements = counted_sim
G.GaussianProcessRegressor.fit = counted_fit
state = {"res": None, "err": None, "done": False}
prof = cProfile.Profile()


def work():
    try:
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
       "peak_device_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "profile_top": buf.getvalue()[:8000]}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps({k: rec[k] for k in ("finished", "seconds", "results", "sampler_stats")})[:1500], flush=True)
os._exit(0)

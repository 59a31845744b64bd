"""BASELINE configs[0]: the reference's shipped example config at its stated size (budget 5000, d_wp 1.0, 1000
particles, 25 stages) through mripp_b200.main.  Usage: python tools/stock_config.py config/reference_example_setup[_lazy].json out.json"""
import cProfile
import io
import json
import os
import pstats
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch  # noqa: E402

from mripp_b200 import exact as EXACT  # noqa: E402
from mripp_b200 import main as M  # noqa: E402
from mripp_b200.src.estimation import estimation as EST  # noqa: E402

cfg, out = sys.argv[1], sys.argv[2]
prof = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
prof.enable()
res = M.main(["-config", cfg, "--out", out + ".metrics"])
prof.disable()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
buf = io.StringIO()
pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(25)
rec = {"config": cfg, "seconds": dt, "results": res, "sampler_stats": dict(EST.STATS), "selector_stats": dict(EXACT.STATS),
       "peak_device_memory_GB": torch.cuda.max_memory_allocated() / 1e9}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec)[:2000])
print(buf.getvalue()[:6000])

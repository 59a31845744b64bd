"""bench.py's reference arm runs on a CPU-only box and prints the JSON line of the contract (the B200 arm needs a
device and is exercised on the GPU box; its line is committed under profiles/ and checked here for the same keys)."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "grid points/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_committed_b200_line_has_the_contract_keys():
    d = json.loads(open(os.path.join(ROOT, "profiles", "bench_c4_r02.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    # the end-to-end number is the STOCK call (11-restart fit included), with the fixed-theta figure as a sub-key
    assert "STOCK" in d["e2e"]["call"] and d["e2e"]["lml_evals_per_call"] > 100 and d["e2e"]["fixed_theta"]["value"] > d["e2e"]["value"]
    assert d["value_at_theta_star"] > 0 and d["value_dense"] > 0 and d["fit"]["batch_of_11"]["frac_of_fp64_peak"] > 0.5

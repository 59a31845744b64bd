"""How many objective evaluations the UNMODIFIED reference's stock fit (PointSourceField.update -> sklearn
GaussianProcessRegressor.fit, n_restarts_optimizer=10: point_source.py:71,337) spends on bench.py's synthetic
workloads.  TEST / BENCH INFRASTRUCTURE, build container only (imports /root/reference).  The count is a property
of the workload (data + RNG seed; scipy's L-BFGS-B is deterministic), so bench.py --impl reference -- which can only
afford a bounded sample on the GPU box -- multiplies its own measured per-evaluation time by the count recorded here.

Usage:  python oracle/count_reference_fit_evals.py c2 c3 c4   ->  profiles/reference_fit_evals.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden as MG  # noqa: E402
from bench import WORKLOADS, synth_observations  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "reference_fit_evals.json")


def main():
    R, PS, B = MG.import_reference()
    from sklearn.gaussian_process import GaussianProcessRegressor as GPR
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in sys.argv[1:] or ["c2", "c3"]:
        wl = WORKLOADS[name]
        X, meas, _ = synth_observations(wl["n"], wl["W"], 0)
        sc = PS.PointSourceField(num_sources=3, workspace_size=(0, wl["W"], 0, wl["W"]), seed=0)
        count = {"n": 0}
        orig = GPR.log_marginal_likelihood

        def counted(self, theta=None, eval_gradient=False, clone_kernel=True):
            if theta is not None:
                count["n"] += 1
            return orig(self, theta, eval_gradient, clone_kernel)

        GPR.log_marginal_likelihood = counted
        try:
            np.random.seed(0)
            t = time.time()
            sc.update([X[i] for i in range(len(X))], np.log10(np.maximum(meas, 1e-6)))
            dt = time.time() - t
        finally:
            GPR.log_marginal_likelihood = orig
        # fit() itself evaluates the LML once more at the chosen theta (_gpr.py:340-343 with clone_kernel=False)
        res[name] = {"n": wl["n"], "seed": 0, "objective_evaluations": count["n"], "theta_c_l": np.exp(sc.gp.kernel_.theta).tolist(),
                     "lml": float(sc.gp.log_marginal_likelihood_value_), "fit_seconds_build_container": dt,
                     "cores_build_container": os.cpu_count(), "versions": [str(v) for v in MG.versions()]}
        print(name, res[name], flush=True)
        json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()

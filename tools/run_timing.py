"""Application-level timing: whole strategy.run() on the B200 path, eager (every refit of the reference evaluated)
and lazy (refits nobody reads deferred).  SURVEY.md section 6 has the reference's CPU time for the
same call (MR_IPP, budget 100, d = 2.5, 1 agent: 4.6 s on 8 cores, 85 % of it in gp.fit)."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt import rrt as R
warnings.simplefilter("ignore")
for agents, budget in [(1, 100), (3, 150)]:
    for lazy in (False, True):
        outs = []
        for rep in range(2):  # first repetition warms allocations / module load
            sc = PointSourceField(num_sources=3, workspace_size=(0, 40, 0, 40), seed=5)
            sc.gp.lazy = lazy
            np.random.seed(11)
            st = R.MR_IPP(sc, budget=budget, d_waypoint_distance=2.5, num_agents=agents, budget_iter=10, n_samples=200,
                          s_stages=10, max_sources=3)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            Z, std = st.run()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            outs.append((dt, np.asarray(st.obs_wp).copy(), sc.gp.n_lml_evals_))
        print(f"MR_IPP agents={agents} budget={budget} lazy={lazy}: run() {outs[1][0]:.3f} s, {len(outs[1][1])} observations, "
              f"{outs[1][2]} LML evaluations, grid {Z.shape}", flush=True)
    # the waypoint sequence does not depend on the mode

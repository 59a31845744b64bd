"""cProfile of one whole strategy run on the B200 path (lazy refits), to see which host functions are left once the
GP and the scoring are on the device.  Run under gpurun."""
import os, sys, cProfile, pstats, io, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt import rrt as R
warnings.simplefilter("ignore")
agents, budget, W = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3, 150, 40)
for rep in range(2):
    sc = PointSourceField(num_sources=3, workspace_size=(0, W, 0, W), seed=5)
    sc.gp.lazy = True
    np.random.seed(11)
    st = R.MR_IPP(sc, budget=budget, d_waypoint_distance=2.5, num_agents=agents, budget_iter=10, n_samples=200,
                  s_stages=10, max_sources=3)
    pr = cProfile.Profile()
    pr.enable(); Z, std = st.run(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(f"agents={agents} budget={budget} W={W} observations={len(st.obs_wp)}")
print(s.getvalue()[:9000])

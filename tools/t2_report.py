"""T2 (re-optimised fit) deviations of every whole-run golden on the device: max |log10 Z - ref| and |std - ref| with the
device's own theta*, next to theta* of both.  Feeds the per-case tolerances of tests/test_gpu_parity.py."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt import rrt as R
from mripp_b200.src.boustrophedon import boustrophedon as B
from tests.test_gpu_parity import RUN_COMMON
from tests.test_r2_cpu import MULTI, MULTI_COMMON
from oracle import gp_oracle as O
g = np.load(os.path.join(ROOT, "tests", "golden", "runs_golden.npz"))
g2 = np.load(os.path.join(ROOT, "tests", "golden", "runs_golden_r2.npz"))
warnings.simplefilter("ignore")
CASES = [("mripp_1a", "MR_IPP", dict(num_agents=1, stage_lambda=0.5)), ("mripp_2a", "MR_IPP", dict(num_agents=2, stage_lambda=0.5)),
         ("rig_dist", "RRTRIG_PointSourceInformative_Distance_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
         ("rig_nopen", "RRTRIG_PointSourceInformative_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
         ("rig_rot", "RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
         ("biasbeta", "RRT_BiasBetaInformative_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.5)),
         ("rrt_random", "RRT_Random_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.0)),
         ("rrtstar_random", "RRTStar_Random_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.0))]
def report(tag, gg, sc, st, Z, std, stride):
    Xo, ylog = np.asarray(st.obs_wp, float), np.log10(np.maximum(np.asarray(st.measurements, float), 1e-6))
    yn, _, _ = O.normalise_y(ylog)
    v_ref, v_dev = O.lml(gg[f"{tag}_theta"], Xo, yn, False), O.lml(sc.gp.kernel_.theta, Xo, yn, False)
    dz = np.abs(np.log10(Z[::stride, ::stride]) - np.log10(gg[f"{tag}_Z"])).max()
    ds = np.abs(std.reshape(Z.shape)[::stride, ::stride] - gg[f"{tag}_std"]).max()
    print(f"{tag:16s} n={len(Xo):4d} dlogZ {dz:.3e} dstd {ds:.3e} theta_dev {sc.gp.kernel_.theta} theta_ref {gg[f'{tag}_theta']} "
          f"lml_dev-lml_ref {v_dev - v_ref:+.3e}", flush=True)
for tag, cls, extra in CASES:
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21); sc.gp.lazy = True
    np.random.seed(1234)
    st = getattr(R, cls)(sc, **RUN_COMMON, **extra)
    Z, std = st.run()
    report(tag, g, sc, st, Z, std, 4)
for tag, cls, kw in [("boust", "Boustrophedon", dict(budget=60)), ("mr_boust", "MR_Boustrophedon", dict(budget=40, num_agents=2))]:
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21)
    np.random.seed(99)
    st = getattr(B, cls)(sc, d_waypoint_distance=2.5, **kw)
    Z, std = st.run()
    report(tag, g, sc, st, Z, std, 4)
for tag, obstacles in MULTI:
    seed, agents, budget = int(g2[f"{tag}_seed"]), int(g2[f"{tag}_num_agents"]), int(g2[f"{tag}_budget"])
    sc = PointSourceField(num_sources=2, workspace_size=(0, 40, 0, 40), obstacles=obstacles, seed=seed); sc.gp.lazy = True
    np.random.seed(2000 + seed)
    st = R.MR_IPP(sc, num_agents=agents, budget=budget, **MULTI_COMMON)
    Z, std = st.run()
    report(tag, g2, sc, st, Z, std, 8)

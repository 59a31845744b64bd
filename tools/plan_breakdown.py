"""Where a planning step spends its time (host wall clock, CUDA-synchronised).  Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_observations, synth_tree
from mripp_b200.scoring import GpuScorer, ScoreContext
from mripp_b200.src.rrt.rrt_utils import TreeArrays
n, W, agents, per = 5000, 100, 8, 1024
X, meas, src = synth_observations(n, W, 0)
obs = [list(X[a * (n // agents):(a + 1) * (n // agents)]) for a in range(agents)]
xy, par = synth_tree(per, W, 1)
tr = TreeArrays(xy[0], capacity=per + 1)
for k in range(1, per + 1):
    tr.add(xy[k], int(par[k]))
sc = GpuScorer()
ctx = ScoreContext(src, tr.point(0), obs, 0, 1.0, (0, W, 0, W), [])
leaves = tr.leaves_in_order(); par_l = tr.parent[leaves]
leaf_xy = np.ascontiguousarray(tr.points()[leaves]); parent_xy = np.ascontiguousarray(tr.points()[np.where(par_l >= 0, par_l, 0)])
aobs = np.asarray(obs[0], dtype=np.float64).reshape(-1, 2)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
print("leaves", len(leaves), "agent obs", len(aobs))
print("tree_information   %.3f ms" % t(lambda: sc.tree_information(3, tr, ctx)))
print("  ctx.all_obs()    %.3f ms" % t(lambda: ctx.all_obs()))
print("  _upload_tree     %.3f ms" % t(lambda: sc._upload_tree(3, tr, ctx)))
print("leaf_scores        %.3f ms" % t(lambda: sc.leaf_scores(leaf_xy, parent_xy, par_l >= 0, ctx, 5.0, 1.0, 1.0)))
print("  mutual_info      %.3f ms" % t(lambda: sc.mutual_information(leaf_xy, aobs, 1.0, 1.0, 5.0)))
print("  close_counts     %.3f ms" % t(lambda: sc.close_counts(leaf_xy, ctx.all_obs(), 1.0)))

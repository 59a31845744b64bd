"""Real-NCCL check of the multi-GPU plumbing (run under torchrun on >= 2 GPUs):
row-slab sharded lattice posterior == single-GPU posterior bit for bit; sharded branch-score argmax ==
np.argmax; chosen-path broadcast.  Prints one line per rank-0 check and exits non-zero on mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mripp_b200 import dist as D
from mripp_b200 import gp as G
from mripp_b200.src.point_source.point_source import PointSourceField
from bench import synth_observations

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, n = 30, 700
X, meas, _ = synth_observations(n, W, 0)
field = PointSourceField(num_sources=3, workspace_size=(0, W, 0, W), seed=1)
field.gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True)
Z1, s1 = field.predict_spatial_field(list(X), meas)            # every rank: whole lattice on its own GPU
field.grid_group = dist.group.WORLD
Z2, s2 = field.predict_spatial_field(list(X), meas)            # row slabs + all_gather
ok = np.array_equal(Z1, Z2) and np.array_equal(s1, s2)
scores = np.random.RandomState(3).randn(8192); scores[[100, 5000]] = scores.max() + 1.0   # a tie across shards
lo, hi = D.slab_bounds(len(scores), world, rank)
k = D.sharded_argmax(torch.from_numpy(scores[lo:hi]).cuda(), lo)
ok = ok and k == int(np.argmax(scores))
scores[7000] = np.nan
k2 = D.sharded_argmax(torch.from_numpy(scores[lo:hi]).cuda(), lo)
ok = ok and k2 == int(np.argmax(scores))
path = D.broadcast_path(np.array([[1.0, 2.0], [3.5, 4.5], [6.0, 7.0]]) if rank == world - 1 else None, src=world - 1, device="cuda")
ok = ok and np.array_equal(path, np.array([[1.0, 2.0], [3.5, 4.5], [6.0, 7.0]]))
# ---- strong scaling of ONE C4 lattice posterior by row slab (device time, max over ranks) ----------------------
import time
W4, n4 = 100, 5000
X4, meas4, _ = synth_observations(n4, W4, 0)
gp4 = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 2.0), optimizer=None, normalize_y=True).fit(X4, np.log10(meas4))
def slab_step(balanced=True):
    bounds = (D.lattice_slabs(gp4, (0, W4, 0, W4), 1000, 1000, world) if balanced
              else [D.row_slab(1000, 1000, world, r) for r in range(world)])   # the split is inside the timed step
    m_, s_, z_ = gp4.predict_grid_device(0, W4, 1000, 0, W4, 1000, g_begin=bounds[rank][0], g_end=bounds[rank][1])
    sizes = [b - a for a, b in bounds]
    return D.gather_slabs(z_, sizes), D.gather_slabs(s_, sizes)
zb, sb = slab_step(True); ze, se = slab_step(False)
mfull, sfull, zfull = gp4.predict_grid_device(0, W4, 1000, 0, W4, 1000)
ok = ok and bool(torch.allclose(sb, sfull, rtol=1e-9, atol=1e-12)) and bool(torch.allclose(se, sfull, rtol=1e-9, atol=1e-12))
if rank == 0:
    print("dist_check balanced slabs:", D.lattice_slabs(gp4, (0, W4, 0, W4), 1000, 1000, world),
          "max |std - single GPU|", float((sb - sfull).abs().max()), flush=True)
for _ in range(3):
    slab_step()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    zf, sf = slab_step()
e1.record(); dist.barrier(); torch.cuda.synchronize()
tms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda"); dist.all_reduce(tms, op=dist.ReduceOp.MAX)
sc8 = torch.randn(8192, dtype=torch.float64, device="cuda")
lo8, hi8 = D.slab_bounds(8192, world, rank)
for _ in range(3):
    D.sharded_argmax(sc8[lo8:hi8], lo8)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    D.sharded_argmax(sc8[lo8:hi8], lo8)
torch.cuda.synchronize(); t_arg = (time.perf_counter() - t0) / 20
if rank == 0:
    print(f"dist_check strong scaling: one C4 lattice (10^6 points, n=5000) by row slab over {world} GPUs incl. all_gather of "
          f"Z and std: {float(tms):.1f} ms per posterior ({1e6 / float(tms) * 1e3:.3e} points/s); sharded argmax of 8192 "
          f"scores (all_gather of one pair per rank): {t_arg * 1e6:.0f} us", flush=True)
# ---- 8192 candidate branches of one tree sharded over the ranks (BASELINE configs[3]) ------------------
from bench import synth_tree
from mripp_b200.scoring import GpuScorer, ScoreContext
from mripp_b200.src.rrt.rrt_utils import TreeArrays
txy, tpar = synth_tree(8192, 100, 1)
tr = TreeArrays(txy[0], capacity=8193)
for k_ in range(1, 8193):
    tr.add(txy[k_], int(tpar[k_]))
Xb, measb, srcb = synth_observations(5000, 100, 0)
obs_b = [list(Xb[a * 625:(a + 1) * 625]) for a in range(8)]
ctx_b = ScoreContext(srcb, tr.point(0), obs_b, 0, 1.0, (0, 100, 0, 100), [])
sc1, scP = GpuScorer(), GpuScorer()
scP.branch_group = True
info1 = sc1.tree_information(3, tr, ctx_b)
infoP = scP.tree_information(3, tr, ctx_b)
same_b = np.array_equal(info1, infoP, equal_nan=True)
ok = ok and same_b
def t_info(sc_):
    for _ in range(3): sc_.tree_information(3, tr, ctx_b)
    dist.barrier(); torch.cuda.synchronize(); t0_ = time.perf_counter()
    for _ in range(10): sc_.tree_information(3, tr, ctx_b)
    torch.cuda.synchronize(); return (time.perf_counter() - t0_) / 10 * 1e3
t1_, tP_ = t_info(sc1), t_info(scP)
if rank == 0:
    print(f"dist_check sharded branches: information of 8192 branches (point_source_gain_all, 5000 observations) over {world} GPUs "
          f"== one GPU: {same_b}; host call one GPU {t1_:.2f} ms, sharded {tP_:.2f} ms (the walks are 0.04 ms of it: sharding a "
          f"batch this small only adds the all_gather)", flush=True)
# ---- one 11-restart fit with its restarts spread over the ranks ---------------------------------------
for n_fit, W_fit in ((2000, 50), (5000, 100)):
    Xf, measf, _ = synth_observations(n_fit, W_fit, 0)
    res = {}
    for mode in ("one GPU", "sharded"):
        gpf = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), n_restarts_optimizer=10, normalize_y=True)
        gpf.restart_group = True if mode == "sharded" else None
        np.random.seed(5)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        gpf.fit(Xf, np.log10(measf))
        torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[mode] = (float(dt), gpf.kernel_.theta.copy(), gpf.log_marginal_likelihood_value_, gpf.n_lml_evals_, np.random.uniform(size=2))
    same = (np.array_equal(res["one GPU"][1], res["sharded"][1]) and res["one GPU"][2] == res["sharded"][2]
            and np.array_equal(res["one GPU"][4], res["sharded"][4]))
    ok = ok and same
    ev = torch.tensor([float(res["sharded"][3])], device="cuda"); dist.all_reduce(ev, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dist_check sharded restarts n={n_fit}: 11-restart fit on one GPU {res['one GPU'][0]:.3f} s ({res['one GPU'][3]} evaluations), "
              f"restarts over {world} GPUs {res['sharded'][0]:.3f} s (at most {int(ev.item())} evaluations on a rank); "
              f"theta*, LML and RNG position identical: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: lattice {Z1.shape} slab-sharded == single GPU: {np.array_equal(Z1, Z2) and np.array_equal(s1, s2)}; "
          f"argmax {k} (tie) / {k2} (nan); path broadcast ok; ALL_RANKS_OK={bool(flag.item())}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)

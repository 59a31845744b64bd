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
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: lattice {Z1.shape} slab-sharded == single GPU: {np.array_equal(Z1, Z2) and np.array_equal(s1, s2)}; "
          f"argmax {k} (tie) / {k2} (nan); path broadcast ok; ALL_RANKS_OK={bool(flag.item())}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)

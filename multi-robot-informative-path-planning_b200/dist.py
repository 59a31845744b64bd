"""Multi-GPU sharding of the hot path: one process per GPU over torch.distributed (NCCL on the
box, gloo in the CPU tests).  The path is embarrassingly parallel:

  * grid posterior   -> contiguous row slabs of the lattice; every rank holds the (replicated)
                        fitted state, computes its slab, and one all_gather assembles the field;
  * branch / leaf scoring -> contiguous candidate batches; the only exchange is an all_gather of one
                        (best value, global index) pair per rank, reduced with np.argmax semantics
                        (first maximum, first NaN wins) so the choice is rank-count independent;
  * independent rounds -> round r runs on rank r mod P, no data-path collective.

No collective moves bulk data on the critical path: the gathers are G*16/P bytes and 16 bytes.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def slab_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of range(total) into `world` near-equal slabs."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_slab(nx: int, ny: int, world: int, rank: int) -> Tuple[int, int]:
    """Flat lattice index range [g0, g1) of this rank's slab of whole rows (index = iy * nx + ix)."""
    r0, r1 = slab_bounds(ny, world, rank)
    return r0 * nx, r1 * nx


def balanced_row_slabs(row_work: Sequence[float], nx: int, world: int) -> List[Tuple[int, int]]:
    """Flat index ranges of `world` contiguous slabs of whole lattice rows with near-equal WORK: with the
    block-sparse posterior the rows at the workspace border have fewer observations in reach than the interior, so
    equal row counts are unequal work.  row_work[iy] >= 0 is any per-row cost estimate (gp.lattice_work)."""
    w = np.maximum(np.asarray(row_work, dtype=np.float64), 0.0) + 1e-12
    ny = len(w)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target))
        k = min(max(k, cuts[-1] + (1 if ny - cuts[-1] > world - r else 0)), ny - (world - r))
        cuts.append(max(k, cuts[-1]))
    cuts.append(ny)
    return [(cuts[r] * nx, cuts[r + 1] * nx) for r in range(world)]


def lattice_slabs(gp, ws, nx: int, ny: int, world: int) -> List[Tuple[int, int]]:
    """The row-slab split every rank uses for one lattice posterior: equal rows for an estimator without a work model,
    else balanced on the per-row work estimate of the block-sparse kernel (gp.lattice_row_work).  Every rank holds the same fitted state,
    so every rank derives the same split without communication."""
    if world <= 1 or not hasattr(gp, "lattice_row_work"):
        return [row_slab(nx, ny, world, r) for r in range(world)]
    return balanced_row_slabs(gp.lattice_row_work(ws[0], ws[1], nx, ws[2], ws[3], ny), nx, world)


def rounds_for_rank(n_rounds: int, world: int, rank: int) -> List[int]:
    return [r for r in range(n_rounds) if r % world == rank]


def _dist():
    import torch.distributed as dist

    return dist


def gather_slabs(local: "torch.Tensor", sizes: Sequence[int], group=None) -> "torch.Tensor":
    """all_gather of unequal 1-D slabs (padded to the largest) -> concatenated full vector."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    m = max(sizes)
    buf = torch.zeros(m, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def gather_field(send: "torch.Tensor", sizes: Sequence[int], group=None, pinned=None):
    """One collective for both output fields of a slab-sharded lattice posterior.  `send` is this rank's (2, m) buffer
    -- row 0 zpred, row 1 std, m = the largest slab, the tail beyond this rank's slab unused -- gathered into a
    pre-sized (world, 2, m) device buffer with all_gather_into_tensor (no list of padded tensors, no torch.cat), read
    back through ONE pinned staging buffer, and cut to the slab sizes on the host.  Returns (zpred, std) as numpy."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    m = send.shape[1]
    recv = torch.empty((world, 2, m), dtype=send.dtype, device=send.device)
    try:
        dist.all_gather_into_tensor(recv.view(-1), send.reshape(-1), group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the flat form (older gloo)
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send.contiguous(), group=group)
        recv = torch.stack(parts)
    if recv.is_cuda:
        host = pinned(recv.numel()) if pinned is not None else torch.empty(recv.numel(), dtype=recv.dtype).pin_memory()
        host.copy_(recv.view(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = host.numpy().reshape(world, 2, m)
    else:
        h = recv.numpy()
    z = np.concatenate([h[r, 0, :sz] for r, sz in enumerate(sizes)])
    s = np.concatenate([h[r, 1, :sz] for r, sz in enumerate(sizes)])
    return z, s


def sharded_grid_posterior(gp, ws, nx: int, ny: int, group=None, compute_slab: Optional[Callable] = None):
    """(zpred, std) over the whole lattice, each rank computing one row slab.

    `compute_slab(g0, g1) -> (zpred, std)` torch tensors; defaults to the fused device posterior of
    the fitted `gp` (mripp_b200.gp.GaussianProcessRegressor.predict_grid_device), which writes straight into the
    send buffer of the gather."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = lattice_slabs(gp, ws, nx, ny, world) if compute_slab is None else [row_slab(nx, ny, world, r) for r in range(world)]
    g0, g1 = bounds[rank]
    sizes = [b - a for a, b in bounds]
    m, n_loc = max(sizes), g1 - g0
    if compute_slab is None and hasattr(gp, "_ws"):  # the device estimator: its kernel writes the send buffer itself
        send = torch.empty((2, m), dtype=torch.float64, device="cuda")
        mean = torch.empty(max(n_loc, 1), dtype=torch.float64, device="cuda")
        gp.predict_grid_device(ws[0], ws[1], nx, ws[2], ws[3], ny, g_begin=g0, g_end=g1, want_zpred=True,
                               out=(mean[:n_loc], send[1, :n_loc], send[0, :n_loc]))
        pinned = lambda numel: gp._ws.pinned("gather", numel)
    else:
        if compute_slab is None:  # an estimator with the slab interface only (the CPU test double)
            def compute_slab(a, b):
                _, std_d, z_d = gp.predict_grid_device(ws[0], ws[1], nx, ws[2], ws[3], ny, g_begin=a, g_end=b, want_zpred=True)
                return z_d, std_d
        z_loc, s_loc = compute_slab(g0, g1)
        send = torch.zeros((2, m), dtype=z_loc.dtype, device=z_loc.device)
        send[0, :n_loc] = z_loc
        send[1, :n_loc] = s_loc
        pinned = None
    return gather_field(send, sizes, group, pinned)


def sharded_restarts(run_one: Callable[[List[int]], List[Tuple[np.ndarray, float]]], n_starts: int, dim: int, group=None,
                     device="cpu") -> List[Tuple[np.ndarray, float]]:
    """The optimiser restarts of one fit (sklearn _gpr.py:312-337) spread over the ranks: start i runs on rank
    i mod P, then one all_gather of (theta*, f*) -- (dim + 1) doubles per start -- gives every rank the full list in
    start order, so that np.argmin (first minimum wins) picks the same optimum everywhere.  Every rank must hold the
    same data and the same start points (same RNG state at fit time).  `run_one(indices)` returns the optima of this
    rank's starts, in the order given."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = rounds_for_rank(n_starts, world, rank)
    local = run_one(mine) if mine else []
    per = (n_starts + world - 1) // world
    buf = torch.full((per, dim + 1), float("nan"), dtype=torch.float64)
    for slot, (theta, fval) in enumerate(local):
        buf[slot, :dim] = torch.as_tensor(np.asarray(theta, dtype=np.float64))
        buf[slot, dim] = float(fval)
    buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    parts = [p.cpu().numpy() for p in parts]
    return [(parts[i % world][i // world, :dim].copy(), float(parts[i % world][i // world, dim])) for i in range(n_starts)]


def combine_argmax(values: np.ndarray, indices: np.ndarray) -> int:
    """np.argmax over the concatenation, given each shard's (local best value, GLOBAL index):
    a NaN wins at the smallest index, otherwise the largest value at the smallest index.
    indices < 0 mark empty shards."""
    best = -1
    for v, i in zip(values, indices):
        if i < 0:
            continue
        if best < 0:
            best, bv = int(i), v
            continue
        if np.isnan(bv) or np.isnan(v):
            take = np.isnan(v) and (not np.isnan(bv) or i < best)
        else:
            take = v > bv or (v == bv and i < best)
        if take:
            best, bv = int(i), v
    return best


def sharded_argmax(local_scores: "torch.Tensor", offset: int, group=None) -> int:
    """Global np.argmax of a score vector sharded in contiguous batches (this rank holds
    [offset, offset + len))."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    if local_scores.numel():
        v = local_scores.detach().double().cpu().numpy()
        k = int(np.argmax(v))  # numpy semantics on the shard
        pair = torch.tensor([float(v[k]), float(offset + k)], dtype=torch.float64, device=local_scores.device)
    else:
        pair = torch.tensor([0.0, -1.0], dtype=torch.float64, device=local_scores.device)
    parts = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(parts, pair, group=group)
    vals = np.array([float(p[0]) for p in parts])
    idx = np.array([int(p[1]) for p in parts])
    return combine_argmax(vals, idx)


def broadcast_path(path: Optional[np.ndarray], src: int, group=None, device="cpu") -> np.ndarray:
    """Broadcast the chosen waypoint list (k x 2 doubles) from rank `src`."""
    import torch

    dist = _dist()
    n = torch.tensor([0 if path is None else len(path)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    buf = torch.zeros((int(n.item()), 2), dtype=torch.float64, device=device)
    if dist.get_rank(group) == src and path is not None:
        buf.copy_(torch.as_tensor(np.asarray(path, dtype=np.float64).reshape(-1, 2)))
    dist.broadcast(buf, src=src, group=group)
    return buf.cpu().numpy()

"""GaussianProcessRegressor-compatible estimator whose arithmetic runs on a B200.

Mirrors the slice of scikit-learn's API that the reference touches through ``PointSourceField.gp``
(src/point_source/point_source.py:71,337,349,351; src/rrt/rrt.py:447,452; src/rrt/rrt_thread.py:375):

    gp = GaussianProcessRegressor(kernel=C(1.0, (1e-3, 50)) * RBF(1.0, (1e-2, 50)),
                                  n_restarts_optimizer=10, normalize_y=True)
    gp.fit(X, y); gp.predict(X, return_std=True); gp.kernel(X[, Y]); gp.kernel_; gp.alpha_; gp.L_
    gp.log_marginal_likelihood(theta, eval_gradient=True); gp.log_marginal_likelihood_value_

Host code keeps everything that is sequential or consumes the RNG (y normalisation bookkeeping, the
``n_restarts_optimizer`` uniform draws from the global ``np.random`` stream, scipy's L-BFGS-B
driver); every dense operation (Gram, Cholesky, triangular inverse, LML + gradient, posterior) is
a hand-written sm_100a kernel behind the C ABI of include/ipp_b200.h.  torch only owns the device
buffers and the stream.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import warnings

import numpy as np
from scipy.optimize import minimize

from . import _lib

__all__ = ["ConstantRBFKernel", "ConstantKernel", "RBF", "Matern", "GaussianProcessRegressor"]


def _chol_flags():
    """Measurement switches of the dataflow Cholesky: IPP_CHOL_HINTS=0 (no chain hand-over), IPP_CHOL_STATS=1 (per-CTA
    task / cycle counters left in the scheduling scratch), IPP_CHOL_WS=0 (the executor without warp specialisation),
    IPP_CHOL_RED=0 (tile updates written back by load / subtract / store instead of reductions at the L2)."""
    st = os.environ.get("IPP_CHOL_STATS", "0")
    return ((1 if os.environ.get("IPP_CHOL_HINTS", "1") == "0" else 0) | (2 if st == "1" else 0) | (4 if st == "2" else 0)
            | (8 if os.environ.get("IPP_CHOL_EAGER_SCAN", "0") == "1" else 0)
            | (16 if os.environ.get("IPP_CHOL_WS", "1") == "0" else 0)
            | (32 if os.environ.get("IPP_CHOL_RED", "1") == "0" else 0))


def chol_trace(ws, R):
    """Per-task events of the last dataflow factorisation on `ws` (IPP_CHOL_STATS=2): rows (cta, queue, problem, task
    index, start_ns, end_ns), times relative to the first event."""
    plan = ws.plan.cpu().numpy()
    base = R * int(plan[4]) + 8 * R + 8 + 16 * 512
    ev = ws.sched[base:base + 4 * 448 * 160].cpu().numpy().reshape(160, 448, 4)
    rows = []
    for cta in range(160):
        for e in ev[cta]:
            if e[3] == 0 and e[2] == 0:
                break
            rows.append((cta, int(e[0]) & 0xff, int(e[0]) >> 8, int(e[1]), int(e[2]), int(e[3])))
    a = np.array(rows, dtype=np.int64).reshape(-1, 6)
    if len(a):
        t0 = a[:, 4].min()
        a[:, 4:] -= t0
    return a


def chol_stats(ws, R):
    """Aggregate of the per-CTA counters of the last dataflow factorisation run on workspace `ws` (IPP_CHOL_STATS=1)."""
    plan = ws.plan.cpu().numpy()
    base = R * int(plan[4]) + 8 * R + 8
    st = ws.sched[base:base + 16 * 512].cpu().numpy().reshape(512, 16)
    st = st[st[:, 7] > 0]
    cyc = st[:, 3:8].astype(np.float64) * 64.0
    return {"ctas": int(len(st)), "tasks_F_S_U": st[:, 0:3].sum(0).tolist(),
            "busy_frac_F_S_U": (cyc[:, 0:3].sum(0) / cyc[:, 4].sum()).tolist(), "acquire_frac": float(cyc[:, 3].sum() / cyc[:, 4].sum()),
            "mean_task_us_F_S_U_at_1965MHz": (cyc[:, 0:3].sum(0) / np.maximum(st[:, 0:3].sum(0), 1) / 1965.0).tolist(),
            "kernel_us": float(cyc[:, 4].max() / 1965.0), "polls_per_task": float(st[:, 8].sum() / max(st[:, 0:3].sum(), 1))}


def _check_watchdog(info):
    if info == -7777.0:
        raise _lib.IppError("dataflow Cholesky: a CTA found no runnable task for ~2 s (scheduling fault); "
                            "set IPP_CHOL=coop to use the cooperative kernel")


# ------------------------------------------------------------------------------------------
# kernel objects (host-side description of c * RBF(l); the arithmetic is ipp_gram_rbf)
# ------------------------------------------------------------------------------------------
class _Factor:
    def __init__(self, value, bounds):
        self.value = float(value)
        self.bounds = bounds

    def __mul__(self, other):
        if isinstance(self, ConstantKernel) and isinstance(other, RBF):
            return ConstantRBFKernel(self.value, other.value, self.bounds, other.bounds, nu=getattr(other, "nu", np.inf))
        if isinstance(self, RBF) and isinstance(other, ConstantKernel):
            return ConstantRBFKernel(other.value, self.value, other.bounds, self.bounds, nu=getattr(self, "nu", np.inf))
        raise TypeError("only ConstantKernel * RBF / Matern is on the accelerated path")


class ConstantKernel(_Factor):
    """sklearn.gaussian_process.kernels.ConstantKernel look-alike (factor of the product only)."""

    def __init__(self, constant_value=1.0, constant_value_bounds=(1e-5, 1e5)):
        super().__init__(constant_value, constant_value_bounds)


class RBF(_Factor):
    """sklearn.gaussian_process.kernels.RBF look-alike (isotropic, factor of the product only)."""

    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        super().__init__(length_scale, length_scale_bounds)


class Matern(RBF):
    """sklearn.gaussian_process.kernels.Matern look-alike for nu in {1.5, 2.5, inf} (factor of the product only).
    The kernel-matrix builder (`kernel(X[, Y])`, ipp_gram_kernel) covers it; the fit and the posterior are RBF-only,
    like the reference (point_source.py:86 builds C * RBF and nothing else)."""

    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5), nu=1.5):
        super().__init__(length_scale, length_scale_bounds)
        if nu not in (1.5, 2.5, np.inf):
            raise ValueError("Matern is implemented for nu = 1.5, 2.5 and inf (= RBF)")
        self.nu = nu


class ConstantRBFKernel:
    """``ConstantKernel(c, c_bounds) * RBF(l, l_bounds)`` -- the only kernel the reference builds
    (point_source.py:82-87).  theta = log([c, l]) (sklearn kernels.py:291-312)."""

    requires_vector_input = True
    n_dims = 2

    def __init__(self, constant_value=1.0, length_scale=1.0, constant_value_bounds=(1e-3, 50.0),
                 length_scale_bounds=(1e-2, 50.0), nu=np.inf):
        self.constant_value = float(constant_value)
        self.length_scale = float(length_scale)
        self.constant_value_bounds = constant_value_bounds
        self.length_scale_bounds = length_scale_bounds
        self.nu = nu  # inf: RBF; 1.5 / 2.5: Matern (kernel-matrix builder only)

    # -- sklearn Kernel protocol -----------------------------------------------------------
    @property
    def theta(self):
        return np.log(np.array([self.constant_value, self.length_scale], dtype=float))

    @theta.setter
    def theta(self, theta):
        theta = np.asarray(theta, dtype=float)
        self.constant_value = float(np.exp(theta[0]))
        self.length_scale = float(np.exp(theta[1]))

    @property
    def bounds(self):
        return np.log(np.array([self.constant_value_bounds, self.length_scale_bounds], dtype=float))

    def clone_with_theta(self, theta):
        k = ConstantRBFKernel(self.constant_value, self.length_scale, self.constant_value_bounds,
                              self.length_scale_bounds, nu=self.nu)
        k.theta = theta
        return k

    def clone(self):
        return ConstantRBFKernel(self.constant_value, self.length_scale, self.constant_value_bounds,
                                 self.length_scale_bounds, nu=self.nu)

    def diag(self, X):
        return np.full(len(np.asarray(X)), self.constant_value, dtype=float)

    def is_stationary(self):
        return True

    def __call__(self, X, Y=None, eval_gradient=False):
        """Dense Gram matrix on the device, returned as a host ndarray (rrt.py:447,452 consume it
        with numpy).  Use :func:`gram_device` to keep it resident."""
        if eval_gradient:
            raise NotImplementedError("the fused LML kernel never materialises dK; use "
                                      "GaussianProcessRegressor.log_marginal_likelihood")
        return gram_device(X, Y, self.constant_value, self.length_scale, kind={np.inf: 0, 1.5: 1, 2.5: 2}[self.nu]).cpu().numpy()

    def __repr__(self):
        # same text sklearn prints for C(c) * RBF(l)
        return "{0:.3g}**2 * RBF(length_scale={1:.3g})".format(np.sqrt(self.constant_value), self.length_scale)


def _as_points(X):
    A = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if A.ndim == 1:
        A = A.reshape(-1, 2)
    if A.ndim != 2 or A.shape[1] != 2:
        raise ValueError(f"expected an (n, 2) array of waypoints, got shape {A.shape}")
    return A


def morton_order(X):
    """Stable argsort of the points along a Z-order curve over their own bounding box (16 bits per axis): the
    order in which the final factorisation sees the observations, so that 16 consecutive ones form a compact
    cluster and whole 16-observation slabs of k* vanish for far-away lattice tiles."""
    X = np.asarray(X, dtype=np.float64)
    lo, hi = X.min(axis=0), X.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((X - lo) / span * 65536.0).astype(np.uint64), np.uint64(65535))

    def spread(v):
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v

    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)), kind="stable")


def clustered_order(X, slab=16, cluster_order="maximin"):
    """Order of the observations for the final factorisation when the lattice kernel is to skip slabs: the points are
    cut into spatial clusters of `slab` consecutive points along the Z-order curve (so that a slab of k* is one
    compact cluster), and the CLUSTERS are then taken coarse to fine -- farthest-point (maximin) order of their
    centroids -- instead of along the curve.  Eliminating in curve order makes every leading block a dense,
    badly conditioned neighbourhood, and the explicit inverse W = L^-1 (built from products of such leading
    blocks) loses 3-4 digits of the variance next to the data when there are more than ~2 observations per
    l x l cell; coarse-to-fine order keeps the leading blocks as well conditioned as a random order does."""
    m = morton_order(X)
    n = len(m)
    blocks = [m[i:i + slab] for i in range(0, n - n % slab, slab)]
    tail = m[n - n % slab:]  # the short cluster stays last, so that every slab of 16 rows is exactly one cluster
    if len(blocks) <= 2 or cluster_order == "curve":
        return m
    cent = np.array([X[b].mean(axis=0) for b in blocks])
    if cluster_order == "shuffled":
        order = np.random.RandomState(len(blocks)).permutation(len(blocks))
    else:
        order = np.empty(len(blocks), dtype=np.int64)
        order[0] = 0
        dmin = np.hypot(*(cent - cent[0]).T)
        dmin[0] = -1.0
        for k in range(1, len(blocks)):
            j = int(np.argmax(dmin))
            order[k] = j
            dmin = np.minimum(dmin, np.hypot(*(cent - cent[j]).T))
            dmin[j] = -1.0
    return np.concatenate([blocks[i] for i in order] + [tail])


def slab_bounding_boxes(Xs, NP, slab=16):
    """{xmin, xmax, ymin, ymax} of each `slab` consecutive rows of Xs, padded to NP rows; slabs that hold only
    padding get an empty box infinitely far away."""
    n = Xs.shape[0]
    ns = NP // slab
    out = np.empty((ns, 4), dtype=np.float64)
    out[:, 0::2] = 1e300
    out[:, 1::2] = 1e300
    full = n // slab
    if full:
        blk = Xs[: full * slab].reshape(full, slab, 2)
        out[:full, 0], out[:full, 1] = blk[:, :, 0].min(1), blk[:, :, 0].max(1)
        out[:full, 2], out[:full, 3] = blk[:, :, 1].min(1), blk[:, :, 1].max(1)
    if n > full * slab:
        rest = Xs[full * slab:]
        out[full] = (rest[:, 0].min(), rest[:, 0].max(), rest[:, 1].min(), rest[:, 1].max())
    return out


def as_lattice(Xq):
    """(x0, x1, nx, y0, y1, ny) when the query points are exactly the raveled 'xy' meshgrid of two np.linspace
    axes in row-major order -- what the reference builds for every prediction (point_source.py:65-67, 350) --
    else None.  Lets `predict(r)` of the UNMODIFIED reference reach the fused lattice kernel."""
    G = Xq.shape[0]
    if G < 256:
        return None
    y = Xq[:, 1]
    nx = int(np.argmax(y != y[0])) if y[-1] != y[0] else G
    if nx < 128 or G % nx:
        return None
    ny = G // nx
    xs, ys = Xq[:nx, 0], y[::nx]
    if ny < 2 or not (np.array_equal(xs, np.linspace(xs[0], xs[-1], nx)) and np.array_equal(ys, np.linspace(ys[0], ys[-1], ny))):
        return None
    grid = Xq.reshape(ny, nx, 2)
    if not (np.array_equal(grid[:, :, 0], np.broadcast_to(xs, (ny, nx))) and
            np.array_equal(grid[:, :, 1], np.broadcast_to(ys[:, None], (ny, nx)))):
        return None
    return float(xs[0]), float(xs[-1]), nx, float(ys[0]), float(ys[-1]), ny


def gram_device(X, Y, c, l, jitter=0.0, kind=0):
    """K(X, Y) as a device-resident torch tensor (n x m, fp64).  kind 0: RBF, 1 / 2: Matern nu = 1.5 / 2.5."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    Xd = torch.from_numpy(_as_points(X)).cuda()
    n = Xd.shape[0]
    if Y is None:
        Yd, m = None, n
    else:
        Yd = torch.from_numpy(_as_points(Y)).cuda()
        m = Yd.shape[0]
    out = torch.empty((n, m), dtype=torch.float64, device="cuda")
    if n and m:
        _lib.check(lib.ipp_gram_kernel(int(kind), _lib.ptr(Xd), n, _lib.ptr(Yd), m, float(c), float(l), float(jitter),
                                       _lib.ptr(out), m, _lib.stream_ptr()), "ipp_gram_kernel")
    return out


# ------------------------------------------------------------------------------------------
# device workspace of one fitted GP
# ------------------------------------------------------------------------------------------
def chol_mode():
    """'queue' (default): the dataflow Cholesky driven by a task plan (csrc/ipp_cholq.cuh); 'coop': the cooperative
    kernel with grid barriers (one problem per launch; kept for A/B measurements).  IPP_CHOL overrides."""
    return os.environ.get("IPP_CHOL", "queue")


_PLANS = {}
_PLAN_LOCK = threading.Lock()
NEAR_SLACK = 0


def chol_plan(n):
    """Device copy of the task plan for n observations (depends on NP and the SM count only; cached)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.cuda.current_device()
    NP = (int(n) + 63) // 64 * 64
    near = int(os.environ.get("IPP_CHOL_NEAR", str(NEAR_SLACK)))  # tiles within near + 1 super-panels of the front: chain priority
    with _PLAN_LOCK:
        plan = _PLANS.get((dev, NP, near))
        if plan is None:
            need = C.c_longlong(0)
            _lib.check(lib.ipp_host_chol_plan(int(n), near, None, 0, C.byref(need)), "ipp_host_chol_plan")
            host = torch.empty(need.value, dtype=torch.int32).pin_memory()
            _lib.check(lib.ipp_host_chol_plan(int(n), near, C.cast(host.data_ptr(), C.POINTER(C.c_int)), need.value,
                                              C.byref(need)), "ipp_host_chol_plan")
            plan = host.cuda()
            _PLANS[(dev, NP, near)] = plan
    return plan


class _Workspace:
    """Caller-owned device buffers of the C ABI (ipp_gp_layout), reused while NP does not change.  `slots` > 1:
    one buffer set per optimiser restart, slot-strided (ipp_gp_lml_grad_batch)."""

    def __init__(self):
        self.np_ = -1
        self.slots = 0
        self.K = self.W = self.T = self.v = self.sched = None

    def ensure(self, n, slots=1):
        torch = _lib.require_cuda()
        lib = _lib.load()
        lay = (C.c_longlong * 10)()
        _lib.check(lib.ipp_gp_layout(int(n), lay), "ipp_gp_layout")
        self.layout = [int(v) for v in lay]
        NP = self.layout[0]
        need = C.c_longlong(0)
        _lib.check(lib.ipp_chol_sched_ints(int(n), max(int(slots), self.slots, 1), C.byref(need)), "ipp_chol_sched_ints")
        fits = (self.slots >= slots and self.K is not None and self.K.numel() >= self.slots * self.layout[3]
                and self.v.numel() >= self.slots * self.layout[6] and self.sched.numel() >= need.value)
        if not fits:
            dev = torch.device("cuda")
            S = max(int(slots), self.slots)
            # sized for the next multiple of 512 observations: a run whose observation count grows step by step (the
            # per-step refits of rrt.py:339,360,749) re-allocates every 512 observations, not every 64
            cap = (C.c_longlong * 10)()
            _lib.check(lib.ipp_gp_layout((int(n) + 511) // 512 * 512, cap), "ipp_gp_layout")
            self.K = self.W = self.T = self.v = None  # release before allocating the new set
            self.K = torch.empty(S * int(cap[3]), dtype=torch.float64, device=dev)
            self.W = torch.empty(S * int(cap[4]), dtype=torch.float64, device=dev)
            self.T = torch.empty(S * int(cap[5]), dtype=torch.float64, device=dev)
            self.v = torch.zeros(S * int(cap[6]), dtype=torch.float64, device=dev)
            self.hyp = torch.zeros(4 * S, dtype=torch.float64, device=dev)
            self.hyp_host = torch.zeros(4 * S, dtype=torch.float64).pin_memory()
            self.res_host = torch.empty(8 * S, dtype=torch.float64).pin_memory()
            self.res_dev = torch.empty(8 * S, dtype=torch.float64, device=dev)
            self.slot_ids = torch.zeros(S, dtype=torch.int32, device=dev)
            self.slot_ids_host = torch.zeros(S, dtype=torch.int32).pin_memory()
            _lib.check(lib.ipp_chol_sched_ints((int(n) + 511) // 512 * 512, S, C.byref(need)), "ipp_chol_sched_ints")
            self.sched = torch.zeros(need.value, dtype=torch.int32, device=dev)
            self.slots = S
        self.np_ = NP
        self.plan = chol_plan(n) if chol_mode() == "queue" else None
        return self

    @property
    def result(self):
        off = self.layout[8]
        return self.v[off:off + 8]

    def pinned(self, name, numel):
        """Reusable page-locked staging buffer (fp64) of at least `numel` elements."""
        torch = _lib.require_cuda()
        buf = getattr(self, "_pin_" + name, None)
        if buf is None or buf.numel() < numel:
            buf = torch.empty(max(int(numel), 1), dtype=torch.float64).pin_memory()
            setattr(self, "_pin_" + name, buf)
        return buf[:numel]


def point_table_rows(m, NP, free_bytes, have_numel=0, sms=148):
    """Queries per chunk of the tabled explicit-point posterior (a multiple of 128): everything at once when the table
    fits a quarter of the free device memory (at most 4 GiB) or the scratch already held, else whole waves of 128-query
    tiles (one tile per SM), else whatever fits -- never less than one tile.  The scratch of `rows` queries is
    rows * NP + NP + 4 * (rows // 128) doubles (include/ipp_b200.h)."""
    rows_all = (int(m) + 127) // 128 * 128
    have_rows = max(0, (int(have_numel) - NP) // (NP + 1) // 128 * 128)
    budget = min(4 << 30, int(free_bytes) // 4)
    fit = max(int(budget // (8 * NP)), have_rows)
    if rows_all <= fit:
        return max(128, rows_all)
    wave = 128 * int(sms)
    return max(128, fit // wave * wave if fit >= wave else fit // 128 * 128)


class GaussianProcessRegressor:
    """Drop-in for sklearn.gaussian_process.GaussianProcessRegressor on the reference's call sites."""

    def __init__(self, kernel=None, *, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0,
                 normalize_y=False, copy_X_train=True, n_targets=None, random_state=None):
        if kernel is None:
            kernel = ConstantRBFKernel(1.0, 1.0, "fixed", "fixed")
        if not isinstance(kernel, ConstantRBFKernel):
            raise TypeError("the B200 path implements ConstantKernel * RBF (point_source.py:86)")
        if kernel.nu != np.inf:
            raise NotImplementedError("Matern is available in the kernel-matrix builder (kernel(X[, Y])); the fit and the "
                                      "posterior implement the reference's C * RBF")
        if np.iterable(alpha):
            raise TypeError("per-sample alpha is not used by the reference and is not implemented")
        self.kernel = kernel
        self.alpha = float(alpha)
        self.optimizer = optimizer
        self.n_restarts_optimizer = int(n_restarts_optimizer)
        self.normalize_y = normalize_y
        self.copy_X_train = copy_X_train
        self.n_targets = n_targets
        self.random_state = random_state
        self._ws = _Workspace()
        self._count_lock = threading.Lock()
        self._restart_slots = []
        # optimiser restarts are independent once their start points are drawn: run them on this many streams
        # (None = automatic: 6 streams, 3 when the Cholesky alone fills the GPU; 1 = serial)
        self.restart_streams = None
        # a torch.distributed group (or True for the default group): spread the restarts of one fit over its ranks
        # (start i on rank i mod P, dist.sharded_restarts); every rank must call fit with the same data and RNG state
        self.restart_group = None
        # block sparsity of the lattice posterior: entries of k* below skip_below * c are not computed (0: dense)
        self.skip_below = 1e-30
        # order in which the optimiser's objective sees the observations: "zcurve" (block-sparse gradient contraction)
        # or "given"
        self.objective_order = os.environ.get("IPP_OBJECTIVE_ORDER", "zcurve")
        # ... when the kernel's reach (the skip cutoff) is below this fraction of the data extent: beyond it every
        # tile reaches (nearly) every slab and the clustered order buys nothing
        self.skip_reach_max = 0.8
        self.cluster_order = "maximin"  # order of the spatial clusters in the sparse state (see clustered_order)
        self._scratch = None
        self._batch_ws = None
        self.n_lml_evals_ = 0
        self.lazy = False
        self._pending = False

    # -- helpers ------------------------------------------------------------------------------
    def _rng_state(self):
        rs = self.random_state
        if rs is None or rs is np.random:
            return np.random.mtrand._rand  # sklearn.utils.check_random_state(None): the global stream
        if isinstance(rs, (int, np.integer)):
            return np.random.RandomState(rs)
        return rs

    def _upload_training(self, X, yn):
        torch = _lib.require_cuda()
        n = X.shape[0]
        stage = self._ws.pinned("train", 3 * n)  # one pinned buffer, one H2D copy: [X (n x 2) | y (n)]
        stage[: 2 * n].view(n, 2).copy_(torch.from_numpy(X))
        stage[2 * n:].copy_(torch.from_numpy(np.ascontiguousarray(yn)))
        dev = torch.empty(3 * n, dtype=torch.float64, device="cuda")
        dev.copy_(stage, non_blocking=True)
        self._Xd = dev[: 2 * n].view(n, 2)
        self._yd = dev[2 * n:]
        self.h2d_bytes_ = 24 * n
        self._perm = np.arange(n)
        self._Xd_fit, self._yd_fit, self._bbox_d = self._Xd, self._yd, None
        # The OBJECTIVE sees the observations along a Z-order curve (the LML and its gradient do not depend on the order
        # of the observations; the rounding does, deterministically): 128 consecutive ones are then a compact patch
        # of the workspace, and the gradient contraction skips the tile pairs whose patches are out of each other's
        # reach (ipp_gp_lml_grad_batch, tile_bbox).  Small problems keep the given order (one or two tiles anyway).
        self._Xd_obj, self._yd_obj, self._tile_bbox_d = self._Xd, self._yd, None
        if self.objective_order == "zcurve" and n >= 1024:
            order = morton_order(X)
            od = torch.from_numpy(order).cuda()
            self._Xd_obj = self._Xd.index_select(0, od).contiguous()
            self._yd_obj = self._yd.index_select(0, od).contiguous()
            NP = (n + 63) // 64 * 64
            tile = 64 if NP <= 1536 else 128
            self._tile_bbox_d = torch.from_numpy(slab_bounding_boxes(X[order], (NP + tile - 1) // tile * tile, slab=tile)).cuda()

    def _eval_device(self, c, l, with_grad, final=False, ws=None, max_ctas=0):
        """Launch one LML (+gradient) evaluation or the final factorisation on the calling thread's
        current stream; returns the 8-double result block on the host."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        ws = self._ws if ws is None else ws
        n = self.X_train_.shape[0]
        ws.hyp_host[0] = c
        ws.hyp_host[1] = l
        ws.hyp_host[2] = self.alpha
        ws.hyp.copy_(ws.hyp_host, non_blocking=True)
        st = _lib.stream_ptr()
        flags = _chol_flags()
        if final:
            Xf, yf = (self._Xd_fit, self._yd_fit) if final == "morton" else (self._Xd, self._yd)
            rc = lib.ipp_gp_fit_final_plan(_lib.ptr(Xf), _lib.ptr(yf), n, _lib.ptr(ws.hyp), _lib.ptr(ws.K),
                                           _lib.ptr(ws.W), _lib.ptr(ws.T), _lib.ptr(ws.v), _lib.ptr(ws.plan),
                                           _lib.ptr(ws.sched), flags, st)
            _lib.check(rc, "ipp_gp_fit_final_plan")
        elif ws.plan is not None:
            rc = lib.ipp_gp_lml_grad_batch(_lib.ptr(self._Xd_obj), _lib.ptr(self._yd_obj), n, _lib.ptr(ws.hyp), _lib.ptr(ws.K),
                                           _lib.ptr(ws.W), _lib.ptr(ws.T), _lib.ptr(ws.v), None, 1, 1 if with_grad else 0,
                                           _lib.ptr(ws.plan), _lib.ptr(ws.sched), flags | (int(max_ctas) << 8),
                                           _lib.ptr(self._tile_bbox_d), st)
            _lib.check(rc, "ipp_gp_lml_grad_batch")
        else:
            rc = lib.ipp_gp_lml_grad(_lib.ptr(self._Xd), _lib.ptr(self._yd), n, _lib.ptr(ws.hyp), _lib.ptr(ws.K),
                                     _lib.ptr(ws.W), _lib.ptr(ws.T), _lib.ptr(ws.v), 1 if with_grad else 0, int(max_ctas), st)
            _lib.check(rc, "ipp_gp_lml_grad")
        ws.res_host[:8].copy_(ws.result, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = ws.res_host[:8].numpy().copy()
        _check_watchdog(res[3])
        return res

    def _eval_batch(self, ws, batch, thetas, with_grad):
        """One LML (+ gradient) evaluation of every restart in `batch` (slot ids) as ONE launch sequence
        (ipp_gp_lml_grad_batch); returns the 8-double result block of each."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        n = self.X_train_.shape[0]
        R = len(batch)
        hh = ws.hyp_host.numpy()
        for slot, th in zip(batch, thetas):
            hh[4 * slot:4 * slot + 3] = (float(np.exp(th[0])), float(np.exp(th[1])), self.alpha)
        ws.slot_ids_host.numpy()[:R] = batch
        ws.hyp.copy_(ws.hyp_host, non_blocking=True)
        ws.slot_ids.copy_(ws.slot_ids_host, non_blocking=True)
        flags = _chol_flags()
        rc = lib.ipp_gp_lml_grad_batch(_lib.ptr(self._Xd_obj), _lib.ptr(self._yd_obj), n, _lib.ptr(ws.hyp), _lib.ptr(ws.K),
                                       _lib.ptr(ws.W), _lib.ptr(ws.T), _lib.ptr(ws.v), _lib.ptr(ws.slot_ids), R,
                                       1 if with_grad else 0, _lib.ptr(ws.plan), _lib.ptr(ws.sched), flags,
                                       _lib.ptr(self._tile_bbox_d), _lib.stream_ptr())
        _lib.check(rc, "ipp_gp_lml_grad_batch")
        off, vt, S = ws.layout[8], ws.layout[6], ws.slots
        torch.index_select(ws.v[: S * vt].view(S, vt)[:, off:off + 8], 0, ws.slot_ids[:R].long(), out=ws.res_dev.view(S, 8)[:R])
        ws.res_host[:8 * R].copy_(ws.res_dev[:8 * R], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = ws.res_host[:8 * R].numpy().reshape(R, 8).copy()
        for row in res:
            _check_watchdog(row[3])
        with self._count_lock:
            self.n_lml_evals_ += R
        return res

    # -- sklearn API ---------------------------------------------------------------------------
    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        """_gpr.py:541-656.  A failed Cholesky gives (-inf, zeros) like the reference (:590-593)."""
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        self._ensure()
        theta = np.asarray(theta, dtype=float)
        if not clone_kernel:
            self._kernel_fitted.theta = theta
        return self._lml(theta, eval_gradient, ws=self._scratch_ws())  # never the buffers the posterior reads

    def _scratch_ws(self):
        if self._scratch is None:
            self._scratch = _Workspace()
        return self._scratch.ensure(self.X_train_.shape[0])

    def _lml(self, theta, eval_gradient, ws=None, max_ctas=0):
        theta = np.asarray(theta, dtype=float)
        c, l = float(np.exp(theta[0])), float(np.exp(theta[1]))
        res = self._eval_device(c, l, eval_gradient, ws=ws, max_ctas=max_ctas)
        with self._count_lock:
            self.n_lml_evals_ += 1
        if res[3] != 0.0 or not np.isfinite(res[0]):
            return (-np.inf, np.zeros_like(theta)) if eval_gradient else -np.inf
        if eval_gradient:
            return float(res[0]), np.array([res[1], res[2]])
        return float(res[0])

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        """_gpr.py:658-676: scipy L-BFGS-B with its defaults, or a user callable."""
        if self.optimizer == "fmin_l_bfgs_b":
            res = minimize(obj_func, initial_theta, method="L-BFGS-B", jac=True, bounds=bounds)
            return res.x, res.fun
        if callable(self.optimizer):
            return self.optimizer(obj_func, initial_theta, bounds=bounds)
        raise ValueError(f"Unknown optimizer {self.optimizer}.")

    def fit(self, X, y):
        """_gpr.py:233-368 (single-output).

        The restart start points are drawn from the RNG here, in the reference's order
        (_gpr.py:329-333; nothing else in fit consumes random numbers).  With `self.lazy = True`
        the optimisation itself is deferred until fitted state is first read (`kernel_`, `alpha_`,
        `L_`, `predict`, ...): a later `fit` supersedes a pending one, which is what makes the
        per-step refits of rrt.py:339,360,749 -- whose results the reference never reads -- free."""
        X = _as_points(X)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        if X.shape[0] != y.shape[0]:
            raise ValueError(f"X and y have inconsistent lengths: {X.shape[0]} != {y.shape[0]}")
        if X.shape[0] == 0:
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required.")
        self._kernel_fitted = self.kernel.clone()
        self._rng = self._rng_state()
        if self.normalize_y:
            self._y_train_mean = np.mean(y, axis=0)
            std = np.std(y, axis=0)
            # sklearn's _handle_zeros_in_scale, scalar branch (y is 1-D here): only an exactly zero spread becomes 1
            self._y_train_std = 1.0 if std == 0.0 else std
            y = (y - self._y_train_mean) / self._y_train_std
        else:
            self._y_train_mean = 0.0
            self._y_train_std = 1.0
        self.X_train_ = np.copy(X) if self.copy_X_train else X
        self.y_train_ = np.copy(y) if self.copy_X_train else y
        self._L_host = None
        self._alpha_host = None
        k = self._kernel_fitted
        fixed = isinstance(k.constant_value_bounds, str) or isinstance(k.length_scale_bounds, str)
        self._starts = None
        if self.optimizer is not None and not fixed:
            bounds = k.bounds
            if self.n_restarts_optimizer > 0 and not np.isfinite(bounds).all():
                raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all "
                                 "bounds are finite.")
            self._starts = [k.theta] + [self._rng.uniform(bounds[:, 0], bounds[:, 1])
                                        for _ in range(self.n_restarts_optimizer)]
        self._pending = True
        if not self.lazy:
            self._ensure()
        return self

    def _ensure(self):
        """Run the deferred optimisation + final factorisation (no-op when up to date)."""
        if not getattr(self, "_pending", False):
            return
        self._pending = False
        self._ws.ensure(self.X_train_.shape[0])
        self._upload_training(self.X_train_, self.y_train_)
        if self._starts is not None:
            def obj_func(theta, eval_gradient=True):
                if eval_gradient:
                    lml, grad = self._lml(theta, True)
                    return -lml, -grad
                return -self._lml(theta, False)

            bounds = self._kernel_fitted.bounds

            def optimise(starts):
                if len(starts) > 1 and self.restart_streams is None and chol_mode() == "queue":
                    return self._optimise_lockstep(starts, bounds)
                workers = self._restart_workers(len(starts))
                if workers > 1:
                    return self._optimise_concurrently(starts, bounds, workers)
                return [self._constrained_optimization(obj_func, t0, bounds) for t0 in starts]

            group = getattr(self, "restart_group", None)
            if group is not None and len(self._starts) > 1:
                # start i on rank i mod P, one all_gather of (theta*, f*); every rank then factors at theta*
                from . import dist as D
                optima = D.sharded_restarts(lambda idx: optimise([self._starts[i] for i in idx]), len(self._starts),
                                            len(self._starts[0]), group=None if group is True else group, device="cuda")
            else:
                optima = optimise(self._starts)
            lml_values = [o[1] for o in optima]
            self._kernel_fitted.theta = optima[int(np.argmin(lml_values))][0]
            self._lml_value = -np.min(lml_values)
            self._finalize(set_lml=False)
        else:
            self._finalize(set_lml=True)

    def _restart_workers(self, n_starts):
        if n_starts <= 1:
            return 1
        if self.restart_streams is not None:
            return max(1, min(int(self.restart_streams), n_starts))
        return min(n_starts, 6 if self._ws.layout[0] <= 2048 else 3)  # measured best on a B200 (tools/fit_tune.py)

    def _optimise_lockstep(self, starts, bounds):
        """The restarts of _gpr.py:312-333 in lockstep: every L-BFGS-B run lives in its own host thread, and whenever
        each run that is still iterating has asked for its next objective value, ONE batched launch sequence
        (ipp_gp_lml_grad_batch) evaluates them all -- the dataflow Cholesky interleaves the factorisations, so the
        dependency chain of one fills the idle SMs of the others.  Each run sees exactly the values it would see
        alone (per-problem arithmetic does not depend on the batch), so optima, evaluation counts and the
        first-minimum choice are those of the serial loop."""
        torch = _lib.require_cuda()
        n = self.X_train_.shape[0]
        R = len(starts)
        if self._batch_ws is None:
            self._batch_ws = _Workspace()
        # one buffer set (K, W, T: 24 NP^2 bytes) per restart in flight: all of them when they fit in half of the free
        # device memory (n = 5000: 11 x 0.6 GB), else the restarts go in groups
        NP = (n + 63) // 64 * 64
        per_slot = 3 * 8 * NP * (NP + 64)
        have = self._batch_ws.slots if self._batch_ws.np_ == NP else 0  # (already allocated: not counted against the free memory)
        room = int(0.5 * torch.cuda.mem_get_info()[0] // per_slot) + have
        if room < R:
            group = max(1, room)
            return [o for lo in range(0, R, group) for o in self._optimise_lockstep(starts[lo:lo + group], bounds)]
        ws = self._batch_ws.ensure(n, slots=R)
        device = torch.cuda.current_device()
        stream = torch.cuda.current_stream()
        cond = threading.Condition()
        requests, results, live, outs, errors = {}, {}, set(range(R)), [None] * R, []
        self.lockstep_rounds_ = []

        def run(i):
            def obj_func(theta, eval_gradient=True):
                with cond:
                    requests[i] = (np.array(theta, dtype=float), bool(eval_gradient))
                    cond.notify_all()
                    while i not in results:
                        cond.wait()
                    res = results.pop(i)
                if isinstance(res, BaseException):
                    raise res
                bad = res[3] != 0.0 or not np.isfinite(res[0])
                if eval_gradient:
                    return (np.inf, np.zeros_like(theta)) if bad else (-float(res[0]), -np.array([res[1], res[2]]))
                return np.inf if bad else -float(res[0])

            try:
                outs[i] = self._constrained_optimization(obj_func, starts[i], bounds)
            except BaseException as e:  # noqa: BLE001 -- re-raised by the coordinating thread
                errors.append(e)
            finally:
                with cond:
                    live.discard(i)
                    cond.notify_all()

        threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(R)]
        for th in threads:
            th.start()
        while True:
            with cond:
                while live and len(requests) < len(live):
                    cond.wait()
                if not live:
                    break
                batch = sorted(requests)
                reqs = [requests.pop(i) for i in batch]
            self.lockstep_rounds_.append(len(batch))  # batch size of every launch sequence (reported by bench.py)
            try:
                with torch.cuda.device(device), torch.cuda.stream(stream):
                    res = self._eval_batch(ws, batch, [r[0] for r in reqs], any(r[1] for r in reqs))
                out = {i: res[j] for j, i in enumerate(batch)}
            except BaseException as e:  # noqa: BLE001 -- hand the failure to every waiting run
                out = {i: e for i in batch}
            with cond:
                results.update(out)
                cond.notify_all()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        return outs

    def _optimise_concurrently(self, starts, bounds, workers):
        """The restarts of _gpr.py:312-333 side by side: one host thread, one CUDA stream and one device
        buffer set per worker.  Every optimisation sees exactly the objective values it would see alone
        (same kernels, same launch geometry), so the optima -- and the first-minimum choice among them,
        taken in start order -- are those of the serial loop; only the wall time changes."""
        import queue
        from concurrent.futures import ThreadPoolExecutor

        torch = _lib.require_cuda()
        n = self.X_train_.shape[0]
        device = torch.cuda.current_device()
        while len(self._restart_slots) < workers:
            self._restart_slots.append((_Workspace(), torch.cuda.Stream()))
        slots = queue.SimpleQueue()
        for ws, stream in self._restart_slots[:workers]:
            ws.ensure(n)
            slots.put((ws, stream))
        uploaded = torch.cuda.Event()
        uploaded.record()
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        cap = max(8, sms // workers)

        def run(t0):
            torch.cuda.set_device(device)
            ws, stream = slots.get()
            try:
                with torch.cuda.stream(stream):
                    stream.wait_event(uploaded)

                    def obj_func(theta, eval_gradient=True):
                        if eval_gradient:
                            lml, grad = self._lml(theta, True, ws=ws, max_ctas=cap)
                            return -lml, -grad
                        return -self._lml(theta, False, ws=ws, max_ctas=cap)

                    return self._constrained_optimization(obj_func, t0, bounds)
            finally:
                slots.put((ws, stream))

        with ThreadPoolExecutor(max_workers=workers) as pool:
            return list(pool.map(run, starts))

    # fitted attributes resolve a pending lazy fit first
    @property
    def kernel_(self):
        self._ensure()
        return self._kernel_fitted

    @property
    def log_marginal_likelihood_value_(self):
        self._ensure()
        return self._lml_value

    def _finalize(self, set_lml):
        k = self._kernel_fitted
        self._choose_order(k.length_scale)
        res = self._eval_device(k.constant_value, k.length_scale, False, final="morton")
        if res[3] != 0.0:
            raise np.linalg.LinAlgError(
                f"The kernel, {k}, is not returning a positive definite matrix. Try gradually "
                "increasing the 'alpha' parameter of your GaussianProcessRegressor estimator.",
                f"{int(res[3])}-th leading minor of the array is not positive definite")
        if set_lml:
            self._lml_value = float(res[0])
        self._fitted_c = k.constant_value
        self._fitted_l = k.length_scale

    def _skip_cutoff(self, length_scale, skip=None):
        """Distance beyond which every entry of k* is below skip_below * c (0.0: dense)."""
        skip = self.skip_below if skip is None else skip
        return float(length_scale) * float(np.sqrt(2.0 * np.log(1.0 / skip))) if (skip and 0.0 < skip < 1.0) else 0.0

    def _choose_order(self, length_scale):
        """Order of the observations in the FINAL factorisation.  When the kernel's reach (the skip cutoff) is
        well inside the extent of the data, the observations are put in Morton order so that 16-observation slabs
        of k* are spatial clusters the lattice kernel can skip; alpha_ / L_ undo the permutation.  Otherwise the
        given order is kept: nothing could be skipped, and a spatially sorted elimination order is the less
        accurate one for the explicit inverse when the kernel is wide (every row nearly repeats its predecessor)."""
        torch = _lib.require_cuda()
        X = self.X_train_
        n = X.shape[0]
        cutoff = self._skip_cutoff(length_scale)
        extent = float(np.max(X.max(axis=0) - X.min(axis=0))) if n > 1 else 0.0
        if cutoff > 0.0 and cutoff < self.skip_reach_max * extent and n >= 256:
            self._perm = clustered_order(X, cluster_order=self.cluster_order)
            pd = torch.from_numpy(self._perm).cuda()
            self._Xd_fit = self._Xd.index_select(0, pd).contiguous()
            self._yd_fit = self._yd.index_select(0, pd).contiguous()
            self._bbox_h = slab_bounding_boxes(X[self._perm], self._ws.layout[0])
            self._bbox_d = torch.from_numpy(self._bbox_h).cuda()
        else:
            self._perm = np.arange(n)
            self._Xd_fit, self._yd_fit, self._bbox_d = self._Xd, self._yd, None

    # -- fitted attributes (lazy device -> host) ------------------------------------------------
    @property
    def L_(self):
        self._ensure()
        if self._L_host is None:
            # sklearn's L_ is the factor in the ORIGINAL order of the observations: factor once more, on demand
            n = self.X_train_.shape[0]
            ws = self._scratch_ws()
            res = self._eval_device(self._fitted_c, self._fitted_l, False, final="original", ws=ws)
            if res[3] != 0.0:
                raise np.linalg.LinAlgError(f"{int(res[3])}-th leading minor of the array is not positive definite")
            NP, ld = ws.layout[0], ws.layout[2]
            self._L_host = ws.K[: NP * ld].view(NP, ld)[:n, :n].cpu().numpy()
        return self._L_host

    @property
    def alpha_(self):
        self._ensure()
        if self._alpha_host is None:
            n = self.X_train_.shape[0]
            off = self._ws.layout[7]
            a = np.empty(n)
            a[self._perm] = self._ws.v[off:off + n].cpu().numpy()  # device state is in Morton order
            self._alpha_host = a
        return self._alpha_host

    # -- prediction -----------------------------------------------------------------------------
    def _posterior_args(self):
        self._ensure()
        return (_lib.ptr(self._ws.W), _lib.ptr(self._ws.v), int(self.X_train_.shape[0]), float(self._fitted_c),
                float(self._fitted_l), float(self._y_train_mean), float(self._y_train_std))

    def _warn_negative(self, counter):
        if int(counter.item()) > 0:
            warnings.warn("Predicted variances smaller than 0. Setting those variances to 0.")

    def predict(self, X, return_std=False, return_cov=False):
        """_gpr.py:370-500.  Unfitted: GP prior (mean 0, std sqrt(c))."""
        if return_cov:
            raise NotImplementedError("return_cov is not used by the reference and is not implemented")
        Xq = _as_points(X)
        if not hasattr(self, "X_train_"):
            mean = np.zeros(Xq.shape[0])
            return (mean, np.sqrt(self.kernel.diag(Xq))) if return_std else mean
        torch = _lib.require_cuda()
        lat = as_lattice(Xq)
        if lat is not None:  # the reference's own call shape: gp.predict(column_stack(X.ravel(), Y.ravel()))
            mean_d, std_d, _ = self.predict_grid_device(*lat, want_zpred=False)
            self._warn_negative(self._neg_counter)
        else:
            mean_d, std_d, _ = self.predict_points_device(torch.from_numpy(Xq).cuda())
        mean = mean_d.cpu().numpy()
        return (mean, std_d.cpu().numpy()) if return_std else mean

    # explicit points: from this many queries on, the cross-covariance goes through a table and the lattice kernel's
    # pipeline (ipp_gp_posterior_points_tabled); below, the generic kernel (one tile or two: latency either way)
    TABLED_MIN = 256

    def _table_scratch(self, need):
        torch = _lib.require_cuda()
        buf = getattr(self._ws, "tables", None)
        if buf is None or buf.numel() < need:
            self._ws.tables = None
            buf = torch.empty(int(need), dtype=torch.float64, device="cuda")
            self._ws.tables = buf
        return buf

    def _point_table_rows(self, m):
        """Queries per chunk of the tabled explicit-point posterior (point_table_rows) for this device's free memory."""
        torch = _lib.require_cuda()
        have = getattr(self._ws, "tables", None)
        return point_table_rows(m, self._ws.layout[0], torch.cuda.mem_get_info()[0], 0 if have is None else have.numel(),
                                torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count)

    @staticmethod
    def _morton_order_device(Qd):
        """morton_order() of device-resident points (same keys: 16 bits per axis over the points' own bounding box)."""
        import torch
        lo, hi = Qd.min(dim=0).values, Qd.max(dim=0).values
        span = torch.where(hi > lo, hi - lo, torch.ones_like(hi))
        q = ((Qd - lo) / span * 65536.0).to(torch.int64).clamp_(max=65535)

        def spread(v):
            v = (v | (v << 8)) & 0x00FF00FF
            v = (v | (v << 4)) & 0x0F0F0F0F
            v = (v | (v << 2)) & 0x33333333
            v = (v | (v << 1)) & 0x55555555
            return v

        return torch.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), stable=True)

    def predict_points_device(self, Qd, want_zpred=False, tabled=None, skip_below=None, table_rows=None):
        """Posterior at explicit device-resident points (m x 2); returns device tensors.  tabled (default: from
        TABLED_MIN queries on): the cross-covariance table + lattice pipeline; with a state fitted in clustered order
        the queries are processed along the Z-order curve so that the 16-observation slabs out of a tile's reach are
        skipped (skip_below=0: dense), and the results are returned in the caller's order."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        m = int(Qd.shape[0])
        mean = torch.empty(m, dtype=torch.float64, device="cuda")
        std = torch.empty(m, dtype=torch.float64, device="cuda")
        zp = torch.empty(m, dtype=torch.float64, device="cuda") if want_zpred else None
        neg = torch.zeros(1, dtype=torch.int32, device="cuda")
        if tabled is None:
            tabled = m >= self.TABLED_MIN
        if m and not tabled:
            rc = lib.ipp_gp_posterior_points(*self._posterior_args(), _lib.ptr(Qd), m, _lib.ptr(mean), _lib.ptr(std),
                                             _lib.ptr(zp), _lib.ptr(neg), 0, _lib.stream_ptr())
            _lib.check(rc, "ipp_gp_posterior_points")
            self._warn_negative(neg)
        elif m:
            args = self._posterior_args()
            Qd = Qd.contiguous()
            cutoff = self._skip_cutoff(self._fitted_l, skip_below) if self._bbox_d is not None else 0.0
            order = self._morton_order_device(Qd) if cutoff > 0.0 else None
            if order is not None:
                Qd = Qd.index_select(0, order)
                out = [torch.empty_like(mean), torch.empty_like(std), torch.empty_like(zp) if want_zpred else None]
            else:
                out = [mean, std, zp]
            NP = self._ws.layout[0]
            rows = int(table_rows) if table_rows else self._point_table_rows(m)
            tables = self._table_scratch(rows * NP + NP + 4 * (rows // 128))
            rc = lib.ipp_gp_posterior_points_tabled(*args, _lib.ptr(Qd), m, _lib.ptr(out[0]), _lib.ptr(out[1]),
                                                    _lib.ptr(out[2]), _lib.ptr(neg), _lib.ptr(tables), rows,
                                                    _lib.ptr(self._bbox_d) if cutoff > 0.0 else None, float(cutoff), 0,
                                                    _lib.stream_ptr())
            _lib.check(rc, "ipp_gp_posterior_points_tabled")
            if order is not None:
                mean[order] = out[0]
                std[order] = out[1]
                if want_zpred:
                    zp[order] = out[2]
            self._warn_negative(neg)
        return mean, std, zp

    def predict_grid_host(self, x0, x1, nx, y0, y1, ny, g_begin=0, g_end=None):
        """Fused lattice posterior returned as fresh host arrays (zpred, std): device outputs are
        read back through one pinned staging buffer (2 * m doubles, one D2H copy)."""
        torch = _lib.require_cuda()
        g_end = nx * ny if g_end is None else g_end
        m = int(g_end - g_begin)
        out = torch.empty(3 * m, dtype=torch.float64, device="cuda")  # [std | zpred | mean]
        self.predict_grid_device(x0, x1, nx, y0, y1, ny, g_begin, g_end, want_zpred=True,
                                 out=(out[2 * m:], out[:m], out[m:2 * m]))
        stage = self._ws.pinned("grid", 2 * m)
        stage.copy_(out[: 2 * m], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._warn_negative(self._neg_counter)
        host = stage.numpy()
        self.d2h_bytes_ = 16 * m
        return host[m:].copy(), host[:m].copy()

    def lattice_work(self, x0, x1, nx, y0, y1, ny, g_begin=0, g_end=None, skip_below=None, per_tile=False):
        """(executed, dense) counts of 128-query x 128-row x 16-observation DMMA slabs of the lattice kernel for
        this fitted state: the host restatement of the kernel's own slab test (lattice_slab_mask in
        csrc/ipp_posterior.cu), used by bench.py to turn a launch time into executed flop/s."""
        self._ensure()
        n = self.X_train_.shape[0]
        g_end = nx * ny if g_end is None else g_end
        first = n % 128
        rlims = ([first] if first else []) + [first + 128 * j for j in range(1, n // 128 + 1)] if first else \
            [128 * (j + 1) for j in range(n // 128)]
        kslabs = np.array([(r + 15) // 16 for r in rlims])
        nt = (g_end - g_begin + 127) // 128
        dense = int(kslabs.sum()) * int(nt)
        cutoff = self._skip_cutoff(self._fitted_l, skip_below) if self._bbox_d is not None else 0.0
        if cutoff <= 0.0:
            return np.full(nt, int(kslabs.sum()), dtype=np.int64) if per_tile else (dense, dense)
        bb = self._bbox_d.cpu().numpy()
        xstep = (x1 - x0) / (nx - 1) if nx > 1 else 0.0
        ystep = (y1 - y0) / (ny - 1) if ny > 1 else 0.0
        gq0 = g_begin + 128 * np.arange(nt)
        nvalid = np.minimum(128, g_end - gq0)
        iy0, ix0 = gq0 // nx, gq0 % nx
        ixl = ix0 + nvalid - 1
        same = ixl < nx
        qx0 = np.where(same, x0 + ix0 * xstep, x0)
        qx1 = np.where(same, x0 + ixl * xstep, x1)
        qy0 = y0 + iy0 * ystep
        qy1 = np.where(same, qy0, y0 + (iy0 + 1) * ystep)
        executed = 0
        tiles = np.zeros(nt, dtype=np.int64)
        for lo in range(0, nt, 2048):  # chunked: tiles x slabs
            sl = slice(lo, lo + 2048)
            # (padding slabs sit at 1e300: their gap squares to inf, i.e. inactive)
            gx = np.maximum(0.0, np.maximum(bb[None, :, 0] - qx1[sl, None], qx0[sl, None] - bb[None, :, 1]))
            gy = np.maximum(0.0, np.maximum(bb[None, :, 2] - qy1[sl, None], qy0[sl, None] - bb[None, :, 3]))
            with np.errstate(over="ignore"):
                cum = np.cumsum(gx * gx + gy * gy <= cutoff * cutoff, axis=1)
            tiles[sl] = cum[:, kslabs - 1].sum(axis=1)
            executed += int(tiles[sl].sum())
        return tiles if per_tile else (executed, dense)

    def lattice_row_work(self, x0, x1, nx, y0, y1, ny, skip_below=None):
        """Cheap per-row estimate (ny values, arbitrary unit) of the lattice kernel's executed slabs, for splitting one
        lattice over GPUs by work (dist.lattice_slabs): a slab counts for a row when its box is within the cutoff in y,
        weighted by the fraction of the row within the cutoff in x and by the row blocks of W that contain it.
        ny x n_slabs host operations (about 1 ms for C4) against the exact per-tile count of lattice_work."""
        self._ensure()
        n = self.X_train_.shape[0]
        cutoff = self._skip_cutoff(self._fitted_l, skip_below) if self._bbox_d is not None else 0.0
        if cutoff <= 0.0:
            return np.ones(int(ny))
        bb = self._bbox_h                                              # host copy: no device sync before the launch
        real = bb[:, 0] < 1e299
        bb = bb[real]
        first = n % 128
        slab_lo = 16 * np.arange(len(real))[real]                      # first observation of the slab
        rlims = np.array(([first] if first else []) + [first + 128 * j for j in range(1, n // 128 + 1)])
        mult = (rlims[None, :] > slab_lo[:, None]).sum(axis=1).astype(np.float64)   # row blocks of W reading the slab
        span = max(float(x1 - x0), 1e-300)
        wx = np.clip(np.minimum(x1, bb[:, 1] + cutoff) - np.maximum(x0, bb[:, 0] - cutoff), 0.0, None) / span
        ys = np.linspace(y0, y1, int(ny))
        gy = np.maximum(0.0, np.maximum(bb[None, :, 2] - ys[:, None], ys[:, None] - bb[None, :, 3]))
        return (gy <= cutoff) @ (wx * mult) + 1e-9

    def _lattice_tables(self, nx, ny):
        """Scratch for the lattice fast path of ipp_gp_posterior_grid: (nx + ny) * NP doubles."""
        return self._table_scratch((int(nx) + int(ny)) * self._ws.layout[0])

    def predict_grid_device(self, x0, x1, nx, y0, y1, ny, g_begin=0, g_end=None, want_zpred=True, out=None,
                            lattice=True, skip_below=None):
        """Fused posterior over flat lattice indices [g_begin, g_end) (index = iy * nx + ix).
        Returns device tensors (mean, std, zpred) of length g_end - g_begin."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        g_end = nx * ny if g_end is None else g_end
        m = int(g_end - g_begin)
        if out is None:
            mean = torch.empty(m, dtype=torch.float64, device="cuda")
            std = torch.empty(m, dtype=torch.float64, device="cuda")
            zp = torch.empty(m, dtype=torch.float64, device="cuda") if want_zpred else None
        else:
            mean, std, zp = out
        neg = torch.zeros(1, dtype=torch.int32, device="cuda")
        if m:
            args = self._posterior_args()
            tables = self._lattice_tables(nx, ny) if (lattice and nx >= 128) else None
            cutoff = self._skip_cutoff(self._fitted_l, skip_below) if self._bbox_d is not None else 0.0
            rc = lib.ipp_gp_posterior_grid(*args, float(x0), float(x1), int(nx), float(y0), float(y1),
                                           int(ny), int(g_begin), int(g_end), _lib.ptr(mean), _lib.ptr(std),
                                           _lib.ptr(zp), _lib.ptr(neg), _lib.ptr(tables),
                                           _lib.ptr(self._bbox_d) if cutoff > 0.0 else None, float(cutoff), 0,
                                           _lib.stream_ptr())
            _lib.check(rc, "ipp_gp_posterior_grid")
        self._neg_counter = neg
        return mean, std, zp

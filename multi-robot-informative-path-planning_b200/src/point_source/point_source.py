"""Point-source radiation field + the GP estimator: drop-in for the reference's
src/point_source/point_source.py (`PointSourceField`).

Hot path: `update`, `predict_spatial_field`, `.gp`, `.gp.kernel` run on the
B200 through mripp_b200.gp.  `predict_spatial_field` never builds the (G x n) cross-covariance or
the (n x G) triangular solve of sklearn's predict: the lattice is generated inside the posterior
kernel from (x0, x1, nx, y0, y1, ny), and mean / std / 10**mean come back as three flat arrays.

Everything else (random source layout, analytic ground truth, simulated measurements) is host
numpy with the reference's expressions and RNG order, because the measurement VALUES feed every
later decision (point_source.py:89-97, 111-212, 339-344).
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Tuple

import numpy as np

from mripp_b200.gp import ConstantKernel, GaussianProcessRegressor, RBF

AIR_ABSORPTION = 0.001205       # point_source.py:130,169
CONCRETE_ABSORPTION = 2.35      # point_source.py:131,170


def _segments_intersect(x1, y1, x2, y2, x3, y3, x4, y4) -> bool:
    """point_source.py:279-297."""
    d1x, d1y = x2 - x1, y2 - y1
    d2x, d2y = x4 - x3, y4 - y3
    den = d1x * d2y - d1y * d2x
    if den == 0:
        return False
    t = ((x3 - x1) * d2y - (y3 - y1) * d2x) / den
    u = ((x3 - x1) * d1y - (y3 - y1) * d1x) / den
    return 0 <= t <= 1 and 0 <= u <= 1


class DeviceField:
    """A posterior field that lives on the device and behaves like the ndarray the reference returns: any numpy
    consumer (np.asarray, ufuncs, indexing, reshape) gets the host copy, made once on first use; a consumer that knows
    about it (main.py's metrics) reads `.tensor` and never triggers the copy."""

    def __init__(self, tensor):
        self.tensor = tensor
        self._host = None

    def host(self):
        if self._host is None:
            self._host = self.tensor.cpu().numpy()
        return self._host

    def __array__(self, dtype=None, copy=None):
        h = self.host()
        return h if dtype is None else h.astype(dtype)

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def __len__(self):
        return int(self.tensor.shape[0])

    def __getitem__(self, key):
        return self.host()[key]

    def reshape(self, *shape):
        return self.host().reshape(*shape)

    def __getattr__(self, name):  # everything else (ravel, max, ...) is the host array's
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.host(), name)


class PointSourceField:
    """Same constructor, attributes and methods as the reference class (point_source.py:17-353)."""

    def __init__(self, num_sources: int = 1, workspace_size: Tuple[int, int, int, int] = (0, 40, 0, 40), obstacles: List = [],
                 intensity_range: Tuple[int, int] = (100000, 1000000), kernel_params: Optional[Dict[str, float]] = None,
                 seed: Optional[int] = None, gp_factory=None, ground_truth_mode: str = "auto"):
        if seed is not None:
            np.random.seed(seed)
        self.validate_inputs(num_sources, workspace_size, intensity_range)
        self.sources = self.generate_sources(num_sources, workspace_size, intensity_range)
        self.r_s = 0.5
        self.r_d = 0.5
        self.T = 100
        self.workspace_size = workspace_size
        self.obstacles = obstacles
        self.intensity_range = intensity_range
        w = workspace_size
        self.x = np.linspace(w[0], w[1], (w[1] - w[0]) * 10)
        self.y = np.linspace(w[2], w[3], (w[3] - w[2]) * 10)
        self.X, self.Y = np.meshgrid(self.x, self.y)
        # "host": the reference's numpy / Python-loop field; "device": ipp_ground_truth_grid; "auto": the device only
        # where the host costs minutes -- rectangle obstacles on a lattice of more than 10^5 points (the Python loop
        # over every lattice point, point_source.py:247-258: 100 us per point and obstacle).  Metrics only.
        self.ground_truth_mode = ground_truth_mode
        self.g_truth = self.ground_truth()
        kernel = self.construct_kernel(kernel_params)
        make = gp_factory if gp_factory is not None else GaussianProcessRegressor
        self.gp = make(kernel=kernel, n_restarts_optimizer=10, normalize_y=True)
        self.grid_group = None  # torch.distributed group (True: the default group): shard the lattice by row slab (mripp_b200.dist)
        # True: predict_spatial_field returns DeviceField views (numpy-compatible, read back only when a host consumer
        # touches them), so that a metrics-only caller (main.py with save = false) never copies the 16 B per lattice
        # point to the host: the metrics are reduced on the device (ipp_grid_metrics)
        self.device_outputs = False

    # -- construction helpers ---------------------------------------------------------------------
    def validate_inputs(self, num_sources, workspace_size, intensity_range) -> None:
        if num_sources < 1:
            raise ValueError("Number of sources must be at least 1.")
        if intensity_range[0] <= 0 or intensity_range[1] <= 0 or intensity_range[0] >= intensity_range[1]:
            raise ValueError("Intensity range must be positive and the lower bound must be less than the upper bound.")

    def construct_kernel(self, kernel_params: Optional[Dict[str, float]]):
        """point_source.py:82-87: C(sigma, (1e-3, 50)) * RBF(length_scale, (1e-2, 50))."""
        if kernel_params is None:
            kernel_params = {"sigma": 1.0, "length_scale": 1.0}
        return ConstantKernel(kernel_params["sigma"], (1e-3, 50)) * RBF(kernel_params["length_scale"], (1e-2, 50))

    def generate_sources(self, num_sources, workspace_size, intensity_range) -> List[List[float]]:
        """point_source.py:89-97: three uniform draws, then the global stream is RE-SEEDED FROM ENTROPY."""
        xs = np.random.uniform(workspace_size[0], workspace_size[1], num_sources)
        ys = np.random.uniform(workspace_size[2], workspace_size[3], num_sources)
        amps = np.random.uniform(*intensity_range, num_sources)
        np.random.seed(None)
        return np.column_stack((xs, ys, amps)).tolist()

    def update_source(self, source_index: int, new_x: float, new_y: float, new_A: float) -> None:
        if source_index < len(self.sources):
            self.sources[source_index] = [new_x, new_y, new_A]
            self.recalculate_ground_truth()
        else:
            print("Source index out of range.")

    def recalculate_ground_truth(self) -> None:
        self.g_truth = self.ground_truth()

    # -- physics (host) -----------------------------------------------------------------------------
    def intensity(self, waypoints) -> np.ndarray:
        """point_source.py:111-161: inverse-square law with air (and concrete) absorption."""
        wp = np.array(waypoints)
        if wp.ndim == 1:
            wp = wp[np.newaxis, :]
        src = np.array(self.sources)
        pos, amp = src[:, :2], src[:, 2]
        dist = np.linalg.norm(wp[:, np.newaxis, :] - pos[np.newaxis, :, :], axis=-1)
        dist = np.maximum(dist, 1e-6)
        absorb = np.exp(-AIR_ABSORPTION * dist)
        for ob in self.obstacles:
            if ob["type"] != "rectangle":
                continue
            for i in range(wp.shape[0]):
                for j in range(pos.shape[0]):
                    hit, length = self.compute_path_length_in_obstacle(
                        np.array([complex(wp[i][0], wp[i][1])]), complex(pos[j][0], pos[j][1]), ob)
                    if hit[0]:
                        absorb[i, j] *= np.exp(-CONCRETE_ABSORPTION * length[0])
        per_source = np.where(dist <= self.r_s, amp / (4 * np.pi * self.r_s ** 2), amp / (4 * np.pi * dist ** 2) * absorb)
        return np.sum(per_source, axis=-1)

    def ground_truth(self):
        """point_source.py:163-212 (see ground_truth_mode for where it is evaluated)."""
        mode = getattr(self, "ground_truth_mode", "host")
        if mode not in ("auto", "host", "device"):
            raise ValueError(f"ground_truth_mode must be 'auto', 'host' or 'device', not {mode!r}")
        rect_only = bool(self.obstacles) and all(o.get("type") == "rectangle" for o in self.obstacles)
        if mode == "device" or (mode == "auto" and rect_only and self.X.size > 100_000):
            return self.ground_truth_device()
        return self.ground_truth_host()

    def ground_truth_host(self):
        """point_source.py:163-212, the reference's own numpy expressions."""
        R = self.X + 1j * self.Y
        field = np.zeros(R.shape)
        for source in self.sources:
            A_n = source[2]
            R_n = source[0] + 1j * source[1]
            dist = np.abs(R - R_n)
            theta = np.arcsin(np.minimum(self.r_d / dist, 1))
            absorb = np.exp(-AIR_ABSORPTION * dist)
            for ob in self.obstacles:
                if ob["type"] == "rectangle":
                    hit, length = self.compute_path_length_in_obstacle(R, R_n, ob)
                    absorb[hit] *= np.exp(-CONCRETE_ABSORPTION * length[hit])
            inten = np.where(dist <= self.r_s, A_n / (4 * np.pi * self.r_s ** 2), A_n / (4 * np.pi * dist ** 2) * absorb)
            resp = np.where(dist <= self.r_d, 0.5 * A_n, 0.5 * A_n * (1 - np.cos(theta)) * absorb)
            field += inten + resp
        return field

    def ground_truth_device(self, as_tensor: bool = False):
        """The same field from the device (ipp_ground_truth_grid): one thread per lattice point instead of the
        Python loop over every point that the obstacle branch costs on the host (point_source.py:247-258).  Equal to
        `ground_truth()` to rounding (CUDA vs numpy exp / asin / cos); it feeds metrics only -- the measurements
        come from `intensity()` on the host, bit-exact."""
        from mripp_b200 import _lib

        torch = _lib.require_cuda()
        lib = _lib.load()
        w = self.workspace_size
        nx, ny = self.x.shape[0], self.y.shape[0]
        src = torch.from_numpy(np.ascontiguousarray(np.asarray(self.sources, dtype=np.float64).reshape(-1, 3))).cuda()
        rect = [[o["x"], o["y"], o["width"], o["height"]] for o in self.obstacles if o.get("type") == "rectangle"]
        ob = torch.from_numpy(np.asarray(rect, dtype=np.float64).reshape(-1, 4)).cuda() if rect else None
        out = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
        _lib.check(lib.ipp_ground_truth_grid(_lib.ptr(src), src.shape[0], _lib.ptr(ob), len(rect), float(w[0]), float(w[1]),
                                             nx, float(w[2]), float(w[3]), ny, float(self.r_s), float(self.r_d),
                                             _lib.ptr(out), _lib.stream_ptr()), "ipp_ground_truth_grid")
        return out.view(ny, nx) if as_tensor else out.view(ny, nx).cpu().numpy()

    def compute_path_length_in_obstacle(self, R, R_n, obstacle):
        """point_source.py:214-260: chord of the source->point ray inside a rectangle."""
        x_min, x_max = obstacle["x"], obstacle["x"] + obstacle["width"]
        y_min, y_max = obstacle["y"], obstacle["y"] + obstacle["height"]
        x_n, y_n = R_n.real, R_n.imag
        if np.isscalar(R):
            mask, length, pts = np.zeros(1, dtype=bool), np.zeros(1), [R]
        else:
            mask, length, pts = np.zeros(R.shape, dtype=bool), np.zeros(R.shape), R.flatten()
        for k, p in enumerate(pts):
            x_r, y_r = p.real, p.imag
            if self.line_intersects_rectangle(x_n, y_n, x_r, y_r, x_min, x_max, y_min, y_max):
                mask.flat[k] = True
                ex, ey, ox, oy = self.calculate_entry_exit_points(x_n, y_n, x_r, y_r, x_min, x_max, y_min, y_max)
                length.flat[k] = np.sqrt((ox - ex) ** 2 + (oy - ey) ** 2)
        return mask, length

    def line_intersects_rectangle(self, x1, y1, x2, y2, x_min, x_max, y_min, y_max):
        """point_source.py:262-277."""
        if (x_min <= x1 <= x_max and y_min <= y1 <= y_max) or (x_min <= x2 <= x_max and y_min <= y2 <= y_max):
            return True
        return (self.line_intersects_line(x1, y1, x2, y2, x_min, y_min, x_min, y_max)
                or self.line_intersects_line(x1, y1, x2, y2, x_max, y_min, x_max, y_max)
                or self.line_intersects_line(x1, y1, x2, y2, x_min, y_min, x_max, y_min)
                or self.line_intersects_line(x1, y1, x2, y2, x_min, y_max, x_max, y_max))

    def line_intersects_line(self, x1, y1, x2, y2, x3, y3, x4, y4):
        return _segments_intersect(x1, y1, x2, y2, x3, y3, x4, y4)

    def calculate_entry_exit_points(self, x1, y1, x2, y2, x_min, x_max, y_min, y_max):
        """point_source.py:299-331: first two boundary crossings, x-sides before y-sides."""
        hits = []
        if x1 != x2:
            for xb in (x_min, x_max):
                yy = y1 + (xb - x1) / (x2 - x1) * (y2 - y1)
                if y_min <= yy <= y_max:
                    hits.append((xb, yy))
        if y1 != y2:
            for yb in (y_min, y_max):
                xx = x1 + (yb - y1) / (y2 - y1) * (x2 - x1)
                if x_min <= xx <= x_max:
                    hits.append((xx, yb))
        if len(hits) >= 2:
            return hits[0][0], hits[0][1], hits[1][0], hits[1][1]
        return None, None, None, None

    # -- GP hot path ----------------------------------------------------------------------------------
    def update(self, obs_wp, log_obs_vals):
        """point_source.py:335-337: full refit (11 L-BFGS-B runs on the device objective)."""
        self.gp.fit(obs_wp, log_obs_vals)

    def simulate_measurements(self, waypoints, noise_level: float = 0.001) -> np.ndarray:
        """point_source.py:339-344: one np.random.normal call per invocation."""
        inten = self.intensity(waypoints)
        noise = np.random.normal(0, noise_level * inten)
        return np.maximum(inten + noise, 1e-6)

    def predict_spatial_field(self, waypoints, measurements: np.ndarray, transient: bool = False):
        """point_source.py:346-353: refit on all observations, posterior over the whole lattice.

        Returns (Z_pred (ny, nx), std (ny * nx,)) -- std stays flat and in log10 units like the
        reference's.  `transient` marks the per-step calls whose result the reference never reads
        (rrt.py:749): with a lazy GP (`gp.lazy = True`) those only consume the restart
        draws from the RNG and return (None, None); the default is the eager, faithful evaluation."""
        y_log = np.log10(np.maximum(measurements, 1e-6))
        gp = self.gp
        if transient and getattr(gp, "lazy", False):
            gp.fit(waypoints, y_log)
            return None, None
        gp.fit(waypoints, y_log)
        w = self.workspace_size
        nx, ny = self.x.shape[0], self.y.shape[0]
        if self.grid_group is not None:
            from mripp_b200.dist import sharded_grid_posterior

            z, std = sharded_grid_posterior(gp, w, nx, ny, None if self.grid_group is True else self.grid_group)
        elif self.device_outputs and hasattr(gp, "predict_grid_device"):
            _, std_d, z_d = gp.predict_grid_device(w[0], w[1], nx, w[2], w[3], ny, want_zpred=True)
            return DeviceField(z_d.view(ny, nx)), DeviceField(std_d)
        else:
            z, std = gp.predict_grid_host(w[0], w[1], nx, w[2], w[3], ny)  # one path: the fused lattice posterior
        return z.reshape(self.X.shape), std

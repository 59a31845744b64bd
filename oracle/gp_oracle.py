"""CPU oracle for the GP half of the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (multi-robot-informative-path-planning_b200/) never does.

It is a plain numpy/scipy restatement of what the reference executes for
  PointSourceField.update / predict_spatial_field   (src/point_source/point_source.py:335-353)
through its un-vendored third-party dependency scikit-learn (requirements.txt:3
`scikit_learn>=1.4.1.post1`, nothing pinned; 1.9.0 + scipy 1.18.1 + numpy 2.3.5 in this image):
  GaussianProcessRegressor.fit / predict / log_marginal_likelihood
      sklearn/gaussian_process/_gpr.py:233-368, 370-500, 541-656
  ConstantKernel * RBF                 sklearn/gaussian_process/kernels.py:936-971, 1244-1296, 1530-1587

Parity pinning: the reference ships no golden vectors for this path, so the oracle
is pinned against outputs of the reference itself, run in the build container by
oracle/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.optimize import minimize

ALPHA = 1e-10  # GaussianProcessRegressor default jitter (_gpr.py: alpha=1e-10; point_source.py:71)
C_BOUNDS = (1e-3, 50.0)  # point_source.py:86
L_BOUNDS = (1e-2, 50.0)  # point_source.py:86
N_RESTARTS = 10  # point_source.py:71


def log_bounds():
    """theta = log([c, l]); bounds in log space (kernels.py:291-360)."""
    return np.log(np.array([C_BOUNDS, L_BOUNDS], dtype=float))


def sqdist_scaled(X, Y, length_scale):
    """cdist(X / l, Y / l, 'sqeuclidean') (kernels.py:1561,1569): sum over features of (u - v)^2."""
    A = np.asarray(X, dtype=float) / length_scale
    B = np.asarray(Y, dtype=float) / length_scale
    d = A[:, None, :] - B[None, :, :]
    return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]


def kernel(X, Y=None, c=1.0, length_scale=1.0):
    """(ConstantKernel(c) * RBF(l))(X, Y) -- kernels.py:964 (K1 * K2), 1290, 1561-1570."""
    if Y is None:
        d2 = sqdist_scaled(X, X, length_scale)
        K = np.exp(-0.5 * d2)
        np.fill_diagonal(K, 1.0)  # squareform + fill_diagonal(K, 1): kernels.py:1563-1565
    else:
        K = np.exp(-0.5 * sqdist_scaled(X, Y, length_scale))
    return c * K


def kernel_with_gradient(X, c, length_scale):
    """K and dK/dtheta for theta = log([c, l]) (kernels.py:964-969, 1283-1292, 1577)."""
    d2 = sqdist_scaled(X, X, length_scale)
    R = np.exp(-0.5 * d2)
    np.fill_diagonal(R, 1.0)
    K = c * R
    dK = np.empty(K.shape + (2,))
    dK[..., 0] = K  # d/dlog c
    dK[..., 1] = c * (R * d2)  # d/dlog l: K2 * squareform(dists)
    return K, dK


def normalise_y(y):
    """_gpr.py:275-280 with _handle_zeros_in_scale: y is 1-D, so np.std is a scalar and sklearn's scalar branch
    applies (sklearn/preprocessing/_data.py: `if scale == 0.0: scale = 1.0`), not the 10 eps test of the array branch."""
    y = np.asarray(y, dtype=float)
    mean = np.mean(y, axis=0)
    std = np.std(y, axis=0)
    if std == 0.0:
        std = 1.0
    return (y - mean) / std, float(mean), float(std)


def lml(theta, X, yn, eval_gradient=True, alpha=ALPHA):
    """log_marginal_likelihood(theta, eval_gradient) -- _gpr.py:583-656."""
    c, l = np.exp(theta[0]), np.exp(theta[1])
    if eval_gradient:
        K, dK = kernel_with_gradient(X, c, l)
    else:
        K = kernel(X, None, c, l)
    K[np.diag_indices_from(K)] += alpha
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return (-np.inf, np.zeros(2)) if eval_gradient else -np.inf
    a = cho_solve((L, True), yn, check_finite=False)
    val = -0.5 * float(yn @ a) - np.log(np.diag(L)).sum() - K.shape[0] / 2 * np.log(2 * np.pi)
    if not eval_gradient:
        return val
    inner = np.outer(a, a) - cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    grad = 0.5 * np.einsum("ij,jik->k", inner, dK)
    return val, grad


def fit(X, y, c0=1.0, l0=1.0, n_restarts=N_RESTARTS, rng=None, optimize=True, alpha=ALPHA):
    """GaussianProcessRegressor.fit -- _gpr.py:249-367.  `rng` defaults to the global np.random
    stream exactly like check_random_state(None) (_gpr.py:251)."""
    X = np.asarray(X, dtype=float).reshape(-1, 2)
    yn, y_mean, y_std = normalise_y(y)
    theta0 = np.log(np.array([c0, l0], dtype=float))
    bounds = log_bounds()
    n_evals = 0
    if optimize:
        rs = np.random.mtrand._rand if rng is None else rng

        def obj(theta):
            nonlocal n_evals
            n_evals += 1
            v, g = lml(theta, X, yn, True, alpha)
            return -v, -g

        def run(start):
            res = minimize(obj, start, method="L-BFGS-B", jac=True, bounds=bounds)
            return res.x, res.fun

        optima = [run(theta0)]
        for _ in range(n_restarts):
            optima.append(run(rs.uniform(bounds[:, 0], bounds[:, 1])))
        vals = [o[1] for o in optima]
        theta = optima[int(np.argmin(vals))][0]
        lml_value = -float(np.min(vals))
    else:
        theta = theta0
        lml_value = float(lml(theta, X, yn, False, alpha))
    c, l = float(np.exp(theta[0])), float(np.exp(theta[1]))
    K = kernel(X, None, c, l)
    K[np.diag_indices_from(K)] += alpha
    L = cholesky(K, lower=True, check_finite=False)
    a = cho_solve((L, True), yn, check_finite=False)
    return dict(X=X, theta=np.asarray(theta, float), c=c, l=l, L=L, alpha=a, y_mean=y_mean, y_std=y_std,
                lml=lml_value, n_evals=n_evals, yn=yn)


def predict(state, Xq, chunk=20000):
    """predict(Xq, return_std=True) -- _gpr.py:446-500, chunked over query rows (row independent)."""
    Xq = np.asarray(Xq, dtype=float).reshape(-1, 2)
    mean = np.empty(len(Xq))
    std = np.empty(len(Xq))
    n_neg = 0
    for s in range(0, len(Xq), chunk):
        q = Xq[s:s + chunk]
        Kt = kernel(q, state["X"], state["c"], state["l"])
        mu = Kt @ state["alpha"]
        mean[s:s + chunk] = state["y_std"] * mu + state["y_mean"]
        V = solve_triangular(state["L"], Kt.T, lower=True, check_finite=False)
        var = np.full(len(q), state["c"]) - np.einsum("ij,ji->i", V.T, V)
        neg = var < 0
        n_neg += int(neg.sum())
        var[neg] = 0.0
        std[s:s + chunk] = np.sqrt(var * state["y_std"] ** 2)
    return mean, std, n_neg


def workspace_grid(workspace_size):
    """point_source.py:65-67,350: 10 lattice points per unit length, 'xy' meshgrid, row-major ravel."""
    w = workspace_size
    x = np.linspace(w[0], w[1], int((w[1] - w[0]) * 10))
    y = np.linspace(w[2], w[3], int((w[3] - w[2]) * 10))
    Xg, Yg = np.meshgrid(x, y)
    return x, y, np.column_stack((Xg.ravel(), Yg.ravel())), Xg.shape


def predict_spatial_field(workspace_size, waypoints, measurements, c0=1.0, l0=1.0, rng=None, optimize=True,
                          theta=None):
    """PointSourceField.predict_spatial_field -- point_source.py:346-353."""
    ylog = np.log10(np.maximum(np.asarray(measurements, dtype=float), 1e-6))
    if theta is not None:
        st = fit(waypoints, ylog, float(np.exp(theta[0])), float(np.exp(theta[1])), optimize=False)
    else:
        st = fit(waypoints, ylog, c0, l0, rng=rng, optimize=optimize)
    _, _, r, shape = workspace_grid(workspace_size)
    mean, std, _ = predict(st, r)
    return 10 ** mean.reshape(shape), std, st


def sklearn_reference_gp(c0=1.0, l0=1.0):
    """The estimator object exactly as the reference constructs it (point_source.py:71,86).  Used by
    bench.py --impl reference / cpu_baseline to time the reference's real third-party code path."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C

    k = C(c0, C_BOUNDS) * RBF(l0, L_BOUNDS)
    return GaussianProcessRegressor(kernel=k, n_restarts_optimizer=N_RESTARTS, normalize_y=True)


def sklearn_fixed_theta_gp(c=1.0, l=1.0):
    """Same estimator with the hyper-parameters held at (c, l) (`optimizer=None`): the posterior-only
    leg of bench.py, so the CPU and GPU arms evaluate the posterior at the same theta."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C

    return GaussianProcessRegressor(kernel=C(c, C_BOUNDS) * RBF(l, L_BOUNDS), optimizer=None, normalize_y=True)

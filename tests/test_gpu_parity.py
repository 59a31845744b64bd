"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every device result goes through the C ABI
(libipp_b200.so) and is compared with (i) golden vectors produced by the unmodified reference
(tests/golden/*.npz, oracle/make_golden.py) and (ii) the CPU oracle on seeded inputs.

Tolerances are the tiers of SURVEY.md §8c:
  T1 fixed theta : |d mean| <= 1e-6 max(1,|mean|); |d std| <= 1e-7; |d var| <= 1e-10 c y_std^2 (+ fp noise);
                   LML rel 1e-6 (the quadratic form y^T K^-1 y carries cond(K) ~ 1e11), gradient 1e-6 of its scale
  T2 re-optimised: |lml* - lml*_ref| <= 1e-4 + 1e-7 |lml|; posterior within 1e-2
  T3 decisions   : bit-identical trees / argmax / waypoint sequences; gain values rel 1e-12
"""
import ctypes as C
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gp_oracle as O  # noqa: E402
from oracle import rrt_oracle as T  # noqa: E402

GP_CASES = ["n40", "n131", "n300"]


@pytest.fixture(scope="module")
def G():
    from mripp_b200 import gp

    return gp


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _gp_at(G, X, ylog, theta):
    c, l = float(np.exp(theta[0])), float(np.exp(theta[1]))
    return G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(c, l), optimizer=None, normalize_y=True).fit(X, ylog)


def _make(n, W, seed=0, dup=0.2):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0, W, (n, 2))
    k = int(dup * n)
    if k:
        X[-k:] = X[:k]
    src = np.array([[0.3 * W, 0.6 * W, 5e5], [0.7 * W, 0.2 * W, 8e5], [0.5 * W, 0.5 * W, 2e5]])
    d2 = ((X[:, None, :] - src[None, :, :2]) ** 2).sum(-1)
    inten = (src[None, :, 2] / (4 * np.pi * np.maximum(d2, 0.25))).sum(1)
    meas = np.maximum(inten * (1 + 1e-3 * rng.randn(n)), 1e-6)
    return X, np.log10(meas)


def _lml_tol(v_ref, theta, alpha_sq):
    """The data-fit term y^T K^-1 y moves by alpha^T dK alpha under a perturbation dK of the Gram
    matrix; entry-wise rounding of K (|dK| <= eps * c) therefore already moves it by up to
    eps * c * |alpha|^2 (|alpha| ~ 1e7 with duplicate waypoints).  LAPACK and the device factorisation
    round differently, so that is the floor of any LML comparison -- in the reference itself too."""
    return 1e-9 * abs(v_ref) + 16 * np.finfo(float).eps * float(np.exp(theta[0])) * alpha_sq


def _mean_tol(st, q, mean_ref):
    """Per-point tolerance of the posterior mean.  mu = k*^T K^-1 y moves by w^T dK alpha
    (w = K^-1 k*) when K is perturbed; rounding K entry-wise (|dK| <= eps c) already allows
    eps c |w|_1 |alpha|_1.  LAPACK's own mean sits at up to 0.6 of that bound from an 80-bit
    evaluation (measured at n = 500, l = 2: 1.2e-4 absolute), so two fp64 implementations may differ
    by about twice the bound; the 1e-6 floor is the T1 tier where the problem is well conditioned."""
    from scipy.linalg import cho_solve

    Kt = O.kernel(q, st["X"], st["c"], st["l"])
    w = cho_solve((st["L"], True), Kt.T, check_finite=False)
    bound = np.finfo(float).eps * st["c"] * np.abs(w).sum(0) * np.abs(st["alpha"]).sum() * st["y_std"]
    return 1e-6 * np.maximum(1.0, np.abs(mean_ref)) + 2.0 * bound


def _var_tol(st):
    """sigma^2 = c - |L^-1 k*|^2: the solve carries a relative error of eps * cond(L), so the T1 tier
    (1e-10 c y_std^2, measured at cond(L) ~ 3e5) widens with the conditioning of the stress cases.  The constant is
    8 (two one-sided bounds of 4 eps cond: LAPACK's solve and the device's explicit inverse each sit inside their
    own; measured worst case over the suite 4.2 with the rank-128 grouping of the dataflow Cholesky, 3.6 before)."""
    cond = float(np.linalg.cond(st["L"])) if st["L"].shape[0] > 1 else 1.0
    return max(1e-10, 8 * np.finfo(float).eps * cond) * st["c"] * st["y_std"] ** 2 + 1e-13


GRAD_RTOL = 1e-4  # of max(|grad|, 1e-9 |alpha|^2): LAPACK's own gradient sits 1.4e-5 of that scale from an 80-bit evaluation


# ---------------------------------------------------------------------------------------------------
# G3: Gram build
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(1, 1), (63, 64), (65, 130), (300, 7)])
def test_gram_matches_oracle(G, n, m):
    rng = np.random.RandomState(n + m)
    A, B = rng.uniform(0, 30, (n, 2)), rng.uniform(0, 30, (m, 2))
    k = G.ConstantRBFKernel(2.3, 1.7)
    Ks, Kc = k(A), k(A, B)
    assert Ks.shape == (n, n) and Kc.shape == (n, m)
    # CUDA exp vs glibc exp: <= 1 ulp each
    assert np.abs(Ks - O.kernel(A, None, 2.3, 1.7)).max() <= 4e-16 * 2.3
    assert np.abs(Kc - O.kernel(A, B, 2.3, 1.7)).max() <= 4e-16 * 2.3
    assert np.array_equal(np.diag(Ks), np.full(n, 2.3))  # exact c on the diagonal (kernels.py:1565)


def test_gram_against_reference_golden(G, gold):
    g = gold("select_golden.npz")
    Kpio = g["c_Kpio"]
    pts, nchild = g["c_pts"], g["c_nchild"]
    leaves = pts[nchild == 0]
    K = G.ConstantRBFKernel(1.0, 1.0)(leaves, g["c_obs0"])
    assert K.shape == Kpio.shape
    assert np.abs(K - Kpio).max() <= 4e-16


# ---------------------------------------------------------------------------------------------------
# G2: LML + gradient, G4: final factor
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", GP_CASES)
def test_lml_and_gradient_vs_reference(G, gold, tag):
    g = gold("gp_golden.npz")
    X, ylog = g[f"{tag}_X"], np.log10(np.maximum(g[f"{tag}_meas"], 1e-6))
    for th, v_ref, g_ref in zip(g[f"{tag}_thetas"], g[f"{tag}_lml"], g[f"{tag}_grad"]):
        gp = _gp_at(G, X, ylog, th)
        v, gr = gp.log_marginal_likelihood(th, eval_gradient=True)
        st = O.fit(X, ylog, float(np.exp(th[0])), float(np.exp(th[1])), optimize=False)
        aa = float(st["alpha"] @ st["alpha"])
        assert abs(v - v_ref) <= _lml_tol(v_ref, th, aa)
        scale = max(1.0, float(np.abs(g_ref).max()), 1e-9 * aa)
        assert np.abs(gr - g_ref).max() <= GRAD_RTOL * scale
        assert abs(gp.log_marginal_likelihood_value_ - v_ref) <= _lml_tol(v_ref, th, aa)
        v_only = gp.log_marginal_likelihood(th, eval_gradient=False)
        assert v_only == pytest.approx(v, rel=1e-12)


@pytest.mark.parametrize("tag", GP_CASES)
def test_fixed_theta_factor_and_posterior_vs_reference(G, gold, tag):
    g = gold("gp_golden.npz")
    X, ylog = g[f"{tag}_X"], np.log10(np.maximum(g[f"{tag}_meas"], 1e-6))
    th = g[f"{tag}_theta_fit"]
    gp = _gp_at(G, X, ylog, th)
    c = float(np.exp(th[0]))
    ys = float(g[f"{tag}_ystd"])
    assert gp._y_train_mean == float(g[f"{tag}_ymean"]) and gp._y_train_std == ys
    # pivots of duplicated waypoints are sqrt(2 jitter) ~ 1.4e-5, the residue of a cancellation at
    # the level of c over n terms: compare the squared pivots at n * eps * c, not the pivots relatively
    n = X.shape[0]
    assert np.abs(np.diag(gp.L_) ** 2 - g[f"{tag}_Ldiag"] ** 2).max() <= 64 * n * np.finfo(float).eps * c
    a_ref = g[f"{tag}_alpha"]
    assert np.abs(gp.alpha_ - a_ref).max() <= 1e-4 * np.abs(a_ref).max()  # cond ~ 1e11: not element-exact
    st = O.fit(X, ylog, c, float(np.exp(th[1])), optimize=False)
    mean, std = gp.predict(g[f"{tag}_Q"], return_std=True)
    mref, sref = g[f"{tag}_qmean"], g[f"{tag}_qstd"]
    assert (np.abs(mean - mref) <= _mean_tol(st, g[f"{tag}_Q"], mref)).all()
    assert np.abs(std - sref).max() <= 1e-7
    assert np.abs(std ** 2 - sref ** 2).max() <= 1e-10 * c * ys * ys + 1e-13
    # workspace lattice (fused grid kernel, coordinates generated on the device)
    W = int(g[f"{tag}_W"])
    nx = ny = W * 10
    stride = int(g[f"{tag}_gstride"])
    m_d, s_d, z_d = gp.predict_grid_device(0, W, nx, 0, W, ny)
    gm, gs = m_d.cpu().numpy()[::stride], s_d.cpu().numpy()[::stride]
    _, _, r, _ = O.workspace_grid((0, W, 0, W))
    assert (np.abs(gm - g[f"{tag}_gmean"]) <= _mean_tol(st, r[::stride], g[f"{tag}_gmean"])).all()
    assert np.abs(gs - g[f"{tag}_gstd"]).max() <= 1e-7
    assert np.allclose(z_d.cpu().numpy()[::stride], 10.0 ** g[f"{tag}_gmean"], rtol=1e-5)


def test_predict_spatial_field_reoptimised_vs_reference(gold):
    """T2 + RNG contract: the full 11-restart fit consumes exactly the reference's draws."""
    from mripp_b200.src.point_source.point_source import PointSourceField

    g = gold("gp_golden.npz")
    W = int(g["psf_W"])
    sc = PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=3)
    np.random.seed(77)
    Z, std = sc.predict_spatial_field(list(g["psf_X"]), g["psf_meas"])
    assert np.array_equal(np.random.uniform(size=3), g["psf_rng_after"])
    assert Z.shape == g["psf_Z"].shape and std.shape == g["psf_std"].shape
    assert np.abs(np.log10(Z) - np.log10(g["psf_Z"])).max() <= 1e-2
    assert np.abs(std - g["psf_std"]).max() <= 1e-2
    # LML at the optimum agrees with the oracle's optimum on the same data
    np.random.seed(77)
    _, _, st = O.predict_spatial_field((0, W, 0, W), g["psf_X"], g["psf_meas"])
    lml_ref = st["lml"]
    assert abs(sc.gp.log_marginal_likelihood_value_ - lml_ref) <= 1e-4 + 1e-6 * abs(lml_ref)
    assert sc.gp.n_lml_evals_ > 50


# ---------------------------------------------------------------------------------------------------
# oracle parity on seeded inputs across padding edge cases, duplicates, explicit points
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 128, 129, 500, 1000])
def test_posterior_vs_oracle_sizes(G, n):
    W = 20
    X, y = _make(n, W, seed=n)
    c, l = 1.3, 2.0
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(c, l), optimizer=None, normalize_y=True).fit(X, y)
    st = O.fit(X, y, c, l, optimize=False)
    q = np.random.RandomState(1).uniform(0, W, (1500, 2))
    k = min(200, n)
    q[:k] = X[:k] + 1e-3  # right next to data: the variance cancels to ~1e-10
    mg, sg = gp.predict(q, return_std=True)
    mo, so, _ = O.predict(st, q)
    assert (np.abs(mg - mo) <= _mean_tol(st, q, mo)).all()
    assert np.abs(sg - so).max() <= 1e-7
    assert np.abs(sg ** 2 - so ** 2).max() <= _var_tol(st)
    th = np.log([c, l])
    v, gr = gp.log_marginal_likelihood(th, eval_gradient=True)
    vo, go = O.lml(th, X, st["yn"])
    assert abs(v - vo) <= _lml_tol(vo, th, float(st["alpha"] @ st["alpha"]))
    scale = max(1.0, float(np.abs(go).max()), 1e-9 * float(st["alpha"] @ st["alpha"]))
    assert np.abs(gr - go).max() <= GRAD_RTOL * scale


def test_grid_paths_agree_and_slabs_bit_exact(G, torch):
    """Size-independent properties of the grid entry point:
      * generic path (every k* element by exp(), lattice coordinates generated like np.linspace) is
        bit-identical to the explicit-point entry on the same coordinates;
      * the lattice fast path (separable coordinate tables) agrees with it to rounding;
      * any slab split reproduces the whole-grid result exactly (multi-GPU sharding is result-invariant)."""
    W = 13
    X, y = _make(333, W, seed=5)
    c, l = 0.9, 1.1
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(c, l), optimizer=None, normalize_y=True).fit(X, y)
    st = O.fit(X, y, c, l, optimize=False)
    nx, ny = 130, 97
    xs, ys = np.linspace(0, W, nx), np.linspace(0.5, 9.25, ny)
    XX, YY = np.meshgrid(xs, ys)
    r = np.column_stack((XX.ravel(), YY.ravel()))
    mp, sp, zp = gp.predict_points_device(torch.from_numpy(r).cuda(), want_zpred=True, tabled=False)  # the generic kernel
    mg, sg, zg = gp.predict_grid_device(0, W, nx, 0.5, 9.25, ny, lattice=False)
    assert torch.equal(mg, mp) and torch.equal(sg, sp) and torch.equal(zg, zp)
    m, s, z = gp.predict_grid_device(0, W, nx, 0.5, 9.25, ny)  # lattice fast path (nx >= 128), block-sparse k*
    tol = _mean_tol(st, r, mp.cpu().numpy())
    assert (np.abs((m - mp).cpu().numpy()) <= tol).all()
    assert float((s - sp).abs().max()) <= 1e-9
    assert float((s * s - sp * sp).abs().max()) <= _var_tol(st)
    for lattice in (True, False):
        # (dense k*: with slab skipping the set of < 1e-30 terms a query drops depends on its tile)
        m, s, z = gp.predict_grid_device(0, W, nx, 0.5, 9.25, ny, lattice=lattice, skip_below=0)
        cuts = [0, 1, 127, 128, 129, 5000, 9999, nx * ny]
        parts = [gp.predict_grid_device(0, W, nx, 0.5, 9.25, ny, g_begin=a, g_end=b, lattice=lattice, skip_below=0)
                 for a, b in zip(cuts[:-1], cuts[1:])]
        assert torch.equal(torch.cat([q[0] for q in parts]), m)
        assert torch.equal(torch.cat([q[1] for q in parts]), s)
        assert torch.equal(torch.cat([q[2] for q in parts]), z)
    # empty slab / empty query
    e = gp.predict_grid_device(0, W, nx, 0.5, 9.25, ny, g_begin=77, g_end=77)
    assert e[0].numel() == 0
    mean0 = gp.predict(np.zeros((0, 2)))
    assert mean0.shape == (0,)


@pytest.mark.parametrize("l", [0.4, 0.6])
def test_block_sparse_posterior_equals_dense(G, torch, l):
    """Slabs of k* whose entries are all below 1e-30 c are skipped (observations fitted in Morton order).  The
    result must be the dense one to rounding, including tiles that see no observation at all (prior) and tiles
    that wrap across two lattice rows."""
    W, n = 60, 1500
    X, y = _make(n, W, seed=3)
    X[:700] = X[:700] * 0.3 + 5.0  # a dense cluster and a sparse remainder: some tiles keep few slabs, some none
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.4, l), optimizer=None, normalize_y=True).fit(X, y)
    assert gp._bbox_d is not None and not np.array_equal(gp._perm, np.arange(n))  # Morton order is in use
    nx, ny = 600, 150
    dense = gp.predict_grid_device(0, W, nx, 0, 15.0, ny, skip_below=0)
    sparse = gp.predict_grid_device(0, W, nx, 0, 15.0, ny)
    assert float((dense[0] - sparse[0]).abs().max()) <= 1e-13 * max(1.0, float(dense[0].abs().max()))
    assert float((dense[1] - sparse[1]).abs().max()) <= 1e-13
    fin = torch.isfinite(dense[2])
    assert torch.equal(fin, torch.isfinite(sparse[2]))
    assert float((dense[2][fin] / sparse[2][fin] - 1).abs().max()) <= 1e-12
    # the same posterior from the factorisation in the GIVEN order of the observations (no Morton sort): equal to
    # the rounding noise of a reordered elimination
    ref = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.4, l), optimizer=None, normalize_y=True)
    ref.skip_below = 0
    ref.fit(X, y)
    assert ref._bbox_d is None and np.array_equal(ref._perm, np.arange(n))
    r0 = ref.predict_grid_device(0, W, nx, 0, 15.0, ny)
    st = O.fit(X, y, 1.4, l, optimize=False)
    xs, ys = np.linspace(0, W, nx), np.linspace(0, 15.0, ny)
    pts = np.column_stack((np.tile(xs, 3), np.repeat(ys[[0, 70, 149]], nx)))
    idx = np.concatenate([np.arange(nx) + nx * r for r in (0, 70, 149)])
    tol = _mean_tol(st, pts, r0[0].cpu().numpy()[idx])
    assert (np.abs((sparse[0] - r0[0]).cpu().numpy()[idx]) <= tol).all()
    assert float((sparse[1] - r0[1]).abs().max()) <= 1e-8
    assert np.abs(gp.alpha_ - ref.alpha_).max() <= 1e-4 * np.abs(ref.alpha_).max()
    far = gp.predict_grid_device(1000.0, 1000.0 + W, nx, 0, 15.0, 8)  # nothing within reach: the prior, exactly
    assert float((far[0] - gp._y_train_mean).abs().max()) == 0.0
    assert float((far[1] - np.sqrt(1.4) * gp._y_train_std).abs().max()) <= 1e-15


@pytest.mark.parametrize("l", [2.0, 4.0])
def test_clustered_elimination_order_keeps_the_variance(G, l):
    """The densest regimes in which the block-sparse path switches on (4 and 16 observations per l x l cell; at
    l = 4 the kernel's reach is 0.67 of the workspace, the regime opened by raising the switch from 0.35 to 0.8 of
    the extent): with the clusters taken coarse to fine the variance is as accurate as in the given order
    (eliminating along the Z-order curve loses 3-4 digits of sigma next to the data here: 1.8e-5 at l = 2)."""
    from bench import synth_observations

    n, W = 5000, 70
    X, meas, _ = synth_observations(n, W, 1)
    y = np.log10(meas)
    q = np.random.RandomState(2).uniform(0, W, (600, 2))
    q[:200] = X[:200] + 1e-3
    st = O.fit(X, y, 1.0, l, optimize=False)
    mo, so, _ = O.predict(st, q)
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, l), optimizer=None, normalize_y=True).fit(X, y)
    assert gp._bbox_d is not None  # sparse path on
    m, s = gp.predict(q, return_std=True)
    given = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, l), optimizer=None, normalize_y=True)
    given.skip_below = 0  # dense, observations in the given order
    mg, sg = given.fit(X, y).predict(q, return_std=True)
    e_clu, e_giv = np.abs(s ** 2 - so ** 2).max(), np.abs(sg ** 2 - so ** 2).max()
    print(f"l={l}: |d sigma^2| clustered {e_clu:.2e} given {e_giv:.2e} (c y_std^2 = {st['c'] * st['y_std'] ** 2:.3g}); "
          f"|d sigma| clustered {np.abs(s - so).max():.2e} given {np.abs(sg - so).max():.2e}")
    # measured at l = 2: 3.8e-10 c y_std^2 in the clustered order, 3.3e-10 in the given order (n = 5000 terms per
    # column norm); the clustered order may cost at most a small factor over the given one
    # (measured on the B200: l = 2: 2.5e-10 clustered / 2.0e-10 given; l = 4: 6.8e-11 / 9.3e-12)
    vtol = max(2e-9 * st["c"] * st["y_std"] ** 2, 4.0 * e_giv)
    assert e_clu <= vtol
    # sigma itself: next to the data sigma^2 is a ~1e-10 residue, so an error e in sigma^2 is up to sqrt(e) in
    # sigma (l = 4: 6.4e-6 from 6.8e-11); 1e-6 where the problem is not that singular
    assert np.abs(s - so).max() <= max(1e-6, np.sqrt(vtol))
    assert (np.abs(m - mo) <= _mean_tol(st, q, mo)).all()
    # slabs are exactly the clusters: 16 consecutive rows of the fitted order are spatially compact
    Xs = X[gp._perm]
    w = np.ptp(Xs[: n - n % 16].reshape(-1, 16, 2), axis=1).max(axis=1)
    assert np.median(w) < 0.15 * W and sorted(gp._perm.tolist()) == list(range(n))  # (a few clusters straddle a jump of the curve)


def test_predict_recognises_the_reference_lattice(G, torch):
    """gp.predict(r) with r = the raveled meshgrid the reference builds (point_source.py:350) takes the fused
    lattice kernel; anything else (a permuted or perturbed copy) takes the explicit-point kernel; same numbers."""
    W = 30
    X, y = _make(400, W, seed=8)
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.1, 0.9), optimizer=None, normalize_y=True).fit(X, y)
    xs, ys = np.linspace(0, W, W * 10), np.linspace(0, 12, 120)
    XX, YY = np.meshgrid(xs, ys)
    r = np.column_stack((XX.ravel(), YY.ravel()))
    assert G.as_lattice(r) == (0.0, float(W), W * 10, 0.0, 12.0, 120)
    r2 = r.copy()
    r2[1234, 0] += 1e-9
    assert G.as_lattice(r2) is None and G.as_lattice(r[::-1].copy()) is None and G.as_lattice(r[:100]) is None
    m1, s1 = gp.predict(r, return_std=True)
    md, sd, _ = gp.predict_grid_device(0, W, W * 10, 0, 12, 120)
    assert np.array_equal(m1, md.cpu().numpy()) and np.array_equal(s1, sd.cpu().numpy())
    m2, s2 = gp.predict(r2, return_std=True)
    st = O.fit(X, y, 1.1, 0.9, optimize=False)
    keep = np.arange(len(r)) != 1234
    assert (np.abs(m1 - m2)[keep][::7] <= _mean_tol(st, r[keep][::7], m2[keep][::7])).all()
    assert np.abs(s1 - s2)[keep].max() <= 1e-9


def test_factor_properties_full_size(G, torch):
    """C4-sized factor (n = 5000, 20 % duplicate waypoints): W L = I, alpha solves K alpha = y, the
    posterior interpolates the data and the variance stays in [0, c]."""
    n, W = 5000, 100
    X, y = _make(n, W, seed=11)
    c, l = 1.0, 2.0
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(c, l), optimizer=None, normalize_y=True).fit(X, y)
    ws = gp._ws
    NP, ld = ws.layout[0], ws.layout[2]
    L = ws.K[: NP * ld].view(NP, ld)[:n, :n]  # device state: observations in Morton order (gp._perm)
    Wm = ws.W[: NP * ld].view(NP, ld)[:n, :n]
    assert sorted(gp._perm.tolist()) == list(range(n))
    assert float(torch.triu(Wm, 1).abs().max()) == 0.0
    R = Wm @ L - torch.eye(n, dtype=torch.float64, device="cuda")
    assert float(R.abs().max()) <= 1e-7  # cond(L) ~ 3e5
    Ks = torch.from_numpy(O.kernel(X[gp._perm], None, c, l)).cuda()
    Ks += 1e-10 * torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((L @ L.T - Ks).abs().max()) <= 1e-12
    K = torch.from_numpy(O.kernel(X, None, c, l)).cuda()
    K += 1e-10 * torch.eye(n, dtype=torch.float64, device="cuda")
    alpha = torch.from_numpy(gp.alpha_).cuda()  # original order
    yn = torch.from_numpy(gp.y_train_).cuda()
    resid = K @ alpha - yn
    assert float(resid.abs().max()) <= 1e-5 * float(alpha.abs().max()) * 1e-3
    mean, std = gp.predict(X[:1000], return_std=True)
    assert np.abs(mean - y[:1000]).max() <= 2e-2  # duplicates carry independent noise draws
    assert std.min() >= 0.0 and std.max() <= np.sqrt(c) * gp._y_train_std * (1 + 1e-12)
    far = gp.predict(np.array([[1e4, 1e4]]), return_std=True)
    assert far[0][0] == pytest.approx(gp._y_train_mean) and far[1][0] == pytest.approx(np.sqrt(c) * gp._y_train_std)
    # the whole C4 lattice (10^6 points): block-sparse == dense, variance within the prior, slabs reassemble
    sp = gp.predict_grid_device(0, W, 1000, 0, W, 1000)
    de = gp.predict_grid_device(0, W, 1000, 0, W, 1000, skip_below=0)
    assert float((sp[0] - de[0]).abs().max()) <= 1e-12 * max(1.0, float(de[0].abs().max()))
    assert float((sp[1] - de[1]).abs().max()) <= 1e-12
    assert float(sp[1].min()) >= 0.0 and float(sp[1].max()) <= np.sqrt(c) * gp._y_train_std * (1 + 1e-12)
    assert bool(torch.isfinite(sp[0]).all()) and bool(torch.isfinite(sp[2]).all())
    executed, dense = gp.lattice_work(0, W, 1000, 0, W, 1000)
    assert 0.1 * dense < executed < 0.5 * dense  # theta = (1, 2) on 100 x 100: about a quarter of the slabs
    half = gp.predict_grid_device(0, W, 1000, 0, W, 1000, g_begin=0, g_end=500 * 1000, skip_below=0)
    assert torch.equal(half[1], de[1][: 500 * 1000])
    # one C4-scale row slab of the lattice against the oracle
    m_d, s_d, _ = gp.predict_grid_device(0, W, 1000, 0, W, 1000, g_begin=500 * 1000, g_end=500 * 1000 + 2048)
    st = O.fit(X, y, c, l, optimize=False)
    xs = np.linspace(0, W, 1000)
    r = np.column_stack((np.concatenate([xs, xs, xs[:48]]), np.repeat(xs[500:503], [1000, 1000, 48])))
    mo, so, _ = O.predict(st, r)
    assert (np.abs(m_d.cpu().numpy() - mo) <= _mean_tol(st, r, mo)).all()
    assert np.abs(s_d.cpu().numpy() - so).max() <= 1e-7


def test_not_positive_definite_is_reported(G):
    """Exact duplicates with alpha = 0: the reference raises LinAlgError from fit (_gpr.py:351-361)
    and returns -inf inside the objective (_gpr.py:590-593)."""
    X = np.array([[1.0, 1.0], [2.0, 2.0], [1.0, 1.0], [3.0, 1.0]])
    y = np.array([0.1, 0.2, 0.1, 0.4])
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), optimizer=None, alpha=0.0)
    with pytest.raises(np.linalg.LinAlgError):
        gp.fit(X, y)
    ok = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), optimizer=None, alpha=1e-10).fit(X, y)
    ok.alpha = 0.0
    v, gr = ok.log_marginal_likelihood(np.zeros(2), eval_gradient=True)
    assert v == -np.inf and np.array_equal(gr, np.zeros(2))


def test_concurrent_restart_streams_equal_serial(G):
    """The optimiser restarts run on several streams / host threads by default; the optimum, its LML, the number of
    objective evaluations and the RNG position must be exactly those of the serial loop of _gpr.py:312-337."""
    X, y = _make(300, 15, seed=2)
    out = []
    for streams in (1, 3, None):
        gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), n_restarts_optimizer=5, normalize_y=True)
        gp.restart_streams = streams
        np.random.seed(5)
        gp.fit(X, y)
        out.append((gp.kernel_.theta.copy(), gp.log_marginal_likelihood_value_, gp.n_lml_evals_, np.random.uniform(),
                    gp.predict(X[:50] + 0.01, return_std=True)))
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1] == out[0][1] and o[2] == out[0][2] and o[3] == out[0][3]
        assert np.array_equal(o[4][0], out[0][4][0]) and np.array_equal(o[4][1], out[0][4][1])


def test_unfitted_prior_and_negative_variance_warning(G):
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(2.0, 1.0))
    m, s = gp.predict(np.zeros((3, 2)), return_std=True)
    assert np.array_equal(m, np.zeros(3)) and np.allclose(s, np.sqrt(2.0))


# ---------------------------------------------------------------------------------------------------
# I1-I5 + T: branch gains; I6/I7: stage-1 selector; I8, argmax; count_close
# ---------------------------------------------------------------------------------------------------
def _strategy(cls, W, seed, **kw):
    from mripp_b200.src.point_source.point_source import PointSourceField

    sc = PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=seed)
    sc.gp.lazy = True
    return sc, cls(sc, **kw)


@pytest.mark.parametrize("S", [0, 1, 3])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_tree_information_vs_reference(gold, S, variant):
    from mripp_b200.src.rrt import rrt as R

    gains = [R.point_source_gain_no_penalties, R.point_source_gain_only_distance_penalty,
             R.point_source_gain_distance_rotation_penalty, R.point_source_gain_all]
    g = gold("gain_golden.npz")
    tag = f"S{S}_v{variant}"
    sc, st = _strategy(R.MR_IPP, 40, 5, budget=120, d_waypoint_distance=2.5, num_agents=2, budget_iter=4)
    np.random.seed(int(g[f"{tag}_seed"]))
    st.best_estimates = g[f"{tag}_est"] if S else np.array([])
    st.agents_obs_wp = [list(g[f"{tag}_obs0"]), list(g[f"{tag}_obs1"])]
    start = g[f"{tag}_start"]
    st.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
    st.initialize_trees(start, 0)
    R.rig_tree_generation(st, 120, 0, gain_function=gains[variant])
    tree = st.trees_soa[0]
    assert np.array_equal(tree.points(), g[f"{tag}_pts"])
    assert np.array_equal(tree.parent[: tree.n], g[f"{tag}_parent_final"])
    info, ref = tree.info[: tree.n], g[f"{tag}_info"]
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(info)) and np.array_equal(info[~fin], ref[~fin])
    assert np.abs(info[fin] - ref[fin]).max() <= 1e-12 * max(1.0, np.abs(ref[fin]).max())
    assert int(np.argmax(info)) == int(np.argmax(ref))
    assert st.scorer.launches > 0


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_selectors_vs_reference(gold, tag):
    from mripp_b200.src.rrt import rrt as R

    g = gold("select_golden.npz")
    sc, st = _strategy(R.MR_IPP, 40, 5, budget=150, d_waypoint_distance=2.5, num_agents=2, budget_iter=4,
                       beta_t=float(g[f"{tag}_beta"]))
    np.random.seed(int(g[f"{tag}_seed"]))
    start = g[f"{tag}_start"]
    st.agents_obs_wp = [list(g[f"{tag}_obs0"]), list(g[f"{tag}_obs1"])]
    st.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
    st.best_estimates = g[f"{tag}_est"]
    st.initialize_trees(start, 0)
    R.rig_tree_generation(st, 150, 0, gain_function=R.point_source_gain_all)
    assert np.array_equal(st.trees_soa[0].points(), g[f"{tag}_pts"])
    p1 = R.bias_beta_path_selection(st, 0, start, 30.0)
    assert np.array_equal(np.array(p1), g[f"{tag}_path_bias"])
    p2 = R.informative_source_metric_path_selection(st, 0, start, 30.0)
    assert np.array_equal(np.array(p2), g[f"{tag}_path_info"])
    obs0 = g[f"{tag}_obs0"]
    if obs0.shape[0]:
        pts, nchild = g[f"{tag}_pts"], g[f"{tag}_nchild"]
        leaves = np.ascontiguousarray(pts[nchild == 0])
        mi = st.scorer.mutual_information(leaves, np.ascontiguousarray(obs0), 1.0, 1.0, float(g[f"{tag}_beta"]))
        assert mi == pytest.approx(float(g[f"{tag}_mi"]), rel=1e-10)


def test_mutual_information_both_orders_and_overflow():
    from mripp_b200.scoring import GpuScorer

    sc = GpuScorer()
    rng = np.random.RandomState(0)
    for L, na in [(5, 200), (200, 5), (70, 70), (300, 130)]:
        leaves, obs = rng.uniform(0, 20, (L, 2)), rng.uniform(0, 20, (na, 2))
        K = O.kernel(leaves, obs, 1.0, 1.0)
        ref = T.mutual_information(K, L, 5.0)
        assert sc.mutual_information(leaves, obs, 1.0, 1.0, 5.0) == pytest.approx(ref, rel=1e-10)
    # numpy.linalg.det overflows to inf for large L: same value here
    leaves, obs = rng.uniform(0, 200, (1500, 2)), rng.uniform(0, 200, (1500, 2))
    assert sc.mutual_information(leaves, obs, 1.0, 1.0, 50.0) == np.inf
    assert sc.mutual_information(leaves, np.zeros((0, 2)), 1.0, 1.0, 5.0) == 0


def test_count_close_and_argmax_semantics():
    from mripp_b200.scoring import GpuScorer

    sc = GpuScorer()
    rng = np.random.RandomState(4)
    P, obs = rng.uniform(0, 30, (777, 2)), rng.uniform(0, 30, (2100, 2))
    obs[:50] = P[:50] + np.array([2.5, 0.0])  # knife-edge pairs at exactly d (up to rounding)
    counts, dev = sc.close_counts(P, obs, 2.5)
    got = counts if counts is not None else dev.cpu().numpy()
    ref = np.array([sum(1 for o in obs if np.linalg.norm(p - o) < 2.5) for p in P[:120]])
    assert np.array_equal(np.asarray(got[:120], dtype=np.int64), ref)
    # the pruned kernel (observations in Z-order + tile boxes) counts the same, knife-edge pairs included
    from mripp_b200.scoring import spatial_tiles

    srt, bbox = spatial_tiles(obs)
    assert bbox is not None and bbox.shape == ((len(obs) + 63) // 64, 4)
    counts_t, dev_t = sc.close_counts(P, srt, 2.5, bbox)
    got_t = counts_t if counts_t is not None else dev_t.cpu().numpy()
    full = counts if counts is not None else dev.cpu().numpy()
    assert np.array_equal(np.asarray(got_t, dtype=np.int64), np.asarray(full, dtype=np.int64))
    for dd in (0.05, 40.0):  # nothing in reach / everything in reach
        a, da = sc.close_counts(P[:200], obs, dd)
        b, db = sc.close_counts(P[:200], srt, dd, bbox)
        assert np.array_equal(np.asarray(a if a is not None else da.cpu().numpy(), dtype=np.int64),
                              np.asarray(b if b is not None else db.cpu().numpy(), dtype=np.int64))
    for v in [np.array([1.0, 3.0, 3.0, 2.0]), np.array([1.0, np.nan, 5.0, np.nan]), np.array([-np.inf, -np.inf]),
              np.array([np.inf, 1.0, np.inf]), rng.randn(100000)]:
        i, _ = sc.argmax_first(v)
        assert i == int(np.argmax(v))


def test_branch_gain_with_obstacles_and_large_tree_vs_oracle():
    """8192 candidate branches (C4's batch) on a seeded tree with rectangle obstacles, all variants."""
    from mripp_b200.scoring import GpuScorer, ScoreContext
    from mripp_b200.src.rrt.rrt_utils import TreeArrays

    rng = np.random.RandomState(9)
    W = 100
    tree = TreeArrays(np.array([50.0, 50.0]))
    nodes = [T.Node(np.array([50.0, 50.0]))]
    N = 8193
    for k in range(1, N):
        par = int(rng.randint(max(0, k - 400), k))
        p = tree.point(par) + rng.uniform(-1.0, 1.0, 2)
        tree.add(p, par)
        T.attach(nodes, T.Node(p), nodes[par])
    est = np.array([[30.0, 40.0, 5e5], [60.0, 55.0, 7e5], [52.0, 48.0, 3e5]])
    obs = [list(rng.uniform(40, 60, (300, 2))), list(rng.uniform(40, 60, (200, 2)))]
    obstacles = [{"type": "rectangle", "x": 51.0, "y": 51.0, "width": 2.0, "height": 1.5}]
    ctx = ScoreContext(est, np.array([50.0, 50.0]), obs, 0, 1.0, (0, W, 0, W), obstacles)
    octx = T.GainContext(est, np.array([50.0, 50.0]), obs, 0, 2, 1.0, (0, W, 0, W), obstacles)
    sc = GpuScorer()
    sample = rng.choice(np.arange(1, N), 300, replace=False)
    for variant in range(4):
        info = sc.tree_information(variant, tree, ctx)
        ref = np.array([T.branch_gain(variant, octx, nodes[i]) for i in sample])
        got = info[sample]
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(got)) and np.array_equal(got[~fin], ref[~fin])
        assert np.abs(got[fin] - ref[fin]).max() <= 1e-12 * max(1.0, np.abs(ref[fin]).max())
    assert (~np.isfinite(sc.tree_information(3, tree, ctx))).any()  # the obstacle is hit by some branch


# ---------------------------------------------------------------------------------------------------
# T3: whole strategy.run() on the device path
# ---------------------------------------------------------------------------------------------------
RUN_COMMON = dict(budget=40, d_waypoint_distance=2.5, budget_iter=2, n_samples=60, s_stages=4, max_sources=2, lambda_b=1,
                  beta_t=5.0)


@pytest.mark.parametrize("tag,cls_name,extra", [
    ("mripp_1a", "MR_IPP", dict(num_agents=1, stage_lambda=0.5)),
    ("mripp_2a", "MR_IPP", dict(num_agents=2, stage_lambda=0.5)),
    ("rig_dist", "RRTRIG_PointSourceInformative_Distance_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
    ("rig_nopen", "RRTRIG_PointSourceInformative_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
    ("rig_rot", "RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_PathPlanning", dict(num_agents=1, stage_lambda=0.6)),
    ("biasbeta", "RRT_BiasBetaInformative_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.5)),
    ("rrt_random", "RRT_Random_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.0)),
    ("rrtstar_random", "RRTStar_Random_GP_PathPlanning", dict(num_agents=1, stage_lambda=0.0)),
])
def test_run_waypoints_identical_on_device(gold, tag, cls_name, extra):
    from mripp_b200.src.rrt import rrt as R

    from mripp_b200.src.point_source.point_source import PointSourceField

    g = gold("runs_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21)
    sc.gp.lazy = True  # per-step refits the reference never reads only consume their RNG draws
    assert np.array_equal(np.array(sc.sources), g[f"{tag}_sources"])
    np.random.seed(1234)
    st = getattr(R, cls_name)(sc, **RUN_COMMON, **extra)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Z, std = st.run()
    assert np.array_equal(np.asarray(st.obs_wp, float), g[f"{tag}_obs_wp"])
    assert np.array_equal(np.asarray(st.measurements, float), g[f"{tag}_meas"])
    for i, p in enumerate(st.agents_full_path):
        assert np.array_equal(np.array(p, float).reshape(-1, 2), g[f"{tag}_path{i}"])
    assert np.array_equal(np.random.uniform(size=3), g[f"{tag}_rng_after"])
    if f"{tag}_best_estimates" in g.files:  # the source estimates of the device-likelihood sampler: the reference's bits
        assert np.array_equal(np.asarray(st.best_estimates, float).reshape(-1, 3), g[f"{tag}_best_estimates"])
    assert Z.shape == (200, 200) and std.shape == (40000,)
    # T2 (re-optimised): the LML surface is flat and its gradient is noise limited, so L-BFGS-B may stop
    # at a different but equally good theta (permuting the data moves theta* by 6 %).
    # Criterion: the reference's own LML, evaluated by the oracle, agrees at the two optima.
    Xo, ylog = np.asarray(st.obs_wp, float), np.log10(np.maximum(np.asarray(st.measurements, float), 1e-6))
    yn, _, _ = O.normalise_y(ylog)
    v_ref = O.lml(g[f"{tag}_theta"], Xo, yn, False)
    v_gpu = O.lml(sc.gp.kernel_.theta, Xo, yn, False)
    # ... up to the evaluation noise of that LML itself: re-measured waypoints (duplicate rows) put
    # 1/(2 jitter) ~ 5e9 into K^-1, and moving theta by 1e-9 already moves the reference's LML by O(1)
    # on these runs (range 1.6 for rig_dist, 0.06 for biasbeta).
    prng = np.random.RandomState(0)
    noise = np.ptp([O.lml(g[f"{tag}_theta"] + 1e-9 * prng.randn(2), Xo, yn, False) for _ in range(24)])
    assert v_gpu >= v_ref - (1e-3 + 2.0 * noise), (v_gpu, v_ref, noise, sc.gp.kernel_.theta, g[f"{tag}_theta"])
    # T2 field tier: 1e-2 wherever the two optimisers stop in the same place.  Measured on the B200
    # (tools/t2_report.py, gpurun_out/t2_report.log): 1.9e-4 .. 8.9e-3 for six of the nine cases.  The exceptions are the
    # 28-observation rig_* runs (the same observations in all three: theta* = (3.61, 3.62) here vs (4.16, 3.71) in the
    # reference, log10 Z off by 5.9e-2) and rrtstar_random (1.3e-2): their LML is flat in log c to within its own
    # evaluation noise (`noise` above: O(1) against an LML difference of 1.55 / 0.29), so L-BFGS-B stops where the
    # rounding of its last line search puts it.  The T1 check below pins the field at the reference's theta.
    t2 = {"rig_dist": 1e-1, "rig_nopen": 1e-1, "rig_rot": 1e-1, "rrtstar_random": 3e-2}.get(tag, 1e-2)
    assert np.abs(np.log10(Z[::4, ::4]) - np.log10(g[f"{tag}_Z"])).max() <= t2
    assert np.abs(std.reshape(Z.shape)[::4, ::4] - g[f"{tag}_std"]).max() <= t2
    # T1 on the same run: at the REFERENCE's theta the device field matches the reference's field
    th = g[f"{tag}_theta"]
    sc.gp.kernel = type(sc.gp.kernel)(float(np.exp(th[0])), float(np.exp(th[1])))
    sc.gp.optimizer, sc.gp.lazy = None, False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Z1, std1 = sc.predict_spatial_field(st.obs_wp, np.asarray(st.measurements, float))
    assert np.abs(np.log10(Z1[::4, ::4]) - np.log10(g[f"{tag}_Z"])).max() <= 1e-4
    assert np.abs(std1.reshape(Z.shape)[::4, ::4] - g[f"{tag}_std"]).max() <= 1e-6


@pytest.mark.parametrize("tag,cls_name,kw", [("boust", "Boustrophedon", dict(budget=60)),
                                             ("mr_boust", "MR_Boustrophedon", dict(budget=40, num_agents=2))])
def test_boustrophedon_on_device(gold, tag, cls_name, kw):
    """The lawn-mower baselines: same waypoints and RNG position as the reference, field within T2 / T1."""
    from mripp_b200.src.boustrophedon import boustrophedon as B
    from mripp_b200.src.point_source.point_source import PointSourceField

    g = gold("runs_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21)
    np.random.seed(99)
    st = getattr(B, cls_name)(sc, d_waypoint_distance=2.5, **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Z, std = st.run()
    assert np.array_equal(np.asarray(st.obs_wp, float), g[f"{tag}_obs_wp"])
    assert np.array_equal(np.asarray(st.full_path, float), g[f"{tag}_full_path"])
    assert np.array_equal(np.random.uniform(size=3), g[f"{tag}_rng_after"])
    assert np.abs(np.log10(Z[::4, ::4]) - np.log10(g[f"{tag}_Z"])).max() <= 1e-2  # T2 (measured: 1e-14, same optimum)
    th = g[f"{tag}_theta"]
    sc.gp.kernel = type(sc.gp.kernel)(float(np.exp(th[0])), float(np.exp(th[1])))
    sc.gp.optimizer = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Z1, std1 = sc.predict_spatial_field(st.obs_wp, np.asarray(st.measurements, float))
    assert np.abs(np.log10(Z1[::4, ::4]) - np.log10(g[f"{tag}_Z"])).max() <= 1e-4
    assert np.abs(std1.reshape(Z1.shape)[::4, ::4] - g[f"{tag}_std"]).max() <= 1e-6


def test_main_config_flow_on_device(tmp_path):
    """The reference's CLI flow (main.py: JSON -> scenarios -> strategies by name -> run -> metrics) end to end on the
    device, with the strategy names of the reference's shipped config (aliases of MR_Boustrophedon / MR_IPP)."""
    import json

    from mripp_b200 import main as M
    from tests.conftest import ROOT

    out = tmp_path / "metrics.json"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = M.main(["-config", os.path.join(ROOT, "config", "small_example_setup.json"), "--out", str(out)])
    assert len(res) == 2 and {r["strategy"] for r in res} == {
        "MultiAgentBoustrophedon", "RRTRIG_PointSourceInformative_All_SourceMetric_PathPlanning"}
    for r in json.load(open(out)):
        assert np.isfinite(r["rmse"]) and np.isfinite(r["wrmse"]) and np.isfinite(r["entropy"]) and r["n_obs"] > 10
        assert r["rmse"] < 2.0  # log10 units: the field is reconstructed, not garbage


@pytest.mark.parametrize("tag", ["free", "obst"])
def test_ground_truth_and_metrics_on_device(gold, torch, tag):
    """the analytic field (with a rectangle obstacle) and the fused end-of-run metrics on the
    device against the reference's golden field and the host expressions."""
    from mripp_b200.src.point_source.point_source import PointSourceField
    from mripp_b200.src.visualization.plot_helper import calculate_differential_entropy, field_metrics_device

    g = gold("field_golden.npz")
    obst = [{"type": "rectangle", "x": 4, "y": 5, "width": 3, "height": 2}] if tag == "obst" else []
    sc = PointSourceField(num_sources=2, workspace_size=(0, 12, 0, 12), obstacles=obst, seed=8)
    ref = g[f"{tag}_gtruth"]
    dev = sc.ground_truth_device()
    assert dev.shape == ref.shape
    assert np.abs(dev / ref - 1.0).max() <= 1e-12
    # metrics: a perturbed copy of the field as "prediction", a positive std with a few exact zeros
    rng = np.random.RandomState(3)
    zp = ref * np.exp(0.3 * rng.randn(*ref.shape))
    sd = np.abs(rng.randn(ref.size)) * 0.2
    sd[::97] = 0.0
    err = np.log10(ref + 1) - np.log10(zp + 1)
    rmse, wrmse = np.sqrt(np.mean(err ** 2)), np.sqrt(np.sum(ref * err ** 2) / np.sum(ref))
    ent = calculate_differential_entropy(sd.copy())
    got = field_metrics_device(torch.from_numpy(zp).cuda(), torch.from_numpy(sd).cuda(), torch.from_numpy(ref).cuda())
    assert got[0] == pytest.approx(rmse, rel=1e-12) and got[1] == pytest.approx(wrmse, rel=1e-12)
    assert got[2] == pytest.approx(ent, rel=1e-12)


def test_ground_truth_device_many_obstacles_matches_host(torch):
    from mripp_b200.src.point_source.point_source import PointSourceField

    obst = [{"type": "rectangle", "x": 2.0 + 3 * i, "y": 1.5 + 2.5 * (i % 3), "width": 1.2, "height": 0.9} for i in range(5)]
    sc = PointSourceField(num_sources=4, workspace_size=(0, 18, 0, 9), obstacles=obst, seed=2)
    host, dev = sc.g_truth, sc.ground_truth_device()
    assert np.abs(dev / host - 1.0).max() <= 1e-12
    # ground_truth_mode: small lattices stay on the reference's host expressions, large obstacle lattices (where the
    # host's per-point Python loop takes minutes) go to the device, "host" / "device" force either
    assert np.array_equal(host, sc.ground_truth_host())                       # auto, 16 200 points -> host
    big = PointSourceField(num_sources=2, workspace_size=(0, 40, 0, 30), obstacles=obst[:2], seed=2)   # 120 000 points
    assert big.ground_truth_mode == "auto" and np.array_equal(big.g_truth, big.ground_truth_device())
    sub = (slice(0, 300, 37), slice(0, 400, 41))
    X0, Y0 = big.X, big.Y
    big.X, big.Y = X0[sub], Y0[sub]                                           # the host expressions on a sub-lattice
    small_host = big.ground_truth_host()
    big.X, big.Y = X0, Y0
    assert np.abs(big.g_truth[sub] / small_host - 1.0).max() <= 1e-12
    forced = PointSourceField(num_sources=2, workspace_size=(0, 6, 0, 6), seed=2, ground_truth_mode="device")
    assert np.abs(forced.g_truth / forced.ground_truth_host() - 1.0).max() <= 1e-12


def test_abi_rejects_bad_arguments_on_device(torch):
    from mripp_b200 import _lib

    lib = _lib.load()
    out = (C.c_longlong * 10)()
    assert lib.ipp_gp_layout(-3, out) == -1
    x = torch.zeros(4, 2, dtype=torch.float64, device="cuda")
    assert lib.ipp_gram_rbf(_lib.ptr(x), 4, None, 0, 1.0, -1.0, 0.0, _lib.ptr(x), 4, None) == -1
    assert lib.ipp_branch_gain(7, None, 0, None, None, None, None, None, None, None, 0, 0, None, 0, None, None, 0, None,
                               None, None) == -1


# ---- particle log-likelihoods of the importance sampler -----------------------------------
def _likelihood_case(n, N, M, W, seed):
    rng = np.random.RandomState(seed)
    obs = rng.uniform(0, W, (n, 2))
    truth = np.column_stack([rng.uniform(0, W, M), rng.uniform(0, W, M), rng.uniform(1e5, 1e6, M)])
    d2 = ((obs[:, None, :] - truth[None, :, :2]) ** 2).sum(-1)
    meas = 1.0 + (truth[None, :, 2] / np.maximum(d2, 1e-6)).sum(1)
    meas = meas * (1 + 1e-3 * rng.randn(n))
    thetas = np.column_stack([c for _ in range(M) for c in (rng.uniform(0, W, N), rng.uniform(0, W, N), rng.uniform(1e5, 1e6, N))])
    return obs, meas, thetas


@pytest.mark.parametrize("n,N,M", [(1, 7, 1), (25, 80, 2), (257, 200, 3), (2000, 200, 3), (5000, 333, 2), (300, 50, 8),
                                   (300, 50, 9), (64, 20, 17)])
def test_particle_log_likelihood_within_its_bound(n, N, M):
    """ipp_poisson_loglik against the reference's expression (estimation.py:15-55 through scipy's logpmf): the
    difference stays inside a quarter of the bound the sampler derives from the returned sums, including a particle
    sitting on an observation (the 1e-6 clamp), zero counts and near-truth particles (tiny terms, large pieces)."""
    from mripp_b200.scoring import GpuScorer
    from mripp_b200.src.estimation import estimation as E

    obs, meas, thetas = _likelihood_case(n, N, M, 40.0, 100 * n + M)
    thetas[0, 0:2] = obs[0]                       # d2 = 0 -> clamp
    meas[n // 2] = 0.2                            # count 0
    if N > 2:
        thetas[1] = thetas[2]                     # identical particles -> identical values
    sc = GpuScorer()
    prep = sc.prepare_counts(list(obs), list(meas))
    ll, s1, s2 = sc.particle_log_likelihood(thetas, prep, 1, M)
    want = E.batched_poisson_log_likelihood(thetas, list(obs), list(meas), 1, M)
    bound = E.EPS * (E.C_TERM * s1 + E.c_sum(n) * s2)
    assert np.all(np.isfinite(ll)) and np.all(bound > 0)
    assert np.all(np.abs(ll - want) <= 0.6 * bound), float(np.max(np.abs(ll - want) / bound))
    assert np.all(bound <= 1e-12 * s1 + 1e-300)   # the band stays narrow: about 1e4 eps of the pieces' magnitude
    if N > 2:
        assert ll[1] == ll[2]


def test_particle_rate_is_bit_identical_to_numpy():
    """With one observation of count 0 the log-likelihood is exactly -rate: the kernel's rate (subtractions, squares,
    clamp, division, numpy's source-axis summation order for M < 8 and M >= 8) carries numpy's bits."""
    from mripp_b200.scoring import GpuScorer

    sc = GpuScorer()
    rng = np.random.RandomState(3)
    for M in (1, 2, 3, 7, 8, 9, 15, 16, 23, 32):
        obs = rng.uniform(0, 30, (1, 2))
        thetas = rng.uniform(0, 30, (400, 3 * M))
        thetas[:, 2::3] = rng.uniform(1e5, 1e6, (400, M))
        thetas[5, 0:2] = obs[0]
        prep = sc.prepare_counts(list(obs), [0.3])
        ll, _, _ = sc.particle_log_likelihood(thetas, prep, 1, M)
        src = thetas.reshape(400, M, 3)
        d2 = (obs[None, :, None, 0] - src[:, None, :, 0]) ** 2 + (obs[None, :, None, 1] - src[:, None, :, 1]) ** 2
        rate = np.maximum(1 + np.sum(src[:, None, :, 2] / np.maximum(d2, 1e-6), axis=2), 1e-10)[:, 0]
        assert np.array_equal(ll, -rate), M


def test_resampled_indices_are_the_references():
    """Stage by stage: np.random.choice on the reference's weights vs the device-likelihood sampler from the same RNG
    state -- same indices, same RNG position -- over spread (few observations) and peaked (many) posteriors."""
    from mripp_b200.scoring import GpuScorer
    from mripp_b200.src.estimation import estimation as E

    sc = GpuScorer()
    before = dict(E.STATS)
    stages = 0
    for case, (n, N, M) in enumerate([(3, 200, 1), (5, 200, 2), (10, 300, 3), (40, 200, 3), (400, 200, 2), (2000, 200, 3)]):
        obs, meas, thetas = _likelihood_case(n, N, M, 40.0, 7 + case)
        prep = sc.prepare_counts(list(obs), list(meas))
        rng = np.random.RandomState(case)
        for rep in range(12):
            P = thetas if rep == 0 else thetas[rng.randint(N)] * (1 + 1e-3 * rng.randn(N, 3 * M))   # spread / clustered cloud
            if rep == 1:
                P = np.tile(thetas[rng.randint(N)], (N, 1))                                          # collapsed cloud
            for gamma in (0.1, 0.5, 1.0):
                np.random.seed(1000 * case + rep)
                w = E.stage_weights(E.batched_poisson_log_likelihood(P, list(obs), list(meas), 1, M), gamma, N)
                want = np.random.choice(N, size=N, p=w)
                after = np.random.uniform(size=2)
                np.random.seed(1000 * case + rep)
                got = E.resample_indices(P, gamma, N, list(obs), list(meas), 1, M, sc, prep)
                assert np.array_equal(want, got) and np.array_equal(after, np.random.uniform(size=2))
                stages += 1
    redone = E.STATS["redecided"] - before["redecided"]
    assert E.STATS["device_evaluations"] - before["device_evaluations"] == stages
    assert redone <= 0.05 * stages, (redone, stages)   # the band is hit rarely; when it is, the stage is re-decided


def test_source_estimator_on_device_matches_reference(gold):
    """estimate_sources_bayesian with the device likelihood: estimate, model order, BIC and the RNG position after
    are the reference's own (golden from the unmodified reference, oracle/make_golden.py gold_field)."""
    from mripp_b200.src.estimation import estimation as E
    from mripp_b200.src.point_source.point_source import PointSourceField

    g = gold("field_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21)
    before = dict(E.STATS)
    np.random.seed(13)
    est, M, bic = E.estimate_sources_bayesian(list(g["est_wp"]), list(g["est_meas"]), 1, 2, 80, 5, sc)
    assert M == int(g["est_M"]) and bic == float(g["est_bic"])
    assert np.array_equal(est, g["est_theta"])
    assert np.array_equal(np.random.uniform(size=3), g["est_rng_after"])
    assert E.STATS["device_evaluations"] - before["device_evaluations"] == 10


def test_work_balanced_lattice_slabs(G):
    """dist.lattice_slabs: the cheap per-row work model splits one lattice so that the EXACT executed slab counts of
    the block-sparse kernel (gp.lattice_work, the host restatement of the kernel's own test) are near equal across
    8 ranks, and closer than equal rows; per-tile counts add up to the total."""
    from mripp_b200 import dist as D

    X, ylog = _make(3000, 80.0, seed=4)
    gp = _gp_at(G, X, ylog, np.log([1.0, 2.0]))
    nx = ny = 800
    ws = (0, 80, 0, 80)
    total, dense = gp.lattice_work(ws[0], ws[1], nx, ws[2], ws[3], ny)
    assert 0 < total < 0.6 * dense                                   # the sparse path is on at this length scale
    tiles = gp.lattice_work(ws[0], ws[1], nx, ws[2], ws[3], ny, per_tile=True)
    assert int(tiles.sum()) == total
    def spread(bounds):
        w = np.array([gp.lattice_work(ws[0], ws[1], nx, ws[2], ws[3], ny, g_begin=a, g_end=b)[0] for a, b in bounds], float)
        return w.max() / w.mean()
    balanced = D.lattice_slabs(gp, ws, nx, ny, 8)
    equal = [D.row_slab(nx, ny, 8, r) for r in range(8)]
    assert balanced[0][0] == 0 and balanced[-1][1] == nx * ny and all(b[1] > b[0] for b in balanced)
    assert spread(balanced) <= 1.06 and spread(balanced) < spread(equal)

"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, which is absent on
the GPU box); the fixtures it writes are committed.  The reference ships no golden vectors for
this path -- these are outputs of the reference's own code:
  src/point_source/point_source.py (PointSourceField + sklearn GaussianProcessRegressor),
  src/rrt/rrt.py (gain functions, tree generators, selectors, strategy classes),
  src/boustrophedon/boustrophedon.py.

Shims applied from the outside (never by editing the reference):
  * matplotlib / imageio import stubs (oracle/_stubs), because neither is installed;
  * src.rrt.rrt.Thread -> a synchronous stand-in, so multi-agent runs are deterministic (agents
    0..N-1 in order within each outer iteration -- one legal interleaving of the reference);
  * np.random.seed(s) after constructing PointSourceField (its ctor re-seeds with None);
  * best_estimates = np.array([]) for the GP-named classes that forget to set it.

Usage:  python oracle/make_golden.py [--only gp|gain|select|runs|field]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("IPP_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True


def import_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
    sys.path.insert(0, REF)
    import src.rrt.rrt as R  # noqa
    import src.point_source.point_source as PS  # noqa
    import src.boustrophedon.boustrophedon as B  # noqa

    class SyncThread:
        def __init__(self, target=None, args=(), kwargs=None):
            self._t, self._a, self._k = target, args, kwargs or {}

        def start(self):
            self._t(*self._a, **self._k)

        def join(self):
            return None

    R.Thread = SyncThread
    R.tqdm = _QuietTqdm
    return R, PS, B


class _QuietTqdm:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):
        pass

    @staticmethod
    def write(*a, **k):
        pass


def versions():
    import scipy
    import sklearn

    return np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}", f"sklearn {sklearn.__version__}"])


def synth_obs(PS, W, n, num_sources, seed, dup=0.2):
    """SURVEY.md §8d synthetic inputs: uniform waypoints with 20 % exact duplicates, measurements
    from the reference's own simulator."""
    sc = PS.PointSourceField(num_sources=num_sources, workspace_size=(0, W, 0, W), seed=seed)
    rng = np.random.RandomState(0)
    X = rng.uniform(0, W, (n, 2))
    k = int(dup * n)
    if k:
        X[-k:] = X[:k]
    np.random.seed(0)
    meas = sc.simulate_measurements(X)
    return sc, X, meas


# ------------------------------------------------------------------------------------------------
def gold_gp(PS):
    """(X, y, theta) -> (lml, grad); fitted state; posterior at points and on the grid."""
    out = {"versions": versions()}
    for tag, W, n, S, seed in [("n40", 10, 40, 2, 3), ("n131", 40, 131, 2, 7), ("n300", 20, 300, 3, 11)]:
        sc, X, meas = synth_obs(PS, W, n, S, seed)
        ylog = np.log10(np.maximum(meas, 1e-6))
        np.random.seed(123)
        t = time.time()
        sc.update(X, ylog)  # point_source.py:335-337 -> 11 L-BFGS-B runs
        gp = sc.gp
        thetas = np.array([[0.0, 0.0], [0.5, 1.0], [-1.0, -0.5], gp.kernel_.theta])
        lmls, grads = [], []
        for th in thetas:
            v, g = gp.log_marginal_likelihood(th, eval_gradient=True)
            lmls.append(v)
            grads.append(g)
        rngq = np.random.RandomState(5)
        Q = rngq.uniform(0, W, (400, 2))
        Q[:30] = X[:30] + 1e-3
        qm, qs = gp.predict(Q, return_std=True)
        r = np.column_stack((sc.X.ravel(), sc.Y.ravel()))
        gm, gs = gp.predict(r, return_std=True)
        stride = max(1, len(r) // 20000)
        out.update({
            f"{tag}_W": W, f"{tag}_X": X, f"{tag}_meas": meas, f"{tag}_thetas": thetas, f"{tag}_lml": np.array(lmls),
            f"{tag}_grad": np.array(grads), f"{tag}_theta_fit": gp.kernel_.theta.copy(),
            f"{tag}_lml_fit": gp.log_marginal_likelihood_value_, f"{tag}_alpha": gp.alpha_.copy(),
            f"{tag}_Ldiag": np.diag(gp.L_).copy(), f"{tag}_ymean": gp._y_train_mean, f"{tag}_ystd": gp._y_train_std,
            f"{tag}_Q": Q, f"{tag}_qmean": qm, f"{tag}_qstd": qs, f"{tag}_gstride": stride,
            f"{tag}_gmean": gm[::stride].copy(), f"{tag}_gstd": gs[::stride].copy(),
            f"{tag}_x": sc.x.copy(), f"{tag}_y": sc.y.copy(),
        })
        print(f"gp {tag}: fit {time.time() - t:.1f}s theta*={gp.kernel_.theta}")
    # whole predict_spatial_field on a small workspace, global RNG seeded (restart draws)
    sc, X, meas = synth_obs(PS, 10, 60, 2, 3)
    np.random.seed(77)
    Z, std = sc.predict_spatial_field(X, meas)
    out.update(psf_W=10, psf_X=X, psf_meas=meas, psf_Z=Z, psf_std=std, psf_theta=sc.gp.kernel_.theta.copy(),
               psf_rng_after=np.random.uniform(size=3))
    np.savez_compressed(os.path.join(GOLD, "gp_golden.npz"), **out)


# ------------------------------------------------------------------------------------------------
def snapshot_tree(nodes):
    idx = {id(nd): i for i, nd in enumerate(nodes)}
    pts = np.array([nd.point for nd in nodes], dtype=float)
    parent = np.array([idx.get(id(nd.parent), -1) if nd.parent is not None else -1 for nd in nodes])
    nchild = np.array([len(nd.children) for nd in nodes])
    info = np.array([float(getattr(nd, "information", 0.0)) for nd in nodes])
    return pts, parent, nchild, info


def gold_gain(R, PS):
    """Trees grown by the reference's rig_tree_generation with each gain function; per-node
    `information` as the reference computed it at insertion time."""
    out = {"versions": versions()}
    gains = [R.point_source_gain_no_penalties, R.point_source_gain_only_distance_penalty,
             R.point_source_gain_distance_rotation_penalty, R.point_source_gain_all]
    W = 40
    for S in (0, 1, 3):
        for v, gfun in enumerate(gains):
            sc = PS.PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=5)
            np.random.seed(100 + 10 * S + v)
            strat = R.MR_IPP(sc, budget=120, d_waypoint_distance=2.5, num_agents=2, budget_iter=4)
            est_rng = np.random.RandomState(S)
            strat.best_estimates = (np.column_stack([est_rng.uniform(0, W, S), est_rng.uniform(0, W, S),
                                                     est_rng.uniform(1e5, 1e6, S)]) if S else np.array([]))
            obs_rng = np.random.RandomState(9)
            strat.agents_obs_wp = [list(obs_rng.uniform(0, W, (30, 2))), list(obs_rng.uniform(0, W, (25, 2)))]
            # a root child sits exactly one step from an observed root: the `< d_wp` knife edge
            start = np.array([12.5, 17.25])
            strat.agents_obs_wp[0].append(start.copy())
            strat.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
            strat.initialize_trees(start, 0)
            R.rig_tree_generation(strat, 120, 0, gain_function=gfun)
            pts, parent, nchild, info = snapshot_tree(strat.tree_nodes[0])
            tag = f"S{S}_v{v}"
            out.update({f"{tag}_pts": pts, f"{tag}_parent_final": parent, f"{tag}_nchild": nchild, f"{tag}_info": info,
                        f"{tag}_est": np.asarray(strat.best_estimates, float).reshape(-1, 3),
                        f"{tag}_obs0": np.array(strat.agents_obs_wp[0]), f"{tag}_obs1": np.array(strat.agents_obs_wp[1]),
                        f"{tag}_seed": 100 + 10 * S + v, f"{tag}_start": start})
            print(f"gain {tag}: {len(pts)} nodes, info range [{info.min():.4g}, {info.max():.4g}]")
    np.savez_compressed(os.path.join(GOLD, "gain_golden.npz"), **out)


# ------------------------------------------------------------------------------------------------
def gold_select(R, PS):
    """bias_beta_path_selection / mutual_information_gain / informative_source_metric_path_selection."""
    out = {"versions": versions()}
    W = 40
    for tag, n_obs0, n_obs1, beta in [("a", 0, 0, 5.0), ("b", 12, 7, 5.0), ("c", 40, 33, 0.5)]:
        sc = PS.PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=5)
        np.random.seed(41)
        strat = R.MR_IPP(sc, budget=150, d_waypoint_distance=2.5, num_agents=2, budget_iter=4, beta_t=beta)
        rng = np.random.RandomState(3)
        start = np.array([20.0, 20.0])
        strat.agents_obs_wp = [list(rng.uniform(5, 35, (n_obs0, 2))), list(rng.uniform(5, 35, (n_obs1, 2)))]
        if n_obs0:
            strat.agents_obs_wp[0].append(start.copy())
        strat.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
        strat.best_estimates = np.array([[10.0, 30.0, 5e5], [28.0, 12.0, 7e5]])
        strat.initialize_trees(start, 0)
        R.rig_tree_generation(strat, 150, 0, gain_function=R.point_source_gain_all)
        nodes = strat.tree_nodes[0]
        pts, parent, nchild, info = snapshot_tree(nodes)
        leaves = [nd for nd in nodes if not nd.children]
        leaf_pts = np.array([nd.point for nd in leaves])
        obs0 = np.array(strat.agents_obs_wp[0]).reshape(-1, 2)
        K_pi = sc.gp.kernel(leaf_pts)
        K_pio = sc.gp.kernel(leaf_pts, obs0) if obs0.size else np.zeros((len(leaf_pts), 0))
        mi = R.mutual_information_gain(K_pi, K_pio, beta)
        path1 = R.bias_beta_path_selection(strat, 0, start, 30.0)
        path2 = R.informative_source_metric_path_selection(strat, 0, start, 30.0)
        out.update({f"{tag}_pts": pts, f"{tag}_parent_final": parent, f"{tag}_nchild": nchild, f"{tag}_info": info,
                    f"{tag}_obs0": obs0, f"{tag}_obs1": np.array(strat.agents_obs_wp[1]).reshape(-1, 2), f"{tag}_beta": beta,
                    f"{tag}_mi": float(mi), f"{tag}_Kpio": K_pio, f"{tag}_path_bias": np.array(path1),
                    f"{tag}_path_info": np.array(path2), f"{tag}_start": start, f"{tag}_seed": 41,
                    f"{tag}_est": strat.best_estimates})
        print(f"select {tag}: L={len(leaves)} mi={mi} |path_bias|={len(path1)} |path_info|={len(path2)}")
    np.savez_compressed(os.path.join(GOLD, "select_golden.npz"), **out)


# ------------------------------------------------------------------------------------------------
def gold_runs(R, PS, B):
    """Whole strategy.run() trajectories for fixed seeds (T3: identical waypoint sequences)."""
    out = {"versions": versions()}
    W = 20
    common = dict(budget=40, d_waypoint_distance=2.5, budget_iter=2, n_samples=60, s_stages=4, max_sources=2, lambda_b=1,
                  beta_t=5.0)
    cases = [
        ("mripp_1a", R.MR_IPP, dict(num_agents=1, stage_lambda=0.5), False),
        ("mripp_2a", R.MR_IPP, dict(num_agents=2, stage_lambda=0.5), False),
        ("rig_nopen", R.RRTRIG_PointSourceInformative_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6), False),
        ("rig_dist", R.RRTRIG_PointSourceInformative_Distance_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6), False),
        ("rig_rot", R.RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6), False),
        ("biasbeta", R.RRT_BiasBetaInformative_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.5), True),
        ("rrt_random", R.RRT_Random_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.0), True),
        ("rrtstar_random", R.RRTStar_Random_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.0), True),
    ]
    for tag, cls, extra, shim in cases:
        sc = PS.PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=21)
        np.random.seed(1234)
        strat = cls(sc, **common, **extra)
        if shim:
            strat.best_estimates = np.array([])
        t = time.time()
        Z, std = strat.run()
        out.update({f"{tag}_obs_wp": np.asarray(strat.obs_wp, float), f"{tag}_meas": np.asarray(strat.measurements, float),
                    f"{tag}_Z": Z[::4, ::4].copy(), f"{tag}_std": std.reshape(Z.shape)[::4, ::4].copy(),
                    f"{tag}_theta": sc.gp.kernel_.theta.copy(), f"{tag}_best_estimates": np.asarray(strat.best_estimates, float).reshape(-1, 3),
                    f"{tag}_rng_after": np.random.uniform(size=3), f"{tag}_sources": np.array(sc.sources)})
        for i, p in enumerate(strat.agents_full_path):
            out[f"{tag}_path{i}"] = np.array(p, float).reshape(-1, 2)
        print(f"run {tag}: {time.time() - t:.1f}s, {len(strat.obs_wp)} obs")
    for tag, cls, kw in [("boust", B.Boustrophedon, dict(budget=60)),
                         ("mr_boust", B.MR_Boustrophedon, dict(budget=40, num_agents=2))]:
        sc = PS.PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=21)
        np.random.seed(99)
        strat = cls(sc, d_waypoint_distance=2.5, **kw)
        Z, std = strat.run()
        out.update({f"{tag}_obs_wp": np.asarray(strat.obs_wp, float), f"{tag}_Z": Z[::4, ::4].copy(),
                    f"{tag}_std": std.reshape(Z.shape)[::4, ::4].copy(), f"{tag}_full_path": np.asarray(strat.full_path, float), f"{tag}_theta": sc.gp.kernel_.theta.copy(),
                    f"{tag}_rng_after": np.random.uniform(size=3)})
        print(f"run {tag}: {len(strat.obs_wp)} obs")
    np.savez_compressed(os.path.join(GOLD, "runs_golden.npz"), **out)


# ------------------------------------------------------------------------------------------------
def gold_field(PS):
    """Ground truth / simulated measurements (with and without a rectangular obstacle) and the
    Bayesian source estimator, all from the reference's own host code."""
    import src.estimation.estimation as E

    out = {"versions": versions()}
    W = 12
    obst = [{"type": "rectangle", "x": 4, "y": 5, "width": 3, "height": 2}]
    for tag, obstacles in [("free", []), ("obst", obst)]:
        sc = PS.PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), obstacles=obstacles, seed=8)
        rng = np.random.RandomState(2)
        wp = rng.uniform(0, W, (25, 2))
        np.random.seed(4)
        meas = sc.simulate_measurements(wp)
        out.update({f"{tag}_sources": np.array(sc.sources), f"{tag}_gtruth": sc.g_truth, f"{tag}_wp": wp,
                    f"{tag}_intensity": sc.intensity(wp), f"{tag}_meas": meas, f"{tag}_single": sc.intensity(wp[3])})
    sc = PS.PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21)
    rng = np.random.RandomState(6)
    wp = rng.uniform(0, 20, (30, 2))
    np.random.seed(12)
    meas = sc.simulate_measurements(wp)
    np.random.seed(13)
    est, M, bic = E.estimate_sources_bayesian(list(wp), list(meas), 1, 2, 80, 5, sc)
    thetas = np.random.RandomState(1).uniform([0, 0, 1e5] * 2, [20, 20, 1e6] * 2, (7, 6))
    ll = np.array([E.poisson_log_likelihood(t, wp, meas, 1, 2) for t in thetas])
    out.update(est_wp=wp, est_meas=meas, est_theta=est, est_M=M, est_bic=bic, est_rng_after=np.random.uniform(size=3),
               ll_thetas=thetas, ll_values=ll)
    np.savez_compressed(os.path.join(GOLD, "field_golden.npz"), **out)
    print(f"field: estimate M={M} bic={bic:.3f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    R, PS, B = import_reference()
    if a.only in (None, "gp"):
        gold_gp(PS)
    if a.only in (None, "gain"):
        gold_gain(R, PS)
    if a.only in (None, "select"):
        gold_select(R, PS)
    if a.only in (None, "runs"):
        gold_runs(R, PS, B)
    if a.only in (None, "field"):
        gold_field(PS)


if __name__ == "__main__":
    main()

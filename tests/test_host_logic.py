"""CPU tests of the product's HOST logic against reference-generated goldens: array-backed tree
growth (bit-exact coordinates / parents), versioned-parent scoring order, selectors, RNG order and
the whole strategy.run() waypoint sequences.  Device backends are replaced by the oracle doubles."""
import numpy as np
import pytest

from mripp_b200.scoring import ScoreContext
from mripp_b200.src.boustrophedon.boustrophedon import Boustrophedon, MR_Boustrophedon
from mripp_b200.src.estimation import estimation as E
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt import rrt as R
from mripp_b200.src.rrt import rrt_utils as U
from mripp_b200.src.rrt.rrt_utils import TreeArrays
from tests._doubles import OracleGP, OracleScorer

GAINS = [R.point_source_gain_no_penalties, R.point_source_gain_only_distance_penalty,
         R.point_source_gain_distance_rotation_penalty, R.point_source_gain_all]


class FastGP(OracleGP):
    full = False


def _strategy(cls, W, seed, gp=FastGP, **kw):
    sc = PointSourceField(num_sources=2, workspace_size=(0, W, 0, W), seed=seed, gp_factory=gp)
    return sc, (lambda **extra: cls(sc, scorer=OracleScorer(), **kw, **extra))


@pytest.mark.parametrize("S", [0, 1, 3])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_rig_tree_bit_exact_and_information(gold, S, variant):
    g = gold("gain_golden.npz")
    tag = f"S{S}_v{variant}"
    sc, make = _strategy(R.MR_IPP, 40, 5, budget=120, d_waypoint_distance=2.5, num_agents=2, budget_iter=4)
    np.random.seed(int(g[f"{tag}_seed"]))
    st = make()
    st.best_estimates = g[f"{tag}_est"] if S else np.array([])
    st.agents_obs_wp = [list(g[f"{tag}_obs0"]), list(g[f"{tag}_obs1"])]
    start = g[f"{tag}_start"]
    st.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
    st.initialize_trees(start, 0)
    R.rig_tree_generation(st, 120, 0, gain_function=GAINS[variant])
    tree = st.trees_soa[0]
    assert np.array_equal(tree.points(), g[f"{tag}_pts"])
    assert np.array_equal(tree.parent[: tree.n], g[f"{tag}_parent_final"])
    assert np.array_equal(tree.nchild[: tree.n], g[f"{tag}_nchild"])
    assert np.array_equal(tree.info[: tree.n], g[f"{tag}_info"])
    nodes = st.tree_nodes[0]  # reference-shaped objects
    assert len(nodes) == tree.n and nodes[0].parent is None
    assert [len(nd.children) for nd in nodes] == list(g[f"{tag}_nchild"])


def test_history_csr_orders_events_per_node():
    t = TreeArrays(np.array([0.0, 0.0]))
    for k in range(1, 6):
        t.add(np.array([float(k), 0.0]), k - 1)
    t.rewires = [(3, 1, 3), (5, 1, 5), (4, 2, 4)]
    off, tm, par = t.history_csr()
    assert list(off) == [0, 0, 2, 3, 3, 3, 3]
    assert list(tm) == [3, 5, 4] and list(par) == [3, 5, 4]


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_selectors_bit_exact(gold, tag):
    g = gold("select_golden.npz")
    sc, make = _strategy(R.MR_IPP, 40, 5, budget=150, d_waypoint_distance=2.5, num_agents=2, budget_iter=4,
                         beta_t=float(g[f"{tag}_beta"]))
    np.random.seed(int(g[f"{tag}_seed"]))
    st = make()
    start = g[f"{tag}_start"]
    st.agents_obs_wp = [list(g[f"{tag}_obs0"]), list(g[f"{tag}_obs1"])]
    st.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
    st.best_estimates = g[f"{tag}_est"]
    st.initialize_trees(start, 0)
    R.rig_tree_generation(st, 150, 0, gain_function=R.point_source_gain_all)
    assert np.array_equal(st.trees_soa[0].points(), g[f"{tag}_pts"])
    assert np.array_equal(st.trees_soa[0].info[: st.trees_soa[0].n], g[f"{tag}_info"])
    p1 = R.bias_beta_path_selection(st, 0, start, 30.0)
    assert np.array_equal(np.array(p1), g[f"{tag}_path_bias"])
    p2 = R.informative_source_metric_path_selection(st, 0, start, 30.0)
    assert np.array_equal(np.array(p2), g[f"{tag}_path_info"])
    Kpio = g[f"{tag}_Kpio"]
    if Kpio.size:
        assert R.mutual_information_gain(np.zeros((Kpio.shape[0],) * 2), Kpio, float(g[f"{tag}_beta"])) == pytest.approx(
            float(g[f"{tag}_mi"]), rel=1e-13)


RUNS = [("mripp_1a", R.MR_IPP, dict(num_agents=1, stage_lambda=0.5)),
        ("mripp_2a", R.MR_IPP, dict(num_agents=2, stage_lambda=0.5)),
        ("rig_nopen", R.RRTRIG_PointSourceInformative_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6)),
        ("rig_dist", R.RRTRIG_PointSourceInformative_Distance_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6)),
        ("rig_rot", R.RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_PathPlanning, dict(num_agents=1, stage_lambda=0.6)),
        ("biasbeta", R.RRT_BiasBetaInformative_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.5)),
        ("rrt_random", R.RRT_Random_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.0)),
        ("rrtstar_random", R.RRTStar_Random_GP_PathPlanning, dict(num_agents=1, stage_lambda=0.0))]
COMMON = dict(budget=40, d_waypoint_distance=2.5, budget_iter=2, n_samples=60, s_stages=4, max_sources=2, lambda_b=1, beta_t=5.0)


@pytest.mark.parametrize("tag,cls,extra", RUNS, ids=[r[0] for r in RUNS])
def test_run_waypoint_sequences_identical(gold, tag, cls, extra):
    """T3: chosen waypoints, simulated measurements, source estimates and the RNG position after the
    run are bit-identical to the reference for a fixed seed."""
    g = gold("runs_golden.npz")
    sc, make = _strategy(cls, 20, 21, **COMMON, **extra)
    assert np.array_equal(np.array(sc.sources), g[f"{tag}_sources"])
    np.random.seed(1234)
    st = make()
    Z, std = st.run()
    assert np.array_equal(np.asarray(st.obs_wp, float), g[f"{tag}_obs_wp"])
    assert np.array_equal(np.asarray(st.measurements, float), g[f"{tag}_meas"])
    for i, p in enumerate(st.agents_full_path):
        assert np.array_equal(np.array(p, float).reshape(-1, 2), g[f"{tag}_path{i}"])
    assert np.array_equal(np.asarray(st.best_estimates, float).reshape(-1, 3), g[f"{tag}_best_estimates"])
    assert np.array_equal(np.random.uniform(size=3), g[f"{tag}_rng_after"])
    assert Z.shape == (200, 200) and std.shape == (40000,)


def test_run_field_reoptimised_tier(gold):
    """T2: with the full optimiser the field agrees with the reference within the loose tier."""
    g = gold("runs_golden.npz")
    tag = "rig_nopen"
    sc, make = _strategy(R.RRTRIG_PointSourceInformative_SourceMetric_PathPlanning, 20, 21, gp=OracleGP, **COMMON,
                         num_agents=1, stage_lambda=0.6)
    np.random.seed(1234)
    Z, std = make().run()
    assert np.abs(np.log10(Z[::4, ::4]) - np.log10(g[f"{tag}_Z"])).max() <= 2e-2
    assert np.abs(std.reshape(Z.shape)[::4, ::4] - g[f"{tag}_std"]).max() <= 2e-2


@pytest.mark.parametrize("tag,cls,kw", [("boust", Boustrophedon, dict(budget=60)),
                                        ("mr_boust", MR_Boustrophedon, dict(budget=40, num_agents=2))])
def test_boustrophedon_paths(gold, tag, cls, kw):
    g = gold("runs_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21, gp_factory=FastGP)
    np.random.seed(99)
    st = cls(sc, d_waypoint_distance=2.5, **kw)
    st.run()
    assert np.array_equal(np.asarray(st.obs_wp, float), g[f"{tag}_obs_wp"])
    assert np.array_equal(np.asarray(st.full_path, float), g[f"{tag}_full_path"])
    assert np.array_equal(np.random.uniform(size=3), g[f"{tag}_rng_after"])


@pytest.mark.parametrize("tag,cls,kw", [("boust_c1", Boustrophedon, {}), ("mr_boust_8a_c4", MR_Boustrophedon, dict(num_agents=8))])
def test_boustrophedon_at_the_config_shapes(gold, tag, cls, kw):
    """The lawnmower baselines at the shapes of configs[0] (one robot, 40 x 40, the constructor's defaults: 131
    measurements) and configs[3] (8 robots, 100 x 100) against the unmodified reference (make_golden_r2.py --only boust)."""
    from tests.test_r2_cpu import DrawsOnlyGP

    g = gold("runs_golden_boust_r2.npz")
    W = int(g[f"{tag}_W"])
    sc = PointSourceField(num_sources=len(g[f"{tag}_sources"]), workspace_size=(0, W, 0, W), seed=int(g[f"{tag}_seed"]),
                          gp_factory=DrawsOnlyGP)
    assert np.array_equal(np.array(sc.sources), g[f"{tag}_sources"])
    np.random.seed(2000 + int(g[f"{tag}_seed"]))
    st = cls(sc, **kw)
    st.run()
    assert np.array_equal(np.asarray(st.obs_wp, float), g[f"{tag}_obs_wp"])
    assert np.array_equal(np.asarray(st.full_path, float), g[f"{tag}_full_path"])
    if g[f"{tag}_meas"].size:   # (the reference's MR_Boustrophedon keeps no `measurements` attribute)
        assert np.array_equal(np.asarray(st.measurements, float), g[f"{tag}_meas"])
    assert np.array_equal(np.random.uniform(size=3), g[f"{tag}_rng_after"])


@pytest.mark.parametrize("tag", ["free", "obst"])
def test_field_physics_bit_exact(gold, tag):
    g = gold("field_golden.npz")
    obst = [{"type": "rectangle", "x": 4, "y": 5, "width": 3, "height": 2}] if tag == "obst" else []
    sc = PointSourceField(num_sources=2, workspace_size=(0, 12, 0, 12), obstacles=obst, seed=8, gp_factory=FastGP)
    assert np.array_equal(np.array(sc.sources), g[f"{tag}_sources"])
    assert np.array_equal(sc.g_truth, g[f"{tag}_gtruth"])
    assert np.array_equal(sc.intensity(g[f"{tag}_wp"]), g[f"{tag}_intensity"])
    assert np.array_equal(sc.intensity(g[f"{tag}_wp"][3]), g[f"{tag}_single"])
    np.random.seed(4)
    assert np.array_equal(sc.simulate_measurements(g[f"{tag}_wp"]), g[f"{tag}_meas"])


def test_source_estimator_bit_exact(gold):
    g = gold("field_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21, gp_factory=FastGP)
    ll = E.batched_poisson_log_likelihood(g["ll_thetas"], g["est_wp"], g["est_meas"], 1, 2)
    assert np.array_equal(ll, g["ll_values"])
    np.random.seed(13)
    est, M, bic = E.estimate_sources_bayesian(list(g["est_wp"]), list(g["est_meas"]), 1, 2, 80, 5, sc, scorer=OracleScorer())
    assert M == int(g["est_M"]) and bic == float(g["est_bic"])
    assert np.array_equal(est, g["est_theta"])
    assert np.array_equal(np.random.uniform(size=3), g["est_rng_after"])


def test_choice_from_draws_is_numpy_choice():
    """The sampler draws the uniforms itself and replays RandomState.choice's index computation: same indices, same
    RNG position, for spread, peaked, zero-holding and uniform weights."""
    rng = np.random.RandomState(5)
    for case in range(40):
        n = int(rng.choice([1, 2, 7, 80, 200, 1000]))
        w = rng.rand(n) ** rng.choice([1, 8, 40])
        if case % 3 == 0:
            w[rng.rand(n) < 0.5] = 0.0
            w[rng.randint(n)] = 1.0
        if case % 7 == 0:
            w = np.full(n, 1 / n)
        w = w / w.sum()
        np.random.seed(case)
        want = np.random.choice(n, size=n, p=w)
        after = np.random.uniform(size=2)
        np.random.seed(case)
        _, got = E.choice_from_draws(w, np.random.random_sample(n))
        assert np.array_equal(want, got) and np.array_equal(after, np.random.uniform(size=2))


@pytest.mark.parametrize("scale", [1.0, 1e4, 1e7])
def test_source_estimator_exact_under_likelihood_error(gold, scale):
    """A likelihood backend that is off by as much as its reported bound allows (and far more than a device is) must
    not change one resampled index: estimate, BIC and RNG position stay the reference's; wide bands are re-decided."""
    from tests._doubles import NoisyLikelihoodScorer
    g = gold("field_golden.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 20, 0, 20), seed=21, gp_factory=FastGP)
    before = dict(E.STATS)
    np.random.seed(13)
    est, M, bic = E.estimate_sources_bayesian(list(g["est_wp"]), list(g["est_meas"]), 1, 2, 80, 5, sc,
                                              scorer=NoisyLikelihoodScorer(scale, seed=3))
    assert M == int(g["est_M"]) and bic == float(g["est_bic"])
    assert np.array_equal(est, g["est_theta"])
    assert np.array_equal(np.random.uniform(size=3), g["est_rng_after"])
    stages = E.STATS["stages"] - before["stages"]
    redone = E.STATS["redecided"] - before["redecided"]
    assert stages == 10 and (redone == 0 if scale == 1.0 else True) and (redone > 0 if scale == 1e7 else True)


@pytest.mark.parametrize("backend", ["exact", "device-like", "wide-bands"])
def test_source_estimator_at_the_stock_config_size(gold, backend):
    """estimate_sources_bayesian with the particle / stage counts of the reference's shipped example config (1000
    particles, 25 stages, M = 1..3) on its scenario, against the unmodified reference (make_golden_r2.py --only estimator):
    estimate, model order, BIC and RNG position bit-identical -- with the reference's own likelihoods, with a backend as
    wrong as a device is allowed to be, and with bands so wide that stages are re-decided on their distinct particles."""
    from tests._doubles import NoisyLikelihoodScorer
    g = gold("estimator_golden_r2.npz")
    sc = PointSourceField(num_sources=2, workspace_size=(0, 40, 0, 40), seed=95789, gp_factory=FastGP)
    sc.update_source(0, 5, 35, 1000000)
    sc.update_source(1, 35, 35, 1000000)
    assert np.array_equal(np.array(sc.sources), g["sources"])
    scorer = OracleScorer() if backend == "exact" else NoisyLikelihoodScorer(1.0 if backend == "device-like" else 1e3, seed=5)
    before = dict(E.STATS)
    np.random.seed(17)
    est, M, bic = E.estimate_sources_bayesian(list(g["wp"]), list(g["meas"]), 1, 3, 1000, 25, sc, scorer=scorer)
    assert M == int(g["M"]) and bic == float(g["bic"]) and np.array_equal(np.asarray(est, float), g["est"])
    assert np.array_equal(np.random.uniform(size=3), g["rng_after"])
    redone = E.STATS["redecided"] - before["redecided"]
    assert E.STATS["stages"] - before["stages"] == 75
    if backend == "wide-bands":
        assert redone > 0 and E.STATS["distinct_rows"] - before["distinct_rows"] < E.STATS["rows"] - before["rows"]


def test_draws_inside_band():
    cdf = np.array([0.0, 0.25, 0.25, 0.7, 1.0])
    u = np.array([0.1, 0.5, 0.9])
    idx = cdf.searchsorted(u, side="right")
    assert not E.draws_inside_band(cdf, u, idx, 1e-3)
    assert E.draws_inside_band(cdf, np.array([0.2501]), cdf.searchsorted([0.2501], side="right"), 1e-3)
    assert E.draws_inside_band(cdf, np.array([0.6995]), cdf.searchsorted([0.6995], side="right"), 1e-3)
    assert E.draws_inside_band(cdf, np.array([1e-9]), cdf.searchsorted([1e-9], side="right"), 1e-3)   # the step at 0
    assert not E.draws_inside_band(cdf, np.array([0.9999999]), cdf.searchsorted([0.9999999], side="right"), 1e-3)  # cdf[-1] is exact


def test_score_context_helpers():
    ctx = ScoreContext(np.array([[1.0, 1.0, 5.0], [4.0, 4.0, 6.0]]), np.array([3.9, 3.9]), [[np.zeros(2)], []], 1, 2.5,
                       (0, 10, 0, 10), [{"type": "rectangle", "x": 1, "y": 2, "width": 3, "height": 4}])
    assert ctx.closest_estimate_index() == 1
    assert ctx.all_obs().shape == (1, 2) and not ctx.agent_has_obs()
    assert ctx.obstacle_rows().tolist() == [[1.0, 2.0, 3.0, 4.0]]


# ---- native tree growth (csrc/ipp_tree_host.cu) vs the Python loops -----------------------
def test_native_norm_is_numpy_norm():
    """The calibrated rounding form of ipp_host_norm2 reproduces np.linalg.norm on fresh vectors."""
    from mripp_b200 import _lib
    variant = U.native_norm_variant()
    assert variant in (0, 1, 2), "no rounding form of the native norm matches np.linalg.norm on this machine"
    rng = np.random.RandomState(99)
    xy = np.ascontiguousarray(rng.uniform(-50, 50, (20000, 2)) * 10.0 ** rng.uniform(-2, 2, (20000, 1)))
    out = np.empty(len(xy))
    assert _lib.load().ipp_host_norm2(variant, xy.ctypes.data, len(xy), out.ctypes.data) == 0
    assert np.array_equal(out, np.array([np.linalg.norm(v) for v in xy]))


@pytest.mark.parametrize("gen,budgets,d", [("rig", [120, 37.5, 300], 2.5), ("rig", [2000], 1.0), ("rrt", [120, 60], 2.5),
                                           ("rrt_star", [80, 80, 40], 2.5), ("rig", [0, 5], 2.5)])
def test_native_growth_equals_python_loops(monkeypatch, gen, budgets, d):
    """Same seeded calls through both routes: coordinates, insertion-time and current parents, edge lengths, child
    counts, the rewire log, node information, the per-iteration root registrations and the RNG position are
    identical; successive calls keep growing the same tree (several blocks of pre-drawn samples at budget 2000)."""
    res = []
    for native in (False, True):
        monkeypatch.setattr(U, "NATIVE_GROWTH", native)
        sc, make = _strategy(R.MR_IPP, 40, 5, budget=120, d_waypoint_distance=d, num_agents=2, budget_iter=4)
        np.random.seed(77)
        st = make()
        st.best_estimates = np.array([[10.0, 30.0, 5e5], [28.0, 12.0, 7e5]])
        start = np.array([3.0, 4.0])
        st.agents_obs_wp = [[start.copy()], [np.array([30.0, 5.0])]]
        st.agents_full_path = [[start.copy()], [np.array([30.0, 5.0])]]
        st.initialize_trees(start, 0)
        for b in budgets:
            if gen == "rig":
                R.rig_tree_generation(st, b, 0, gain_function=R.point_source_gain_all)
            elif gen == "rrt":
                R.rrt_tree_generation(st, b, 0)
            else:
                R.rrt_star_tree_generation(st, b, 0)
        t = st.trees_soa[0]
        res.append((t.points().copy(), t.parent0[: t.n].copy(), t.parent[: t.n].copy(), t.edge[: t.n].copy(),
                    t.nchild[: t.n].copy(), list(t.rewires), t.info[: t.n].copy(), len(st.agents_trees[0].trees),
                    np.random.uniform(size=3), [t.point(i).copy() for i in range(t.n)]))
    a, b = res
    assert len(a[0]) == len(b[0]) and len(a[0]) > (1 if budgets != [0, 5] else 0)
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
    assert a[5] == b[5] and np.array_equal(a[6], b[6]) and a[7] == b[7] and np.array_equal(a[8], b[8])
    assert all(np.array_equal(p, q) for p, q in zip(a[9], b[9]))
    if gen != "rrt" and budgets[0] >= 80:
        assert len(a[5]) > 0  # the case exercises rewires


def test_main_configuration_helpers(tmp_path):
    """main.py:16-60 of the reference: configuration loading errors and constructor-argument routing by introspection."""
    from mripp_b200 import main as M

    with pytest.raises(SystemExit, match="not found"):
        M.load_configuration(str(tmp_path / "missing.json"))
    bad = tmp_path / "bad.json"
    bad.write_text("{not json")
    with pytest.raises(SystemExit, match="decoding"):
        M.load_configuration(str(bad))
    args = {"budget": 50, "num_agents": 2, "beta_t": 3.0, "line_spacing": 4, "unknown_key": 1, "stage_lambda": 0.25}
    picked = M.get_constructor_args(R.MR_IPP, args)
    assert picked["budget"] == 50 and picked["num_agents"] == 2 and picked["stage_lambda"] == 0.25 and "unknown_key" not in picked
    picked_b = M.get_constructor_args(MR_Boustrophedon, args)
    assert picked_b == {"budget": 50, "num_agents": 2, "line_spacing": 4}
    cfg = {"scenarios": [{"class": "PointSourceField", "params": {"num_sources": 2, "workspace_size": [12, 8]},
                          "specific_params": {"1": [3, 4, 500000]}}], "strategies": [], "args": {"seed": 5, "lazy_gp": True}}
    (sc,) = M.initialize_scenarios(cfg, gp_factory=FastGP)
    assert sc.workspace_size == (0, 12, 0, 8) and list(sc.sources[0]) == [3, 4, 500000]      # 2-tuple widened; source 1 pinned


# ---------------------------------------------------------------------------------------------------
# round 2, late: the host loops that bound the reference's own example config
# ---------------------------------------------------------------------------------------------------
def test_native_leaf_order_and_leaf_filter_equal_the_python_forms(monkeypatch):
    """ipp_host_leaves_dfs against the Python traversal of rrt.py:389-395 on a grown tree with rewires (the DFS runs over
    the insertion-time children lists), and the vectorised leaf filter against the reference's list comprehension."""
    np.random.seed(5)
    t = U.TreeArrays(np.array([20.0, 20.0]))
    assert t.grow_native(2, (0, 40, 0, 40), 1.0, 1.0, 700.0) is not None
    assert t.n > 256 and len(t.rewires) > 0
    fast = t.leaves_dfs()
    monkeypatch.setattr(U, "NATIVE_GROWTH", False)
    slow = t.leaves_dfs()
    assert fast == slow and all(isinstance(i, int) for i in fast)
    assert sorted(fast) == t.leaves_in_order().tolist()
    for pos in (t.point(fast[3]).copy(), np.array([1.0, 2.0]), np.array([np.nan, 2.0])):
        want = [i for i in fast if not np.array_equal(t.point(i), pos)]
        assert R._drop_leaves_at(t, fast, pos) == want
    assert len(R._drop_leaves_at(t, fast, t.point(fast[3]))) == len(fast) - 1


def test_dense_tree_growth_with_grid_and_cached_costs_equals_python_loops(monkeypatch):
    """The grid-accelerated nearest / near scans and the bounded path-cost shortcut of ipp_host_tree_grow against the
    reference's all-pairs loops on a DENSE tree (radius and step 1.0 in a 12 x 12 workspace: ~10 nodes per unit area,
    tens of near nodes per insertion, hundreds of rewires), continued over several calls."""
    res = []
    for native in (False, True):
        monkeypatch.setattr(U, "NATIVE_GROWTH", native)
        sc, make = _strategy(R.MR_IPP, 12, 5, budget=120, d_waypoint_distance=1.0, num_agents=1, budget_iter=4)
        np.random.seed(3)
        st = make()
        st.best_estimates = np.array([[3.0, 9.0, 5e5]])
        start = np.array([6.0, 6.0])
        st.agents_obs_wp = [[start.copy()]]
        st.agents_full_path = [[start.copy()]]
        st.initialize_trees(start, 0)
        for b in (400.0, 250.0):
            R.rig_tree_generation(st, b, 0, gain_function=R.point_source_gain_all)
        t = st.trees_soa[0]
        res.append((t.points().copy(), t.parent0[: t.n].copy(), t.parent[: t.n].copy(), t.edge[: t.n].copy(),
                    t.nchild[: t.n].copy(), list(t.rewires), np.random.uniform(size=3)))
    a, b = res
    assert len(a[0]) == len(b[0]) > 1200 and len(a[5]) > 300
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
    assert a[5] == b[5] and np.array_equal(a[6], b[6])


def test_host_likelihood_fast_path_keeps_scipy_bits():
    """batched_poisson_log_likelihood (blocks of particles, threads, the bare xlogy - gammaln - mu expression) against the
    reference's per-particle scipy call (estimation.py:15-55), bit for bit, including a zero count and the rate floor."""
    from scipy.stats import poisson
    from mripp_b200.src.estimation import estimation as E

    rng = np.random.RandomState(2)
    n, N, M = 700, 420, 2     # 420 x 700 >= 2^18: the threaded path
    obs = rng.uniform(0, 40, (n, 2))
    vals = rng.poisson(30.0, n).astype(float)
    vals[:3] = 0.0
    th = rng.uniform(0, 40, (N, 3 * M))
    th[:, 2::3] = rng.uniform(1e3, 1e6, (N, M))
    th[0, 2::3] = 0.0          # no signal at all: rate = lambda_b
    assert E._logpmf_direct_ok()
    got = E.batched_poisson_log_likelihood(th, obs, vals, 1.0, M)

    def reference_one(theta):   # the reference's own function body
        sources = theta.reshape(M, 3)
        d2 = (obs[:, None, 0] - sources[:, 0]) ** 2 + (obs[:, None, 1] - sources[:, 1]) ** 2
        rate = 1.0 + np.sum(sources[:, 2] / np.maximum(d2, 1e-6), axis=1)
        return np.sum(poisson.logpmf(np.round(vals).astype(int), np.maximum(rate, 1e-10)))

    for i in (0, 1, 17, 255, 256, 419):
        assert got[i] == reference_one(th[i]), i
    # a negative count falls back to scipy's own masking (-inf), not to the bare expression
    vals2 = vals.copy()
    vals2[5] = -1.0
    assert np.all(np.isneginf(E.batched_poisson_log_likelihood(th[:4], obs, vals2, 1.0, M)))


def test_redecided_stage_evaluates_distinct_rows_only():
    """A re-decided sampler stage evaluates the reference's expression once per DISTINCT particle (the stages that hit
    the band are nearly collapsed clouds): same values as the full evaluation bit for bit, same resampled indices as
    np.random.choice on the reference's weights, and a cloud of all-distinct or non-finite rows takes the full route."""
    from mripp_b200.src.estimation import estimation as E

    rng = np.random.RandomState(8)
    n, N, M = 600, 500, 3
    obs = rng.uniform(0, 40, (n, 2))
    vals = rng.poisson(50.0, n).astype(float)
    rows = rng.uniform(0, 40, (7, 3 * M))
    rows[:, 2::3] = rng.uniform(1e5, 1e6, (7, M))
    rows[3, 0] = -0.0
    parts = rows[rng.randint(0, 7, N)]
    before = dict(E.STATS)
    got = E.distinct_rows_log_likelihood(parts, obs, vals, 1.0, M)
    full = E.batched_poisson_log_likelihood(parts, obs, vals, 1.0, M)
    assert np.array_equal(got, full)
    assert E.STATS["rows"] - before["rows"] == N and E.STATS["distinct_rows"] - before["distinct_rows"] == 7
    distinct = rng.uniform(1, 39, (40, 3 * M))
    assert np.array_equal(E.distinct_rows_log_likelihood(distinct, obs, vals, 1.0, M),
                          E.batched_poisson_log_likelihood(distinct, obs, vals, 1.0, M))
    bad = parts.copy()
    bad[5, 1] = np.nan
    a, b = E.distinct_rows_log_likelihood(bad, obs, vals, 1.0, M), E.batched_poisson_log_likelihood(bad, obs, vals, 1.0, M)
    assert np.array_equal(a, b, equal_nan=True) and np.isnan(a[5])

    class LooseBound:   # device values off by 1e-7 relative with a bound the band logic must reject: forces the re-decision
        def particle_log_likelihood(self, thetas, prepared, lambda_b, M_):
            ll = E.batched_poisson_log_likelihood(thetas, obs, vals, lambda_b, M_)
            return ll * (1 + 1e-7), np.full(len(ll), 1e30), np.full(len(ll), 1e30)

    gamma = 0.4
    np.random.seed(21)
    before = dict(E.STATS)
    idx = E.resample_indices(parts, gamma, N, obs, vals, 1.0, M, LooseBound(), None)
    after_state = np.random.get_state()[2]
    assert E.STATS["redecided"] - before["redecided"] == 1 and E.STATS["distinct_rows"] - before["distinct_rows"] == 7
    np.random.seed(21)
    want = np.random.choice(N, size=N, p=E.stage_weights(full, gamma, N))   # estimation.py:117
    assert np.array_equal(idx, want) and np.random.get_state()[2] == after_state


def test_exact_close_counts_many_points_equal_the_scalar_loop():
    """count_close_exact_many (one array pass per branch, observation lists converted once) against the reference's
    double loop `np.linalg.norm(point - obs) < d` (rrt.py:246-252), with observations exactly at, one ulp inside and one
    ulp outside the threshold; the cached conversion follows a growing list."""
    from mripp_b200 import exact as E

    rng = np.random.RandomState(4)
    d = 2.5
    pts = rng.uniform(0, 30, (150, 2))
    obs_a = [rng.uniform(0, 30, 2) for _ in range(400)]
    obs_b = []
    for p in pts[:40]:   # threshold cases: a 3-4-5 triangle scaled to d, and its neighbours in floating point
        q = p + np.array([1.5, 2.0])
        obs_b += [q, np.nextafter(q, q + 1), np.nextafter(q, q - 1)]
    lists = [obs_a, [], obs_b]
    want = np.array([sum(1 for ob in lists for o in ob if np.linalg.norm(p - o) < d) for p in pts])
    assert np.array_equal(E.count_close_exact_many(pts, lists, d), want)
    assert E.count_close_exact(pts[7], lists, d) == want[7]
    obs_a.append(pts[3] + 0.1)           # the list grows: the cached array must not be reused
    assert E.count_close_exact(pts[3], lists, d) == want[3] + 1


def test_rewire_log_behaves_like_the_list_of_tuples():
    """RewireLog (array-backed rewire events of natively grown trees) against the plain list it replaces: length,
    iteration, indexing, equality in both directions, mixed Python appends and native blocks in time order, and the
    scorer's CSR (history_csr) from either form."""
    from mripp_b200.src.rrt.rrt_utils import RewireLog, TreeArrays

    rng = np.random.RandomState(4)
    ev = [(int(t), int(rng.randint(1, 40)), int(t)) for t in sorted(rng.randint(2, 60, 50))]
    log = RewireLog()
    for e in ev[:7]:
        log.append(e)
    log.extend_array(np.array(ev[7:30], dtype=np.int64))
    log.append(np.array(ev[30]))            # numpy integers are stored as Python ints
    log.extend_array(np.zeros((0, 3), dtype=np.int64))
    log.extend_array(np.array(ev[31:], dtype=np.int64))
    assert len(log) == 50 and bool(log) and not RewireLog()
    assert list(log) == ev and log == ev and ev == log and log == RewireLog(ev) and log != ev[:-1]
    assert log[3] == ev[3] and log[-1] == ev[-1] and log[5:9] == ev[5:9]
    assert all(type(v) is int for e in log for v in e)
    assert log.as_array().shape == (50, 3) and RewireLog().as_array().shape == (0, 3)
    a, b = TreeArrays(np.zeros(2)), TreeArrays(np.zeros(2))
    for k in range(1, 60):
        for t in (a, b):
            t.add(np.array([float(k), 0.5 * k]), int(rng.randint(0, k)) if t is a else int(a.parent0[k]))
    a.rewires, b.rewires = log, list(ev)    # the list form (tests assign one) still works
    for x, y in zip(a.history_csr(), b.history_csr()):
        assert x.dtype == y.dtype == np.int32 and np.array_equal(x, y)
    assert [p.tolist() for p in a.chain_at(41)] == [p.tolist() for p in b.chain_at(41)]


def test_point_table_chunking_policy(monkeypatch):
    """Chunk size of the tabled explicit-point posterior: a multiple of 128, never below one tile, the whole set when it
    fits, whole waves of one tile per SM when it does not, and a scratch that is already held is always reused."""
    import types
    import torch
    from mripp_b200 import _lib
    from mripp_b200 import gp as G

    NP = 5056
    for m in (1, 127, 128, 129, 256, 37888, 10 ** 6):
        for free in (0, 1 << 20, 1 << 30, 40 << 30, 170 << 30):
            for have in (0, 130 * NP, 20000 * NP + NP + 4 * 200):
                rows = G.point_table_rows(m, NP, free, have)
                assert rows >= 128 and rows % 128 == 0
                scratch = rows * NP + NP + 4 * (rows // 128)
                assert rows <= max(128, (m + 127) // 128 * 128)
                if rows > 128:   # beyond the one-tile floor the scratch respects the budget or the buffer already held
                    assert 8 * rows * NP <= max(min(4 << 30, free // 4), 8 * have) and (scratch <= have or 8 * rows * NP <= free // 4)
                if rows < (m + 127) // 128 * 128 and rows >= 128 * 148:
                    assert rows % (128 * 148) == 0
    assert G.point_table_rows(37888, NP, 170 << 30) == 37888
    assert G.point_table_rows(10 ** 6, NP, 170 << 30) == 5 * 148 * 128
    assert G.point_table_rows(10 ** 6, NP, 170 << 30, sms=132) == (4 << 30) // (8 * NP) // (132 * 128) * 132 * 128
    # the method asks the device for the two numbers and nothing else
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (8 << 30, 180 << 30))
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))
    gp = G.GaussianProcessRegressor.__new__(G.GaussianProcessRegressor)
    gp._ws = types.SimpleNamespace(layout=(NP,), tables=None)
    assert gp._point_table_rows(10 ** 6) == G.point_table_rows(10 ** 6, NP, 8 << 30, 0, 148) == 2 * 148 * 128
    gp._ws.tables = torch.empty(300 * NP)
    assert gp._point_table_rows(200) == 256


def test_device_morton_order_equals_the_host_order():
    """The Z-order permutation of explicit query points computed with tensor ops (predict_points_device) is the host's
    morton_order -- same 16-bit keys, same stable sort -- including duplicate points and a degenerate axis."""
    import torch
    from mripp_b200 import gp as G

    rng = np.random.RandomState(6)
    for Q in (rng.uniform(-3, 47, (5000, 2)), np.repeat(rng.uniform(0, 10, (40, 2)), 7, axis=0),
              np.column_stack([rng.uniform(0, 1, 300), np.full(300, 2.5)]), np.zeros((5, 2))):
        got = G.GaussianProcessRegressor._morton_order_device(torch.from_numpy(Q)).numpy()
        assert np.array_equal(got, G.morton_order(Q))


def test_reported_metrics_are_the_references(gold):
    """What main.py reports per run -- RMSE and source-weighted RMSE of log10(Z + 1) (main.py:113-114), the differential
    entropy of std (plot_helper.py:383-398, zeros replaced in place) and calculate_source_errors
    (path_planning_utils.py:125-141) -- against the reference's own expressions and functions, bit for bit."""
    from mripp_b200.src.utils.path_planning_utils import calculate_source_errors
    from mripp_b200.src.visualization.plot_helper import calculate_differential_entropy

    g = gold("metrics_golden_r2.npz")
    sc = PointSourceField(num_sources=3, workspace_size=(0, 30, 0, 30), seed=int(g["seed"]), gp_factory=FastGP,
                          ground_truth_mode="host")
    assert np.array_equal(np.array(sc.sources), g["sources"])
    Z_true, Z_pred = sc.ground_truth(), g["Z_pred"]
    err = np.log10(Z_true + 1) - np.log10(Z_pred + 1)          # the host branch of mripp_b200.main
    assert np.sqrt(np.mean(err ** 2)) == float(g["rmse"])
    assert np.sqrt(np.sum(Z_true * err ** 2) / np.sum(Z_true)) == float(g["wrmse"])
    std = g["std"].copy()
    assert calculate_differential_entropy(std) == float(g["entropy"]) and not np.any(std == 0)
    errs = calculate_source_errors(sc.sources, g["est"])
    for k in ("x_error", "y_error", "intensity_error"):
        assert np.array_equal(np.array(errs[k]), g[k])
    assert calculate_source_errors(sc.sources, np.zeros((0, 3))) == {"x_error": [], "y_error": [], "intensity_error": []}

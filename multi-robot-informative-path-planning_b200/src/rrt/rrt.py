"""RRT informative path planning strategies: drop-in for the reference's src/rrt/rrt.py.

Same public names, constructor keywords, `run() -> (Z_pred, std)` contract and attributes
 -- but the per-node Python gain functions, the selector penalty loops and the GP
are batched onto the B200:

  reference (per inserted node, Python)                      here (per tree, one launch each)
  -------------------------------------------------------    ------------------------------------
  gain_function(self, node, i) walks node -> root and        tree grown on the host (sequential, RNG
  recomputes every per-node term, incl. an O(#obs) scan      ordered), then ipp_node_terms +
  (rrt.py:51-259, 321)                                       ipp_count_close + ipp_branch_gain score
                                                             every node against versioned parents
  bias_beta_path_selection: L x (L x L gram + det)           ipp_mi_logdet once + ipp_leaf_penalty
  (rrt.py:438-510)
  predict_spatial_field: sklearn fit + predict               mripp_b200.gp (fused LML / posterior)

Host code keeps everything sequential or RNG-consuming, in the reference's order (SURVEY.md
Appendix B), so tree indexing and waypoint sequences are identical for a fixed seed.  Agents run
one after another inside each outer iteration (the deterministic interleaving of the reference's
per-agent threads, rrt.py:699-711).
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Tuple

import numpy as np
from scipy.spatial import distance

from mripp_b200.exact import exact_branch_gain, exact_leaf_scores, integer_scores, near_tie_candidates
from mripp_b200.scoring import (VARIANT_ALL, VARIANT_DISTANCE, VARIANT_DISTANCE_ROTATION, VARIANT_NO_PENALTIES,
                                ScoreContext)
from mripp_b200.src.estimation.estimation import estimate_sources_bayesian
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt.rrt_utils import (InformativeTreeNode, TreeArrays, TreeCollection, TreeNode, add_node,
                                          choose_parent, cost, line_cost, near, nearest, node_selection_key_distance,
                                          obstacle_free, rewire, steer, steer_point, trace_path_to_root)

_norm = np.linalg.norm


# ---- scoring helpers --------------------------------------------------------------------------------
def _context(self, agent_idx: int) -> ScoreContext:
    path = self.agents_full_path[agent_idx]
    cache = self.__dict__.setdefault("_obs_array_cache", {})  # observation lists only ever grow (rrt.py:727-737)
    return ScoreContext(best_estimates=getattr(self, "best_estimates", np.array([])),
                        last_point=path[-1] if len(path) else None, agents_obs_wp=self.agents_obs_wp, agent_idx=agent_idx,
                        d_wp=self.d_waypoint_distance, workspace=self.scenario.workspace_size,
                        obstacles=list(self.scenario.obstacles), obs_cache=cache)


def _chain_tree(node: TreeNode) -> TreeArrays:
    """A TreeArrays holding just the root->node chain of a reference-shaped node object."""
    chain = []
    cur = node
    while cur is not None:
        chain.append(cur.point)
        cur = cur.parent
    chain.reverse()
    tree = TreeArrays(chain[0])
    for k in range(1, len(chain)):
        tree.add(chain[k], k - 1)
    return tree


def _single_gain(self, node: TreeNode, agent_idx: int, variant: int) -> float:
    tree = _chain_tree(node)
    return float(self.scorer.tree_information(variant, tree, _context(self, agent_idx))[tree.n - 1])


def calculate_suppression_factor(node_point: np.ndarray, source: np.ndarray, other_sources: List[np.ndarray]) -> float:
    """rrt.py:32-47 (host scalar form; the batched form is ipp_node_terms)."""
    xt, yt = node_point
    xk, yk, _ = source
    total = 0
    for other in other_sources:
        if np.array_equal(source, other):
            continue
        mx, my = (xk + other[0]) / 2, (yk + other[1]) / 2
        d = _norm([xt - mx, yt - my])
        c_dist = 2 - 1 / (1 + np.exp((d - 2) / 16))
        total += (1 - c_dist + c_dist / (1 + np.exp((2 - d) / 16)))
    return total


def point_source_gain_no_penalties(self, node: InformativeTreeNode, agent_idx: int) -> float:
    """rrt.py:51-70."""
    return _single_gain(self, node, agent_idx, VARIANT_NO_PENALTIES)


def point_source_gain_only_distance_penalty(self, node: InformativeTreeNode, agent_idx: int) -> float:
    """rrt.py:72-96."""
    return _single_gain(self, node, agent_idx, VARIANT_DISTANCE)


def point_source_gain_distance_rotation_penalty(self, node: InformativeTreeNode, agent_idx: int) -> float:
    """rrt.py:98-128."""
    return _single_gain(self, node, agent_idx, VARIANT_DISTANCE_ROTATION)


def point_source_gain_all(self, node: InformativeTreeNode, agent_idx: int) -> float:
    """rrt.py:130-259."""
    return _single_gain(self, node, agent_idx, VARIANT_ALL)


point_source_gain_no_penalties.variant = VARIANT_NO_PENALTIES
point_source_gain_only_distance_penalty.variant = VARIANT_DISTANCE
point_source_gain_distance_rotation_penalty.variant = VARIANT_DISTANCE_ROTATION
point_source_gain_all.variant = VARIANT_ALL


def _score_tree(self, agent_idx: int, gain_function: Optional[Callable]) -> None:
    """node.information for every node of the agent's tree, as of each node's insertion time."""
    tree = self.trees_soa[agent_idx]
    variant = getattr(gain_function, "variant", None)
    tree.variant = variant
    if gain_function is None:
        return
    # The reference computes node.information once, when the node is inserted, and never refreshes it: a tree that
    # survives into another tree_generation call (the selector returned an empty path, so initialize_trees was
    # skipped) keeps the values its old nodes got under the estimates / observations of that time.  Only the nodes
    # added since the last scoring pass are written.
    lo = max(1, int(getattr(tree, "scored_upto", 1)))
    if lo >= tree.n:
        return
    if variant is not None:
        tree.info[lo: tree.n] = self.scorer.tree_information(variant, tree, _context(self, agent_idx))[lo:]
    else:  # user-supplied callable: evaluated per node on reference-shaped objects (final parents)
        nodes = tree.materialise()
        for i in range(lo, tree.n):
            tree.info[i] = gain_function(self, nodes[i], agent_idx)
    tree.scored_upto = tree.n


def _sample(self) -> np.ndarray:
    ws = self.scenario.workspace_size
    return np.random.rand(2) * (ws[1] - ws[0], ws[3] - ws[2]) + (ws[0], ws[2])


# ---- tree generation (rrt.py:263-327) ---------------------------------------------------------------
def _no_penalties_gain_error(self, agent_idx: int, gain_function: Optional[Callable]) -> None:
    """The reference's error, kept where it occurs: point_source_gain_no_penalties measures the estimates from
    agents_full_path[agent_idx][-1] (rrt.py:57); an agent that has not moved yet has no such point once another agent's
    measurements produced estimates, and the first node its generator inserts -- one sample into the loop
    (rrt.py:266,274 / :310,321) -- dies with this IndexError."""
    if (getattr(gain_function, "variant", None) == VARIANT_NO_PENALTIES and not self.agents_full_path[agent_idx]
            and np.size(getattr(self, "best_estimates", ())) > 0):
        _sample(self)
        raise IndexError("list index out of range")


def rrt_tree_generation(self, budget_portion: float, agent_idx: int) -> None:
    tree = self.trees_soa[agent_idx]
    travelled = 0
    if travelled < budget_portion:
        _no_penalties_gain_error(self, agent_idx, point_source_gain_no_penalties)
    samples = tree.grow_native(0, self.scenario.workspace_size, self.d_waypoint_distance, 0.0, budget_portion)
    if samples is not None:             # grown natively (csrc/ipp_tree_host.cu): same nodes, same RNG position
        travelled = budget_portion
        # the reference's loop registers the root once per non-final iteration
        self.agents_trees[agent_idx].add_repeated(self.agents_roots[agent_idx], samples - 1)
    while travelled < budget_portion:
        target = _sample(self)
        nn = tree.nearest(target)
        point = steer_point(tree.point(nn), target, self.d_waypoint_distance)
        new = tree.add(point, nn)
        travelled += _norm(tree.point(new) - tree.point(nn))
        if travelled >= budget_portion:
            break
        self.agents_trees[agent_idx].add(self.agents_roots[agent_idx])
    _score_tree(self, agent_idx, point_source_gain_no_penalties)
    self._invalidate_nodes(agent_idx)


def rrt_star_tree_generation(self, budget_portion: float, agent_idx: int) -> None:
    tree = self.trees_soa[agent_idx]
    travelled = 0
    if tree.grow_native(1, self.scenario.workspace_size, 1.0, self.d_waypoint_distance, budget_portion) is not None:
        travelled = budget_portion
    while travelled < budget_portion:
        target = _sample(self)
        nn = tree.nearest(target)
        point = steer_point(tree.point(nn), target, 1.0)  # the literal step of rrt.py:285
        nb = tree.near(point, self.d_waypoint_distance)
        best, best_cost = nn, tree.path_cost(nn) + _norm(tree.point(nn) - point)
        for q in nb:
            c = tree.path_cost(q) + _norm(tree.point(q) - point)
            if c < best_cost:
                best, best_cost = q, tree.path_cost(q) + _norm(tree.point(q) - point)
        new = tree.add(point, best)
        travelled += _norm(tree.point(new) - tree.point(best))
        tree.rewire(nb, new)
        if travelled >= budget_portion:
            break
    self.agents_trees[agent_idx].add(self.agents_roots[agent_idx])
    self._invalidate_nodes(agent_idx)


def rig_tree_generation(self, budget_portion: float, agent_idx: int, gain_function: Callable) -> None:
    tree = self.trees_soa[agent_idx]
    travelled = 0
    if travelled < budget_portion:
        _no_penalties_gain_error(self, agent_idx, gain_function)
    samples = tree.grow_native(2, self.scenario.workspace_size, self.d_waypoint_distance, self.d_waypoint_distance,
                               budget_portion)
    if samples is not None:
        travelled = budget_portion
        self.agents_trees[agent_idx].add_repeated(self.agents_roots[agent_idx], samples - 1)
    while travelled < budget_portion:
        target = _sample(self)
        nn = tree.nearest(target)
        point = steer_point(tree.point(nn), target, self.d_waypoint_distance)
        nb = tree.near(point, self.d_waypoint_distance)
        new = tree.add(point, nn)  # add_node overrides choose_parent: parent = nearest (App. D.3)
        travelled += _norm(tree.point(new) - tree.point(nn))
        tree.rewire(nb, new)       # information is evaluated against pre-rewire parents: see _score_tree
        if travelled >= budget_portion:
            break
        self.agents_trees[agent_idx].add(self.agents_roots[agent_idx])
    _score_tree(self, agent_idx, gain_function)
    self._invalidate_nodes(agent_idx)


# ---- information updates (rrt.py:331-369) ---------------------------------------------------------------
def _gather(self) -> None:
    self.measurements = sum((m for m in self.agents_measurements), [])
    self.obs_wp = sum((w for w in self.agents_obs_wp), [])
    self.full_path = sum((p for p in self.agents_full_path), [])


def gp_information_update(self) -> None:
    _gather(self)
    if len(self.measurements) > 0 and len(self.measurements) % 5 == 0:
        k = min(len(self.measurements), len(self.obs_wp))
        self.measurements = self.measurements[:k]
        self.obs_wp = self.obs_wp[:k]
        self.scenario.update(self.obs_wp, np.log10(np.maximum(self.measurements, 1e-6)))
        if hasattr(self, "best_estimates") and self.best_estimates.size > 0:
            estimates, _, bic = estimate_sources_bayesian(self.obs_wp, self.measurements, self.lambda_b, self.max_sources,
                                                          self.n_samples, self.s_stages, self.scenario, scorer=self.scorer)
            if bic > self.best_bic:
                self.best_bic = bic
                self.best_estimates = estimates.reshape((-1, 3))


def source_metric_information_update(self) -> None:
    _gather(self)
    if len(self.measurements) > 0 and len(self.measurements) % 5 == 0:
        k = min(len(self.measurements), len(self.obs_wp))
        self.measurements = self.measurements[:k]
        self.obs_wp = self.obs_wp[:k]
        self.scenario.update(self.obs_wp, np.log10(np.maximum(self.measurements, 1e-6)))
        estimates, _, bic = estimate_sources_bayesian(self.obs_wp, self.measurements, self.lambda_b, self.max_sources,
                                                      self.n_samples, self.s_stages, self.scenario, scorer=self.scorer)
        self.best_bic = bic
        self.best_estimates = estimates.reshape((-1, 3))


# ---- path selection (rrt.py:373-526) --------------------------------------------------------------------
def _truncate(self, possible_path, current_position, current_budget):
    """rrt.py:426-436 / 500-510 (the running budget never decreases: App. D.6)."""
    if current_budget is None or current_position is None:
        return possible_path
    path = []
    for p in possible_path:
        if current_budget - _norm(p - current_position) < self.d_waypoint_distance:
            path.append(p)
            break
        path.append(p)
        current_budget -= _norm(p - current_position) if not path else _norm(p - path[-1])
    return path


def random_path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None) -> List[np.ndarray]:
    """rrt.py:373-378: one np.random.choice over the leaves in insertion order."""
    tree = self.trees_soa[agent_idx]
    leaves = tree.leaves_in_order()
    if leaves.size == 0:
        return []
    k = np.random.choice(leaves.size)
    return tree.path_to_root(int(leaves[k]))


def _drop_leaves_at(tree, leaves: List[int], position) -> List[int]:
    """[i for i in leaves if not np.array_equal(point(i), position)] of rrt.py:398-399 / 420 as one comparison (two
    coordinates equal; a NaN coordinate is never equal, as in array_equal)."""
    position = np.asarray(position)
    if position.shape != (2,) or len(leaves) == 0:
        return [i for i in leaves if not np.array_equal(tree.point(i), position)]
    idx = np.asarray(leaves, dtype=np.int64)
    pts = tree.points()[idx]
    keep = ~((pts[:, 0] == position[0]) & (pts[:, 1] == position[1]))
    return idx[keep].tolist()


def informative_source_metric_path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None,
                                             current_budget: Optional[float] = None) -> List[np.ndarray]:
    """rrt.py:380-436: argmax of node.information over the DFS-ordered leaves (first maximum wins)."""
    if not self.agents_roots[agent_idx]:
        return []
    tree = self.trees_soa[agent_idx]
    leaves = tree.leaves_dfs()
    if not leaves:
        return []
    if current_position is not None:
        leaves = _drop_leaves_at(tree, leaves, current_position)
        if not leaves:
            return [current_position]
    if self.best_estimates.size == 0:
        pick = leaves[np.random.choice(len(leaves))]
        while tree.info[pick] == -np.inf:
            pick = leaves[np.random.choice(len(leaves))]
    else:
        vals = tree.info[leaves]
        pick = leaves[int(np.argmax(vals))]  # max(..., key=information): the first maximum
        cand = near_tie_candidates(vals)
        if cand is not None and getattr(tree, "variant", None) is not None:  # see mripp_b200.exact
            ctx = _context(self, agent_idx)
            exact = [exact_branch_gain(tree.variant, ctx, tree.chain_at(leaves[k])) for k in cand]
            pick = leaves[int(cand[int(np.argmax(exact))])]
    possible_path = tree.path_to_root(pick)
    if len(possible_path) == 1 and np.array_equal(possible_path[0], current_position):
        leaves = _drop_leaves_at(tree, leaves, tree.point(pick))
        if leaves:
            pick = leaves[np.random.choice(len(leaves))]
            possible_path = tree.path_to_root(pick)
    return _truncate(self, possible_path, current_position, current_budget)


def mutual_information_gain(K_pi, K_pio, beta_t: float) -> float:
    """rrt.py:512-526 for callers that already hold host matrices (K_pi enters through its shape)."""
    K_pio = np.asarray(K_pio, dtype=float)
    if K_pio.shape[1] == 0:
        return 0
    modified = np.eye(np.asarray(K_pi).shape[0]) + beta_t * np.dot(K_pio, K_pio.T)
    return 0.5 * np.log(np.linalg.det(modified))


def bias_beta_path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None,
                             current_budget: Optional[float] = None) -> List[np.ndarray]:
    """rrt.py:438-510.  One device evaluation of the (leaf-independent) mutual information with the
    PRIOR kernel (rrt.py:447,452 use gp.kernel, not kernel_), one penalty launch, host argmax with
    numpy's tie / NaN rules."""
    tree = self.trees_soa[agent_idx]
    leaves = tree.leaves_in_order()
    leaf_xy = np.ascontiguousarray(tree.points()[leaves])
    par = tree.parent[leaves]
    has_parent = par >= 0
    parent_xy = np.ascontiguousarray(tree.points()[np.where(has_parent, par, 0)])
    kern = self.scenario.gp.kernel
    ctx = _context(self, agent_idx)
    scores, mi = self.scorer.leaf_scores(leaf_xy, parent_xy, has_parent, ctx, self.beta_t, kern.constant_value,
                                         kern.length_scale)
    if not ctx.agent_has_obs():
        # mutual_information_gain returned the int 0: the reference's scores are int64 (mripp_b200.exact)
        best = int(np.argmax(integer_scores(scores)))
    else:
        best = int(np.argmax(scores))
        cand = near_tie_candidates(scores)
        if cand is not None:
            pts = [tree.point(int(i)) for i in leaves]
            pars = [tree.point(int(q)) if q >= 0 else None for q in par]
            exact = exact_leaf_scores(cand, pts, pars, ctx, self.beta_t, kern.constant_value, kern.length_scale, mi_dev=mi)
            best = int(cand[int(np.argmax(exact))])
    selected = int(leaves[best])
    possible_path = tree.path_to_root(selected)
    if len(possible_path) == 1 and np.array_equal(possible_path[0], current_position):
        # rrt.py:493-497 reads two names it never assigned
        raise UnboundLocalError("cannot access local variable 'current_leafs' where it is not associated with a value")
    return _truncate(self, possible_path, current_position, current_budget)


def ucb_path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None,
                       current_budget: Optional[float] = None) -> List[np.ndarray]:
    """The posterior-driven selector of the reference's alternative module (rrt_thread.py:371-401)."""
    tree = self.trees_soa[agent_idx]
    leaves = tree.leaves_in_order()
    leaf_xy = np.ascontiguousarray(tree.points()[leaves])
    acq, mean_std = self.scorer.ucb_acquisition(self.scenario.gp, leaf_xy, self.agents_full_path[agent_idx], self.beta_t,
                                                self.d_waypoint_distance)
    self.uncertainty_reduction.append(mean_std)
    node = int(leaves[int(np.argmax(acq))])
    path = []
    while node >= 0 and current_budget is not None and current_budget > 0:
        path.append(tree.point(node))
        if len(path) > 1:
            current_budget -= _norm(path[-1] - path[-2])
        if current_budget <= self.d_waypoint_distance:
            break
        node = int(tree.parent[node])
    path.reverse()
    return path


# ---- strategies -----------------------------------------------------------------------------------------
class _QuietBar:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):
        pass


class _NodeLists:
    """`strategy.tree_nodes[i]`: reference-shaped node objects, built from the arrays on demand."""

    def __init__(self, owner):
        self._owner = owner

    def __getitem__(self, i):
        return self._owner._nodes(i)

    def __len__(self):
        return self._owner.num_agents

    def __iter__(self):
        return (self._owner._nodes(i) for i in range(self._owner.num_agents))


class InformativeRRTBaseClass:
    """rrt.py:531-799."""

    def __init__(self, scenario: PointSourceField, beta_t: float = 5.0, budget: float = 375, d_waypoint_distance: float = 2.5,
                 num_agents: int = 1, n_samples: int = 25, s_stages: int = 10, lambda_b: float = 1, max_sources: int = 3,
                 budget_iter: int = 10, stage_lambda: float = 0.875, scorer=None, progress: bool = False, **kwargs):
        self.scenario = scenario
        self.beta_t = beta_t
        self.budget = [budget] * num_agents
        self.d_waypoint_distance = d_waypoint_distance
        self.num_agents = num_agents
        self.budget_iter = budget_iter
        self.name = None
        self.trees = TreeCollection()
        self.uncertainty_reduction = []
        self.time_taken = 0
        self.new_scenario = None
        self.Z_pred = None
        self.best_bic = -np.inf
        self.n_samples = n_samples
        self.s_stages = s_stages
        self.lambda_b = lambda_b
        self.max_sources = max_sources
        self.assignments = [-1] * num_agents
        self.stage_lambda = stage_lambda
        # The reference leaves best_estimates unset on the three GP-named classes, which then die in
        # run() (rrt.py:567); an empty estimate set is what makes them runnable.
        self.best_estimates = np.array([])
        ws = scenario.workspace_size
        if num_agents > 1:
            self.agent_positions = [np.array([ws[0] + 0.5 + i * (ws[1] - 1) / (num_agents - 1), ws[2] + 0.5])
                                    for i in range(num_agents)]
        else:
            self.agent_positions = [np.array([0.5, 0.5])]
        self.agents_trees = [TreeCollection() for _ in range(num_agents)]
        self.agents_obs_wp = [[] for _ in range(num_agents)]
        self.agents_measurements = [[] for _ in range(num_agents)]
        self.agents_full_path = [[] for _ in range(num_agents)]
        self.trees_soa: List[Optional[TreeArrays]] = [None] * num_agents
        self._node_cache: List[Optional[list]] = [None] * num_agents
        self.tree_nodes = _NodeLists(self)
        self.agents_roots = [None] * num_agents
        self._scorer = scorer
        self._progress = progress

    # -- device scoring backend (constructed lazily; raises without CUDA) ------------------------------
    @property
    def scorer(self):
        if self._scorer is None:
            from mripp_b200.scoring import GpuScorer

            self._scorer = GpuScorer()
        return self._scorer

    def _nodes(self, agent_idx):
        if self._node_cache[agent_idx] is None:
            tree = self.trees_soa[agent_idx]
            self._node_cache[agent_idx] = tree.materialise() if tree is not None else []
        return self._node_cache[agent_idx]

    def _invalidate_nodes(self, agent_idx):
        self._node_cache[agent_idx] = None

    # -- reference methods ---------------------------------------------------------------------------
    def assign_agents_to_sources(self):
        """rrt.py:566-591 (its result is not read by any selector of this module: App. D.9)."""
        if self.best_estimates.size == 0:
            return
        src = [s[:2] for s in self.best_estimates]
        if len(src) >= len(self.agent_positions):
            self.assignments = np.argmin(distance.cdist(self.agent_positions, src), axis=1)
        else:
            order = np.argsort(distance.cdist(self.agent_positions, src), axis=1)
            taken = set()
            self.assignments = [-1] * len(self.agent_positions)
            for a in range(len(self.agent_positions)):
                for s in order[a]:
                    if s not in taken:
                        self.assignments[a] = s
                        taken.add(s)
                        break

    def initialize_trees(self, start_position: np.ndarray, agent_idx: int) -> None:
        self.trees_soa[agent_idx] = TreeArrays(start_position)
        self._invalidate_nodes(agent_idx)
        self.agents_roots[agent_idx] = InformativeTreeNode(start_position)
        self.trees.add(self.agents_roots[agent_idx])

    def plot_current_state(self, iteration: int, save_dir: str) -> None:
        return None  # plotting is outside the accelerated path (and its call is commented out, rrt.py:714)

    def agent_thread(self, agent_idx: int, budget_portion: float, max_budget: float) -> None:
        """rrt.py:655-676."""
        current_budget = budget_portion
        while current_budget > self.d_waypoint_distance and self.budget[agent_idx] > 0:
            self.tree_generation(max_budget, agent_idx)  # the WHOLE budget, not the portion (App. D.1)
            budget_factor = self.budget[agent_idx] / max_budget
            here = self.agents_full_path[agent_idx][-1] if self.agents_full_path[agent_idx] else None
            if budget_factor > self.stage_lambda:
                path = bias_beta_path_selection(self, agent_idx, here, current_budget)
            else:
                path = self.path_selection(agent_idx, here, current_budget)
            spent = self.calculate_budget_spent(path)
            self.update_observations_and_model(path, agent_idx)
            self.agents_full_path[agent_idx].extend(path)
            self.budget[agent_idx] -= spent
            current_budget -= spent
            if len(self.agents_obs_wp[agent_idx % self.num_agents]) > 0:
                self.information_update() if budget_factor <= self.stage_lambda else gp_information_update(self)
            if len(path) > 0 and self.budget[agent_idx] > 0:
                self.initialize_trees(path[-1], agent_idx)

    def run(self) -> Tuple[np.ndarray, np.ndarray]:
        """rrt.py:678-721."""
        portion = [b / self.budget_iter for b in self.budget]
        max_budget = max(self.budget)
        for i in range(self.num_agents):
            self.initialize_trees(self.agent_positions[i], i)
        start = time.time()
        bar = _QuietBar()
        if self._progress:
            from tqdm import tqdm

            bar = tqdm(total=sum(self.budget), desc="Running " + str(self.num_agents) + " Agent " + str(self.name))
        with bar as pbar:
            while any(b > 0 for b in self.budget):
                self.assign_agents_to_sources()
                for i in range(self.num_agents):
                    if self.budget[i] > 0:
                        self.agent_thread(i, portion[i], max_budget)
                pbar.update(sum(portion))
        self.time_taken = time.time() - start
        return self.finalize_all_agents()

    def update_observations_and_model(self, path: List[np.ndarray], agent_idx: int) -> None:
        """rrt.py:722-749: one simulated measurement (one normal draw) per path point, then the field."""
        if not path:
            return
        new_meas, new_wp = [], []
        for p in path:
            new_meas.append(self.scenario.simulate_measurements([p])[0])
            new_wp.append(p)
        self.agents_measurements[agent_idx].extend(new_meas)
        self.agents_obs_wp[agent_idx].extend(new_wp)
        _gather(self)
        self.obs_wp = np.array(self.obs_wp)
        self.full_path = np.array(self.full_path).reshape(-1, 2).T
        self.Z_pred, _ = self.scenario.predict_spatial_field(self.obs_wp, np.array(self.measurements), transient=True)

    def calculate_budget_spent(self, path: List[np.ndarray]) -> float:
        if not path:
            return 0
        arr = np.array(path)
        if arr.ndim > 1:
            return np.sum(_norm(np.diff(arr, axis=0), axis=1))
        return 0

    def finalize_all_agents(self) -> Tuple[np.ndarray, np.ndarray]:
        _gather(self)
        self.obs_wp = np.array(self.obs_wp)
        self.full_path = np.array(self.full_path).reshape(-1, 2).T
        self.Z_pred, std = self.scenario.predict_spatial_field(self.obs_wp, np.array(self.measurements))
        return self.Z_pred, std

    def tree_generation(self, budget_portion: float, agent_idx: int) -> None:
        raise NotImplementedError("tree_generation method must be implemented in the subclass")

    def path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None) -> List[np.ndarray]:
        raise NotImplementedError("path_selection method must be implemented in the subclass")

    def information_update(self) -> None:
        raise NotImplementedError("information_update method must be implemented in the subclass")

    def get_current_node(self, agent_idx: int):
        nodes = self.tree_nodes[agent_idx]
        return nodes[-1] if nodes else None

    def get_chosen_branch(self, agent_idx: int):
        node = self.get_current_node(agent_idx)
        branch = []
        while node is not None:
            branch.append(node.point)
            node = node.parent
        branch.reverse()
        return branch


class RRT_Random_GP_PathPlanning(InformativeRRTBaseClass):
    """rrt.py:801-813 ("Naive RRT")."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.name = "RRT_Random_GP_Path"

    def tree_generation(self, budget_portion: float, agent_idx: int) -> None:
        rrt_tree_generation(self, budget_portion, agent_idx)

    def path_selection(self, agent_idx: int, *_unused) -> List[np.ndarray]:
        # the reference declares one argument but is called with three (rrt.py:664,809): accept them
        return random_path_selection(self, agent_idx)

    def information_update(self) -> None:
        gp_information_update(self)


class RRTStar_Random_GP_PathPlanning(InformativeRRTBaseClass):
    """rrt.py:815-827."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.name = "RRTStar_Random_GP_Path"

    def tree_generation(self, budget_portion: float, agent_idx: int) -> None:
        rrt_star_tree_generation(self, budget_portion, agent_idx)

    def path_selection(self, agent_idx: int, *_unused) -> List[np.ndarray]:
        # the reference declares one argument but is called with three (rrt.py:664,809): accept them
        return random_path_selection(self, agent_idx)

    def information_update(self) -> None:
        gp_information_update(self)


class _SourceMetricStrategy(InformativeRRTBaseClass):
    gain_function: Callable = None
    label: str = ""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.best_estimates = np.array([])
        self.name = self.label

    def tree_generation(self, budget_portion: float, agent_idx: int) -> None:
        rig_tree_generation(self, budget_portion, agent_idx, gain_function=type(self).gain_function)

    def path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None,
                       current_budget: Optional[float] = None) -> List[np.ndarray]:
        return informative_source_metric_path_selection(self, agent_idx, current_position, current_budget)

    def information_update(self) -> None:
        source_metric_information_update(self)


class RRTRIG_PointSourceInformative_SourceMetric_PathPlanning(_SourceMetricStrategy):
    """rrt.py:829-843 ("Strategic RRT")."""
    gain_function = staticmethod(point_source_gain_no_penalties)
    label = "RRTRIG_PointSourceInformative_SourceMetric_Path"


class RRTRIG_PointSourceInformative_Distance_SourceMetric_PathPlanning(_SourceMetricStrategy):
    """rrt.py:845-859."""
    gain_function = staticmethod(point_source_gain_only_distance_penalty)
    label = "RRTRIG_PointSourceInformative_Distance_SourceMetric_Path"


class RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_PathPlanning(_SourceMetricStrategy):
    """rrt.py:861-875."""
    gain_function = staticmethod(point_source_gain_distance_rotation_penalty)
    label = "RRTRIG_PointSourceInformative_DistanceRotation_SourceMetric_Path"


class MR_IPP(_SourceMetricStrategy):
    """rrt.py:877-891 ("Adaptive RRT": two-stage lambda switch + every penalty)."""
    gain_function = staticmethod(point_source_gain_all)
    label = "MR_IPP"


# the name the shipped config and rrt_thread.py:731 use for the same strategy
RRTRIG_PointSourceInformative_All_SourceMetric_PathPlanning = MR_IPP


class RRT_BiasBetaInformative_GP_PathPlanning(InformativeRRTBaseClass):
    """rrt.py:893-906 ("Bias Beta RRT")."""

    def __init__(self, *args, directional_bias: Optional[float] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.directional_bias = directional_bias
        self.name = "RRT_BiasBetaInformative_GP_Path"

    def path_selection(self, agent_idx: int, current_position: Optional[np.ndarray] = None,
                       current_budget: Optional[float] = None) -> List[np.ndarray]:
        return bias_beta_path_selection(self, agent_idx, current_position)

    def tree_generation(self, budget_portion: float, agent_idx: int) -> None:
        rrt_tree_generation(self, budget_portion, agent_idx)

    def information_update(self) -> None:
        gp_information_update(self)

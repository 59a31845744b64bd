"""CPU oracle for the information-gain / tree half of the hot path -- TEST INFRASTRUCTURE ONLY.

Plain-Python restatement (node objects, scalar numpy calls, no vectorisation) of
  src/rrt/rrt_utils.py:10-114   tree node model, nearest / steer / near / rewire / trace
  src/rrt/rrt.py:32-259         suppression factor + the four point-source gain functions
  src/rrt/rrt.py:263-327        rrt / rrt* / rig tree generators
  src/rrt/rrt.py:380-526        stage-2 selector, stage-1 (bias-beta) selector, mutual information
  src/rrt/rrt_thread.py:371-401 UCB selector on the GP posterior
It is pinned against reference runs by tests/test_oracle_golden.py (fixtures from
oracle/make_golden.py).  The product never imports it.
"""
from __future__ import annotations

import numpy as np

norm = np.linalg.norm


class Node:
    """rrt_utils.py:10-35."""

    __slots__ = ("point", "parent", "children", "information", "index")

    def __init__(self, point, parent=None):
        self.point = point
        self.parent = parent
        self.children = []
        self.information = 0
        self.index = -1


def attach(nodes, child, parent):
    """add_node -> add_child (rrt_utils.py:26-29,103-105): the nearest node always becomes the parent."""
    parent.children.append(child)
    child.parent = parent
    child.index = len(nodes)
    nodes.append(child)


def path_cost(node):
    """rrt_utils.py:71-76, leaf-to-root accumulation order."""
    total = 0
    while node.parent is not None:
        total += norm(node.point - node.parent.point)
        node = node.parent
    return total


def nearest(nodes, p):
    return min(nodes, key=lambda nd: norm(nd.point - p))  # first minimum wins (rrt_utils.py:100-101)


def steer(node, target, step):
    """rrt_utils.py:93-98."""
    delta = target - node.point
    dist = norm(delta)
    delta = delta / dist if dist > 0 else delta
    dist = min(dist, step)
    return node.point + delta * dist


def near(new_node, nodes, radius):
    return [nd for nd in nodes if norm(new_node.point - nd.point) < radius]


def rewire(neighbours, new_node):
    """rrt_utils.py:85-88: parent pointer only, children lists untouched."""
    for nb in neighbours:
        if path_cost(new_node) + norm(new_node.point - nb.point) < path_cost(nb):
            nb.parent = new_node


def root_path(leaf):
    out = []
    while leaf is not None:
        out.append(leaf.point)
        leaf = leaf.parent
    out.reverse()
    return out


def sample_point(ws):
    """One np.random.rand(2) per attempt (rrt.py:266,283,310)."""
    return np.random.rand(2) * (ws[1] - ws[0], ws[3] - ws[2]) + (ws[0], ws[2])


# ---- gains (rrt.py:32-259) --------------------------------------------------------------------------
def suppression(point, source, sources):
    xt, yt = point
    xk, yk, _ = source
    F = 0
    for other in sources:
        if np.array_equal(source, other):
            continue
        mid = [(xk + other[0]) / 2, (yk + other[1]) / 2]
        d = norm([xt - mid[0], yt - mid[1]])
        C = 2 - 1 / (1 + np.exp((d - 2) / 16))
        F += (1 - C + C / (1 + np.exp((2 - d) / 16)))
    return F


class GainContext:
    """Everything a gain function reads from the strategy object (rrt.py:51-259)."""

    def __init__(self, best_estimates, last_point, agents_obs_wp, agent_idx, n_agents, d_wp, workspace, obstacles=()):
        self.est = np.asarray(best_estimates, dtype=float)
        self.last_point = last_point
        self.agents_obs_wp = agents_obs_wp
        self.agent_idx = agent_idx
        self.n_agents = n_agents
        self.d_wp = d_wp
        self.ws = workspace
        self.obstacles = list(obstacles)


def _segments_cross(x1, y1, x2, y2, x3, y3, x4, y4):
    d1 = (x2 - x1, y2 - y1)
    d2 = (x4 - x3, y4 - y3)
    den = d1[0] * d2[1] - d1[1] * d2[0]
    if den == 0:
        return False
    t = ((x3 - x1) * d2[1] - (y3 - y1) * d2[0]) / den
    u = ((x3 - x1) * d1[1] - (y3 - y1) * d1[0]) / den
    return 0 <= t <= 1 and 0 <= u <= 1


def _workspace_penalty(ctx, node, parent):
    """rrt.py:167-242."""
    p, ws = node.point, ctx.ws
    if p[0] < ws[0] or p[0] > ws[1] or p[1] < ws[2] or p[1] > ws[3]:
        return np.inf
    for ob in ctx.obstacles:
        if ob["type"] == "rectangle":
            x0, y0 = ob["x"], ob["y"]
            x1, y1 = x0 + ob["width"], y0 + ob["height"]
            if x0 <= p[0] <= x1 and y0 <= p[1] <= y1:
                return np.inf
    if parent:
        a, b = parent.point, node.point
        for ob in ctx.obstacles:
            if ob["type"] != "rectangle":
                continue
            x0, y0 = ob["x"], ob["y"]
            x1, y1 = x0 + ob["width"], y0 + ob["height"]
            if (_segments_cross(a[0], a[1], b[0], b[1], x0, y0, x0, y1) or _segments_cross(a[0], a[1], b[0], b[1], x1, y0, x1, y1)
                    or _segments_cross(a[0], a[1], b[0], b[1], x0, y0, x1, y0)
                    or _segments_cross(a[0], a[1], b[0], b[1], x0, y1, x1, y1)):
                return np.inf
    return 0


def _close_count(ctx, point):
    """rrt.py:246-252 / 471-477: observations of ALL agents closer than d_wp."""
    n = 0
    for i in range(ctx.n_agents):
        n += len([o for o in ctx.agents_obs_wp[i] if norm(point - o) < ctx.d_wp])
    return n


def node_term(variant, ctx, node):
    """One summand of the branch gain (the nested sources_gain closures of rrt.py:52-63,73-84,99-110,142-160)."""
    xt, yt = node.point
    have = ctx.est.size > 0
    others = [s for s in ctx.est]
    if variant == 3:
        if not have:
            return -_workspace_penalty(ctx, node, node.parent)
        g = 0
        for s in ctx.est:
            d = norm([xt - s[0], yt - s[1]])
            g += (np.exp(-(d - 2) ** 2 / (32))) * suppression(node.point, s, others)
        dist_pen = np.exp(0.5 * norm(node.point - node.parent.point) ** 2) if node.parent else 0
        expl = np.exp(0.05 * _close_count(ctx, node.point) ** 2) if len(ctx.agents_obs_wp[ctx.agent_idx]) > 0 else 0
        return g - dist_pen - expl - _workspace_penalty(ctx, node, node.parent)
    if not have:
        return 0
    g = 0
    if variant == 0:
        lp = ctx.last_point
        s = min(ctx.est, key=lambda src: norm([lp[0] - src[0], lp[1] - src[1]]))
        d = norm([xt - s[0], yt - s[1]])
        g += (1 + np.exp(-(d - 2) ** 2 / 2 * 16)) * suppression(node.point, s, others)
        return g
    for s in ctx.est:
        d = norm([xt - s[0], yt - s[1]])
        g += (1 + np.exp(-(d - 2) ** 2 / 2 * 16)) * suppression(node.point, s, others)
    g = g * (np.exp(0.5 * norm(node.point - node.parent.point)) if node.parent else 1)
    if variant == 2:
        if node.parent:
            th = np.arctan2(node.point[1] - node.parent.point[1], node.point[0] - node.parent.point[0])
            g = g * np.exp(0.05 * (th ** 2) / 0.1)
    return g


def branch_gain(variant, ctx, node):
    """node -> root walk (rrt.py:65-70, 91-96, 123-128, 254-259); the root contributes nothing."""
    total = 0
    cur = node
    while cur.parent:
        total += node_term(variant, ctx, cur)
        cur = cur.parent
    return total if variant == 3 else max(total, 0)


# ---- tree generators (rrt.py:263-327) ------------------------------------------------------------
def grow_rig(nodes, ws, budget, d_wp, gain=None):
    """rig_tree_generation; `gain(node)` is evaluated at insertion, before the rewire (rrt.py:321-323)."""
    travelled = 0
    while travelled < budget:
        target = sample_point(ws)
        nn = nearest(nodes, target)
        node = Node(steer(nn, target, d_wp))
        nb = near(node, nodes, d_wp)
        attach(nodes, node, nn)  # choose_parent's pick is overwritten by add_node
        travelled += norm(node.point - nn.point)
        if gain is not None:
            node.information = gain(node)
        rewire(nb, node)
        if travelled >= budget:
            break
    return nodes


def grow_rrt(nodes, ws, budget, d_wp, gain=None):
    """rrt_tree_generation (rrt.py:263-278)."""
    travelled = 0
    while travelled < budget:
        target = sample_point(ws)
        nn = nearest(nodes, target)
        node = Node(steer(nn, target, d_wp), nn)
        attach(nodes, node, nn)
        travelled += norm(node.point - nn.point)
        if gain is not None:
            node.information = gain(node)
        if travelled >= budget:
            break
    return nodes


def grow_rrt_star(nodes, ws, budget, d_wp):
    """rrt_star_tree_generation (rrt.py:280-305): steer step is the literal 1.0."""
    travelled = 0
    while travelled < budget:
        target = sample_point(ws)
        nn = nearest(nodes, target)
        node = Node(steer(nn, target, 1.0))
        nb = near(node, nodes, d_wp)
        best, best_cost = nn, path_cost(nn) + norm(nn.point - node.point)
        for cand in nb:
            if path_cost(cand) + norm(cand.point - node.point) < best_cost:
                best_cost = path_cost(cand) + norm(cand.point - node.point)
                best = cand
        attach(nodes, node, best)
        travelled += norm(node.point - best.point)
        rewire(nb, node)
        if travelled >= budget:
            break
    return nodes


# ---- selectors --------------------------------------------------------------------------------------
def mutual_information(K_pio, n_leaves, beta):
    """rrt.py:512-526 (K_pi enters only through its shape)."""
    if K_pio.shape[1] == 0:
        return 0
    return 0.5 * np.log(np.linalg.det(np.eye(n_leaves) + beta * np.dot(K_pio, K_pio.T)))


def truncate_by_budget(path, position, budget, d_wp):
    """rrt.py:426-436 / 500-510 (the budget never actually decreases)."""
    if budget is None or position is None:
        return path
    out = []
    for p in path:
        if budget - norm(p - position) < d_wp:
            out.append(p)
            break
        out.append(p)
        budget -= norm(p - position) if not out else norm(p - out[-1])
    return out


def bias_beta_scores(nodes, kernel_fn, agents_obs_wp, agent_idx, n_agents, beta, d_wp, ws):
    """Scores of rrt.py:443-482 for every leaf (in tree order)."""
    leaves = [nd for nd in nodes if not nd.children]
    pts = np.array([nd.point for nd in leaves])
    obs = np.array(agents_obs_wp[agent_idx])
    K_pio = kernel_fn(pts, obs) if obs.size > 0 else np.zeros((len(pts), 0))
    mi = mutual_information(K_pio, len(pts), beta)
    # NB: no dtype -- when the agent has no observations mi is the int 0 and the reference's score
    # array is int64, so every `-=` below truncates toward zero (and -inf casts to INT64_MIN).
    scores = np.array([mi for _ in leaves])
    for i, nd in enumerate(leaves):
        scores[i] -= np.exp(0.5 * norm(nd.point - nd.parent.point) ** 2) if nd.parent else 0
        p = nd.point
        scores[i] -= np.inf if (p[0] < ws[0] or p[0] > ws[1] or p[1] < ws[2] or p[1] > ws[3]) else 0
        if len(agents_obs_wp[agent_idx]) > 0:
            n = 0
            for a in range(n_agents):
                n += len([o for o in agents_obs_wp[a] if norm(nd.point - o) < d_wp])
            scores[i] -= np.exp(0.05 * n ** 2)
    return leaves, scores, mi


def bias_beta_select(nodes, kernel_fn, agents_obs_wp, agent_idx, n_agents, beta, d_wp, ws, position=None, budget=None):
    leaves, scores, _ = bias_beta_scores(nodes, kernel_fn, agents_obs_wp, agent_idx, n_agents, beta, d_wp, ws)
    path = root_path(leaves[int(np.argmax(scores))])
    return truncate_by_budget(path, position, budget, d_wp)


def dfs_leaves(root):
    """rrt.py:389-395: leaves through `children` lists, depth first, in insertion order."""
    out = []

    def visit(nd):
        if not nd.children:
            out.append(nd)
        for ch in nd.children:
            visit(ch)

    visit(root)
    return out


def informative_select(root, have_estimates, position=None, budget=None, d_wp=None):
    """informative_source_metric_path_selection (rrt.py:380-436)."""
    leaves = dfs_leaves(root)
    if not leaves:
        return []
    if position is not None:
        leaves = [nd for nd in leaves if not np.array_equal(nd.point, position)]
        if not leaves:
            return [position]
    if not have_estimates:
        pick = np.random.choice(leaves)
        while pick.information == -np.inf:
            pick = np.random.choice(leaves)
    else:
        pick = max(leaves, key=lambda nd: nd.information)
    path = root_path(pick)
    if len(path) == 1 and np.array_equal(path[0], position):
        leaves = [nd for nd in leaves if not np.array_equal(nd.point, pick.point)]
        if leaves:
            path = root_path(np.random.choice(leaves))
    return truncate_by_budget(path, position, budget, d_wp)


def ucb_select(nodes, predict_fn, visited, beta, d_wp, budget):
    """rrt_thread.py:371-401."""
    leaves = [nd for nd in nodes if not nd.children]
    pts = np.array([nd.point for nd in leaves])
    mu, sd = predict_fn(pts)
    mu_n = (mu - np.mean(mu)) / np.std(mu) if np.std(mu) != 0 else mu
    sd_n = (sd - np.mean(sd)) / np.std(sd) if np.std(sd) != 0 else sd
    acq = mu_n + beta * sd_n
    for i, nd in enumerate(leaves):
        for v in visited:
            d = norm(nd.point - v)
            if d < d_wp:
                acq[i] -= 0.1 * (d_wp - d) / d_wp
    leaf = leaves[int(np.argmax(acq))]
    path = []
    while leaf is not None and budget is not None and budget > 0:
        path.append(leaf.point)
        if len(path) > 1:
            budget -= norm(path[-1] - path[-2])
        if budget <= d_wp:
            break
        leaf = leaf.parent
    path.reverse()
    return path, acq

"""Oracle-backed stand-ins for the two device backends, so the HOST logic of the product (tree
growth, RNG order, selection, budget bookkeeping, config flow) is testable without a GPU.
Test infrastructure only; the product never imports this."""
import numpy as np

from oracle import gp_oracle as O
from oracle import rrt_oracle as T


class OracleKernel:
    def __init__(self, c, l):
        self.constant_value, self.length_scale = float(c), float(l)

    def __call__(self, X, Y=None):
        return O.kernel(np.asarray(X, float).reshape(-1, 2), None if Y is None else np.asarray(Y, float).reshape(-1, 2),
                        self.constant_value, self.length_scale)

    @property
    def theta(self):
        return np.log([self.constant_value, self.length_scale])


class OracleGP:
    """sklearn-shaped estimator on the numpy oracle.  full=False only consumes the restart draws
    (enough for the decision tests: the reference's selectors never read the fitted state)."""
    full = True

    def __init__(self, kernel=None, n_restarts_optimizer=0, normalize_y=False, **_):
        self.kernel = OracleKernel(kernel.constant_value, kernel.length_scale)
        self.n_restarts_optimizer = n_restarts_optimizer
        self.state = None

    def fit(self, X, y):
        X = np.asarray(X, float).reshape(-1, 2)
        if self.full:
            self.state = O.fit(X, y, self.kernel.constant_value, self.kernel.length_scale, n_restarts=self.n_restarts_optimizer)
            self.kernel_ = OracleKernel(self.state["c"], self.state["l"])
        else:
            b = O.log_bounds()
            for _ in range(self.n_restarts_optimizer):
                np.random.mtrand._rand.uniform(b[:, 0], b[:, 1])
            self.state = O.fit(X, y, self.kernel.constant_value, self.kernel.length_scale, optimize=False)
            self.kernel_ = self.kernel
        self.X_train_ = X
        return self

    def predict_grid_host(self, x0, x1, nx, y0, y1, ny):
        """The estimator interface PointSourceField.predict_spatial_field calls, in the reference's own three lines
        (point_source.py:350-352) on the numpy oracle -- the product itself has no CPU branch."""
        Xg, Yg = np.meshgrid(np.linspace(x0, x1, nx), np.linspace(y0, y1, ny))
        z_log, std = self.predict(np.column_stack((Xg.ravel(), Yg.ravel())), return_std=True)
        return 10 ** z_log, std

    def predict(self, X, return_std=False):
        X = np.asarray(X, float).reshape(-1, 2)
        if self.state is None:
            return (np.zeros(len(X)), np.sqrt(np.full(len(X), self.kernel.constant_value))) if return_std else np.zeros(len(X))
        m, s, _ = O.predict(self.state, X)
        return (m, s) if return_std else m


class OracleScorer:
    """The scorer interface of mripp_b200.scoring on top of oracle/rrt_oracle.py."""

    def __init__(self):
        self.launches = 0

    @staticmethod
    def _ctx(ctx):
        return T.GainContext(ctx.best_estimates, ctx.last_point, ctx.agents_obs_wp, ctx.agent_idx, len(ctx.agents_obs_wp),
                             ctx.d_wp, ctx.workspace, ctx.obstacles)

    def tree_information(self, variant, tree, ctx):
        """Replay the insertions in order, applying each rewire right after the node that caused it
        was scored -- i.e. the reference's own evaluation order (rrt.py:321-323)."""
        g = self._ctx(ctx)
        nodes = [T.Node(tree.point(0))]
        nodes[0].index = 0
        events = {}
        for t, q, p in tree.rewires:
            events.setdefault(t, []).append((q, p))
        info = np.zeros(tree.n)
        for i in range(1, tree.n):
            nd = T.Node(tree.point(i))
            T.attach(nodes, nd, nodes[tree.parent0[i]])
            info[i] = T.branch_gain(variant, g, nd)
            for q, p in events.get(i, []):
                nodes[q].parent = nodes[p]
        return info

    def prepare_counts(self, obs_wp, obs_vals):
        return {"obs_wp": obs_wp, "obs_vals": obs_vals}

    def particle_log_likelihood(self, thetas, prepared, lambda_b, M):
        """The reference's own values (S1 = S2 = None tells the sampler there is no band to check)."""
        from mripp_b200.src.estimation.estimation import batched_poisson_log_likelihood
        return batched_poisson_log_likelihood(thetas, prepared["obs_wp"], prepared["obs_vals"], lambda_b, M), None, None

    def leaf_scores(self, leaf_xy, parent_xy, has_parent, ctx, beta, kernel_c, kernel_l):
        nodes = []
        for k in range(len(leaf_xy)):
            nd = T.Node(leaf_xy[k], T.Node(parent_xy[k]) if has_parent[k] else None)
            nodes.append(nd)
        kfn = lambda A, B: O.kernel(A, B, kernel_c, kernel_l)
        _, scores, mi = T.bias_beta_scores(nodes, kfn, ctx.agents_obs_wp, ctx.agent_idx, len(ctx.agents_obs_wp), beta, ctx.d_wp,
                                           ctx.workspace)
        return scores, mi

    def ucb_acquisition(self, gp, leaf_xy, visited_xy, beta, d_wp):
        nodes = [T.Node(p) for p in leaf_xy]
        _, acq = T.ucb_select(nodes, lambda P: gp.predict(P, return_std=True), list(visited_xy), beta, d_wp, None)
        _, sd = gp.predict(leaf_xy, return_std=True)
        return acq, float(np.mean(sd))


class NoisyLikelihoodScorer(OracleScorer):
    """Stands in for the device likelihood on a CPU box: the reference's log-likelihoods plus a seeded error of up
    to `scale` times the bound a real device evaluation reports, with S1 / S2 inflated by the same factor -- i.e. a
    backend that is as wrong as it is allowed to be.  Lets the error-band logic of the sampler run without a GPU."""

    def __init__(self, scale=1.0, seed=0, extreme=False):
        super().__init__()
        self.scale, self.rng, self.extreme = float(scale), np.random.RandomState(seed), bool(extreme)

    def particle_log_likelihood(self, thetas, prepared, lambda_b, M):
        from scipy.special import gammaln
        from mripp_b200.src.estimation import estimation as E
        ll, _, _ = super().particle_log_likelihood(thetas, prepared, lambda_b, M)
        obs = np.array(prepared["obs_wp"], dtype=float).reshape(-1, 2)
        k = np.round(prepared["obs_vals"]).astype(int)
        src = np.asarray(thetas, float).reshape(len(thetas), M, 3)
        d2 = (obs[None, :, None, 0] - src[:, None, :, 0]) ** 2 + (obs[None, :, None, 1] - src[:, None, :, 1]) ** 2
        rate = np.maximum(lambda_b + np.sum(src[:, None, :, 2] / np.maximum(d2, 1e-6), axis=2), 1e-10)
        kl, g = k[None, :] * np.log(rate), gammaln(k + 1)[None, :]
        s1 = (np.abs(kl) + g + rate).sum(axis=1) * self.scale
        s2 = np.abs(kl - g - rate).sum(axis=1) * self.scale
        bound = E.EPS * (E.C_TERM * s1 + E.c_sum(len(k)) * s2)
        err = np.sign(self.rng.uniform(-1.0, 1.0, len(ll))) if self.extreme else self.rng.uniform(-1.0, 1.0, len(ll))
        return ll + err * bound, s1, s2

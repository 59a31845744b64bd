/*
 * ipp_b200.h -- C ABI of the B200 (sm_100a) GP-estimation / information-gain hot path.
 *
 * The reference (franfrancisco9/Multi-Robot-Informative-Path-Planning) is pure Python and has no
 * FFI of its own; its boundary for this path is the duck-typed attribute PointSourceField.gp
 * (an sklearn GaussianProcessRegressor, src/point_source/point_source.py:71) plus the module-level
 * gain / selection functions in src/rrt/rrt.py.  Each entry point below names the reference
 * call it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to caller-owned memory unless the name ends in _host;
 *     the library never allocates, frees or synchronises;
 *   - all matrices are row-major fp64; `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, <0 on bad argument (-1) or CUDA launch error (-2).  Numerical failure
 *     (Gram matrix not positive definite) is reported LAPACK-style through the `info` slot of the
 *     result block: info = k > 0 means the leading minor of order k is not positive definite
 *     (the caller raises numpy.linalg.LinAlgError in fit, returns (-inf, 0) inside the optimiser
 *     objective -- sklearn/gaussian_process/_gpr.py:590-593, 351-361);
 *   - no global mutable state: re-entrant, one stream per calling thread.
 */
#ifndef IPP_B200_H
#define IPP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version, bumped on any signature change. */
int ipp_abi_version(void);

/* ---- GP fit state ------------------------------------------------------------------------
 * Buffer sizes for n observations.  out[0..9] = {NP, rows, ld, Kbuf doubles, Wbuf doubles,
 * Tbuf doubles, vbuf doubles, offset of alpha_ in vbuf, offset of the 8-double result block in
 * vbuf, vector slot length}.  NP = round_up(n, 64); Kbuf is (NP+64) x NP, Wbuf and Tbuf NP x NP.
 * vbuf slots of length out[9]: 0 = x/l, 1 = y/l, 2 = (unused), 3 = back-substitution work,
 * 4 = alpha_.  Result block: {lml, dlml/dlog c, dlml/dlog l, info, z.z, sum log diag L, -, -}. */
int ipp_gp_layout(int n, long long* out_host);

/* K = c * exp(-|x-y|^2 / (2 l^2)), n x m, leading dimension ldk.  Y == NULL: symmetric Gram of X
 * with the diagonal set to exactly c + jitter.
 * Replaces gp.kernel(X) / gp.kernel(X, Y): src/rrt/rrt.py:447,452 ->
 * sklearn kernels.py:936-971 (Product), 1244-1296 (ConstantKernel), 1530-1587 (RBF). */
int ipp_gram_rbf(const double* X, int n, const double* Y, int m, double c, double l, double jitter, double* K,
                 long long ldk, void* stream);

/* The same builder for the other stationary kernels scikit-learn offers for this slot: kind 0 = RBF, 1 = Matern
 * nu = 1.5, 2 = Matern nu = 2.5 (sklearn kernels.py:1723-1735), each times the constant c.  The reference only ever
 * builds C * RBF (point_source.py:86); fit and posterior are RBF-only. */
int ipp_gram_kernel(int kind, const double* X, int n, const double* Y, int m, double c, double l, double jitter, double* K,
                    long long ldk, void* stream);

/* One log-marginal-likelihood (+ gradient w.r.t. log c, log l) evaluation.
 * X: n x 2, y: n (already normalised), hyp: DEVICE {c, l, jitter}.  On return the result block
 * in vbuf holds lml / gradient / info; Kbuf holds L (lower, row NP = z = L^-1 y), Wbuf = L^-1.
 * max_ctas > 1 caps the grid of the cooperative Cholesky kernel, so that several evaluations (the
 * independent optimiser restarts, one stream and one buffer set each) can be resident at once; 0 = whole GPU.
 * Replaces GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True):
 * sklearn _gpr.py:541-656, the objective of the 11 L-BFGS-B runs of fit (_gpr.py:302-333). */
int ipp_gp_lml_grad(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                    double* Tbuf, double* vbuf, int with_grad, int max_ctas, void* stream);

/* Final factorisation at the chosen theta: L_ in Kbuf, W = L_^-1 in Wbuf, alpha_ = L_^-T L_^-1 y
 * in vbuf (offset out[7]), scaled coordinates in slots 0/1, lml in the result block.
 * Replaces sklearn _gpr.py:349-367 (kernel_(X), cholesky, cho_solve). */
int ipp_gp_fit_final(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                     double* Tbuf, double* vbuf, void* stream);

/* ---- batched fit: the optimiser restarts of one fit evaluated together ---------------------------------
 * The dataflow Cholesky (csrc/ipp_cholq.cuh) executes a TASK PLAN: the list of block tasks (factor / panel solve /
 * rank-128 tile update) with the flags each one waits for.  ipp_host_chol_plan writes the plan for n observations
 * into out_host (host ints; *need = its length; out_host == NULL only sizes it); `near_slack` >= 0: tile updates
 * within near_slack + 1 super-panels of the factorisation front get the priority of the dependency chain (0: only
 * the tiles the chain waits for at once).  The caller uploads the plan once per size.
 * Host function: no device work, no stream. */
int ipp_host_chol_plan(int n, int near_slack, int* out_host, long long capacity, long long* need_host);

/* ints of device scratch (`sched`: dependency flags + queue heads) the dataflow Cholesky needs for R problems. */
int ipp_chol_sched_ints(int n, int R, long long* out_host);

/* ipp_gp_lml_grad for R independent hyper-parameter vectors at once over the SAME (X, y): problem r uses slot
 * s = slots[r] (slots == NULL: s = r) of slot-strided buffers -- Kbuf + s * out[3], Wbuf + s * out[4],
 * Tbuf + s * out[5], vbuf + s * out[6] (ipp_gp_layout) and hyp + 4 s = {c, l, jitter, -}.  One launch sequence; the
 * factorisations share one persistent kernel and fill each other's dependency stalls.  Per-problem arithmetic does
 * not depend on R or on the slot.  plan: DEVICE copy of ipp_host_chol_plan; sched: ipp_chol_sched_ints(n, R) ints.
 * flags: bit 0 = do not hand the critical chain over by hints (measurement aid); bit 1 / bit 2 = leave per-CTA task /
 * cycle counters / a per-task trace at the end of sched; bit 3 = advance the ready frontiers after every chain task;
 * bit 4 = the executor without warp specialisation (every warp claims, copies and multiplies in turn: A/B aid);
 * bit 5 = tile updates written back by load / subtract / store instead of RED.ADD.F64 (A/B aid; same bits);
 * bits 8.. = cap on the CTAs of the factorisation kernel (0: whole GPU).  info = -7777 in the result block: the kernel's watchdog fired.
 * tile_bbox (nullable): {xmin, xmax, ymin, ymax} (unscaled) of every block of T consecutive observations in the order
 * given (T = 64 for NP <= 1536, else 128; NP / T boxes).  When present the gradient contraction skips tile pairs
 * whose boxes are farther apart than 13.6 l (every dK entry there is below 2e-38 c): worth giving when the caller has
 * put the observations in a spatial order (the Python layer: Z-order curve).
 * Replaces the objective evaluations of the n_restarts_optimizer + 1 L-BFGS-B runs of fit (sklearn _gpr.py:302-333),
 * which are independent once their start points are drawn. */
int ipp_gp_lml_grad_batch(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                          double* Tbuf, double* vbuf, const int* slots, int R, int with_grad, const int* plan, int* sched,
                          int flags, const double* tile_bbox, void* stream);

/* ipp_gp_fit_final with the dataflow Cholesky (plan / sched as above; plan == NULL: the cooperative kernel). */
int ipp_gp_fit_final_plan(const double* X, const double* y, int n, const double* hyp, double* Kbuf, double* Wbuf,
                          double* Tbuf, double* vbuf, const int* plan, int* sched, int flags, void* stream);

/* ---- GP posterior --------------------------------------------------------------------------
 * mean = y_std * (k*^T alpha) + y_mean ; std = y_std * sqrt(max(0, c - |W k*|^2)) for the flat
 * grid indices [g_begin, g_end) of the nx x ny lattice x = linspace(x0, x1, nx),
 * y = linspace(y0, y1, ny), index = iy * nx + ix (point_source.py:65-67,350).  Outputs are indexed
 * from 0 (slab-local).  zpred (nullable) receives 10**mean (point_source.py:352); neg_count
 * (nullable, int*) counts variances clamped at 0 (_gpr.py:485-491).
 * tables (nullable): caller-owned scratch of (nx + ny) * NP doubles (NP = ipp_gp_layout out[0]).  When
 * given and nx >= 128 the lattice fast path is used: the RBF cross-covariance factorises over the
 * lattice axes, so two coordinate tables replace the G x n exponentials; without it every element of
 * k* is evaluated with exp() like the explicit-point entry below.
 * slab_bbox (nullable) + cutoff: block sparsity of k*.  slab_bbox[s] = {xmin, xmax, ymin, ymax} (unscaled
 * coordinates) of observations 16 s .. 16 s + 15 in the order the state was fitted in; a slab whose box is farther
 * than `cutoff` from every query of a 128-query tile is skipped.  The caller chooses cutoff = l * sqrt(2 ln(1/eps))
 * so that every skipped entry of k* is below eps * c (the Python layer: eps = 1e-30, observations fitted in
 * Morton order so that slabs are compact).  NULL or cutoff <= 0: dense.
 * Replaces gp.predict(r, return_std=True) on the workspace grid: point_source.py:350-351 ->
 * sklearn _gpr.py:446-500, without materialising K_trans (G x n) or V (n x G). */
int ipp_gp_posterior_grid(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                          double y_std, double x0, double x1, int nx, double y0, double y1, int ny,
                          long long g_begin, long long g_end, double* mean, double* std, double* zpred,
                          int* neg_count, double* tables, const double* slab_bbox, double cutoff, int max_ctas,
                          void* stream);

/* Same posterior at m explicit query points Q (m x 2).
 * Replaces gp.predict(leaf_points, return_std=True): src/rrt/rrt_thread.py:375. */
int ipp_gp_posterior_points(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                            double y_std, const double* Q, long long m, double* mean, double* std, double* zpred,
                            int* neg_count, int max_ctas, void* stream);

/* The same posterior at explicit points through a cross-covariance TABLE: for chunks of `table_rows` queries (a
 * multiple of 128) the library first writes T[q][k] = exp(-|q - x_k|^2 / 2 l^2) -- one exp() per (query, observation)
 * where the entry above re-evaluates it once per 128-row block of W -- and then runs the chunk on the lattice kernel's
 * warp-specialised pipeline as a one-row lattice.  tables: caller-owned scratch of
 * table_rows * NP + NP + 4 * (table_rows / 128) doubles.  slab_bbox / cutoff as for ipp_gp_posterior_grid; the box of
 * a 128-query tile is that of its own queries, so skipping needs the caller to pass spatially sorted queries
 * (the Python layer sorts along the Z-order curve and undoes the permutation).  Results agree with
 * ipp_gp_posterior_points to rounding (same products; the row blocks of W are aligned from the other end), and do
 * not depend on the chunking or on the order of the queries when nothing is skipped.
 * Replaces gp.predict(X, return_std=True) on arbitrary point sets: sklearn _gpr.py:446-500. */
int ipp_gp_posterior_points_tabled(const double* Wbuf, const double* vbuf, int n, double c, double l, double y_mean,
                                   double y_std, const double* Q, long long m, double* mean, double* std, double* zpred,
                                   int* neg_count, double* tables, long long table_rows, const double* slab_bbox,
                                   double cutoff, int max_ctas, void* stream);

/* ---- information gain / selectors ---------------------------------------------------------
 * 0.5 * log det(I_L + beta * K K^T), K = kernel(leaf_xy (L x 2), obs_xy (na x 2)) with the PRIOR
 * hyper-parameters (c, l).  Factorises the smaller of the L- and na-order forms (Sylvester).
 * Buffers: Gtmp >= round_up(L,64) * round_up(na,64) doubles; Kbuf/Wbuf/vbuf sized by
 * ipp_gp_layout(min(L, na)).  Result block slot [5] = the value, [3] = info.
 * Replaces mutual_information_gain(K_pi, K_pio, beta_t): src/rrt/rrt.py:512-526 (and the two
 * gp.kernel calls feeding it, rrt.py:447,452). */
int ipp_mi_logdet(const double* leaf_xy, int L, const double* obs_xy, int na, double c, double l, double beta,
                  double* Gtmp, double* Kbuf, double* Wbuf, double* vbuf, void* stream);

/* For each of np points: number of observations strictly closer than d (`sure`), and number
 * within a few ulp of d (`amb`, to be re-decided by the host with numpy's own norm); amb_total
 * accumulates the sum of amb (caller zeroes it).
 * tile_bbox (nullable): obs is grouped in tiles of 64 consecutive rows with bounding boxes
 * tile_bbox[t] = {xmin, xmax, ymin, ymax} (the count is order independent, so the caller may sort the observations
 * spatially); each point then only visits the tiles within d of it.
 * Replaces the inner list comprehension of exploitation_penalty: rrt.py:246-252, 471-477. */
int ipp_count_close(const double* P, int np, const double* obs, int no, double d, int* sure, int* amb, int* amb_total,
                    const double* tile_bbox, void* stream);

/* Phase 1 of branch scoring: per-node static terms of gain variant
 *   0 point_source_gain_no_penalties (rrt.py:51-70; only source `closest_idx`)
 *   1 ..._only_distance_penalty (rrt.py:72-96)   2 ..._distance_rotation_penalty (rrt.py:98-128)
 *   3 point_source_gain_all (rrt.py:130-259; also wpen[N] = workspace/obstacle penalty of the node)
 * sources: S x 3 rows [x, y, A] (best_estimates); ws4 = workspace_size; obstacles: n_obst x 4
 * rows [x, y, width, height].  Includes calculate_suppression_factor (rrt.py:32-47). */
int ipp_node_terms(int variant, const double* node_xy, int N, const double* sources, int S, int closest_idx,
                   const double* ws4, const double* obstacles, int n_obst, double* src_gain, double* wpen,
                   void* stream);

/* Phase 2: gain of B branches.  Branch b starts at node eval_node[b] and follows parent pointers
 * to the root as they were when node eval_time[b] was inserted: parent0[q] is the insertion-time
 * parent (-1 = root) and (hist_off, hist_time, hist_parent) is a CSR log of later rewire events
 * (rrt_utils.py:85-88) applied when hist_time < eval_time[b].  hist_off / eval_time may be NULL
 * (no rewires / final snapshot).  term0: N doubles of scratch -- the per-step term of every node against its
 * insertion-time parent is computed once (phase 1b) and the walks gather it; a step is recomputed only where a
 * rewire changed the parent.  Replaces the `while current_node.parent` walks of rrt.py:65-70,
 * 91-96, 123-128, 254-259 as called from rig_tree_generation (rrt.py:321). */
int ipp_branch_gain(int variant, const double* node_xy, int N, const int* parent0, const int* hist_off,
                    const int* hist_time, const int* hist_parent, const double* src_gain, const double* wpen,
                    const int* n_close, int agent_has_obs, int have_estimates, const double* obstacles, int n_obst,
                    const int* eval_node, const int* eval_time, int B, double* term0, double* gain_out, void* stream);

/* Stage-1 selector scores ((mi - distance) - workspace) - exploitation for L leaves.
 * Replaces the penalty loop of bias_beta_path_selection: rrt.py:460-482. */
int ipp_leaf_penalty(const double* leaf_xy, const double* parent_xy, const int* has_parent, int L, const int* n_close,
                     int agent_has_obs, const double* ws4, double mi, double* scores, void* stream);

/* acq[i] -= 0.1 * (d - dist) / d for every path point closer than d, in path order.
 * Replaces src/rrt/rrt_thread.py:383-387. */
int ipp_ucb_visit_penalty(const double* leaf_xy, int L, const double* path_xy, int P, double d, double* acq,
                          void* stream);

/* np.argmax semantics (first maximum; the first NaN wins).  rrt.py:486, rrt_thread.py:389. */
int ipp_argmax_first(const double* v, int n, double* best_val, int* best_idx, void* stream);

/* ---- callers either side of the hot path -----------------------------------
 * Analytic ground-truth field over the workspace lattice (index = iy * nx + ix, np.linspace axes):
 * inverse-square intensity + detector response with air absorption, and concrete absorption along the chord of
 * the source -> point ray inside every rectangle obstacle.  sources: S x 3 rows [x, y, A]; obstacles: n_obst x 4
 * rows [x, y, width, height].  Replaces PointSourceField.ground_truth and compute_path_length_in_obstacle:
 * src/point_source/point_source.py:163-260 (a Python loop over every lattice point when obstacles are present). */
int ipp_ground_truth_grid(const double* sources, int S, const double* obstacles, int n_obst, double x0, double x1,
                          int nx, double y0, double y1, int ny, double r_s, double r_d, double* field, void* stream);

/* out[0..3] = {RMSE, source-weighted RMSE of log10(Z + 1), differential entropy of std in bits, sum of z_true} over
 * G lattice points, one fused pass over device-resident arrays.  partial: 4 * 592 doubles of scratch.
 * Replaces main.py:113-114 and calculate_differential_entropy, src/visualization/plot_helper.py:383-398. */
int ipp_grid_metrics(const double* zpred, const double* std, const double* ztrue, long long G, double* partial,
                     double* out, void* stream);

/* Poisson log-likelihood of N particles (rows of 3*M doubles [x, y, A] per source) against n count observations:
 *   out[3i]   = sum_j [ k_j log(rate_ij) - lgam_j - rate_ij ],  rate_ij = max(lambda_b + sum_m A_im / max(d2_ijm, 1e-6), 1e-10)
 *   out[3i+1] = sum_j (|k_j log rate_ij| + |lgam_j| + rate_ij),  out[3i+2] = sum_j |term_ij|   (error-bound sums)
 * counts: the rounded measurements as doubles; lgam: lgamma(counts + 1) from the caller (one per observation,
 * shared by all particles); obs_xy must be 16-byte aligned (read as pairs).  rate is bit-identical to the numpy expression (no contraction, numpy's source-axis
 * summation order); M <= 32.  Replaces the per-particle loop of calc_weights, src/estimation/estimation.py:97,
 * over poisson_log_likelihood, src/estimation/estimation.py:15-55. */
int ipp_poisson_loglik(const double* particles, int N, int M, const double* obs_xy, const double* counts,
                       const double* lgam, int n, double lambda_b, double* out, void* stream);

/* ---- host-side entry points (no device work, no stream): SURVEY.md section 8f row 4 -----------------
 * np.linalg.norm of m 2-vectors in the rounding form `variant` (0: sqrt(x*x + y*y), 1: sqrt(fma(y, y, x*x)),
 * 2: sqrt(fma(x, x, y*y))).  Used once, to find the form that equals np.linalg.norm with the BLAS numpy links. */
int ipp_host_norm2(int variant, const double* xy, int m, double* out);

/* Grow one agent's tree in place on the structure-of-arrays tree (xy[capacity][2], parent0 = parent at insertion,
 * parent = current parent, edge = |point - current parent|, nchild) from K pre-drawn uniform pairs (np.random.rand(2)
 * per sample) until the travelled length reaches budget_portion, the pairs run out, or capacity / rewire_cap is
 * reached.  mode 0: rrt_tree_generation (rrt.py:263-277), 1: rrt_star_tree_generation (:280-305),
 * 2: rig_tree_generation (:307-328); primitives nearest / steer / near / cost / rewire / add_node of
 * src/rrt/rrt_utils.py:59-105 with their first-minimum and strict-inequality semantics.  ws4 = {x0, x1, y0, y1};
 * step = steer's d_max_step, radius = near's distance.  rewires: rows {time = id of the inserted node, node,
 * new parent} appended from *n_rewires_io; scratch: capacity ints (the library keeps a uniform grid and per-node path
 * costs of its own for the duration of the call).  Outputs: pairs consumed, node count,
 * finished (1 when the budget portion is spent).  Host pointers. */
int ipp_host_tree_grow(int mode, int norm_variant, double* xy, long long* parent0, long long* parent, double* edge,
                       long long* nchild, int n, int capacity, const double* uniforms, int K, const double* ws4,
                       double step, double radius, double budget_portion, double* travelled_io, long long* rewires,
                       int rewire_cap, int* n_rewires_io, int* scratch, int* consumed, int* n_out, int* finished);

/* Leaves of a tree in the depth-first order of the reference's traversal (rrt.py:389-395): from the root (node 0)
 * through the children lists as they were at insertion (parent0; children in insertion order), a node without children
 * is a leaf.  out: n ints; *n_leaves receives the count.  Host pointers. */
int ipp_host_leaves_dfs(const long long* parent0, int n, int* out, int* n_leaves);

#ifdef __cplusplus
}
#endif
#endif /* IPP_B200_H */

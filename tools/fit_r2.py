"""Round-2 fit experiments on the GPU box: dataflow (queue) Cholesky against the cooperative kernel.

  python tools/fit_r2.py check            correctness: queue vs coop (LML, gradient, factor), batch == single, bit for bit
  python tools/fit_r2.py time [n ...]     single LML+grad latency, batched throughput, full 11-restart fits
Every section prints as it goes (a later failure keeps the earlier numbers)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import synth_observations  # noqa: E402
from mripp_b200 import _lib  # noqa: E402
from mripp_b200 import gp as G  # noqa: E402

WS = {131: 40, 500: 20, 2000: 50, 5000: 100}


def data(n):
    X, meas, _ = synth_observations(n, WS.get(n, max(10, int(np.sqrt(n) * 1.4))), 0)
    return X, np.log10(meas)


def fixed_gp(X, y, mode, c=1.0, l=2.0):
    os.environ["IPP_CHOL"] = mode
    gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(c, l), optimizer=None, normalize_y=True)
    gp.skip_below = 0  # given order in both modes
    return gp.fit(X, y)


def section(name):
    print(f"\n==== {name}", flush=True)


def check():
    ok = True
    for n in (40, 64, 100, 300, 640, 1000, 1536, 1600, 2000, 3000, 5000):
        try:
            X, y = data(n)
            out = {}
            for mode in ("coop", "queue"):
                gp = fixed_gp(X, y, mode)
                th = np.log([1.3, 1.7])
                lml, gr = gp.log_marginal_likelihood(th, eval_gradient=True)
                lml0 = gp.log_marginal_likelihood(th, eval_gradient=False)
                L = gp.L_
                al = gp.alpha_.copy()
                out[mode] = (lml, gr, lml0, L, al)
            a, b = out["coop"], out["queue"]
            dl = abs(a[0] - b[0]) / max(1.0, abs(a[0]))
            dg = np.abs(a[1] - b[1]).max() / max(1.0, np.abs(a[1]).max())
            dL = np.abs(a[3] - b[3]).max() / np.abs(a[3]).max()
            da = np.abs(a[4] - b[4]).max() / np.abs(a[4]).max()
            # residual of the factor itself: L L^T against the Gram matrix (host, fp64)
            K = G.gram_device(X, None, 1.0, 2.0, 1e-10).cpu().numpy()
            res = [np.abs(o[3] @ o[3].T - K).max() for o in (a, b)]
            good = dl < 1e-9 and dg < 1e-6 and b[0] == b[2] and res[1] <= 4 * res[0] + 1e-12
            ok &= good
            print(f"n={n}: lml coop {a[0]:.10g} queue {b[0]:.10g} rel {dl:.2e}; grad rel {dg:.2e}; L rel {dL:.2e}; alpha rel {da:.2e}; "
                  f"|LL^T-K| coop {res[0]:.2e} queue {res[1]:.2e}; lml(no grad)==lml {b[0] == b[2]} -> {'OK' if good else 'FAIL'}", flush=True)
        except Exception:
            ok = False
            traceback.print_exc()
    section("warp-specialised executor == plain executor, bit for bit")
    for n in (300, 1000, 2000, 5000):
        try:
            X, y = data(n)
            got = {}
            for wsf in ("0", "1", "red"):
                os.environ["IPP_CHOL_WS"] = "0" if wsf == "0" else "1"
                os.environ["IPP_CHOL_RED"] = "1" if wsf == "red" else "0"  # "1": warp-specialised, read-modify-write
                gp = fixed_gp(X, y, "queue")
                lml, gr = gp.log_marginal_likelihood(np.log([1.3, 1.7]), eval_gradient=True)
                got[wsf] = (lml, gr, gp.L_)
            os.environ["IPP_CHOL_WS"] = "1"
            os.environ.pop("IPP_CHOL_RED")
            same = all(got["0"][0] == got[k][0] and np.array_equal(got["0"][1], got[k][1]) and np.array_equal(got["0"][2], got[k][2])
                       for k in ("1", "red"))
            ok &= same
            print(f"n={n}: identical {same} (lml {got['1'][0]:.12g})", flush=True)
        except Exception:
            ok = False
            traceback.print_exc()
    section("batch == single, bit for bit")
    os.environ["IPP_CHOL"] = "queue"
    for n, R in ((300, 5), (1000, 3), (2000, 4), (5000, 3)):
        try:
            X, y = data(n)
            gp = fixed_gp(X, y, "queue")
            rng = np.random.RandomState(1)
            thetas = [np.array([rng.uniform(-2, 2), rng.uniform(-1, 2)]) for _ in range(R)]
            single = [gp._lml(t, True) for t in thetas]
            ws = G._Workspace().ensure(n, slots=R + 2)
            slots = list(range(R + 1, 1, -1))[:R]  # not the identity mapping
            res = gp._eval_batch(ws, slots, thetas, True)
            same = all(res[j][0] == single[j][0] and res[j][1] == single[j][1][0] and res[j][2] == single[j][1][1] for j in range(R))
            ok &= same
            print(f"n={n} R={R}: batch == single: {same}  (lml {[float(r[0]) for r in res][:3]})", flush=True)
        except Exception:
            ok = False
            traceback.print_exc()
    section("positive-definiteness failure is reported (info > 0), no hang")
    try:
        X = np.array([[1.0, 1.0], [2.0, 2.0], [1.0, 1.0], [3.0, 1.0]])
        y = np.array([0.1, 0.2, 0.1, 0.4])
        g = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), optimizer=None, alpha=1e-10).fit(X, y)
        g.alpha = 0.0
        v, gr = g.log_marginal_likelihood(np.zeros(2), eval_gradient=True)
        print("singular:", v, gr, flush=True)
        ok &= (v == -np.inf)
    except Exception:
        ok = False
        traceback.print_exc()
    print("\nCHECK", "PASSED" if ok else "FAILED", flush=True)
    return ok


def time_evals(gp, th, grad, reps):
    for _ in range(2):
        gp.log_marginal_likelihood(th, eval_gradient=grad)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        gp.log_marginal_likelihood(th, eval_gradient=grad)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t) / reps


def timing(ns, fits=True):
    th = np.log([1.0, 2.0])
    for n in ns:
        X, y = data(n)
        NP = (n + 63) // 64 * 64
        section(f"n = {n} (NP = {NP})")
        for mode, hints in (("coop", "1"), ("queue", "1")):
            try:
                os.environ["IPP_CHOL_HINTS"] = hints
                gp = fixed_gp(X, y, mode)
                reps = 5 if n >= 2000 else 20
                t_ng, t_g = time_evals(gp, th, False, reps), time_evals(gp, th, True, reps)
                print(f"  {mode:5s} hints={hints}: Gram+Cholesky {t_ng:.3f} ms ({NP**3 / 3 / t_ng / 1e9:.2f} TFLOP/s), LML+grad {t_g:.3f} ms "
                      f"({NP**3 / t_g / 1e9:.2f} TFLOP/s)", flush=True)
            except Exception:
                traceback.print_exc()
        os.environ["IPP_CHOL_HINTS"] = "1"
        os.environ["IPP_CHOL"] = "queue"
        try:
            os.environ["IPP_CHOL_STATS"] = "1"
            gp = fixed_gp(X, y, "queue")
            gp.log_marginal_likelihood(th, eval_gradient=False)
            print("  stats R=1:", G.chol_stats(gp._scratch_ws(), 1), flush=True)
            for R in (4,):
                ws = G._Workspace().ensure(n, slots=R)
                rng = np.random.RandomState(1)
                thetas = [np.array([rng.uniform(-1, 1), rng.uniform(0, 1.5)]) for _ in range(R)]
                gp._eval_batch(ws, list(range(R)), thetas, False)
                print(f"  stats R={R}:", G.chol_stats(ws, R), flush=True)
                del ws
        except Exception:
            traceback.print_exc()
        os.environ["IPP_CHOL_STATS"] = "0"
        try:
            gp = fixed_gp(X, y, "queue")
            for R in (2, 4, 11):
                ws = G._Workspace().ensure(n, slots=R)
                rng = np.random.RandomState(1)
                thetas = [np.array([rng.uniform(-1, 1), rng.uniform(0, 1.5)]) for _ in range(R)]
                for grad in (False, True):
                    for _ in range(2):
                        gp._eval_batch(ws, list(range(R)), thetas, grad)
                    torch.cuda.synchronize()
                    t = time.perf_counter()
                    reps = 3 if n >= 2000 else 10
                    for _ in range(reps):
                        gp._eval_batch(ws, list(range(R)), thetas, grad)
                    torch.cuda.synchronize()
                    ms = 1e3 * (time.perf_counter() - t) / reps
                    fl = NP ** 3 * (1.0 if grad else 1 / 3)
                    print(f"  batch R={R:2d} grad={int(grad)}: {ms:.3f} ms per launch, {ms / R:.3f} ms per problem "
                          f"({R * fl / ms / 1e9:.2f} TFLOP/s)", flush=True)
                del ws
        except Exception:
            traceback.print_exc()
        if not fits:
            continue
        section(f"full 11-restart fit, n = {n}")
        for label, mode, streams in (("lockstep batch (queue)", "queue", None), ("queue, serial", "queue", 1),
                                     ("coop, streams (round 1)", "coop", None)):
            try:
                os.environ["IPP_CHOL"] = mode
                gp = G.GaussianProcessRegressor(kernel=G.ConstantRBFKernel(1.0, 1.0), n_restarts_optimizer=10, normalize_y=True)
                gp.restart_streams = streams
                np.random.seed(0)
                gp.fit(X, y)  # warm-up (allocations)
                ev0 = gp.n_lml_evals_
                np.random.seed(0)
                torch.cuda.synchronize()
                t = time.perf_counter()
                gp.fit(X, y)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t
                print(f"  {label:26s}: {dt:.3f} s, {gp.n_lml_evals_ - ev0} evals, theta {gp.kernel_.theta}, "
                      f"lml {gp.log_marginal_likelihood_value_:.9f}", flush=True)
                del gp
                torch.cuda.empty_cache()
            except Exception:
                traceback.print_exc()
        os.environ["IPP_CHOL"] = "queue"


def chol_only(n):
    """Batched factorisation alone (no gradient), R = 1, 4, 11: the number the scheduling switches move."""
    X, y = data(n)
    NP = (n + 63) // 64 * 64
    gp = fixed_gp(X, y, "queue")
    out = []
    for R in (1, 4, 11):
        ws = G._Workspace().ensure(n, slots=R)
        rng = np.random.RandomState(1)
        thetas = [np.array([rng.uniform(-1, 1), rng.uniform(0, 1.5)]) for _ in range(R)]
        for _ in range(2):
            gp._eval_batch(ws, list(range(R)), thetas, False)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(4):
            gp._eval_batch(ws, list(range(R)), thetas, False)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t) / 4
        out.append(f"R={R}: {ms / R:.3f} ms/problem ({R * NP ** 3 / 3 / ms / 1e9:.2f} TF)")
        del ws
    print(f"n={n} WS={os.environ.get('IPP_CHOL_WS', '1')} RED={os.environ.get('IPP_CHOL_RED', '1')} NEAR={os.environ.get('IPP_CHOL_NEAR', 'default')} EAGER_SCAN={os.environ.get('IPP_CHOL_EAGER_SCAN', '0')} "
          f"MB={os.environ.get('IPP_GEMM_MB', '1')}: " + "; ".join(out), flush=True)


if __name__ == "__main__":
    _lib.require_cuda()
    cmd = sys.argv[1] if len(sys.argv) > 1 else "check"
    if cmd == "check":
        sys.exit(0 if check() else 1)
    if cmd == "chol":
        chol_only(int(sys.argv[2]) if len(sys.argv) > 2 else 5000)
        sys.exit(0)
    timing([int(a) for a in sys.argv[2:]] or [500, 2000, 5000], fits=(cmd != "evals"))

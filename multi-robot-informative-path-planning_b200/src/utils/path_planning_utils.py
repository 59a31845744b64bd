"""Reporting helpers (subset of the reference's src/utils/path_planning_utils.py used by main)."""
import os

import numpy as np


def run_number_from_folder() -> str:
    """path_planning_utils.py:12-27: next free integer under ./images."""
    base = "./images"
    if not os.path.exists(base):
        return "1"
    taken = [int(d) for d in os.listdir(base) if d.isdigit()]
    return str(max(taken) + 1) if taken else "1"


def calculate_source_errors(actual_sources, estimated_locs):
    """Per-coordinate RMSE between true and estimated sources, nearest-first pairing."""
    actual = np.asarray(actual_sources, dtype=float).reshape(-1, 3)
    est = np.asarray(estimated_locs, dtype=float).reshape(-1, 3)
    if est.size == 0 or actual.size == 0:
        return {"x_error": None, "y_error": None, "intensity_error": None}
    k = min(len(actual), len(est))
    d = np.linalg.norm(actual[:, None, :2] - est[None, :, :2], axis=-1)
    pairs, used = [], set()
    for i in np.argsort(d.min(axis=1))[:k]:
        j = next(j for j in np.argsort(d[i]) if j not in used)
        used.add(j)
        pairs.append((i, j))
    a = actual[[p[0] for p in pairs]]
    e = est[[p[1] for p in pairs]]
    rm = lambda c: float(np.sqrt(np.mean((a[:, c] - e[:, c]) ** 2)))
    return {"x_error": rm(0), "y_error": rm(1), "intensity_error": rm(2)}


def save_run_info(run_number, rmse, wrmse, entropy, sources, times, args, scenarios, folder="./runs_review"):
    """Plain-text summary of a sweep (path_planning_utils.py:29-123), one block per scenario/strategy."""
    os.makedirs(folder, exist_ok=True)
    name = os.path.join(folder, f"run_{run_number}_{args.get('stage_lambda', 'na')}.txt")
    with open(name, "w") as f:
        f.write(f"Run {run_number}\nargs: {args}\n")
        for sc_name, per_strategy in rmse.items():
            f.write(f"\n{sc_name}\n")
            for strat, vals in per_strategy.items():
                t = [v for v in times[sc_name][strat] if v is not None]
                f.write(f"  {strat}: RMSE {np.mean(vals):.6f} +- {np.std(vals):.6f}; "
                        f"WRMSE {np.mean(wrmse[sc_name][strat]):.6f}; entropy {np.mean(entropy[sc_name][strat]):.3f}; "
                        f"time {np.mean(t) if t else float('nan'):.3f}s; "
                        f"n_sources {sources[sc_name][strat]['n_sources']}\n")
    return name

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
    """path_planning_utils.py:125-141: absolute x / y / intensity errors of the estimates paired with the true sources in
    the order given (zip: the shorter list decides), as three lists."""
    errors = {"x_error": [], "y_error": [], "intensity_error": []}
    for actual, estimated in zip(actual_sources, estimated_locs):
        errors["x_error"].append(abs(actual[0] - estimated[0]))
        errors["y_error"].append(abs(actual[1] - estimated[1]))
        errors["intensity_error"].append(abs(actual[2] - estimated[2]))
    return errors


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

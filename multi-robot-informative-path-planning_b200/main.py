"""Experiment driver: the reference's main.py flow (JSON -> scenarios -> strategies -> run -> metrics)
on top of the B200 path.

    python -m mripp_b200.main -config path/to/setup.json [--sweep] [--out results.json]

Kept from the reference (main.py:16-150): the config keys, `seed == -1 -> None`, construction of a
strategy as `cls(scenario, **kwargs)` with kwargs routed by introspection of the constructors'
argument names, the per-run metrics (log10 RMSE, source-weighted RMSE, differential entropy, time).
Extension (BASELINE configs[4]): `--shard-rounds` gives every round its own RNG seed
(`seed + round`, so a round's result does not depend on which process runs it) and, under torchrun, runs round r on
rank r mod P -- one process per GPU, no data-path collective -- with the per-run metrics gathered on every rank.
Deliberate differences: the hard-coded sweep over num_agents x stage_lambda (main.py:167-183) runs
only with --sweep; figures are skipped (matplotlib is not a dependency); a 2-tuple workspace_size
from the shipped example config is widened to (0, W, 0, H) so the example runs (it raises
IndexError in the reference).
"""
from __future__ import annotations

import argparse
import json
from typing import Callable, Dict, List

import numpy as np

from mripp_b200.src.boustrophedon.boustrophedon import Boustrophedon, MR_Boustrophedon, MultiAgentBoustrophedon  # noqa
from mripp_b200.src.point_source.point_source import PointSourceField
from mripp_b200.src.rrt.rrt import *  # noqa: F401,F403  (strategy classes are looked up by name)
from mripp_b200.src.utils.path_planning_utils import calculate_source_errors, run_number_from_folder, save_run_info
from mripp_b200.src.point_source.point_source import DeviceField
from mripp_b200.src.visualization.plot_helper import calculate_differential_entropy, field_metrics_device


def load_configuration(config_path: str) -> Dict:
    try:
        with open(config_path, "r") as fh:
            return json.load(fh)
    except FileNotFoundError:
        raise SystemExit(f"Configuration file {config_path} not found.")
    except json.JSONDecodeError:
        raise SystemExit("Error decoding the configuration file. Please check its format.")


def initialize_scenarios(config: Dict, gp_factory=None) -> List[PointSourceField]:
    scenarios = []
    for sc in config["scenarios"]:
        params = dict(sc["params"])
        if "seed" in config["args"]:
            params["seed"] = None if config["args"]["seed"] == -1 else config["args"]["seed"]
        ws = params.get("workspace_size")
        if ws is not None and len(ws) == 2:
            params["workspace_size"] = (0, ws[0], 0, ws[1])
        elif ws is not None:
            params["workspace_size"] = tuple(ws)
        scenario = PointSourceField(**params) if gp_factory is None else PointSourceField(**params, gp_factory=gp_factory)
        # extension of the reference's config: defer the per-step GP refits nobody reads
        if hasattr(scenario.gp, "lazy"):
            scenario.gp.lazy = bool(config["args"].get("lazy_gp", False))
        # nothing is saved or shown: the posterior is only reduced to metrics, which happens on the device
        if gp_factory is None and not config["args"].get("save", False) and not config["args"].get("show", False):
            scenario.device_outputs = True
        for key, value in sc.get("specific_params", {}).items():
            scenario.update_source(int(key) - 1, *value)
        scenarios.append(scenario)
    return scenarios


def get_constructor_args(cls: Callable, args: Dict) -> Dict:
    """main.py:42-60: every config arg whose name appears in a constructor of the class or its bases."""
    picked = {}
    init = getattr(cls, "__init__", None)
    if init is not None and hasattr(init, "__code__"):
        picked.update({k: v for k, v in args.items() if k in init.__code__.co_varnames})
    for base in getattr(cls, "__bases__", []):
        picked.update(get_constructor_args(base, args))
    return picked


def initialize_strategies(config: Dict, args: Dict, scorer_factory=None) -> Dict[str, Callable]:
    makers = {}
    for name in config["strategies"]:
        cls = globals()[name]
        kwargs = get_constructor_args(cls, args)
        if scorer_factory is not None and get_constructor_args(cls, {"scorer": None}):
            kwargs = dict(kwargs, scorer=scorer_factory())
        makers[name] = (lambda scenario, cls=cls, kwargs=kwargs: cls(scenario, **kwargs))
    return makers


def run_simulations(scenarios, makers, args: Dict, debug: bool = False, shard=None) -> Dict:
    """shard = None: the reference's sequential rounds on one RNG stream.  shard = (rank, world): rounds seeded
    individually, round r run by rank (r - 1) mod world only."""
    run_number = run_number_from_folder() if args.get("run_number", -1) == -1 else args["run_number"]
    names = list(makers)
    rmse_all, wrmse_all, ent_all, time_all, src_all = {}, {}, {}, {}, {}
    results = []
    for s_idx, scenario in enumerate(scenarios, start=1):
        key = f"Scenario_{s_idx}"
        rmse_all[key] = {n: [] for n in names}
        wrmse_all[key] = {n: [] for n in names}
        ent_all[key] = {n: [] for n in names}
        time_all[key] = {n: [] for n in names}
        src_all[key] = {n: {"source": [], "n_sources": []} for n in names}
        for round_number in range(1, args["rounds"] + 1):
            if shard is not None:
                if (round_number - 1) % shard[1] != shard[0]:
                    continue
                base = args.get("seed", 0)
                np.random.seed((0 if base in (None, -1) else int(base)) + 7919 * s_idx + round_number)
            for name, make in makers.items():
                strategy = make(scenario)
                Z_pred, std = strategy.run()
                if isinstance(Z_pred, DeviceField) and isinstance(std, DeviceField):
                    # main.py:113-116 in one fused pass over the device-resident posterior (ipp_grid_metrics): the
                    # 16 B per lattice point are never read back
                    rmse, wrmse, entropy = field_metrics_device(Z_pred.tensor, std.tensor, scenario.ground_truth_device(as_tensor=True))
                else:
                    Z_true = scenario.ground_truth()
                    err = np.log10(Z_true + 1) - np.log10(Z_pred + 1)
                    rmse = np.sqrt(np.mean(err ** 2))
                    wrmse = np.sqrt(np.sum(Z_true * err ** 2) / np.sum(Z_true))
                    entropy = calculate_differential_entropy(std)
                est = np.array(getattr(strategy, "best_estimates", [])).reshape(-1, 3)
                rmse_all[key][name].append(rmse)
                wrmse_all[key][name].append(wrmse)
                ent_all[key][name].append(entropy)
                time_all[key][name].append(getattr(strategy, "time_taken", None))
                src_all[key][name]["source"].append(est)
                src_all[key][name]["n_sources"].append(len(est))
                results.append({"scenario": s_idx, "round": round_number, "strategy": name, "rmse": float(rmse),
                                "wrmse": float(wrmse), "entropy": float(entropy), "time": getattr(strategy, "time_taken", None),
                                "n_obs": int(len(strategy.obs_wp)), "source_errors": {k: [float(v) for v in vals] for k, vals in calculate_source_errors(scenario.sources, est).items()}})
                print(f"Round {round_number}/{args['rounds']} scenario {s_idx} {name}: RMSE {rmse:.4f} WRMSE {wrmse:.4f} "
                      f"time {getattr(strategy, 'time_taken', None)}")
    if args.get("save", False) and (shard is None or shard[1] == 1):  # sharded: the gathered JSON (--out) is the record
        save_run_info(run_number, rmse_all, wrmse_all, ent_all, src_all, time_all, args, scenarios)
    return {"run_number": run_number, "results": results}


def main(argv=None):
    ap = argparse.ArgumentParser(description="Run path planning scenarios on the B200 path.")
    ap.add_argument("-config", "--config", required=True)
    ap.add_argument("-debug", "--debug", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="the reference's num_agents x stage_lambda sweep (main.py:167-183)")
    ap.add_argument("--out", default=None, help="write the per-run metrics as JSON")
    ap.add_argument("--shard-rounds", action="store_true",
                    help="seed every round on its own and, under torchrun, run round r on rank r mod P (one process per GPU)")
    a = ap.parse_args(argv)
    return run(a, load_configuration(a.config))


def _init_shard():
    """(rank, world, group or None) for --shard-rounds: torch.distributed when launched by torchrun, else one process."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, False
    import torch
    import torch.distributed as dist
    started = False
    if not dist.is_initialized():
        cuda = torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl" if cuda else "gloo")
        started = True
    return dist.get_rank(), dist.get_world_size(), started


def run(a, config: Dict, gp_factory=None, scorer_factory=None):
    """main() after argument parsing; gp_factory / scorer_factory let the tests drive the flow without a GPU."""
    shard = None
    started = False
    if getattr(a, "shard_rounds", False):
        rank, world, started = _init_shard()
        shard = (rank, world)
    grid = [(config["args"].get("num_agents", 1), config["args"].get("stage_lambda", 0.875))]
    if a.sweep:
        grid = [(n, lam) for n in (1, 2, 3) for lam in (1.0, 0.5, 0.25, 0.0)]
    everything = []
    for num_agents, lam in grid:
        config["args"]["num_agents"] = num_agents
        config["args"]["stage_lambda"] = lam
        scenarios = initialize_scenarios(config, gp_factory)
        makers = initialize_strategies(config, config["args"], scorer_factory)
        out = run_simulations(scenarios, makers, config["args"], a.debug, shard)
        for r in out["results"]:
            r.update(num_agents=num_agents, stage_lambda=lam)
        everything.extend(out["results"])
    if shard is not None and shard[1] > 1:
        import torch.distributed as dist
        parts = [None] * shard[1]
        dist.all_gather_object(parts, everything)          # a few hundred bytes per run: the only exchange
        order = {name: k for k, name in enumerate(config["strategies"])}
        everything = sorted((r for p in parts for r in p),
                            key=lambda r: (r["num_agents"], -r["stage_lambda"], r["scenario"], r["round"], order[r["strategy"]]))
        if started:
            dist.destroy_process_group()
    if a.out and (shard is None or shard[0] == 0):
        with open(a.out, "w") as fh:
            json.dump(everything, fh, indent=1)
    return everything


if __name__ == "__main__":
    main()

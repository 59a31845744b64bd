"""Compare the per-run metrics of `main.py --shard-rounds` from a single process and from a torchrun launch
(gpurun_out/rounds_1.json vs gpurun_out/rounds_N.json) and print the wall times recorded next to them."""
import json, sys
a, b = (json.load(open(p)) for p in sys.argv[1:3])
key = lambda r: (r["scenario"], r["round"], r["strategy"])
same = len(a) == len(b) and all(key(x) == key(y) and x["rmse"] == y["rmse"] and x["wrmse"] == y["wrmse"] and x["entropy"] == y["entropy"]
                                and x["n_obs"] == y["n_obs"] for x, y in zip(a, b))
print(f"rounds_check: {len(a)} runs, metrics identical between the two launches: {same}")
for x in a[:4]:
    print("  ", key(x), f"rmse {x['rmse']:.6f} wrmse {x['wrmse']:.6f} entropy {x['entropy']:.1f} n_obs {x['n_obs']}")
sys.exit(0 if same else 1)

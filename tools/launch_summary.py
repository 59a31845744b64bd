"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean."""
import collections
import csv
import re
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = {}, collections.Counter()
    for r in data:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik])
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1e-6)
        agg[name] = agg.get(name, 0.0) + v
        cnt[name] += 1
    total = sum(agg.values())
    print(f"{sum(cnt.values())} launches, {total:.3f} ms of kernel time")
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"{k[:72]:72s} n={cnt[k]:5d} total {v:11.3f} ms ({100 * v / total:5.1f}%)  mean {1e3 * v / cnt[k]:10.1f} us")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        summarise(p)

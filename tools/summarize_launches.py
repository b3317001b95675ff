#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total ms / share.
usage: summarize_launches.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum" or int(r["ID"]) < skip:
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = name.replace("gpw::", "").replace("Fe<FpParams>", "Fp")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f %% |" % (k, n, ms, 100 * ms / tot))
    print("\nTotal %.1f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main()

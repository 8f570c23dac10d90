#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (and per grid for GEMMs).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys


def main(path, by_grid=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, n = collections.OrderedDict(), 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        if len(k) > 70:
            k = k[:67] + "..."
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u.startswith("n") else v * 1000 if u.startswith("m") else v
        key = (k, row["Grid Size"], row["Block Size"]) if by_grid else (k,)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1000:.2f} ms summed device time (ncu: cold cache, serialised — compare shares)")
    print(f"# {'sum us':>10} {'count':>6} {'avg us':>9} {'share':>6}  kernel" + (" grid block" if by_grid else ""))
    for key, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.1f} {c:6d} {t / c:9.1f} {100 * t / tot:5.1f}%  " + " ".join(key))


if __name__ == "__main__":
    main(sys.argv[1], by_grid=len(sys.argv) > 2 and sys.argv[2] == "--by-grid")

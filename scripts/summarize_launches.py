#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
count, total and mean device time, share of the listed launches.  Usage: summarize_launches.py launches.csv [skip]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("gm::<unnamed>::", "gm::")
        name = re.sub(r"void at::native::", "at::", name)[:70]
        rows.append((int(r["ID"]), name, us, r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, name, us, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | mean us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, (n, t, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% | {grid} | {block} |")
    print(f"\ntotal {total:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

#!/usr/bin/env python
"""Opcode histogram of one kernel from an .ncu-rep (source page, SASS view): executed warp instructions and
stall samples per opcode.  Usage: sass_profile.py report.ncu-rep kernel_name [top]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(rep, kernel, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    agg = defaultdict(lambda: [0.0, 0.0])
    tot = ts = 0.0
    for r in rows[h + 1:]:
        if len(r) <= ie:
            continue
        if r[ie] == "Instructions Executed":      # a second launch of the same kernel follows: keep the first
            break
        n, s = float(r[ie] or 0), float(r[smp] or 0)
        toks = r[src].split()
        op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
        op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "RED", "ATOMG", "SHFL") else "")
        agg[op][0] += n
        agg[op][1] += s
        tot += n
        ts += s
    print(f"{kernel}: {tot:.3e} warp instructions, {ts:.0f} samples")
    for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {op:14s} {100 * n / tot:5.1f}% inst   {100 * s / max(ts, 1):5.1f}% samples")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)

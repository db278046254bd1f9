#!/usr/bin/env python
"""Per-CUDA-source-line executed warp instructions and stall samples of one kernel (ncu report with -lineinfo).
Usage: line_profile.py report.ncu-rep kernel_name [top]"""
import csv
import subprocess
import sys


def main(rep, kernel, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, lines, seen_hdr = "", [], 0
    ie = smp = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            ie, smp = r.index("Instructions Executed"), r.index("# Samples")
            seen_hdr += 1
            continue
        if ie is None or len(r) <= ie or r[0] == "":
            continue       # SASS rows have an empty line number
        try:
            lines.append((float(r[ie] or 0), float(r[smp] or 0), cur_file, r[0], r[1].strip()))
        except ValueError:
            pass
    # several launches of the kernel may be in the report: fold identical lines
    agg = {}
    for n, s, f, ln, txt in lines:
        a = agg.setdefault((f, ln, txt), [0.0, 0.0])
        a[0] += n
        a[1] += s
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"{kernel}: {tot:.3e} warp instructions over the captured launches")
    for (f, ln, txt), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / tot:5.1f}% inst {100 * s / ts:5.1f}% smp  {f}:{ln:>4}  {txt[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)

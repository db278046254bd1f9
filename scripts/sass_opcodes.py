#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of libCudaRasterizer.so (cuobjdump -sass): the evidence that the library is
sm_100a-only code using bulk copies (UBLKCP / SYNCS), packed fp32 (FFMA2 / FMUL2 / FADD2), tensor-core MMA (HMMA),
multicast loads / stores (LDGMC / STG...MC), cp.async (LDGSTS) and programmatic dependent launch (ACQBULK / PREEXIT).
Usage: sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.md"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

LIB = sys.argv[1] if len(sys.argv) > 1 else "gaussianmesh_b200/diff_gaussian_rasterizater/libCudaRasterizer.so"
WATCH = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "HMMA", "LDGMC", "LDGSTS", "REDG", "ATOMG", "ATOMS", "SHFL", "MUFU", "ACQBULK",
         "PREEXIT", "LDS", "STS", "LDG", "STG", "BAR", "VOTE", "MATCH"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print("# Round 2 — static SASS opcode counts per kernel (`cuobjdump -sass`, `scripts/sass_opcodes.py`)\n")
    print(f"Library: `{LIB}`; architectures in the fatbin: {', '.join(arch) or 'n/a'}.\n")
    print("| kernel | instructions | " + " | ".join(WATCH) + " |")
    print("|---|---:|" + "---:|" * len(WATCH))
    for name, c in kernels.items():
        print(f"| `{name[:60]}` | {sum(c.values())} | " + " | ".join(str(c.get(w, 0)) for w in WATCH) + " |")
    print("\nUBLKCP = `cp.async.bulk` (TMA, 1-D), SYNCS = mbarrier operations, FFMA2 / FMUL2 / FADD2 = packed fp32x2, HMMA = "
          "`mma.sync` (TF32), LDGMC = `multimem.ld_reduce`, LDGSTS = `cp.async`, REDG = fire-and-forget global reduction, "
          "ACQBULK / PREEXIT = `griddepcontrol.wait` / `.launch_dependents` (programmatic dependent launch).")


if __name__ == "__main__":
    main()

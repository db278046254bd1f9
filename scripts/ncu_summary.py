#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: duration, DRAM bytes, throughputs, occupancy, top stalls.
Usage: ncu_summary.py report.ncu-rep [--json out.json]   (the JSON maps bench.py stage names to DRAM bytes per launch)"""
import csv
import json
import subprocess
import sys

STAGE_OF = {"blend_backward_kernel": "blend_backward", "blend_forward_kernel": "blend_forward",
            "blend_backward_pairs_kernel": "blend_backward", "blend_backward_cols_kernel": "blend_backward", "blend_backward_mma_kernel": "blend_backward", "blend_forward_pairs_kernel": "blend_forward", "preprocess_kernel": "preprocess",
            "emit_kernel": "emit", "bucket_sort_pack_kernel": "sort_pack", "geometry_backward_kernel": "geometry_backward",
            "l1_kernel": "l1_loss", "tile_scan_kernel": "tile_scan"}
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp instr")]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(rep, json_out=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic, seen = {}, set()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"### `{name}`  grid {r[idx['launch__grid_size']]} x block {r[idx['launch__block_size']]}")
        for key, label in WANT:
            if key in idx:
                print(f"- {label}: {r[idx[key]]} {units[idx[key]]}")
        st = [(float(r[i] or 0), h) for h, i in idx.items() if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
        tot = sum(v for v, _ in st) or 1
        print("- top stalls: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%"
                                          for v, h in sorted(st, reverse=True)[:5]))
        b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        if name in STAGE_OF:
            traffic[STAGE_OF[name]] = b
        print()
    if json_out:
        json.dump(traffic, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)

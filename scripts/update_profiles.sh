#!/bin/bash
# Regenerate profiles/r1_* from the artefacts a `scripts/gpu_check.sh <tag>` run left in gpurun_out/.
tag=${1:?tag}
set -e
cp gpurun_out/launches_${tag}.csv profiles/r1_launches.csv
cp gpurun_out/bench_ours_${tag}.json profiles/r1_bench_ours.json
cp gpurun_out/bench_ref_${tag}.json profiles/r1_bench_reference.json
python scripts/ncu_summary.py gpurun_out/prof_${tag}.ncu-rep --json profiles/dram_traffic.json > /tmp/ncu_${tag}.md
{
echo "# Round 1 — ncu launch list (B200, \`bench.py --steps 3 --warmup 2\`, 1M Gaussians @ 1920x1080)"
echo
echo "Command (scripts/gpu_check.sh): \`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize\`."
echo "Raw list: \`profiles/r1_launches.csv\`.  Times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event stage shares of \`bench.py\` below, not absolutes.  All sections of the bench are in the list (training frames, forward-only frames, edit frames, arena sizing passes)."
echo
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv 0
echo
echo "CUDA-event stage times inside the timed training region of bench.py (\`profiles/r1_bench_ours.json\`, 50 steps):"
echo
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}.json'))
print("| stage | ms / launch | share of step | algorithmic MB | achieved GB/s | frac of measured HBM peak |")
print("|---|---:|---:|---:|---:|---:|")
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['ms_per_launch']):
    ab=v.get('algorithmic_bytes')
    print(f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | {ab/1e6:.0f} | {v['achieved_gbs']:.0f} | {v['frac_of_hbm_peak']:.3f} |" if ab else f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | | | |")
print()
print(f"step {d['ms_per_step']:.4f} ms ({d['value']:.1f} frames/s), with stage events {d['ms_per_step_with_stage_events']:.4f} ms, e2e {d['e2e']['ms_per_step']:.4f} ms ({d['e2e']['value']:.1f} frames/s), forward {d['forward']['ms_per_frame']:.4f} ms, edit {d['edit']['ms_per_frame']:.4f} ms; clocks {d['clocks']}")
r=json.load(open('gpurun_out/bench_ref_${tag}.json'))
it=d.get('train_iteration')
if it:
    print()
    print(f"full training iteration (section 5): {it['ms_per_iteration']:.4f} ms ({it['value']:.1f} iterations/s); reference-style composition {r['train_iteration']['ms_per_iteration']:.4f} ms ({r['train_iteration']['value']:.1f} iterations/s)")
    print()
    print("| iteration stage | ms / launch | share | algorithmic MB | frac of measured HBM peak |")
    print("|---|---:|---:|---:|---:|")
    for k,v in sorted(it['stages'].items(), key=lambda kv:-kv[1]['ms_per_launch']):
        ab=v.get('algorithmic_bytes')
        print(f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | {ab/1e6:.0f} | {v['frac_of_hbm_peak']:.3f} |" if ab else f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | | |")
    print()
print(f"reference arm: step {r['ms_per_step']:.4f} ms ({r['value']:.1f} frames/s), e2e {r['e2e']['value']:.1f} frames/s, forward {r['forward']['ms_per_frame']:.4f} ms, edit {r['edit']['ms_per_frame']:.4f} ms")
PY
} > profiles/r1_launches.md
{
echo "# Round 1 — \`ncu --set full\` per-kernel summary (B200)"
echo
echo "Command: \`ncu --set full --clock-control none --import-source on -k regex:^(blend|emit|geometry|preprocess|bucket_sort|big_bucket|large_tiles|tile_scan|depth_hist|bucket_lut|l1_kernel) -s 32 -c 13 -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize\` (one GPU)."
echo "Extracted with \`scripts/ncu_summary.py\`; DRAM bytes per launch are also in \`profiles/dram_traffic.json\` (read by bench.py for \`roofline.traffic\`).  Percentages are ncu's (of its own peaks)."
echo
cat /tmp/ncu_${tag}.md
if [ -f gpurun_out/prof_iter_${tag}.ncu-rep ]; then
echo "## Kernels of the full training iteration (bench section 5; \`-k regex:^(photometric|adam_kernel|densify_stats|mesh_restrict|mesh_bind) -s 14 -c 7\`)"
echo
python scripts/ncu_summary.py gpurun_out/prof_iter_${tag}.ncu-rep
fi
echo "## Opcode mix of the two blend kernels (scripts/sass_profile.py)"
echo
echo '```'
python scripts/sass_profile.py gpurun_out/prof_${tag}.ncu-rep blend_backward_pairs_kernel 14
python scripts/sass_profile.py gpurun_out/prof_${tag}.ncu-rep blend_forward_pairs_kernel 14
echo '```'
echo
echo "## Hottest source lines (scripts/line_profile.py)"
echo
echo '```'
python scripts/line_profile.py gpurun_out/prof_${tag}.ncu-rep blend_backward_pairs_kernel 12
python scripts/line_profile.py gpurun_out/prof_${tag}.ncu-rep blend_forward_pairs_kernel 10
echo '```'
} > profiles/r1_ncu_summary.md
rm -f profiles/r1_bench_ours_prelim.json
ls -la profiles

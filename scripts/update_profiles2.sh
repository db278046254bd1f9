#!/bin/bash
# Regenerate profiles/r2_* from the artefacts a `scripts/gpu_check2.sh <tag>` run left in gpurun_out/.
tag=${1:?tag}
set -e
cp gpurun_out/launches_${tag}.csv profiles/r2_launches.csv
cp gpurun_out/launches_ref_${tag}.csv profiles/r2_reference_launches.csv
cp gpurun_out/bench_ours_${tag}.json profiles/r2_bench_ours.json
cp gpurun_out/bench_ref_${tag}.json profiles/r2_bench_reference.json
for v in GM_BLEND_BWD_pairs GM_BLEND_FWD_tile GM_BLEND_BWD_mma GM_BLEND_FWD_ring GM_PDL_0; do cp gpurun_out/bench_ours_${tag}_$v.json profiles/r2_bench_ours_$v.json; done
[ -f gpurun_out/bench_ours_r2_n8.json ] && tail -1 gpurun_out/bench_ours_r2_n8.json > profiles/r2_bench_ours_n8.json
[ -f gpurun_out/bench_ours_m2_n2.json ] && tail -1 gpurun_out/bench_ours_m2_n2.json > profiles/r2_bench_ours_n2.json
python scripts/ncu_summary.py gpurun_out/prof_${tag}.ncu-rep --json profiles/dram_traffic.json > /tmp/ncu_${tag}.md
python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.md
{
echo "# Round 2 — ncu launch list (B200, \`bench.py --steps 3 --warmup 2\`, 1M Gaussians @ 1920x1080)"
echo
echo "Command (scripts/gpu_check2.sh): \`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize\`."
echo "Raw list: \`profiles/r2_launches.csv\`.  Times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event stage shares of \`bench.py\` below, not absolutes.  All sections of the bench are in the list (training frames incl. the CUDA-graph replays, forward-only frames, edit frames, the training iteration)."
echo
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv 0
echo
python - <<PY
import csv, json
rows=[]
with open('gpurun_out/launches_${tag}.csv', newline='') as f:
    lines=[l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    us = v/1e3 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1e3)
    rows.append((r['Kernel Name'], us))
starts=[i for i,(n,_) in enumerate(rows) if 'depth_histogram_kernel' in n]
frames=[rows[a:b] for a,b in zip(starts, starts[1:]) if any('blend_backward' in n for n,_ in rows[a:b]) and not any('photometric' in n or 'adam' in n for n,_ in rows[a:b])]
if frames:
    seg=frames[min(2,len(frames)-1)]
    tot=sum(u for _,u in seg)
    d=json.load(open('gpurun_out/bench_ours_${tag}.json')); x=json.load(open('gpurun_out/bench_ours_${tag}_GM_PDL_0.json'))
    print(f"One training frame under ncu: {len(seg)} kernel launches, sum of kernel times {tot/1e3:.3f} ms (serialised, cold cache).  Measured step: {d['ms_per_step']:.3f} ms with programmatic dependent launch, {x['ms_per_step']:.3f} ms with plain stream-ordered launches (\`GM_PDL=0\`); sum of the CUDA-event stage times {sum(v['ms_per_launch'] for v in d['stages'].values()):.3f} ms.")
PY
echo
echo "CUDA-event stage times inside the timed training region of bench.py (\`profiles/r2_bench_ours.json\`, 50 steps):"
echo
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}.json'))
print("| stage | ms / launch | share of step | algorithmic MB | achieved GB/s | frac of measured HBM peak |")
print("|---|---:|---:|---:|---:|---:|")
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['ms_per_launch']):
    ab=v.get('algorithmic_bytes')
    print(f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | {ab/1e6:.0f} | {v['achieved_gbs']:.0f} | {v['frac_of_hbm_peak']:.3f} |" if ab else f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | | | |")
ssum=sum(v['ms_per_launch'] for v in d['stages'].values())
print()
print(f"step {d['ms_per_step']:.4f} ms ({d['value']:.1f} frames/s; per-step median {d['step_ms']['median_ms']:.4f}, p10 {d['step_ms']['p10_ms']:.4f}, p90 {d['step_ms']['p90_ms']:.4f}); sum of the stage times {ssum:.4f} ms; with stage events {d['ms_per_step_with_stage_events']:.4f} ms; as one CUDA graph launch {d['cuda_graph']['ms_per_step']:.4f} ms (host enqueue {d['cuda_graph']['host_enqueue_ms_per_step']*1e3:.0f} us vs {d['host_enqueue_ms_per_step']*1e3:.0f} us eager); e2e {d['e2e']['ms_per_step']:.4f} ms ({d['e2e']['value']:.1f} frames/s, {d['e2e']['h2d_bytes_per_step']/1e6:.1f} MB uploaded per step); forward {d['forward']['ms_per_frame']:.4f} ms, edit {d['edit']['ms_per_frame']:.4f} ms; clocks {d['clocks']}")
for name in ("GM_PDL_0", "GM_BLEND_BWD_pairs", "GM_BLEND_FWD_tile", "GM_BLEND_BWD_mma", "GM_BLEND_FWD_ring"):
    x=json.load(open(f'gpurun_out/bench_ours_${tag}_{name}.json'))
    print(f"- variant {name.replace('_', '=', 1) if name.startswith('GM_PDL') else name.replace('GM_BLEND_BWD_', 'GM_BLEND_BWD=').replace('GM_BLEND_FWD_', 'GM_BLEND_FWD=')}: step {x['ms_per_step']:.4f} ms, blend_forward {x['stages']['blend_forward']['ms_per_launch']:.4f}, blend_backward {x['stages']['blend_backward']['ms_per_launch']:.4f}, e2e {x['e2e']['ms_per_step']:.4f} ms")
r=json.load(open('gpurun_out/bench_ref_${tag}.json'))
it=d.get('train_iteration')
print()
print(f"full training iteration (section 5): {it['ms_per_iteration']:.4f} ms ({it['value']:.1f} iterations/s); reference-style composition {r['train_iteration']['ms_per_iteration']:.4f} ms ({r['train_iteration']['value']:.1f} iterations/s)")
print()
print("| iteration stage | ms / launch | share | algorithmic MB | frac of measured HBM peak |")
print("|---|---:|---:|---:|---:|")
for k,v in sorted(it['stages'].items(), key=lambda kv:-kv[1]['ms_per_launch']):
    ab=v.get('algorithmic_bytes')
    print(f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | {ab/1e6:.0f} | {v['frac_of_hbm_peak']:.3f} |" if ab else f"| {k} | {v['ms_per_launch']:.4f} | {100*v['share']:.1f}% | | |")
print()
b=r['reference_best']
print(f"reference arm: step {r['ms_per_step']:.4f} ms mean ({r['value']:.1f} frames/s; per-step median {r['step_ms']['median_ms']:.4f}, p10 {r['step_ms']['p10_ms']:.4f}, p90 {r['step_ms']['p90_ms']:.4f}, max {r['step_ms']['max_ms']:.2f}), best case (variant i, single-call forward over persistent chunks) {b['ms_per_step']:.4f} ms ({b['value']:.1f} frames/s), e2e {r['e2e']['value']:.1f} frames/s, forward {r['forward']['ms_per_frame']:.4f} ms, edit {r['edit']['ms_per_frame']:.4f} ms; e2e loss {r['e2e']['loss']:.7f} vs ours {d['e2e']['loss']:.7f} (guard: {d['e2e'].get('loss_matches_other_arm')})")
PY
} > profiles/r2_launches.md
{
echo "# Round 2 — ncu launch list of the REFERENCE arm (\`bench.py --impl reference --steps 3 --warmup 2\`)"
echo
echo "Command: \`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ref_${tag}.csv python bench.py --impl reference --steps 3 --warmup 2\`.  Raw list: \`profiles/r2_reference_launches.csv\`."
echo "The process maps only \`oracle/_ref/libRefCudaRasterizer.so\` (the unmodified reference sources rebuilt for sm_100a); every other launch in the list is a torch kernel standing in for the Jittor ops of the reference's glue (zero-fills of the chunks and gradient tensors, the elementwise L1, the caching allocator's work is host side and not in the list)."
echo
python scripts/summarize_launches.py gpurun_out/launches_ref_${tag}.csv 0
echo
python - <<PY
import csv, re, json
rows=[]
with open('gpurun_out/launches_ref_${tag}.csv', newline='') as f:
    lines=[l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    us = v/1e3 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1e3)
    rows.append((int(r['ID']), r['Kernel Name'], us))
# one training frame = from one duplicateWithKeys launch (one per forward) to the next, if it holds a backward
starts=[i for i,(_,n,_) in enumerate(rows) if 'duplicateWithKeys' in n]
frames=[]
for a,b in zip(starts, starts[1:]):
    seg=rows[a:b]
    if any('computeCov2DCUDA' in n for _,n,_ in seg):
        frames.append(seg)
if frames:
    seg=frames[min(2,len(frames)-1)]
    tot=sum(u for _,_,u in seg)
    ras=sum(u for _,n,u in seg if 'cub::' in n or 'renderCUDA' in n or 'preprocessCUDA' in n or 'duplicateWithKeys' in n or 'identifyTileRanges' in n or 'computeCov2DCUDA' in n or 'checkFrustum' in n)
    r=json.load(open('gpurun_out/bench_ref_${tag}.json'))
    med=r['step_ms']['median_ms']
    gap=med-tot/1e3
    tail=(f"so about {gap:.2f} ms per step is not kernel time (the blocking read of num_rendered in the middle of the forward, chunk allocation, launch gaps)."
          if gap > 0.1 else
          "so kernel time accounts for the whole step: the host-side protocol (blocking read of num_rendered, fresh chunks, nine zero-filled gradient tensors) is hidden behind the 6.5 ms of the two renderCUDA launches (ncu's serialised, cold-cache kernel times run slightly longer than in the pipelined run).")
    print(f"One training frame of the reference arm under ncu: {len(seg)} kernel launches, sum of kernel times {tot/1e3:.3f} ms, of which the reference's own CUDA (rasterizer + CUB sort / scan) {ras/1e3:.3f} ms and torch fills / elementwise ops {(tot-ras)/1e3:.3f} ms; the arm's measured step is {med:.3f} ms (median), " + tail)
    b=r['reference_best']
    print()
    print(f"That is also why the reference's best case (variant i: single-call forward over persistent, pre-sized chunks, no chunk memsets, one gradient slab) gains so little: {b['ms_per_step']:.3f} ms against {med:.3f} ms.")
PY
} > profiles/r2_reference_launches.md
{
echo "# Round 2 — \`ncu --set full\` per-kernel summary (B200)"
echo
echo "Command: \`ncu --set full --clock-control none --import-source on -k regex:^(blend|emit|geometry|preprocess|bucket_sort|big_bucket|large_tiles|tile_scan|depth_hist|bucket_lut|l1_kernel) -s 32 -c 13 -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize\` (one GPU)."
echo "Extracted with \`scripts/ncu_summary.py\`; DRAM bytes per launch are also in \`profiles/dram_traffic.json\` (read by bench.py for \`roofline.traffic\`).  Percentages are ncu's (of its own peaks)."
echo
cat /tmp/ncu_${tag}.md
echo "## Kernels of the full training iteration (bench section 5; \`-k regex:^(photometric|adam_kernel|densify_stats|mesh_restrict|mesh_bind) -s 14 -c 7\`)"
echo
python scripts/ncu_summary.py gpurun_out/prof_iter_${tag}.ncu-rep
echo "## The blend kernels that were the defaults until late in round 2 (DESIGN.md 8)"
echo
echo "\`GM_BLEND_BWD=pairs GM_BLEND_FWD=tile\` (backward with the 18-value shuffle butterfly, one 256-thread block per tile in both passes), same capture command with the variables set:"
echo
python scripts/ncu_summary.py gpurun_out/prof_prev_${tag}.ncu-rep
echo "## Opcode mix of the blend kernels (scripts/sass_profile.py): default kernels, then the previous ones"
echo
echo '```'
python scripts/sass_profile.py gpurun_out/prof_${tag}.ncu-rep blend_backward_cols_kernel 16
python scripts/sass_profile.py gpurun_out/prof_prev_${tag}.ncu-rep blend_backward_pairs_kernel 16
python scripts/sass_profile.py gpurun_out/prof_${tag}.ncu-rep blend_forward_pairs_kernel 14
echo '```'
echo
echo "## Hottest source lines (scripts/line_profile.py)"
echo
echo '```'
python scripts/line_profile.py gpurun_out/prof_${tag}.ncu-rep blend_backward_cols_kernel 14
python scripts/line_profile.py gpurun_out/prof_${tag}.ncu-rep blend_forward_pairs_kernel 10
echo '```'
} > profiles/r2_ncu_summary.md
ls -la profiles

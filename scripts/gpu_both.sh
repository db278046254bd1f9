#!/bin/bash
# Both bench arms (reference first, as the driver does) + the new parity tests.  Usage: bash scripts/gpu_both.sh <tag> [pytest -k expr]
tag=${1:-both}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_${tag}.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err
echo "bench ref exit $?"; tail -3 gpurun_out/bench_ref_${tag}.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench ours exit $?"; tail -3 gpurun_out/bench_ours_${tag}.err
python - <<PY
import json
r=json.load(open('gpurun_out/bench_ref_${tag}.json')); d=json.load(open('gpurun_out/bench_ours_${tag}.json'))
print('ref : step', round(r['ms_per_step'],4), 'best', r['reference_best'] and round(r['reference_best']['ms_per_step'],4), 'e2e', r['e2e'], 'fwd', round(r['forward']['ms_per_frame'],4), 'edit', round(r['edit']['ms_per_frame'],4), 'iter', round(r['train_iteration']['ms_per_iteration'],4), 'shard_identical', r.get('shard_identical'))
print('ours: step', round(d['ms_per_step'],4), d['step_ms'], 'e2e', d['e2e'], 'fwd', round(d['forward']['ms_per_frame'],4), 'edit', round(d['edit']['ms_per_frame'],4), 'iter', round(d['train_iteration']['ms_per_iteration'],4), 'shard_identical', d.get('shard_identical'))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['stages'].items()))
print('cpu_baseline', d.get('cpu_baseline'))
PY
python - <<'PY'
# which shared objects did each arm map?
import subprocess, sys
for impl in ("reference",):
    code = "import sys; sys.argv=['bench.py','--impl','reference','--steps','2','--warmup','1','--gaussians','20000']; import runpy\ntry:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit: pass\nprint('MAPS', sorted(set(l.split()[-1] for l in open('/proc/self/maps') if ('Rasterizer' in l))))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print([l for l in out.stdout.splitlines() if l.startswith('MAPS')], out.stderr[-300:])
PY

#!/bin/bash
# Bench only (no tests, no profiler): one JSON + a one-line stage summary.  Usage: bash scripts/gpu_bench_only.sh [tag]
tag=${1:-b}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ${BENCH_FLAGS} > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_ours_${tag}.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}.json'))
print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'fwd', round(d['forward']['ms_per_frame'],4), 'edit', round(d['edit']['ms_per_frame'],4), 'iter', round(d['train_iteration']['ms_per_iteration'],4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['stages'].items()))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['train_iteration']['stages'].items()))
PY

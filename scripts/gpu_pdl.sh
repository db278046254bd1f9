#!/bin/bash
tag=${1:-pdl}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size and not config5_full" > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_${tag}.log
for v in 1 0; do
GM_PDL=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}_pdl$v.json 2> gpurun_out/bench_ours_${tag}_pdl$v.err
echo "bench GM_PDL=$v exit $?"; tail -3 gpurun_out/bench_ours_${tag}_pdl$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}_pdl$v.json'))
print('step', round(d['ms_per_step'],4), d['step_ms'], 'e2e', round(d['e2e']['ms_per_step'],4), 'fwd', round(d['forward']['ms_per_frame'],4), 'edit', round(d['edit']['ms_per_frame'],4), 'iter', round(d['train_iteration']['ms_per_iteration'],4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['stages'].items()), 'sum', round(sum(v['ms_per_launch'] for v in d['stages'].values()),4))
PY
done

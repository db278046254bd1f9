#!/bin/bash
tag=${1:-fw}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not config5_full and not full_size" > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_${tag}.log
for v in ring barrier; do
GM_BLEND_FWD=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}_$v.json 2> gpurun_out/bench_ours_${tag}_$v.err
echo "bench GM_BLEND_FWD=$v exit $?"; tail -3 gpurun_out/bench_ours_${tag}_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}_$v.json'))
print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'fwd', round(d['forward']['ms_per_frame'],4), 'edit', round(d['edit']['ms_per_frame'],4), 'iter', round(d['train_iteration']['ms_per_iteration'],4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['stages'].items()), 'sum', round(sum(v['ms_per_launch'] for v in d['stages'].values()),4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['train_iteration']['stages'].items()))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:blend_forward" -s 2 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu exit $?"

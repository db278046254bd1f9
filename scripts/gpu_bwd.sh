#!/bin/bash
# Backward-blend iteration pass: the gradient parity tests, a bench run, and one ncu capture of the kernel.
tag=${1:-bw}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or full_size or alternative or config2 or bg_render or surface or train_step" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_ours_${tag}.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}.json'))
print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'fwd', round(d['forward']['ms_per_frame'],4), 'edit', round(d['edit']['ms_per_frame'],4), 'iter', round(d['train_iteration']['ms_per_iteration'],4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['stages'].items()))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k,v in d['train_iteration']['stages'].items()))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:blend_backward" -s 2 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu exit $?"

#!/bin/bash
# GPU-box pass for kernel A/B work: parity tests, then bench with the default kernels and with the variants named in
# AB_ENVS (semicolon-separated `VAR=value` lists).  Usage: AB_ENVS="GM_BLEND_BWD=pairs" bash scripts/gpu_ab.sh <tag> [pytest -k expr]
tag=${1:-ab}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_${tag}.log
summ() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'fwd', round(d['forward']['ms_per_frame'], 4),
      'edit', round(d['edit']['ms_per_frame'], 4), 'iter', round(d['train_iteration']['ms_per_iteration'], 4))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d['stages'].items()))
print(' '.join(f"{k}={v['ms_per_launch']:.4f}" for k, v in d['train_iteration']['stages'].items()))
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench (default) exit $?"; tail -3 gpurun_out/bench_ours_${tag}.err; summ gpurun_out/bench_ours_${tag}.json
IFS=';' read -ra VARIANTS <<< "${AB_ENVS}"
i=0
for v in "${VARIANTS[@]}"; do
    [ -z "$v" ] && continue
    i=$((i+1))
    env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}_v${i}.json 2> gpurun_out/bench_ours_${tag}_v${i}.err
    echo "bench ($v) exit $?"; tail -3 gpurun_out/bench_ours_${tag}_v${i}.err; summ gpurun_out/bench_ours_${tag}_v${i}.json
done

#!/bin/bash
# N-GPU pass: view-parallel parity test + the bench at N ranks.  Usage: bash scripts/gpu_multi.sh N [tag]
n=${1:-2}; tag=${2:-m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_${tag}.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_multi_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus $n --steps 30 --warmup 5 \
    > gpurun_out/bench_ours_n${n}_${tag}.json 2> gpurun_out/bench_ours_n${n}_${tag}.err
echo "bench exit $?"; tail -5 gpurun_out/bench_ours_n${n}_${tag}.err | cut -c1-400
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_ours_n${n}_${tag}.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'iter', d['train_iteration']['value'], d['train_iteration']['ms_per_iteration'])
print(json.dumps(d.get('view_parallel'), indent=1))
PY

#!/bin/bash
# 2-GPU pass: the multi-GPU parity test (view-parallel training incl. densification) and a short 2-rank bench.
tag=${1:-m2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${tag}.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_${tag}.log 2>&1
echo "pytest multi exit $?"; tail -15 gpurun_out/pytest_multi_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_ours_${tag}_n2.json 2> gpurun_out/bench_ours_${tag}_n2.err
echo "bench n2 exit $?"; tail -3 gpurun_out/bench_ours_${tag}_n2.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ours_${tag}_n2.json'))
print('value', round(d['value'],1), 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'shard_identical', d['shard_identical'], 'replicas_identical', d['replicas_identical'])
print(json.dumps(d.get('view_parallel'), indent=0)[:1500])
PY

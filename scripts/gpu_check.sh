#!/bin/bash
# One GPU-box pass: parity tests, both bench arms, and the ncu launch list of a short bench run.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${tag}.log
tail -15 gpurun_out/pytest_gpu_${tag}.log
timeout 600 python bench.py --steps 50 --warmup 5 ${BENCH_FLAGS} > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench ours exit $?"; tail -c 3000 gpurun_out/bench_ours_${tag}.json; tail -5 gpurun_out/bench_ours_${tag}.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err
echo "bench ref exit $?"; tail -c 2000 gpurun_out/bench_ref_${tag}.json; tail -5 gpurun_out/bench_ref_${tag}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_bench_${tag}.log 2>&1
echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(blend|emit|geometry|preprocess|bucket_sort|big_bucket|large_tiles|tile_scan|depth_hist|bucket_lut|l1_kernel)" -s 32 -c 13 -f \
    -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(photometric|adam_kernel|densify_stats|mesh_restrict|mesh_bind)" -s 14 -c 7 -f \
    -o gpurun_out/prof_iter_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_iter_${tag}.log 2>&1
echo "ncu iteration kernels exit $?"

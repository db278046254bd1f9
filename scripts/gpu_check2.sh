#!/bin/bash
# Round-2 evidence pass on ONE GPU: full parity suite, both bench arms (reference first, as the driver runs them), ncu launch
# lists of both arms, ncu --set full of the kernels of the frame, SASS opcode histograms, compute-sanitizer.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check2.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${tag}.log
tail -5 gpurun_out/pytest_gpu_${tag}.log
rm -f /tmp/gm_bench_e2e_loss.json
timeout 900 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err
echo "bench ref exit $?"; tail -3 gpurun_out/bench_ref_${tag}.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_ours_${tag}.json 2> gpurun_out/bench_ours_${tag}.err
echo "bench ours exit $?"; tail -3 gpurun_out/bench_ours_${tag}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_bench_${tag}.log 2>&1
echo "ncu launch list (ours) exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ref_${tag}.csv \
    python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/ncu_bench_ref_${tag}.log 2>&1
echo "ncu launch list (reference) exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(blend|emit|geometry|preprocess|bucket_sort|big_bucket|large_tiles|tile_scan|depth_hist|bucket_lut|l1_kernel)" -s 32 -c 13 -f \
    -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(photometric|adam_kernel|densify_stats|mesh_restrict|mesh_bind)" -s 14 -c 7 -f \
    -o gpurun_out/prof_iter_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_iter_${tag}.log 2>&1
echo "ncu iteration kernels exit $?"
GM_BLEND_BWD=pairs GM_BLEND_FWD=tile timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^blend_" -s 4 -c 2 -f \
    -o gpurun_out/prof_prev_${tag} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_prev_${tag}.log 2>&1
echo "ncu previous default blend kernels exit $?"
for v in "GM_BLEND_BWD=pairs" "GM_BLEND_FWD=tile" "GM_BLEND_BWD=mma" "GM_BLEND_FWD=ring" "GM_PDL=0"; do
  env $v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_${tag}_${v//=/_}.json 2> gpurun_out/bench_ours_${tag}_${v//=/_}.err
  echo "bench $v exit $?"
done
bash scripts/sanitize.sh > gpurun_out/sanitize_${tag}.log 2>&1
tail -12 gpurun_out/sanitize_${tag}.log

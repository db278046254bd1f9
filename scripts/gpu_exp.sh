#!/bin/bash
# Kernel-experiment pass: scripts/time_step.py once per variant in EXP_ENVS (semicolon-separated `VAR=value` lists; an
# empty entry = the default kernels).  Usage: EXP_ENVS=";GM_EXP=1" bash scripts/gpu_exp.sh <tag> [pytest -k expr | none]
tag=${1:-exp}
mkdir -p gpurun_out
if [ "${2}" != "none" ]; then
    timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/pytest_gpu_${tag}.log 2>&1
    echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_${tag}.log
fi
IFS=';' read -ra VARIANTS <<< "${EXP_ENVS};"
: > gpurun_out/exp_${tag}.jsonl
for v in "${VARIANTS[@]}" ""; do
    [ -z "$v" ] && [ -n "$done_default" ] && continue
    [ -z "$v" ] && done_default=1
    env $v timeout 300 python scripts/time_step.py --steps 30 ${EXP_ARGS} >> gpurun_out/exp_${tag}.jsonl 2> gpurun_out/exp_${tag}.err
    echo "variant [$v] exit $?"; tail -1 gpurun_out/exp_${tag}.jsonl; tail -2 gpurun_out/exp_${tag}.err
done

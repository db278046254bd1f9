#!/bin/bash
# ncu --set full on the launches of one kernel (regex) inside a short bench run.  Usage: bash scripts/ncu_one.sh <regex> <tag> [skip] [count]
re=${1:?kernel regex}; tag=${2:-one}; skip=${3:-4}; cnt=${4:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${re}" -s ${skip} -c ${cnt} -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-presize > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_${tag}.log

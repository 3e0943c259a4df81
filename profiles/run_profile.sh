#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, one --set full capture of the dominant kernel.
# usage: profiles/run_profile.sh <tag> [pairs_for_ncu] [kernel regex]
TAG=${1:-r01}; NP=${2:-5328}; KRE=${3:-gotoh_packed}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 2500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --pairs $NP --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --pairs $NP --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log

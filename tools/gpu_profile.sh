#!/bin/bash
# Runs on the GPU box under gpurun: bench line, ncu launch list, ncu --set full of the person kernel.
# Each ncu pass only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
SHORT="python bench.py --short --steps 6 --warmup 3"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'person_sweep|global_draw' -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:person_sweep -s 4 -c 1 \
    -o gpurun_out/prof_person_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
ls -la gpurun_out

#!/bin/bash
# Round-2 profiling pass (one GPU): launch list of a short bench run, ncu --set full of the f32 fast person kernel and of the f64 person kernel.
# Each ncu pass only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
SHORT="python bench.py --short --steps 6 --warmup 3"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'person_sweep|global_draw' -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:person_sweep_fast -s 4 -c 1 \
    -o gpurun_out/prof_person_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full f32 rc=$?"
SHORT64="python bench.py --short --dtype f64 --steps 3 --warmup 3"
$SHORT64 > gpurun_out/plain64_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:person_sweep_kernel -s 2 -c 1 \
    -o gpurun_out/prof_person64_${TAG} $SHORT64 > gpurun_out/ncu_full64_${TAG}.log 2>&1
echo "full f64 rc=$?"
ls -la gpurun_out | tail -12

#!/bin/bash
TAG=${1:-r1b}; TPP=${2:-2}
export ERIRT_TPP=$TPP
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "one_sweep or multi_sweep or ragged or graph or variants" 2>&1 | tail -3
SHORT="python bench.py --short --steps 6 --warmup 3"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 && tail -c 600 gpurun_out/plain_${TAG}.log &&
ncu --set full --clock-control none --import-source on -k regex:person_sweep -s 4 -c 1 \
    -o gpurun_out/prof_person_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"

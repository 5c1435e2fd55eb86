#!/bin/bash
# GPU check of a change: selected parity tests (-k expression in $1), then the full bench line into gpurun_out/bench_$2.json
set -u
mkdir -p gpurun_out
TAG=${2:-chk}
timeout 900 python -m pytest tests -m gpu -q -x -k "${1:-checkpoint or y8 or one_sweep or moments or fixture or graph}" 2>&1 | tail -15
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.err; python - <<PY
import json
for l in open("gpurun_out/bench_${TAG}.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"])
PY

#!/bin/bash
# scaling run on one box: bench.py at N = 8, 4, 2 GPUs (persons of the one chain sharded, fused peer-memory exchange), full JSON line
for n in ${1:-8 4 2}; do
  echo "== gpus=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps ${2:-100} --warmup 10 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_scale_${n}.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'roofline frac', round(d['roofline']['frac'], 3), 'kernel ms', round(d['roofline']['kernel_ms'], 4), d['clocks'])
"
done

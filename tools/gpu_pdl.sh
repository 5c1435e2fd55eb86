#!/bin/bash
# A/B of programmatic dependent launch (ERIRT_PDL=1/0): parity subset, C5 timed sweeps, small configs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "${1:-one_sweep or multi_sweep or ragged or graph or variants or checkpoint or fixture or loglik}" 2>&1 | tail -8
short() {
  python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'long', (d.get('long_run') or {}).get('ms_per_step'), 'sm_mhz', d['clocks']['sm_mhz'])
    elif l: print(l[:300])
"
}
for pdl in 1 0 1 0; do
  ERIRT_PDL=$pdl timeout 300 python bench.py --short --steps 200 --warmup 20 2>&1 | short "C5 pdl=$pdl"
done
for pdl in 1 0; do
  echo "== configs pdl=$pdl"
  ERIRT_PDL=$pdl timeout 600 python bench.py --sub configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        for k,v in d.items():
            if isinstance(v, dict): print(k, 'f32', round(v.get('gpu_f32_sweeps_per_s',0)), 'f64', round(v.get('gpu_f64_sweeps_per_s',0)))
"
done

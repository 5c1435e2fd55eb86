#!/bin/bash
# A/B timing of library builds on one box: bench.py --short (timed sweeps only) for every .so given, three runs each
for lib in "$@"; do
  for r in 1 2 3; do
    ERIRT_B200_LIB=$PWD/extendedrtirtmodeling.jl_b200/$lib timeout 300 python bench.py --short --steps 100 --warmup 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'ms/step', round(d['ms_per_step'],4), 'sm_mhz', d['clocks']['sm_mhz'])
    elif l: print(l[:200])
"
  done
done

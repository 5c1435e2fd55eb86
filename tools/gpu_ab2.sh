#!/bin/bash
# A/B timing of library builds with per-variant environment: every argument is "lib.so[,VAR=value...]"; bench.py --short, two runs each
for spec in "$@"; do
  lib=${spec%%,*}; envs=""; [ "$spec" != "$lib" ] && envs=$(echo "${spec#*,}" | tr ',' ' ')
  for r in 1 2; do
    env $envs ERIRT_B200_LIB=$PWD/extendedrtirtmodeling.jl_b200/$lib timeout 300 python bench.py --short --steps 100 --warmup 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$spec', 'ms/step', round(d['ms_per_step'],4), 'sm_mhz', d['clocks']['sm_mhz'])
    elif l and 'Traceback' in l or 'rror' in l: print(l[:300])
"
  done
done

#!/bin/bash
# A/B: an older build of the library vs the current one, same box, TPP=2
for lib in old_f476.so extendedrtirtmodeling.jl_b200/liberirt_b200.so old_f476.so extendedrtirtmodeling.jl_b200/liberirt_b200.so; do
  echo "== $lib"; ERIRT_B200_LIB=$PWD/$lib ERIRT_TPP=2 timeout 300 python bench.py --short --steps 60 --warmup 5 2>&1 | grep -o "\"ms_per_step\": [0-9.]*\|rror.*" | head -3
done

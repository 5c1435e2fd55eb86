#!/bin/bash
# time-attribution builds: what the sweep costs with one phase of the person kernel removed (results are NOT valid samplers)
for lib in extendedrtirtmodeling.jl_b200/liberirt_b200.so diag_1.so diag_2.so diag_4.so diag_7.so; do
  echo "== $lib"; ERIRT_B200_LIB=$PWD/$lib timeout 300 python bench.py --short --steps 40 --warmup 4 2>&1 | grep -o "\"ms_per_step\": [0-9.]*\|rror.*" | head -3
done

#!/usr/bin/env python
"""Build diag_tick.so: the library with clock64() phase timers inserted into person_sweep_kernel (diagnostic only, never shipped).
Phases are delimited by the kernel's own landmarks; read the result with tools/gpu_ticks.py on the GPU box."""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "extendedrtirtmodeling.jl_b200", "csrc")
lines = open(os.path.join(CSRC, "person.cuh")).read().split("\n")

def find(sub, start=0):
    for i in range(start, len(lines)):
        if sub in lines[i]:
            return i
    raise KeyError(sub)

ins = []
i_loop = find("for (int tile = blockIdx.x; tile < A.n_tiles;")
ins.append((i_loop + 1, "TICK_START();"))
i_wait = find("mbar_wait(s_bar, parity);")
ins += [(i_wait, "TICK(0);"), (i_wait + 1, "TICK(1);")]
i_s1 = find("__syncthreads();", i_wait)
ins += [(i_s1, "TICK(2);"), (i_s1 + 1, "TICK(3);")]
i_s2 = find("__syncthreads();", i_s1 + 1)
ins += [(i_s2, "TICK(4);"), (i_s2 + 1, "TICK(5);")]
i_ll = find("acc_ll_bern += (double)ll_tile;")
ins.append((i_ll, "TICK(6);"))
i_s3 = find("__syncthreads();", i_ll)
ins += [(i_s3, "TICK(7);"), (i_s3 + 1, "TICK(8);")]
i_pe = find("// ---- per-item statistics: thread per (item group, person class)")
ins.append((i_pe, "TICK(9);"))
i_fl = find("if ((tiles_done % STAT_FLUSH_TILES) == STAT_FLUSH_TILES - 1) flush_item_stats();")
ins.append((i_fl, "TICK(10);"))
i_fe = find("fence_proxy_async();", i_fl)
ins.append((i_fe, "TICK(11);"))
i_s6 = find("__syncthreads();", i_fe)
ins.append((i_s6 + 1, "TICK(12);"))
for pos, txt in sorted(ins, reverse=True):
    lines.insert(pos, "    " + txt)
out = "\n".join(lines)
hdr = """
#define NTICK 16
__device__ unsigned long long g_ticks[NTICK];
#define TICK_START() long long _t0 = clock64(); long long _t1
#define TICK(n) do { _t1 = clock64(); if ((threadIdx.x & 31) == 0) _tk[n] += (unsigned long long)(_t1 - _t0); _t0 = _t1; } while (0)
"""
out = out.replace("constexpr int STAT_FLUSH_TILES = 8;", hdr + "constexpr int STAT_FLUSH_TILES = 8;")
out = out.replace("  int tiles_done = 0;\n  for (int tile = blockIdx.x;", "  unsigned long long _tk[NTICK] = {0};\n  int tiles_done = 0;\n  for (int tile = blockIdx.x;")
out = out.replace("  flush_item_stats();\n\n  // ---- flush CTA accumulators ----",
                  "  flush_item_stats();\n  if ((threadIdx.x & 31) == 0) for (int n = 0; n < NTICK; ++n) atomicAdd(&g_ticks[n], _tk[n]);\n\n  // ---- flush CTA accumulators ----")
bd = "/tmp/tickbuild"
shutil.rmtree(bd, ignore_errors=True)
shutil.copytree(CSRC, bd)
open(os.path.join(bd, "person.cuh"), "w").write(out)
cu = open(os.path.join(bd, "erirt_b200.cu")).read().replace('#include "../../include/erirt_b200.h"', f'#include "{ROOT}/include/erirt_b200.h"')
cu += """
extern "C" int erirt_diag_ticks(unsigned long long* out, int reset) {
  unsigned long long h[NTICK];
  cudaMemcpyFromSymbol(h, erirt::g_ticks, sizeof(h));
  for (int i = 0; i < NTICK; ++i) out[i] = h[i];
  if (reset) { unsigned long long z[NTICK] = {0}; cudaMemcpyToSymbol(erirt::g_ticks, z, sizeof(z)); }
  return NTICK;
}
"""
open(os.path.join(bd, "erirt_b200.cu"), "w").write(cu)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-o", os.path.join(ROOT, "diag_tick.so"), os.path.join(bd, "erirt_b200.cu"), "-ldl"])
print("built diag_tick.so")

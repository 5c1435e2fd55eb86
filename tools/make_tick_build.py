#!/usr/bin/env python
"""Build diag_tick.so: the library with the clock64() phase timers of person_sweep_fast_kernel compiled in (-DERIRT_TICKS;
diagnostic only, never shipped).  Read the result with tools/gpu_ticks.py on the GPU box."""
import os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "extendedrtirtmodeling.jl_b200", "csrc")
src = open(os.path.join(CSRC, "erirt_b200.cu")).read() + """
#ifdef ERIRT_TICKS
extern "C" int erirt_diag_ticks(unsigned long long* out, int reset) {
  unsigned long long h[PF_NTICK];
  cudaMemcpyFromSymbol(h, erirt::g_ticks, sizeof(h));
  for (int i = 0; i < PF_NTICK; ++i) out[i] = h[i];
  if (reset) { unsigned long long z[PF_NTICK] = {0}; cudaMemcpyToSymbol(erirt::g_ticks, z, sizeof(z)); }
  return PF_NTICK;
}
extern "C" int erirt_diag_gticks(long long* out) {
  cudaMemcpyFromSymbol(out, erirt::g_gticks, 16 * sizeof(long long));
  return 16;
}
#endif
#ifdef ERIRT_TIMELINE
extern "C" int erirt_diag_timeline(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, erirt::g_timeline, sizeof(unsigned long long) * TL_SLOTS * TL_N);
  return TL_SLOTS * TL_N;
}
extern "C" int erirt_diag_cta_timeline(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, erirt::g_tl_cta, sizeof(unsigned long long) * TL_CTAS * 6);
  return TL_CTAS;
}
#endif
"""
tmp = os.path.join(CSRC, "_tick_build.cu")
open(tmp, "w").write(src)
try:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                           ("-DERIRT_CHECKS" if os.environ.get("ERIRT_CHECKS_BUILD") else "-DERIRT_TIMELINE" if os.environ.get("ERIRT_TIMELINE_BUILD") else "-DERIRT_TICKS")] + os.environ.get("ERIRT_NVCC_EXTRA", "").split() +
                          ["-o", os.path.join(ROOT, "diag_checks.so" if os.environ.get("ERIRT_CHECKS_BUILD") else "diag_timeline.so" if os.environ.get("ERIRT_TIMELINE_BUILD") else "diag_tick.so"), tmp, "-ldl"])
finally:
    os.remove(tmp)
print("built")

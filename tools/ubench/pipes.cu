// Micro-benchmark of the sm_100a issue/pipe rates the person kernel depends on (not part of the product):
// FFMA, FFMA2 (fma.rn.f32x2), IMAD.WIDE.U32, LOP3, MUFU.EX2, I2FP and two mixes, as warp-instructions per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a),"f"(b)); return r; }
template <int MODE>
__global__ void bench(float* out, long long* cyc, float s0, float s1, uint32_t m0) {
  float a[8]; uint32_t u[8]; unsigned long long d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x * 977u + i; d[i] = pk(a[i], a[i] + 1.f); }
  const unsigned long long S0 = pk(s0, s0), S1 = pk(s1, s1);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], s0, s1);
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(S0), "l"(S1));
      if (MODE == 2) { uint64_t p = (uint64_t)u[i] * m0; u[i] = (uint32_t)p ^ (uint32_t)(p >> 32); }  // IMAD.WIDE + LOP3
      if (MODE == 3) u[i] = (u[i] ^ m0) + (u[i] >> 3);  // ALU ops
      if (MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 5) { a[i] = (float)u[i]; u[i] = __float_as_uint(a[i]) ^ m0; }  // I2FP + LOP3
      if (MODE == 6) { a[i] = fmaf(a[i], s0, s1); u[i] = (u[i] ^ m0) + (uint32_t)i; }  // FFMA + LOP3/IADD mix
      if (MODE == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], s0, s1); a[i] = fmaf(a[i], s0, s1); a[i] = fmaf(a[i], s0, s1);
                       u[i] = (u[i] ^ m0) + (uint32_t)i; u[i] = (u[i] ^ m0) + (uint32_t)i;}  // 1 MUFU : 3 FFMA : 4 ALU
      if (MODE == 8) { uint64_t p = (uint64_t)u[i] * m0; u[i] = (uint32_t)(p >> 32) ^ m0 ^ (uint32_t)p; }  // IMAD.WIDE + 1 LOP3
      if (MODE == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(S0), "l"(S1)); u[i] = (u[i] ^ m0) + (uint32_t)i; }
    }
  }
  long long t1 = clock64();
  float acc = 0; uint32_t ua = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc += a[i]; ua ^= u[i]; ua ^= (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)ua;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int ops_per_inner, int warps_per_sm) {
  int nsm = 148; float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * nsm * 1024); cudaMalloc(&cyc, sizeof(long long) * nsm);
  bench<MODE><<<nsm, warps_per_sm * 32>>>(out, cyc, 1.0001f, 0.5f, 0x9E3779B9u);
  bench<MODE><<<nsm, warps_per_sm * 32>>>(out, cyc, 1.0001f, 0.5f, 0x9E3779B9u);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
  double winst = (double)ITERS * 8 * ops_per_inner * warps_per_sm;
  printf("%-34s warps/SM=%2d  cycles=%9.0f  warp-inst/clk/SM=%.3f\n", name, warps_per_sm, mean, winst / mean);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 12, 16, 32}) {
    run<0>("FFMA", 1, w); run<1>("FFMA2", 1, w); run<2>("IMAD.WIDE+LOP3 (2)", 2, w); run<8>("IMAD.WIDE+LOP3(3in) (2)", 2, w); run<3>("LOP3+SHF+IADD (3?)", 3, w);
    run<4>("MUFU.EX2", 1, w); run<5>("I2FP+LOP3 (2)", 2, w); run<6>("FFMA+LOP3+IADD (3)", 3, w); run<7>("MUFU+3FFMA+4ALU (8)", 8, w);
    run<9>("FFMA2+LOP3+IADD (3)", 3, w);
  }
  return 0;
}

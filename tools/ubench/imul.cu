// Micro-benchmark: ways to form the 64-bit product of Philox (32x32 -> hi, lo) on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void bench(uint32_t* out, long long* cyc, uint32_t m0, uint32_t m1) {
  uint32_t u[8], v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 977u + i; v[i] = threadIdx.x * 31u + i * 7u; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { uint64_t p = (uint64_t)u[i] * 0xD2511F53u; u[i] = (uint32_t)p; v[i] ^= (uint32_t)(p >> 32); }   // IMAD.WIDE imm + LOP3
      if (MODE == 1) { uint32_t hi = __umulhi(u[i], 0xD2511F53u); uint32_t lo = u[i] * 0xD2511F53u; u[i] = lo; v[i] ^= hi; }  // hi + lo separately
      if (MODE == 2) { u[i] = __umulhi(u[i], 0xD2511F53u) ^ v[i]; }   // IMAD.HI only
      if (MODE == 3) { u[i] = u[i] * 0xD2511F53u + v[i]; }            // IMAD lo only
      if (MODE == 4) { uint64_t p = (uint64_t)u[i] * m0; u[i] = (uint32_t)p; v[i] ^= (uint32_t)(p >> 32); }  // IMAD.WIDE reg
      if (MODE == 5) { asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(u[i]) : "h"((unsigned short)u[i]), "h"((unsigned short)v[i])); v[i] ^= u[i]; }
      if (MODE == 6) { // 4 philox-like rounds flavour: 2 wide + 2 lop3 (true Philox round)
        uint64_t p0 = (uint64_t)0xD2511F53u * u[i]; uint64_t p1 = (uint64_t)0xCD9E8D57u * v[i];
        u[i] = (uint32_t)(p1 >> 32) ^ (uint32_t)p0 ^ m0; v[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p1 ^ m1; }
    }
  }
  long long t1 = clock64();
  uint32_t ua = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) ua ^= u[i] ^ v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = ua;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
  int nsm = 148; uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4 * nsm * 1024); cudaMalloc(&cyc, 8 * nsm);
  bench<MODE><<<nsm, warps_per_sm * 32>>>(out, cyc, 0x9E3779B9u, 0xBB67AE85u);
  bench<MODE><<<nsm, warps_per_sm * 32>>>(out, cyc, 0x9E3779B9u, 0xBB67AE85u);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
  printf("%-40s warps/SM=%2d  cycles per 8-chain step per SMSP = %.2f\n", name, warps_per_sm, mean / ITERS / (warps_per_sm / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 16}) {
    run<0>("IMAD.WIDE imm + LOP3 (x8)", w); run<1>("umulhi + mul lo + LOP3 (x8)", w); run<2>("umulhi + LOP3 (x8)", w);
    run<3>("IMAD lo (x8)", w); run<4>("IMAD.WIDE reg + LOP3 (x8)", w); run<5>("mul.wide.u16 + LOP3 (x8)", w); run<6>("philox round: 2 WIDE + 2 LOP3 (x8)", w);
  }
}

// diagnostics.cuh -- convergence diagnostics of the traces, on the device.
//
// Replaces `Chains(...) |> ess_rhat` in checkConvergence (/root/reference/src/SimTools.jl:419-443; MCMCChains is a dependency that is
// not vendored, its estimator is stated here): rank-normalised split-chain bulk ESS and R-hat of Vehtari, Gelman, Simpson, Carpenter,
// Buerkner (2021).  For one column of draws x[t, chain]:
//   1. every chain is split in two halves (2 n_chain chains of n = floor(n_used / 2) draws);
//   2. the N = 2 n_chain n values are replaced by z = Phi^{-1}((r - 3/8) / (N + 1/4)), r = average rank;
//   3. W = mean within-chain variance, B = n var(chain means), var+ = W (n-1)/n + B/n, R-hat = sqrt(var+ / W);
//   4. rho_t = 1 - (W - mean_chains acov_t) / var+,  Geyer's initial monotone sequence over the pairs rho_2k + rho_2k+1,
//      tau = -1 + 2 sum pairs (>= 1 / log10 N),  ESS = N / tau.
// One CTA per column: bitonic sort of (value, position) in shared memory (global scratch when a column does not fit), the
// autocovariances by direct sums in blocks of blockDim lags, evaluated lazily -- the monotone sequence ends after a few
// autocorrelation times, so usually one block suffices.  The host-side numpy statement of the same estimator
// (diagnostics.py) is the checker of the GPU tests.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace erirt {

constexpr int DIAG_THREADS = 512;

struct DiagArgs {
  const double* x;    // [n_chain][n_cols][n_iter]  (Julia's [nIter, P, nChain] as erirt_get_trace lays it out)
  int64_t n_iter, n_cols, skip, n_used;  // draws [skip, skip + n_used) of every chain are used
  int n_chain;
  int64_t npad;       // power of two >= N = 2 n_chain floor(n_used / 2)
  double* keys;       // scratch [n_cols][npad]  (unused when the column fits in shared memory)
  uint32_t* idx;      // scratch [n_cols][npad]
  double* z;          // scratch [n_cols][N]
  int use_smem;
  double* ess;
  double* rhat;
};

__device__ __forceinline__ double diag_block_sum(double v, double* red) {  // all threads of the CTA call; returns the total to all
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < DIAG_THREADS / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(DIAG_THREADS) ess_rhat_kernel(const DiagArgs A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ double red[DIAG_THREADS / 32];
  __shared__ double s_rho[DIAG_THREADS];
  __shared__ double s_mean[64];
  __shared__ double s_geyer[4];  // tau, prev, done flag, finite/min/max scratch
  __shared__ int s_flag;
  const int tid = threadIdx.x;
  const int64_t col = blockIdx.x;
  const int64_t n = A.n_used / 2, m2 = 2 * (int64_t)A.n_chain, N = n * m2, npad = A.npad;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (n < 4 || m2 > 64) {
    if (tid == 0) { A.ess[col] = nan; A.rhat[col] = nan; }
    return;
  }
  double* keys = A.use_smem ? reinterpret_cast<double*>(dsm) : A.keys + col * npad;
  uint32_t* idx = A.use_smem ? reinterpret_cast<uint32_t*>(dsm + npad * sizeof(double)) : A.idx + col * npad;
  double* gz = A.z + col * N;

  // ---- 1. load the split chains; finite / constant check ----
  if (tid == 0) s_flag = 0;
  __syncthreads();
  double vmin = __longlong_as_double(0x7ff0000000000000LL), vmax = -vmin;
  bool bad = false;
  for (int64_t e = tid; e < npad; e += DIAG_THREADS) {
    double v = __longlong_as_double(0x7ff0000000000000LL);  // padding sorts to the end
    if (e < N) {
      const int64_t s = e / n, t = e % n;
      const int64_t l = s < A.n_chain ? s : s - A.n_chain;
      const int64_t draw = A.skip + (s < A.n_chain ? t : (A.n_used - n) + t);
      v = A.x[draw + A.n_iter * (col + A.n_cols * l)];
      if (!isfinite(v)) bad = true;
      vmin = fmin(vmin, v);
      vmax = fmax(vmax, v);
    }
    keys[e] = v;
    idx[e] = (uint32_t)e;
  }
  for (int o = 16; o; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (bad) atomicOr(&s_flag, 1);
  if ((tid & 31) == 0) { red[tid >> 5] = vmin; s_rho[tid >> 5] = vmax; }
  __syncthreads();
  if (tid == 0) {
    double lo = red[0], hi = s_rho[0];
    for (int w = 1; w < DIAG_THREADS / 32; ++w) { lo = fmin(lo, red[w]); hi = fmax(hi, s_rho[w]); }
    if (!(hi > lo)) s_flag |= 1;  // constant column (or NaN)
  }
  __syncthreads();
  if (s_flag) {
    if (tid == 0) { A.ess[col] = nan; A.rhat[col] = nan; }
    return;
  }

  // ---- 2. bitonic sort of (value, position), ascending ----
  for (int64_t k = 2; k <= npad; k <<= 1)
    for (int64_t j = k >> 1; j > 0; j >>= 1) {
      for (int64_t t = tid; t < npad / 2; t += DIAG_THREADS) {
        const int64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair (bit j clear)
        const int64_t p = i | j;
        const bool up = (i & k) == 0;
        const double a = keys[i], b = keys[p];
        if ((a > b) == up) {
          keys[i] = b; keys[p] = a;
          const uint32_t ia = idx[i]; idx[i] = idx[p]; idx[p] = ia;
        }
      }
      __syncthreads();
    }

  // ---- 3. average ranks -> normal scores, scattered back to (chain, draw) order ----
  for (int64_t i = tid; i < N; i += DIAG_THREADS) {
    const double v = keys[i];
    int64_t lo = i, hi = i;
    while (lo > 0 && keys[lo - 1] == v) --lo;
    while (hi + 1 < N && keys[hi + 1] == v) ++hi;
    const double r = 0.5 * (double)(lo + hi) + 1.0;
    gz[idx[i]] = normcdfinv((r - 0.375) / ((double)N + 0.25));
  }
  __threadfence_block();
  __syncthreads();
  double* z = gz;
  if (A.use_smem) {  // the sorted keys are no longer needed: the scores move into their place
    for (int64_t e = tid; e < N; e += DIAG_THREADS) keys[e] = gz[e];
    z = keys;
    __syncthreads();
  }

  // ---- 4. chain means, centring, W, B, R-hat ----
  double W = 0.0;
  for (int64_t s = 0; s < m2; ++s) {
    double a = 0.0;
    for (int64_t t = tid; t < n; t += DIAG_THREADS) a += z[s * n + t];
    const double mean = diag_block_sum(a, red) / (double)n;
    double q = 0.0;
    for (int64_t t = tid; t < n; t += DIAG_THREADS) {
      const double d = z[s * n + t] - mean;
      z[s * n + t] = d;
      q += d * d;
    }
    W += diag_block_sum(q, red) / (double)(n - 1);
    if (tid == 0) s_mean[s] = mean;
  }
  W /= (double)m2;
  __syncthreads();
  double mm = 0.0;
  for (int64_t s = 0; s < m2; ++s) mm += s_mean[s];
  mm /= (double)m2;
  double B = 0.0;
  for (int64_t s = 0; s < m2; ++s) B += (s_mean[s] - mm) * (s_mean[s] - mm);
  B = (double)n * B / (double)(m2 - 1);
  const double var_plus = W * (double)(n - 1) / (double)n + B / (double)n;
  if (tid == 0) A.rhat[col] = W > 0.0 ? sqrt(var_plus / W) : nan;

  // ---- 5. autocorrelations in blocks of lags, Geyer's initial monotone sequence ----
  if (tid == 0) { s_geyer[0] = -1.0; s_geyer[1] = __longlong_as_double(0x7ff0000000000000LL); s_geyer[2] = 0.0; }
  __syncthreads();
  for (int64_t base = 0; base + 1 < n; base += DIAG_THREADS) {
    const int64_t lag = base + tid;
    double rho = 0.0;
    if (lag < n) {
      double acc = 0.0;
      for (int64_t s = 0; s < m2; ++s) {
        const double* zs = z + s * n;
        double a0 = 0.0, a1 = 0.0;
        int64_t t = 0;
        for (; t + 1 < n - lag; t += 2) { a0 = fma(zs[t], zs[t + lag], a0); a1 = fma(zs[t + 1], zs[t + 1 + lag], a1); }
        if (t < n - lag) a0 = fma(zs[t], zs[t + lag], a0);
        acc += a0 + a1;
      }
      rho = 1.0 - (W - acc / ((double)n * (double)m2)) / var_plus;
      if (lag == 0) rho = 1.0;
    }
    s_rho[tid] = rho;
    __syncthreads();
    if (tid == 0) {
      double tau = s_geyer[0], prev = s_geyer[1];
      bool done = false;
      for (int64_t t = 0; t + 1 < DIAG_THREADS && base + t + 1 < n; t += 2) {
        double pair = s_rho[t] + s_rho[t + 1];
        if (pair < 0.0) { done = true; break; }
        pair = fmin(pair, prev);
        tau += 2.0 * pair;
        prev = pair;
      }
      s_geyer[0] = tau; s_geyer[1] = prev; s_geyer[2] = done ? 1.0 : 0.0;
    }
    __syncthreads();
    if (s_geyer[2] != 0.0) break;
  }
  if (tid == 0) {
    const double tau = fmax(s_geyer[0], 1.0 / log10((double)N));
    A.ess[col] = (double)N / tau;
  }
}

}  // namespace erirt

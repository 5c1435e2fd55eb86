// global.cuh -- item and structural draws from the reduced sufficient statistics (K3/K4/K6/K7 of SURVEY.md 2c).
//
// One CTA, launched after the person kernel P(k).  When persons are sharded over GPUs the all-reduce of the statistics is
// fused into this kernel (one-shot exchange over NVLink peer memory, see the top of global_draw_kernel; an ncclAllReduce
// in front of the launch is the fallback when no peer buffers are attached); every GPU then runs the draws redundantly
// with the same Philox key, so the new parameters need no broadcast.
//   1. log-likelihood of state k: Bernoulli + structural parts arrive in the statistics, the response-time
//      part is evaluated from sufficient statistics (getLogLikelihood*, GibbsRtIrt.pl.jl:262-272)
//   2. parameters of sweep k+1 in the reference's order, every sum over persons expanded into
//      S0,S1,S2,Ky,C, the Gram matrix and the ingest constants T1,T2,K0,X'X:
//        beta   drawSubjCoefficients :380-393 | ...Latent :399-416 | getSubjCoefficientsMlIrt :351-357 | ...LatentQr :446-458
//        Sigma  drawSubjCovariance :499-515 | ...Null :522-535 | ...Latent :563-579 | ...LatentQr :585-606
//        b      drawItemDifficulty :98-105,   a  drawItemDiscrimination :88-93   (MlIrt: a before b)
//        lambda drawItemIntensity :215-220,   sigma2  drawItemTimeResidual :257-262
//   3. trace row k+1 (Post.ra/rt/qr item + structural columns, GibbsRtIrt.pl.jl:319-321), statistics reset.
// All arithmetic is f64.
#pragma once
#include "layout.cuh"

namespace erirt {

constexpr int G_THREADS = 512;  // warp 0: structural block (one lane); warps 1..15: item draws; all warps: raw variates, reductions

// ---- small dense SPD helpers (column-major, executed by one thread on shared-memory scratch).  Every loop is kept rolled
//      (#pragma unroll 1): this code runs once per sweep on one lane, so unrolled straight-line code would only turn into
//      instruction-cache misses (measured: 7 k cycles for a 25-element copy before, see profiles/) ----
__device__ inline bool chol_lower(int n, const double* A, double* Lm) {
  #pragma unroll 1
  for (int t = 0; t < n * n; ++t) Lm[t] = 0.0;
  #pragma unroll 1
  for (int j = 0; j < n; ++j) {
    double d = A[j + n * j];
    #pragma unroll 1
    for (int k = 0; k < j; ++k) d -= Lm[j + n * k] * Lm[j + n * k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    Lm[j + n * j] = d;
    #pragma unroll 1
    for (int i = j + 1; i < n; ++i) {
      double s = A[i + n * j];
      #pragma unroll 1
      for (int k = 0; k < j; ++k) s -= Lm[i + n * k] * Lm[j + n * k];
      Lm[i + n * j] = s / d;
    }
  }
  return true;
}
// Ainv = A^{-1} through the Cholesky factor; Lm and Li are n*n scratch, none of the four may alias
__device__ inline bool spd_inverse(int n, const double* A, double* Ainv, double* Lm, double* Li) {
  if (!chol_lower(n, A, Lm)) return false;
  #pragma unroll 1
  for (int t = 0; t < n * n; ++t) Li[t] = 0.0;
  #pragma unroll 1
  for (int j = 0; j < n; ++j) {
    Li[j + n * j] = 1.0 / Lm[j + n * j];
    #pragma unroll 1
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      #pragma unroll 1
      for (int k = j; k < i; ++k) s -= Lm[i + n * k] * Li[k + n * j];
      Li[i + n * j] = s / Lm[i + n * i];
    }
  }
  #pragma unroll 1
  for (int i = 0; i < n; ++i)
    #pragma unroll 1
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      #pragma unroll 1
      for (int k = (i > j ? i : j); k < n; ++k) s += Li[k + n * i] * Li[k + n * j];
      Ainv[i + n * j] = s;
    }
  return true;
}
__device__ inline bool spd_solve(int n, const double* A, const double* rhs, double* sol, double* Lm, double* y) {
  if (!chol_lower(n, A, Lm)) return false;
  #pragma unroll 1
  for (int i = 0; i < n; ++i) {
    double s = rhs[i];
    #pragma unroll 1
    for (int k = 0; k < i; ++k) s -= Lm[i + n * k] * y[k];
    y[i] = s / Lm[i + n * i];
  }
  #pragma unroll 1
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    #pragma unroll 1
    for (int k = i + 1; k < n; ++k) s -= Lm[k + n * i] * sol[k];
    sol[i] = s / Lm[i + n * i];
  }
  return true;
}
// quadratic form b' M b and bilinear b' v
__device__ inline double quad_form(int n, const double* M, const double* b) {
  double acc = 0.0;
  #pragma unroll 1
  for (int r = 0; r < n; ++r) {
    double x = 0.0;
    #pragma unroll 1
    for (int q = 0; q < n; ++q) x += M[r + n * q] * b[q];
    acc += b[r] * x;
  }
  return acc;
}
__device__ inline double dotn(int n, const double* a, const double* b) {
  double acc = 0.0;
  #pragma unroll 1
  for (int r = 0; r < n; ++r) acc += a[r] * b[r];
  return acc;
}

__device__ inline void cov2one_2x2(double* S) {  // Draw.pl.jl:507-511
  const double r12 = S[2] / sqrt(S[0]) / sqrt(S[3]), r21 = S[1] / sqrt(S[0]) / sqrt(S[3]);
  S[0] = 1.0; S[3] = 1.0; S[2] = r12; S[1] = r21;
}

// rand(InverseWishart(df, Psi)) = inv(rand(Wishart(df, inv(Psi)))), Wishart by Bartlett's decomposition
__device__ inline void inv_wishart2(PhiloxKey key, uint32_t sweep, double df, const double Psi[4], double out[4], const double* graw) {
  const double det = Psi[0] * Psi[3] - Psi[1] * Psi[2];
  const double S[4] = {Psi[3] / det, -Psi[1] / det, -Psi[2] / det, Psi[0] / det};
  const double l11 = sqrt(S[0]), l21 = S[1] / l11, l22 = sqrt(S[3] - l21 * l21);
  const uint32_t site = make_site(DOM_GLOBAL, GK_SIGMAP);
  const double a11 = sqrt(2.0 * gamma_from_raw(key, 0, sweep, site, 0.5 * df, graw[0], graw[1]));
  const double a22 = sqrt(2.0 * gamma_from_raw(key, 1, sweep, site, 0.5 * (df - 1.0), graw[2], graw[3]));
  const double a21 = graw[4];
  const double x11 = l11 * a11, x21 = l21 * a11 + l22 * a21, x22 = l22 * a22;
  const double w11 = x11 * x11, w21 = x21 * x11, w22 = x21 * x21 + x22 * x22;
  const double dw = w11 * w22 - w21 * w21;
  out[0] = w22 / dw; out[1] = -w21 / dw; out[2] = -w21 / dw; out[3] = w11 / dw;
}

#ifdef ERIRT_TICKS
__device__ long long g_gticks[16];
__device__ long long g_g0;
#define G_TICK(n, who) do { if (tid == (who)) g_gticks[n] = clock64() - _g0; } while (0)
#define S_TICK(n) (g_gticks[n] = clock64() - g_g0)
#else
#define G_TICK(n, who)
#define S_TICK(n)
#endif

struct GramView {  // accessors into the upper-triangular Gram of u = [1, X(1..F), theta, zeta, nu]
  const double* g;
  int Dg, F;
  __device__ double at(int r, int c) const { return r <= c ? g[tri_index(r, c, Dg)] : g[tri_index(c, r, Dg)]; }
  __device__ int th() const { return F + 1; }
  __device__ int ze() const { return F + 2; }
  __device__ int nu() const { return F + 3; }
};

// ---- fused one-shot all-reduce over NVLink peer memory (one CTA; all its threads call it at the same point) ----
// Push `count` doubles of `src` into slot [parity][rank] of every GPU's exchange buffer, publish the sequence stamp, wait for the
// peers' stamps, sum the slots in rank order into `dst` (shared or global; bitwise the same result on every GPU).
// Two parities suffice: a GPU can only be two exchanges ahead of a slot it overwrites after the owner has published the stamp of
// the exchange in between, i.e. after the owner has finished reading that slot.
__device__ inline void peer_allreduce_sum(double* const* peer_bufs, uint32_t* xseq, int world, int rank, int S, const double* src,
                                          int count, double* dst, int* status, int tid, int nthreads) {
  const uint32_t seq = *xseq + 1u;
  const int par = (int)(seq & 1u);
  for (int r = 0; r < world; ++r) {
    double* slot = peer_bufs[r] + ((size_t)par * world + rank) * S;
    for (int t = tid; t < count; t += nthreads) slot[t] = src[t];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(peer_bufs[tid] + (size_t)2 * world * S) + par * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(seq) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(peer_bufs[rank] + (size_t)2 * world * S) + par * world + tid;
    uint32_t got;
    const long long t_start = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
      if (got != seq && clock64() - t_start > 20000000000LL) {  // ~10 s: a peer is gone; flag it instead of hanging the GPU
        atomicExch(status, -1000 - tid);
        break;
      }
    } while (got != seq);
  }
  __syncthreads();
  const double* slots = peer_bufs[rank] + (size_t)par * world * S;
  for (int t = tid; t < count; t += nthreads) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += __ldcg(slots + (size_t)r * S + t);
    dst[t] = acc;
  }
  __syncthreads();  // every thread has read *xseq and the slots
  if (tid == 0) *xseq = seq;
}
// the same exchange on its own, for the one-time ingest constants
__global__ void __launch_bounds__(512) peer_allreduce_kernel(double* const* peer_bufs, uint32_t* xseq, int world, int rank, int S,
                                                             double* buf, int count, int* status) {
  if (*status <= -1000) return;
  peer_allreduce_sum(peer_bufs, xseq, world, rank, S, buf, count, buf, status, (int)threadIdx.x, (int)blockDim.x);
}

struct GScratch {
  double M[MAXD * MAXD], V[MAXD * MAXD], Lc[MAXD * MAXD], T[MAXD * MAXD], XX[MAXD * MAXD];
  double rhs[MAXD], mean[MAXD], z[MAXD], beta[MAXD], Sigma[4];
  double zn[MAXD];  // N(0,1) of the global beta site, unit t (precomputed on separate lanes)
  double graw[6];   // attempt-0 raw material of the Sigma site: (normal, uniform) of units 0, 1, 2
};

// mean + chol(V).L * z with z_t ~ N(0,1) at the global beta site; returns false if V is not SPD
__device__ inline bool mvn_draw(const GlobalArgs& A, uint32_t s, int d, GScratch& w) {
  if (!chol_lower(d, w.V, w.Lc)) return false;
  #pragma unroll 1
  for (int r = 0; r < d; ++r) {
    double acc = w.mean[r];
    #pragma unroll 1
    for (int q = 0; q <= r; ++q) acc += w.Lc[r + d * q] * w.zn[q];
    w.beta[r] = acc;
  }
  return true;
}

// w.XX = x'x and w.rhs = x'y of the Latent family (x = [1 X θ]), one element per lane
__device__ inline void latent_prebuild(const GlobalArgs& A, const double* XtX, const GramView& Gm, GScratch& w, int tid, int nthreads) {
  const int F = A.L.F, pb = F + 1, d = F + 2;
  const bool qrm = A.model == M_LATENTQR;
  for (int t = tid; t < d * d + d; t += nthreads) {
    if (t < d * d) {
      const int r = t % d, q = t / d;
      w.XX[r + d * q] = (r < pb && q < pb) ? XtX[r + pb * q] : Gm.at(r, q);
    } else {
      const int r = t - d * d;
      w.rhs[r] = Gm.at(r, Gm.ze()) - (qrm ? A.k1 * Gm.at(r, Gm.nu()) : 0.0);
    }
  }
}

__device__ inline bool structural_draws(const GlobalArgs& A, const double* XtX, uint32_t s, const GramView& Gm, const GramView& Gw, GScratch& w) {
  const Layout& L = A.L;
  const int F = L.F, pb = F + 1, model = A.model;
  const double N = (double)A.n_total;
  const int TH = Gm.th(), ZE = Gm.ze(), NU = Gm.nu();
  const double Sth2 = Gm.at(TH, TH), Sze2 = Gm.at(ZE, ZE), Sthze = Gm.at(TH, ZE);
  const double add = (A.compat & 1) ? 0.0 : 1.0;  // `1/σβ₀^2 .+ M` adds to EVERY element (Draw.pl.jl:386,410; quirk Q1)
  bool ok = true;
  if (model == M_MLIRT) {
    #pragma unroll 1
    for (int r = 0; r < pb; ++r) w.rhs[r] = Gm.at(r, TH);
    ok = spd_solve(pb, XtX, w.rhs, w.beta, w.Lc, w.z);  // getSubjCoefficientsMlIrt
    if (!A.intercept) w.beta[0] = 0.0;
  } else if (model == M_RTIRT) {
    const int d = 2 * pb;
    const double S11 = w.Sigma[0], S12 = w.Sigma[2], S22 = w.Sigma[3], det = S11 * S22 - S12 * S12;
    const double iO[4] = {S22 / det, -S12 / det, -S12 / det, S11 / det};
    #pragma unroll 1
    for (int br = 0; br < 2; ++br)
      #pragma unroll 1
      for (int bc = 0; bc < 2; ++bc)
        #pragma unroll 1
        for (int r = 0; r < pb; ++r)
          #pragma unroll 1
          for (int q = 0; q < pb; ++q) w.M[(br * pb + r) + d * (bc * pb + q)] = iO[br + 2 * bc] * XtX[r + pb * q] + add;
    if (add == 0.0)
      #pragma unroll 1
      for (int t = 0; t < d; ++t) w.M[t + d * t] += 1.0;
    ok = spd_inverse(d, w.M, w.V, w.Lc, w.T);
    #pragma unroll 1
    for (int c = 0; c < 2; ++c)  // vec(x'η invΩ')
      #pragma unroll 1
      for (int r = 0; r < pb; ++r) w.rhs[c * pb + r] = Gm.at(r, TH) * iO[c + 0] + Gm.at(r, ZE) * iO[c + 2];
    #pragma unroll 1
    for (int r = 0; r < d; ++r) {
      double acc = 0.0;
      #pragma unroll 1
      for (int q = 0; q < d; ++q) acc += w.V[r + d * q] * w.rhs[q];
      w.mean[r] = acc;
    }
    ok = ok && mvn_draw(A, s, d, w);
    if (!A.intercept) { w.beta[0] = 0.0; w.beta[pb] = 0.0; }
    // e'e, e = [θ ζ] - xβ   (drawSubjCovariance)
    const double* b1 = w.beta;
    const double* b2 = w.beta + pb;
    double q11 = 0, q12 = 0, q22 = 0, l1t = 0, l1z = 0, l2t = 0, l2z = 0;
    #pragma unroll 1
    for (int r = 0; r < pb; ++r) {
      double x1 = 0, x2 = 0;
      #pragma unroll 1
      for (int q = 0; q < pb; ++q) { x1 += XtX[r + pb * q] * b1[q]; x2 += XtX[r + pb * q] * b2[q]; }
      q11 += b1[r] * x1; q12 += b1[r] * x2; q22 += b2[r] * x2;
      l1t += b1[r] * Gm.at(r, TH); l1z += b1[r] * Gm.at(r, ZE);
      l2t += b2[r] * Gm.at(r, TH); l2z += b2[r] * Gm.at(r, ZE);
    }
    const double E11 = Sth2 - 2.0 * l1t + q11, E12 = Sthze - l1z - l2t + q12, E22 = Sze2 - 2.0 * l2z + q22;
    const double Psi[4] = {E11 + 1.0, E12, E12, E22 + 1.0};
    inv_wishart2(A.key, s, N + 3.0, Psi, w.Sigma, w.graw);
    if (A.cov2one) cov2one_2x2(w.Sigma);
  } else if (model == M_NULL) {
    #pragma unroll 1
    for (int t = 0; t < 2 * pb; ++t) w.beta[t] = 0.0;
    const double Psi[4] = {Sth2 + 1.0, Sthze, Sthze, Sze2 + 1.0};
    inv_wishart2(A.key, s, N + 3.0, Psi, w.Sigma, w.graw);  // drawSubjCovarianceNull
    if (A.cov2one) cov2one_2x2(w.Sigma);
  } else if (model == M_CROSS || model == M_CROSSQR) {
    // drawSubjCovarianceCross, Draw.pl.jl:542-557
    const double parA = 1e-3 + N / 2.0, parB = 1e-3 + Sze2 / 2.0;
    const double sv = parB / gamma_from_raw(A.key, 0, s, make_site(DOM_GLOBAL, GK_SIGMAP), parA, w.graw[0], w.graw[1]);
    w.Sigma[0] = 1.0; w.Sigma[1] = 0.0; w.Sigma[2] = 0.0; w.Sigma[3] = sv;
    if (A.cov2one) cov2one_2x2(w.Sigma);
  } else if (model == M_LATENT || model == M_LATENTQR) {
    const int d = F + 2;
    const bool qrm = model == M_LATENTQR;
    // x = [1 X θ]: x'x (ingest constant bordered by the θ column of the Gram) and x'y, y = ζ (Latent) or ζ - k1 ν (LatentQr),
    // were built by all lanes in latent_prebuild
    S_TICK(7);
    double yy = Sze2;
    if (qrm) yy += -2.0 * A.k1 * Gm.at(ZE, NU) + A.k1 * A.k1 * Gm.at(NU, NU);
    if (!qrm) {  // drawSubjCoefficientsLatent
      const double iO = 1.0 / w.Sigma[3];
      #pragma unroll 1
      for (int t = 0; t < d * d; ++t) w.M[t] = iO * w.XX[t] + add;
      if (add == 0.0)
        #pragma unroll 1
        for (int t = 0; t < d; ++t) w.M[t + d * t] += 1.0;
      ok = spd_inverse(d, w.M, w.V, w.Lc, w.T);
      #pragma unroll 1
      for (int r = 0; r < d; ++r) {
        double acc = 0.0;
        #pragma unroll 1
        for (int q = 0; q < d; ++q) acc += w.V[r + d * q] * (w.rhs[q] * iO);
        w.mean[r] = acc;
      }
      ok = ok && mvn_draw(A, s, d, w);
    } else {  // getSubjCoefficientsLatentQr: the tall Kronecker system collapses to OLS
      ok = spd_solve(d, w.XX, w.rhs, w.beta, w.Lc, w.z);
    }
    S_TICK(8);
    if (!A.intercept) w.beta[0] = 0.0;
    // residual sum of squares r = y - xβ
    const double ss = yy - 2.0 * dotn(d, w.beta, w.rhs) + quad_form(d, w.XX, w.beta);
    S_TICK(9);
    double parA, parB;
    if (!qrm) {
      parA = 1e-3 + N / 2.0;
      parB = 1e-3 + ss / 2.0;
    } else {
      const double Snu = Gm.at(0, NU), Snu2 = Gm.at(NU, NU);
      parA = 1e-3 + N * 3.0 / 2.0;
      double scale;
      if (A.compat & 2) {
        // Σ r_i²/(2 k2 ν_i) from the 1/ν-weighted Gram: y/ν terms
        #pragma unroll 1
        for (int r = 0; r < d; ++r)
          #pragma unroll 1
          for (int q = 0; q < d; ++q) w.T[r + d * q] = Gw.at(r, q);
        #pragma unroll 1
        for (int r = 0; r < d; ++r) w.mean[r] = Gw.at(r, ZE) - A.k1 * Gm.at(0, r);  // Σ x (ζ - k1 ν)/ν
        const double yyw = Gw.at(ZE, ZE) - 2.0 * A.k1 * Gm.at(0, ZE) + A.k1 * A.k1 * Snu;
        const double ssw = yyw - 2.0 * dotn(d, w.beta, w.mean) + quad_form(d, w.T, w.beta);
        scale = ssw / (2.0 * A.k2);
      } else {
        // as written: sum(r.^2 / (2 k2e)) with vector / vector (quirk Q2, Draw.pl.jl:594)
        scale = ss * (2.0 * A.k2 * Snu) / (4.0 * A.k2 * A.k2 * Snu2);
      }
      parB = 1e-3 + scale + Snu;
    }
    S_TICK(10);
    const double sv = parB / gamma_from_raw(A.key, 0, s, make_site(DOM_GLOBAL, GK_SIGMAP), parA, w.graw[0], w.graw[1]);
    w.Sigma[0] = 1.0; w.Sigma[1] = 0.0; w.Sigma[2] = 0.0; w.Sigma[3] = sv;
    if (A.cov2one) cov2one_2x2(w.Sigma);
    S_TICK(11);
  }
  return ok;
}

__global__ void __launch_bounds__(G_THREADS) global_draw_kernel(const GlobalArgs A) {
#ifdef ERIRT_TICKS
  const long long _g0 = clock64();
  if (threadIdx.x == 0) g_g0 = _g0;
#endif
  __shared__ GScratch w;
  __shared__ double sRed[G_THREADS];
  const Layout& L = A.L;
  const int J = L.J, F = L.F, Dg = L.Dg, tid = threadIdx.x;
  const int model = A.model;
  const bool has_rt = model != M_MLIRT;
  if (*A.status <= -1000) return;   // a peer timed out in an earlier exchange: no draws from partial sums, no further 10 s waits
  const uint32_t k = *A.sweep_ctr;  // person launch P(k) just finished
  const uint32_t s = k + 1;         // sweep whose parameters are drawn now
  const double N = (double)A.n_total;
  double* par = A.params;
  // Stage the reduced statistics and X'X in shared memory first: the structural block reads them element by element in
  // dependent loops, and every such read would otherwise be an L2 round trip (the person kernel's atomics left them in L2).
  extern __shared__ double g_dyn[];
  double* st = g_dyn;                     // [s_count + 2] local copy of the statistics (read-only below)
  double* sXtX = g_dyn + L.s_count + 2;   // [(F+1)^2]
  double* sRaw = sXtX + (F + 1) * (F + 1);  // [5][Jp] attempt-0 raw material of the item sites: z_b, z_a, z_lambda, x_sigma2, u_sigma2
  const int Jp = L.Jp;
  if (A.peer_bufs) {
    peer_allreduce_sum(A.peer_bufs, A.xseq, A.world, A.rank, A.xstride, A.stats, L.s_count, st, A.status, tid, G_THREADS);
    if (*reinterpret_cast<volatile int*>(A.status) <= -1000) return;  // timed out in this exchange (uniform: written before the CTA barrier)
  } else {
    for (int t = tid; t < L.s_count; t += G_THREADS) st[t] = A.stats[t];
  }
  for (int t = tid; t < (F + 1) * (F + 1); t += G_THREADS) sXtX[t] = A.XtX[t];
  if (tid < MAXD) w.beta[tid] = par[L.p_beta + tid];  // one lane each: a single lane copying global -> shared pays a round trip per element
  if (tid >= 64 && tid < 68) w.Sigma[tid - 64] = par[L.p_Sigma + tid - 64];
  __syncthreads();
  G_TICK(0, 0);  // staging done
  const GramView Gm{st + L.s_gram, Dg, F};
  const GramView Gw{st + L.s_gramw, Dg, F};
  const double Sth = Gm.at(0, Gm.th()), Sth2 = Gm.at(Gm.th(), Gm.th()), Sthze = Gm.at(Gm.th(), Gm.ze());
  const double Sze = Gm.at(0, Gm.ze()), Sze2 = Gm.at(Gm.ze(), Gm.ze());
  const double LOG2PI = 1.8378770664093454835606594728112;
  const bool cross = model == M_CROSS || model == M_CROSSQR;
  const double mu_lam = has_rt ? A.consts[0] : 0.0, sd_lam = has_rt ? A.consts[1] : 1.0;

  if (A.stage == 1) {
    // ---- G_a of the Cross family: lambda_k, sigma2_k from (theta_k, zeta_{k-1}, rho_k)   Draw.pl.jl:225-231, 267-273 ----
    for (int j = tid; j < J; j += G_THREADS) {
      const double s2 = par[L.p_sigma2 + j], rho = par[L.p_rho + j];
      const double T1 = A.T1[j], T2 = A.T2[j];
      const double pv = 1.0 / (sd_lam * sd_lam);
      double lam, s2n;
      if (model == M_CROSS) {
        const double C = st[L.s_C + j], D = st[L.s_D + j];
        const double parV = 1.0 / (pv + N / s2);
        const double parM = parV * (mu_lam * pv + (T1 + Sze + rho * Sth) / s2);
        lam = site_tnorm_pos(A.key, (uint32_t)j, k, make_site(DOM_ITEM, IK_LAMBDA), parM, sqrt(parV));
        const double Q = T2 - 2.0 * lam * T1 + N * lam * lam + 2.0 * (C - lam * Sze) + Sze2 + rho * rho * Sth2 +
                         2.0 * rho * (D - lam * Sth + Sthze);
        s2n = (1e-3 + Q / 2.0) / site_gamma(A.key, (uint32_t)j, k, make_site(DOM_ITEM, IK_SIGMA2), 1e-3 + N / 2.0);
      } else {
        // CrossQr (Draw.pl.jl:239-251, 278-288): weights w = 1/nu, r = logT + zeta; A0=Σw A1=Σθw A2=Σθrw A3=Σθ²w A4=Σrw A5=Σr²w V1=Σν
        const double A0 = st[L.s_S0 + j], A1 = st[L.s_S1 + j], A2 = st[L.s_S2 + j], A3 = st[L.s_Ky + j], A4 = st[L.s_C + j],
                     A5 = st[L.s_D + j], V1 = st[L.s_V + j];
        const double R1 = T1 + Sze, k1 = A.k1, k2 = A.k2;
        const double parV = 1.0 / (pv + A0 / (s2 * k2));
        const double parM = parV * (mu_lam * pv + (A4 + rho * A1 - k1 * N) / (s2 * k2));
        lam = site_tnorm_pos(A.key, (uint32_t)j, k, make_site(DOM_ITEM, IK_LAMBDA), parM, sqrt(parV));
        const double SS = A5 + lam * lam * A0 + rho * rho * A3 + k1 * k1 * V1 - 2.0 * lam * A4 + 2.0 * rho * A2 - 2.0 * k1 * R1 -
                          2.0 * lam * rho * A1 + 2.0 * lam * k1 * N - 2.0 * rho * k1 * Sth;
        s2n = (1e-3 + SS / (2.0 * k2) + V1) / site_gamma(A.key, (uint32_t)j, k, make_site(DOM_ITEM, IK_SIGMA2), 1e-3 + N * 3.0 / 2.0);
      }
      par[L.p_lambda + j] = lam;
      par[L.p_sigma2 + j] = s2n;
      if (k >= 1 && (int)(k - 1) < A.cap) {
        A.tr_items_rt[(size_t)(k - 1) * 2 * J + j] = lam;
        A.tr_items_rt[(size_t)(k - 1) * 2 * J + J + j] = s2n;
      }
    }
    __syncthreads();
    for (int t = tid; t < L.s_count; t += G_THREADS) A.stats[t] = 0.0;
    return;
  }

  // ---- 1. log-likelihood of state k ----
  if (k >= 1 || A.stage == 3) {
    double part = 0.0;
    if (model == M_CROSSQR) {
      if (tid == 0) part = st[L.s_scal + SC_LL_RT];
    } else if (has_rt || A.kz_from_stats)
      for (int j = tid; j < J; j += G_THREADS) {
        if (has_rt) {
          const double lam = par[L.p_lambda + j], s2 = par[L.p_sigma2 + j];
          double Q = A.T2[j] - 2.0 * lam * A.T1[j] + N * lam * lam + 2.0 * (st[L.s_C + j] - lam * Sze) + Sze2;
          if (cross) {  // residual logT - lambda + zeta + theta rho_j  (getLogLikelihoodRtIrtCross, GibbsRtIrtCross.pl.jl:158-169)
            const double rho = par[L.p_rho + j];
            Q += rho * rho * Sth2 + 2.0 * rho * (st[L.s_D + j] - lam * Sth + Sthze);
          }
          part += -0.5 * N * (LOG2PI + log(s2)) - 0.5 * Q / s2;
        }
        // sum_i kappa_ij z_ij = a_j [(Ky_j - sum_i theta_i / 2) - b_j K0_j] at state k: the part of the Bernoulli term the f32
        // person kernel leaves to its statistics (Ky = sum_i y_ij theta_i, K0 = sum_i kappa_ij)
        if (A.kz_from_stats) part += par[L.p_a + j] * ((st[L.s_Ky + j] - 0.5 * Sth) - par[L.p_b + j] * A.K0[j]);
      }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) sRed[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int t = 0; t < G_THREADS / 32; ++t) tot += sRed[t];
      sRed[0] = tot;
    }
    __syncthreads();
    if (A.stage == 3) {  // evaluation only: report and leave every parameter untouched
      if (tid == 0) *A.ll_out = st[L.s_scal + SC_LL_BERN] + sRed[0] + st[L.s_scal + SC_LL_STRUCT];
      __syncthreads();
      for (int t = tid; t < L.s_count; t += G_THREADS) A.stats[t] = 0.0;
      return;
    }
    if (tid == 0 && (int)(k - 1) < A.cap) A.tr_ll[k - 1] = st[L.s_scal + SC_LL_BERN] + sRed[0] + st[L.s_scal + SC_LL_STRUCT];
  }

  G_TICK(1, 0);  // log-likelihood done
  // ---- 2.0 raw variates of attempt 0 of every site of sweep s, one task per lane: they do not depend on the parameters, and
  //      an f64 Box-Muller (log, sqrt, cospi) is a long dependent chain that would otherwise run 4 + 2 times in a row per lane ----
  for (int t = tid; t < 4 * J + MAXD + 3; t += G_THREADS) {
    if (t < 4 * J) {
      const int j = t >> 2, kind = t & 3;
      const uint32_t ik = kind == 0 ? IK_B : (kind == 1 ? IK_A : (kind == 2 ? IK_LAMBDA : IK_SIGMA2));
      const uint4 wd = philox(A.key, (uint32_t)j, s, make_site(DOM_ITEM, ik), 0);
      if (kind < 3) sRaw[kind * Jp + j] = normal2(wd.x, wd.y);
      else { sRaw[3 * Jp + j] = normal2(wd.x, wd.y); sRaw[4 * Jp + j] = u01d(wd.z); }
    } else if (t < 4 * J + MAXD) {
      const int u = t - 4 * J;
      const uint4 wd = philox(A.key, (uint32_t)u, s, make_site(DOM_GLOBAL, GK_BETA), 0);
      w.zn[u] = normal2(wd.x, wd.y);
    } else {
      const int u = t - 4 * J - MAXD;
      const uint4 wd = philox(A.key, (uint32_t)u, s, make_site(DOM_GLOBAL, GK_SIGMAP), 0);
      w.graw[2 * u] = normal2(wd.x, wd.y);
      w.graw[2 * u + 1] = u01d(wd.z);
    }
  }
  if (model == M_LATENT || model == M_LATENTQR) latent_prebuild(A, sXtX, Gm, w, tid, G_THREADS);
  __syncthreads();
  G_TICK(6, 0);  // raw variates done
  // ---- 2a. structural draws (thread 0), 2b. item draws (warps 1..) ----
  if (tid == 0) {
    if (!structural_draws(A, sXtX, s, Gm, Gw, w)) atomicExch(A.status, (int)s);
    G_TICK(2, 0);  // structural block done
  }
  // item draws on warps 1..: the structural lane of warp 0 runs concurrently.  Two tasks per item, dealt so that a warp holds one
  // kind: (b, a) [MlIrt: (a, b)] and the response-time pair (lambda, sigma2) [Cross family: rho], which do not depend on each other
  for (int t = tid - 32; t >= 0 && t < 2 * J; t += G_THREADS - 32) {
    const int part = t / J, j = t - part * J;
    if (part == 0) {
      const double S0 = st[L.s_S0 + j], S1 = st[L.s_S1 + j], S2 = st[L.s_S2 + j];
      const double K0 = A.K0[j], K1 = st[L.s_Ky + j] - 0.5 * Sth;  // Σκ, Σκθ
      double a = par[L.p_a + j], b = par[L.p_b + j];
      auto draw_b = [&]() {  // drawItemDifficulty
        const double parV = 1.0 / (1.0 + a * a * S0);
        const double parM = parV * (0.0 - (a * K0 - a * a * S1));
        double v = parM + sqrt(parV) * sRaw[0 * Jp + j];
        b = v < -4.0 ? -4.0 : (v > 4.0 ? 4.0 : v);
      };
      auto draw_a = [&]() {  // drawItemDiscrimination
        const double parV = 1.0 / (1.0 + (S2 - 2.0 * b * S1 + b * b * S0));
        const double parM = parV * (1.0 + (K1 - b * K0));
        a = tnorm_pos_from_raw(A.key, (uint32_t)j, s, make_site(DOM_ITEM, IK_A), parM, sqrt(parV), sRaw[1 * Jp + j]);
        if (A.onepl) a = 1.0;
      };
      if (model == M_MLIRT) { draw_a(); draw_b(); }  // GibbsRtIrt.pl.jl:233-237 (quirk Q8)
      else { draw_b(); draw_a(); }
      par[L.p_a + j] = a;
      par[L.p_b + j] = b;
    } else if (cross) {  // drawSubjCorrCross, Draw.pl.jl:463-469 (state k: theta_k, zeta_k, lambda_k, sigma2_k)
      const double s2 = par[L.p_sigma2 + j], lam = par[L.p_lambda + j];
      double parV, parM;
      if (model == M_CROSS) {
        parV = 1.0 / (1.0 + Sth2 / s2);
        parM = parV * (0.0 + (lam * Sth - Sthze - st[L.s_D + j]) / s2);
      } else {  // drawSubjCorrCrossQr, Draw.pl.jl:474-489 with A1'=Σθ/ν, A2'=Σθ(logT+ζ)/ν, A3'=Σθ²/ν at nu_{k+1}
        const double A1 = st[L.s_C + j], A2 = st[L.s_D + j], A3 = st[L.s_V + j];
        parV = 1.0 / (1.0 + A3 / (s2 * A.k2));
        parM = parV * (0.0 + (lam * A1 - A2 + A.k1 * Sth) / (s2 * A.k2));
      }
      par[L.p_rho + j] = parM + sqrt(parV) * site_normal(A.key, (uint32_t)j, s, make_site(DOM_ITEM, IK_RHO));
    } else if (has_rt) {
      const double s2 = par[L.p_sigma2 + j];
      const double T1 = A.T1[j], T2 = A.T2[j], C = st[L.s_C + j];
      // drawItemIntensity
      const double pv = 1.0 / (sd_lam * sd_lam);
      const double parV = 1.0 / (pv + N / s2);
      const double parM = parV * (mu_lam * pv + (T1 + Sze) / s2);
      const double lam = tnorm_pos_from_raw(A.key, (uint32_t)j, s, make_site(DOM_ITEM, IK_LAMBDA), parM, sqrt(parV), sRaw[2 * Jp + j]);
      // drawItemTimeResidual: Σ_i (logT - λ + ζ)²
      const double Q = T2 - 2.0 * lam * T1 + N * lam * lam + 2.0 * (C - lam * Sze) + Sze2;
      const double parA = 1e-3 + N / 2.0, parB = 1e-3 + Q / 2.0;
      const double s2n = parB / gamma_from_raw(A.key, (uint32_t)j, s, make_site(DOM_ITEM, IK_SIGMA2), parA, sRaw[3 * Jp + j], sRaw[4 * Jp + j]);
      par[L.p_lambda + j] = lam;
      par[L.p_sigma2 + j] = s2n;
    }
  }
  G_TICK(3, 32);  // item draws of lane 32 done
  __syncthreads();
  G_TICK(4, 0);
  if (tid < MAXD) par[L.p_beta + tid] = w.beta[tid];
  if (tid >= 64 && tid < 68) par[L.p_Sigma + tid - 64] = w.Sigma[tid - 64];
  __syncthreads();

  // ---- 3. trace row of sweep s, statistics reset ----
  if ((int)k < A.cap) {
    for (int j = tid; j < J; j += G_THREADS) {
      A.tr_items_ra[(size_t)k * 2 * J + j] = par[L.p_a + j];
      A.tr_items_ra[(size_t)k * 2 * J + J + j] = par[L.p_b + j];
      if (has_rt && !cross) {
        A.tr_items_rt[(size_t)k * 2 * J + j] = par[L.p_lambda + j];
        A.tr_items_rt[(size_t)k * 2 * J + J + j] = par[L.p_sigma2 + j];
      }
    }
    const int nb = cross ? J : ((model == M_MLIRT) ? F + 1 : ((model == M_RTIRT || model == M_NULL) ? 2 * (F + 1) : F + 2));
    for (int t = tid; t < A.qw; t += G_THREADS) {
      double v = t < nb ? (cross ? par[L.p_rho + t] : par[L.p_beta + t]) : par[L.p_Sigma + (t - nb)];
      A.tr_qr[(size_t)k * A.qw + t] = v;
    }
  }
  if (tid == 0) {
    A.stats[L.s_count] = st[L.s_scal + SC_PG_DEFER];      // keep the diagnostics of the last sweep
    A.stats[L.s_count + 1] = st[L.s_scal + SC_PG_CELLS];
  }
  __syncthreads();
  for (int t = tid; t < L.s_count; t += G_THREADS) A.stats[t] = 0.0;
  G_TICK(5, 0);  // trace + reset done
  if (tid == 0) *A.sweep_ctr = k + 1;
}

}  // namespace erirt

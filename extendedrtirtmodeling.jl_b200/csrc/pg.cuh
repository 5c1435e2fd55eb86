// pg.cuh -- exact Polya-Gamma PG(1, z) draw on the device.
//
// Replaces rand(PolyaGammaPSWSampler(1, η)) at /root/reference/src/Draw.pl.jl:38 (PolyaGammaSamplers.jl is
// not vendored; algorithm restated from Polson-Scott-Windle 2013 / Devroye 2009, see oracle/pg.c for the
// full statement).  PG(1,z) = J*(1,c)/4, c = |z|/2, truncation point t = 0.64:
//   Method A (c <= 1/t): envelope = exponential tail on (t,inf) [mass p] + truncated-Levy piece on (0,t)
//     [mass q0 = 4 Phic(1/sqrt t), sampled by inverse CDF], the tilt exp(-c^2 x/2) folded into acceptance.
//   Method B (c > 1/t):  envelope = exponential tail + full IG(1/c,1) (Michael-Schucany-Haas), rejected if x >= t.
// A rejected attempt restarts from the mixture choice with fresh words (attempt counter in the Philox
// counter), so every draw is a pure function of (seed, chain, person, item, sweep).
//
// Two code paths share this statement:
//   * pg_draw_exact<R>  -- complete loop in the working precision (f64 everywhere);
//   * pg_fast.cuh       -- f32: attempt 0 of two cells per packed instruction with squeeze tests that only ever say
//     "certainly accepted" / "certainly rejected"; every other cell is replayed with the a_1 term from a work queue.
#pragma once
#include "pg_coeffs.h"
#include "rng.cuh"

namespace erirt {

constexpr double PG_T = 0.64;
constexpr double PG_P0 = 0.10564977366685535;     // Phic(1/sqrt(t))
constexpr double PG_Q0 = 0.4225990946674214;      // 4 * PG_P0
constexpr double PG_CSWITCH = 1.5625;             // 1/t
constexpr double PI_D = 3.14159265358979323846;
constexpr double PG_M2LNP0 = 4.4952513544286345;  // -2 ln(PG_P0)
constexpr double PG_R1MAX_RIGHT = 0.005418509355439241;  // 3 exp(-pi^2 t): bound of a_1/a_0 on x >= t
constexpr double PG_R1MAX_LEFT = 0.005791362408683128;   // 3 exp(-4/t):    bound of a_1/a_0 on x <= t

__device__ __constant__ const float c_xq[ERIRT_XQ_DEG + 1] = ERIRT_XQ_COEFFS;
__device__ __constant__ const float c_l1p[ERIRT_L1P_DEG + 1] = ERIRT_L1P_COEFFS;
__device__ __constant__ const double c_xqd[ERIRT_XQD_DEG + 1] = ERIRT_XQD_COEFFS;

// s/Z as a polynomial in r = 1/s, s = sqrt(-2 ln y), Z = Phic^{-1}(y), y in (0, PG_P0]
__device__ __forceinline__ float xq_poly(float r) {
  const float x = fmaf(r, (float)ERIRT_XQ_A, (float)ERIRT_XQ_B);
  float p = (float)c_xq[ERIRT_XQ_DEG];
#pragma unroll
  for (int k = ERIRT_XQ_DEG - 1; k >= 0; --k) p = fmaf(p, x, c_xq[k]);
  return p;
}
// ln(1 + w), w in [0,1]
__device__ __forceinline__ float log1p_poly(float w) {
  const float x = fmaf(w, (float)ERIRT_L1P_A, (float)ERIRT_L1P_B);
  float p = c_l1p[ERIRT_L1P_DEG];
#pragma unroll
  for (int k = ERIRT_L1P_DEG - 1; k >= 0; --k) p = fmaf(p, x, c_l1p[k]);
  return p * w;
}

// XQD(r) (pg_coeffs.h), degree 24, by Estrin's scheme: the dependent chain is 6 fused multiply-adds deep instead of 24 (the f64
// kernels are bound by the latency of dependent DFMAs at 12 warps per SM)
__device__ __forceinline__ double xqd_poly(double r) {
  static_assert(ERIRT_XQD_DEG == 24, "xqd_poly is written for degree 24");
  const double t = fma(r, ERIRT_XQD_A, ERIRT_XQD_B);
  const double t2 = t * t, t4 = t2 * t2, t8 = t4 * t4, t16 = t8 * t8;
  double q[13];
#pragma unroll
  for (int i = 0; i < 12; ++i) q[i] = fma(c_xqd[2 * i + 1], t, c_xqd[2 * i]);
  q[12] = c_xqd[24];
  const double r0 = fma(q[1], t2, q[0]), r1 = fma(q[3], t2, q[2]), r2 = fma(q[5], t2, q[4]), r3 = fma(q[7], t2, q[6]);
  const double r4 = fma(q[9], t2, q[8]), r5 = fma(q[11], t2, q[10]), r6 = q[12];
  const double s0 = fma(r1, t4, r0), s1 = fma(r3, t4, r2), s2 = fma(r5, t4, r4), s3 = r6;
  const double u0 = fma(s1, t8, s0), u1 = fma(s3, t8, s2);
  return fma(u1, t16, u0);
}

// Phic^{-1}(y) for y in (0, PG_P0]: a polynomial in 1/sqrt(-2 ln y), to the rounding level of the working precision
template <typename R>
__device__ __forceinline__ R inv_normal_tail(R y);
template <>
__device__ __forceinline__ float inv_normal_tail<float>(float y) {
  float r = rsqrtf(-2.0f * logf(y));
  return 1.0f / (r * xq_poly(r));
}
template <>
__device__ __forceinline__ double inv_normal_tail<double>(double y) {
  // s/Z as a degree-24 polynomial in r = 1/s, s = sqrt(-2 ln y), fitted in 40-digit arithmetic (tools/gen_coeffs.py: 5.5e-16 relative
  // error of Z against the exact root when evaluated in Float64, i.e. the rounding level of log and rsqrt; the oracle finds the root by
  // Newton steps on ln Phic to convergence, 1e-15 apart).  Until round 2 the f64 path refined the f32 start by two Newton steps on
  // erfc / exp / log: 40 % of the instructions of the f64 person kernel (profiles/r02i_person_kernel_f64_ncu_breakdown.txt).
  const double r = rsqrt(-2.0 * log(y));
  return 1.0 / (r * xqd_poly(r));
}

// alternating-series acceptance: U <= sum_n (-1)^n a_n(x)/a_0(x) ?
template <typename R>
__device__ __forceinline__ bool series_accept(R U, R x, bool right) {
  const R e = right ? R(-0.5 * PI_D * PI_D) * x : R(-2.0) / x;
  R S = R(1);
#pragma unroll 1
  for (int n = 1; n < 400; ++n) {
    R rn = R(2 * n + 1) * exp(e * R(n * n + n));
    if (n & 1) {
      S -= rn;
      if (U <= S) return true;
    } else {
      S += rn;
      if (U > S) return false;
    }
  }
  return false;
}

// One Method-A attempt up to the alternating-series test: the proposal x, the uniform U the series is compared with and the piece
// (`right`: exponential tail); false = already rejected by the tilt.  The series test itself has ONE call site per attempt (the f64
// person kernel is bound by instruction fetch: every inlined copy of exp / log / erfc counts, profiles/r02h_person_kernel_f64_*).
template <typename R>
__device__ __forceinline__ bool pg_propose_A(R c, uint32_t wa, uint32_t wb, R& x, R& U, bool& right) {
  const R um = u01<R>(wa), up = u01<R>(wb);
  const R K = R(PI_D * PI_D / 8.0) + R(0.5) * c * c;
  const R Rm1 = R(2.0 * PG_Q0 / PI_D) * K * exp(K * R(PG_T));  // q0 / p
  const R v = um * (R(1) + Rm1);
  right = v < R(1);
  if (right) {
    x = R(PG_T) - log(up) / K;
    U = v;
    return true;
  }
  const R Z = inv_normal_tail<R>(up * R(PG_P0));
  x = R(1) / (Z * Z);
  const R ua = (v - R(1)) / Rm1;
  const R tilt = exp(R(-0.5) * c * c * x);
  if (ua >= tilt) return false;
  U = ua / tilt;
  return true;
}
template <typename R>
__device__ __forceinline__ bool pg_attempt_A(R c, uint32_t wa, uint32_t wb, R& X) {
  R x, U;
  bool right;
  if (!pg_propose_A<R>(c, wa, wb, x, U, right)) return false;
  if (!series_accept<R>(U, x, right)) return false;
  X = x;
  return true;
}

template <typename R>
__device__ __forceinline__ bool pg_propose_B(R c, uint4 w, R& x, R& U, bool& right) {
  const R um = u01<R>(w.x);
  const R K = R(PI_D * PI_D / 8.0) + R(0.5) * c * c;
  const R p = (R(PI_D) / (R(2) * K)) * exp(-K * R(PG_T));
  const R ql = R(2) * exp(-c);
  const R Pr = p / (p + ql);
  right = um < Pr;
  if (right) {
    x = R(PG_T) - log(u01<R>(w.y)) / K;
    U = um / Pr;
    return true;
  }
  U = (um - Pr) / (R(1) - Pr);
  const R z = normal2r<R>(w.y, w.z);
  x = ig_msh<R>(R(1) / c, R(1), z, u01<R>(w.w));
  return x < R(PG_T);
}
template <typename R>
__device__ __forceinline__ bool pg_attempt_B(R c, uint4 w, R& X) {
  R x, U;
  bool right;
  if (!pg_propose_B<R>(c, w, x, U, right)) return false;
  if (!series_accept<R>(U, x, right)) return false;
  X = x;
  return true;
}

// ---- Float64 fast path: attempt 0 of one cell without a branch ----
// Both pieces of the Method-A envelope are evaluated and the piece is selected at the end (a warp whose lanes sit in different pieces
// runs one instruction stream), and the alternating series is cut after its first term.  Returns omega >= 0 when attempt 0 is
// accepted at the first partial sum, U <= 1 - a_1/a_0 -- which is also the first exit of series_accept, and a cell that the exact
// loop would carry to the third partial sum instead (a rounding-level difference in U) is accepted there as well, because
// S_3 - S_1 = 5 e^{6e} - 7 e^{12e} > 0 by far more than a rounding error.  Returns -2 otherwise (rejected by the tilt, not decided
// by one term, |z| beyond the attempt-0 range, NaN): those cells are drawn by pg_draw_exact from attempt 0 on.  The accepted
// value is the same expression as in pg_propose_A up to the log of the left piece, ln(up P0) = ln up + ln P0 (one rounding).
constexpr double PG_LN_P0 = -2.2476256772143173;
constexpr double PG_Z0MAX_D_FAST = 16.0;           // == PG_Z0MAX_D (declared below)  // ln(PG_P0) = -PG_M2LNP0 / 2
__device__ __forceinline__ double pg_attempt0_f64(double z, uint32_t wa, uint32_t wb) {
  const double c = 0.5 * fabs(z);
  const double cc = c * c;
  const double um = u01d(wa), up = u01d(wb);
  const double K = (PI_D * PI_D / 8.0) + 0.5 * cc;
  const double Rm1 = (2.0 * PG_Q0 / PI_D) * K * exp(K * PG_T);  // q0 / p
  const double v = um * (1.0 + Rm1);
  const bool right = v < 1.0;
  const double lu = log(up);
  const double r = rsqrt(-2.0 * (lu + PG_LN_P0));
  const double rp = r * xqd_poly(r);                              // 1 / Z
  const double xl = rp * rp;                                      // truncated Levy piece, 1 / Z^2
  const double inv = 1.0 / (K * xl);                              // one division for 1/K and 1/xl
  const double xr = fma(-lu, xl * inv, PG_T);                     // exponential tail, t - ln(up) / K
  const double tilt = exp(-0.5 * cc * xl);
  const double e = right ? (-0.5 * PI_D * PI_D) * xr : -2.0 * (K * inv);
  const double S1 = 1.0 - 3.0 * exp(2.0 * e);                     // first partial sum of the alternating series
  // right: v <= S1;  left: ua = (v - 1) / Rm1 < tilt and ua / tilt <= S1, i.e. v - 1 <= S1 tilt Rm1 (S1 < 1)
  const bool acc = right ? v <= S1 : (v - 1.0) <= S1 * tilt * Rm1;
  const double x = right ? xr : xl;
  return (acc && fabs(z) <= PG_Z0MAX_D_FAST) ? 0.25 * x : -2.0;
}

constexpr uint32_t PG_MAX_ATTEMPTS = 2000u;  // bound on every device loop; acceptance is >= 0.2, so 0.8^2000 never happens

// Complete draw of omega_ij ~ PG(1, z).  Stream layout (identical in oracle/pg.c):
//   attempt 0            Method A from the cell pair's block (site PK_PG, index j/2; words 0,1 for even j, 2,3 for odd j),
//                        evaluated only when |z| <= PG_Z0MAX_D (beyond that e^{Kt} leaves the f32 range of the fast path);
//   retry block r >= 1   (site PK_PG_RETRY, index j, ctr.w = r): c <= 1/t: Method-A attempts 2r-1 (words 0,1) and 2r (words 2,3);
//                        c > 1/t: Method-B attempt r (all four words).
// `skip_attempt0` resumes after an attempt 0 already known to be rejected.
constexpr double PG_Z0MAX_D = 16.0;
static_assert(PG_Z0MAX_D == PG_Z0MAX_D_FAST, "the branch-free Float64 attempt 0 and the exact loop must agree on the attempt-0 range");
template <typename R>
__device__ __forceinline__ R pg_draw_exact(PhiloxKey key, uint32_t person, uint32_t sweep, int j, R z, int skip_attempt0,
                                           uint32_t* n_attempts = nullptr) {
  const R c = R(0.5) * fabs(z);
  R X = R(PG_T);
  uint32_t used = 0;
  if (!(c >= R(0))) {  // NaN in, NaN out: never spin on poisoned state (the reference would throw)
    if (n_attempts) *n_attempts = 0;
    return z;
  }
  // ONE loop over the attempts with one call site per function: the lanes of a warp sit at different attempt numbers of different
  // cells and still run the same instructions, and the code of an attempt exists once (it was inlined three times)
  const bool method_b = !(c <= R(PG_CSWITCH));
  uint32_t r = (!skip_attempt0 && fabs(z) <= R(PG_Z0MAX_D)) ? 0u : 1u;  // block number: 0 = the cell pair's block, r >= 1 = retry block r
  bool second = false;  // Method A: the second attempt of retry block r (words 2, 3)
  bool done = false;
  uint4 w = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  while (!done && r < PG_MAX_ATTEMPTS) {
    if (!second) w = philox(key, person, sweep, r == 0u ? make_site(DOM_PERSON, PK_PG, (uint32_t)(j >> 1)) : make_site(DOM_PERSON, PK_PG_RETRY, (uint32_t)j), r);
    ++used;
    R x, U;
    bool right, alive;
    if (r != 0u && method_b) {
      alive = pg_propose_B<R>(c, w, x, U, right);
      ++r;
    } else {
      const bool hi = r == 0u ? (j & 1) != 0 : second;
      alive = pg_propose_A<R>(c, hi ? w.z : w.x, hi ? w.w : w.y, x, U, right);
      if (r == 0u || second) { ++r; second = false; }
      else second = true;
    }
    if (alive && series_accept<R>(U, x, right)) {
      X = x;
      done = true;
    }
  }
  if (n_attempts) *n_attempts = used;
  return R(0.25) * X;
}

}  // namespace erirt

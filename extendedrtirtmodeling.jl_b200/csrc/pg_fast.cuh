// pg_fast.cuh -- the f32 Polya-Gamma fast path: two cells per instruction with the packed FP32 pipe of sm_100
// (fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2), MUFU approximations for the transcendental steps.
//
// Statement of the draw: pg.cuh / oracle/pg.c (omega_ij ~ PG(1, z), Draw.pl.jl:38).  Attempt 0 of a cell is always a
// Method-A attempt from the cell pair's Philox block; this file evaluates it so that it only ever concludes
//   * "certainly accepted"  (v below the acceptance threshold computed with the UPPER bounds of a_1/a_0), or
//   * "certainly rejected"  (left piece, v above the threshold computed with a_1 = 0),
// and everything in between (about 0.5 % of the cells) is replayed with the a_1 term by pg_resolve_f32.  Neglecting
// a_2/a_0 <= 5 e^{-3 pi^2 t} = 2.9e-8 (x > t) / 5 e^{-12/t} = 3.6e-8 (x <= t) is below the resolution of an f32 uniform.
#pragma once
#include "pg.cuh"

namespace erirt {

typedef unsigned long long u64;

// ---- packed f32x2 helpers (a u64 holds {lo, hi}; ptxas keeps them in aligned register pairs, so packing is free) ----
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 bc2(float s) { return pk2(s, s); }
__device__ __forceinline__ float lo2(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi2(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ex2_2(u64 a) { return pk2(fast_ex2(lo2(a)), fast_ex2(hi2(a))); }
__device__ __forceinline__ u64 lg2_2(u64 a) { return pk2(fast_lg2(lo2(a)), fast_lg2(hi2(a))); }
__device__ __forceinline__ u64 rcp_2(u64 a) { return pk2(fast_rcp(lo2(a)), fast_rcp(hi2(a))); }
__device__ __forceinline__ u64 rsqrt_2(u64 a) { return pk2(fast_rsqrt(lo2(a)), fast_rsqrt(hi2(a))); }
__device__ __forceinline__ u64 u01_2(uint32_t w0, uint32_t w1) {  // same rounding as u01f
  return ffma2(pk2((float)w0, (float)w1), bc2(2.3283064365386963e-10f), bc2(1.1641532182693481e-10f));
}

constexpr float PGF_PI2_8 = 1.2337005501361697f;     // pi^2/8
constexpr float PGF_EA_K = 0.11541560327111708f;     // t log2(e) / 8
constexpr float PGF_EA_B = -0.7550282017651118f;     // (pi^2/8) t log2(e) + log2(2 q0 / pi)
constexpr float PGF_LOG2E = 1.4426950408889634f;
constexpr float PGF_LN2 = 0.6931471805599453f;
constexpr float PGF_PI2LOG2E = 14.238829324987504f;  // pi^2 log2(e)
constexpr float PG_Z0MAX = 16.0f;                    // attempt 0 is evaluated only for |z| <= 16 (finite in f32), see pg.cuh

// 0.5 * s/Z as a polynomial in r (coefficients of xq_poly halved, so that (r * p)^2 = X/4 = omega)
__device__ __forceinline__ u64 xq_half_poly2(u64 r) {
  const u64 x = ffma2(r, bc2((float)ERIRT_XQ_A), bc2((float)ERIRT_XQ_B));
  constexpr float c[ERIRT_XQ_DEG + 1] = ERIRT_XQ_COEFFS;
  u64 p = bc2(0.5f * c[ERIRT_XQ_DEG]);
#pragma unroll
  for (int k = ERIRT_XQ_DEG - 1; k >= 0; --k) p = ffma2(p, x, bc2(0.5f * c[k]));
  return p;
}

// Shared part of an attempt pair: everything up to the proposals and thresholds.
struct PgPair {
  u64 v;     // um (1 + q0/p)
  u64 Xr4;   // right proposal / 4
  u64 Xl4;   // left proposal / 4
  u64 fm;    // (q0/p) * tilt
};
__device__ __forceinline__ PgPair pg_pair_core(u64 zz, u64 K, u64 Rm1, u64 rK, uint32_t wa0, uint32_t wb0, uint32_t wa1, uint32_t wb1) {
  PgPair o;
  const u64 um = u01_2(wa0, wa1), up = u01_2(wb0, wb1);
  o.v = ffma2(um, Rm1, um);
  const u64 L = lg2_2(up);
  o.Xr4 = ffma2(fmul2(L, bc2(-0.25f * PGF_LN2)), rK, bc2(0.25f * (float)PG_T));     // (t - ln(up)/K) / 4
  const u64 Lw = ffma2(L, bc2(-2.0f * PGF_LN2), bc2((float)PG_M2LNP0));             // -2 ln(up P0)
  const u64 r = rsqrt_2(Lw);
  const u64 rz = fmul2(r, xq_half_poly2(r));                                        // 1/(2Z)
  o.Xl4 = fmul2(rz, rz);                                                            // 1/(4 Z^2)
  const u64 tilt = ex2_2(fmul2(fmul2(zz, o.Xl4), bc2(-0.5f * PGF_LOG2E)));          // exp(-c^2 X/2), c^2 = zz/4, X = 4 Xl4
  o.fm = fmul2(Rm1, tilt);
  return o;
}

// Decision of one cell, branch-free (selp / predicated or): certain accept -> omega, else -2 and its bit in dmask; certain
// reject -> its bit in rmask.  right piece: v < 1 - max a_1/a_0;  left piece: 1 <= v < thl;  reject: v >= thj.
template <int BIT>
__device__ __forceinline__ float pg_decide(float v, float xr4, float xl4, float thl, float thj, uint32_t& dmask, uint32_t& rmask) {
  float out;
  asm("{\n"
      ".reg .pred pr, pa, pj;\n"
      ".reg .f32 x, th;\n"
      "setp.lt.f32 pr, %3, 0f3F800000;\n"
      "selp.f32 x, %4, %5, pr;\n"
      "selp.f32 th, %8, %6, pr;\n"
      "setp.lt.f32 pa, %3, th;\n"
      "selp.f32 %0, x, 0fC0000000, pa;\n"
      "@!pa or.b32 %1, %1, %9;\n"
      "setp.ge.f32 pj, %3, %7;\n"
      "@pj or.b32 %2, %2, %9;\n"
      "}"
      : "=f"(out), "+r"(dmask), "+r"(rmask)
      : "f"(v), "f"(xr4), "f"(xl4), "f"(thl), "f"(thj), "f"(1.0f - (float)PG_R1MAX_RIGHT), "n"(1u << BIT));
  return out;
}

// ---- the hot function: attempt 0 of two cells (z2 = {z_a, z_b}), certain outcomes only ----
// out_x: omega (>= 0) when certainly accepted, -2 otherwise; bit BIT / BIT+1 of `dmask` is set for a cell that is not
// certainly accepted, of `rmask` for a cell that is certainly rejected.  `prod` and `sabs` carry the Bernoulli
// log-likelihood pieces: prod *= 1 + e^{-|z|}, sabs += |z|.
template <int BIT>
__device__ __forceinline__ void pg_fast_pair(u64 z2, uint32_t wa0, uint32_t wb0, uint32_t wa1, uint32_t wb1, float& out0, float& out1,
                                             uint32_t& dmask, uint32_t& rmask, u64& prod, float& sabs) {
  const float za = lo2(z2), zb = hi2(z2);
  prod = ffma2(prod, pk2(fast_ex2(-PGF_LOG2E * fabsf(za)), fast_ex2(-PGF_LOG2E * fabsf(zb))), prod);
  sabs += fabsf(za);
  sabs += fabsf(zb);
  const u64 zz = fmul2(z2, z2);
  const u64 K = ffma2(zz, bc2(0.125f), bc2(PGF_PI2_8));
  const u64 Rm1 = fmul2(K, ex2_2(ffma2(zz, bc2(PGF_EA_K), bc2(PGF_EA_B))));         // q0/p = (2 q0/pi) K e^{Kt}
  const PgPair a = pg_pair_core(zz, K, Rm1, rcp_2(K), wa0, wb0, wa1, wb1);
  const u64 thl = ffma2(a.fm, bc2(1.0f - (float)PG_R1MAX_LEFT), bc2(1.0f));         // left: v < 1 + fm (1 - max a_1/a_0)
  const u64 thj = fadd2(a.fm, bc2(1.0f));                                           // v >= 1 + fm: certainly rejected
  out0 = pg_decide<BIT>(lo2(a.v), lo2(a.Xr4), lo2(a.Xl4), lo2(thl), lo2(thj), dmask, rmask);
  out1 = pg_decide<BIT + 1>(hi2(a.v), hi2(a.Xr4), hi2(a.Xl4), hi2(thl), hi2(thj), dmask, rmask);
}

// ---- exact (to the a_1 term) evaluation of two Method-A attempts of ONE cell: words (wa0, wb0) and (wa1, wb1) ----
// Returns omega of the first accepted attempt, or -2 when both are rejected.
__device__ __forceinline__ float pg_exact_pair(float z, uint32_t wa0, uint32_t wb0, uint32_t wa1, uint32_t wb1) {
  const float zz1 = z * z;
  const float K1 = fmaf(zz1, 0.125f, PGF_PI2_8);
  const float Rm11 = K1 * fast_ex2(fmaf(zz1, PGF_EA_K, PGF_EA_B));
  const PgPair a = pg_pair_core(bc2(zz1), bc2(K1), bc2(Rm11), bc2(fast_rcp(K1)), wa0, wb0, wa1, wb1);
  // right: accept iff v <= 1 - 3 exp(-pi^2 X);  left: accept iff v - 1 < fm (1 - 3 exp(-4/X))
  const u64 r1r = ex2_2(fmul2(a.Xr4, bc2(-4.0f * PGF_PI2LOG2E)));
  const u64 r1l = ex2_2(fmul2(rcp_2(a.Xl4), bc2(-PGF_LOG2E)));
  const u64 thr = ffma2(r1r, bc2(-3.0f), bc2(1.0f));
  const u64 thl = ffma2(a.fm, ffma2(r1l, bc2(-3.0f), bc2(1.0f)), bc2(1.0f));
  const float v0 = lo2(a.v), v1 = hi2(a.v);
  const bool right0 = v0 < 1.0f, right1 = v1 < 1.0f;
  const bool acc0 = right0 ? v0 <= lo2(thr) : v0 < lo2(thl);
  const bool acc1 = right1 ? v1 <= hi2(thr) : v1 < hi2(thl);
  const float o0 = right0 ? lo2(a.Xr4) : lo2(a.Xl4), o1 = right1 ? hi2(a.Xr4) : hi2(a.Xl4);
  return acc0 ? o0 : (acc1 ? o1 : -2.0f);
}

// One fast attempt of Method B (c > 1/t) from the four words of a retry block; omega, or -2 when rejected.
__device__ __forceinline__ float pg_fast_attemptB(float z, uint4 w) {
  const float c = 0.5f * fabsf(z);
  const float K = fmaf(0.5f * c, c, PGF_PI2_8);
  const float rK = fast_rcp(K);
  const float p = (float)(PI_D / 2.0) * rK * fast_ex2(-K * (float)(PG_T * 1.4426950408889634));
  const float ql = 2.0f * fast_ex2(-PGF_LOG2E * c);
  const float Pr = p * fast_rcp(p + ql);
  const float um = u01f(w.x);
  if (um < Pr) {
    const float X = fmaf(-PGF_LN2 * fast_lg2(u01f(w.y)), rK, (float)PG_T);
    const float r1 = 3.0f * fast_ex2(-PGF_PI2LOG2E * X);
    return (um <= Pr * (1.0f - r1)) ? 0.25f * X : -2.0f;
  }
  const float ua = (um - Pr) * fast_rcp(1.0f - Pr);
  const float X = ig_msh<float>(fast_rcp(c), 1.0f, normal2f(w.y, w.z), u01f(w.w));
  if (!(X < (float)PG_T)) return -2.0f;
  const float r1 = 3.0f * fast_ex2((-4.0f * PGF_LOG2E) * fast_rcp(X));
  return (ua <= 1.0f - r1) ? 0.25f * X : -2.0f;
}

// Retry rounds of one cell from block r0 on (block r supplies Method-A attempts 2r-1, 2r or Method-B attempt r).
__device__ __forceinline__ float pg_retry_f32_inl(PhiloxKey key, uint32_t gid, uint32_t sweep, int j, float z, uint32_t r0) {
  if (!(z == z)) return z;  // poisoned state: do not spin
  const bool method_b = 0.5f * fabsf(z) > (float)PG_CSWITCH;
#pragma unroll 1
  for (uint32_t r = r0; r < PG_MAX_ATTEMPTS; ++r) {
    const uint4 w = philox(key, gid, sweep, make_site(DOM_PERSON, PK_PG_RETRY, (uint32_t)j), r);
    const float om = method_b ? pg_fast_attemptB(z, w) : pg_exact_pair(z, w.x, w.y, w.z, w.w);
    if (om >= 0.f) return om;
  }
  return 0.25f * (float)PG_T;
}
// Complete resolution of a cell that left the fast path: replay attempt 0 with the a_1 term when it was undecided, then retry.
__device__ __noinline__ float pg_resolve_f32(PhiloxKey key, uint32_t gid, uint32_t sweep, int j, float z, bool replay0) {
  if (replay0 && fabsf(z) <= PG_Z0MAX) {
    const uint4 w = philox(key, gid, sweep, make_site(DOM_PERSON, PK_PG, (uint32_t)(j >> 1)), 0);
    const uint32_t wa = (j & 1) ? w.z : w.x, wb = (j & 1) ? w.w : w.y;
    const float om = pg_exact_pair(z, wa, wb, wa, wb);
    if (om >= 0.f) return om;
  }
  return pg_retry_f32_inl(key, gid, sweep, j, z, 1u);
}

}  // namespace erirt

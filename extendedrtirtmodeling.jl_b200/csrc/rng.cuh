// rng.cuh -- counter-based Philox4x32-10 stream and scalar variates (device side).
//
// Replaces Julia's global RNG, implicit in every rand/randn of /root/reference/src/Draw.pl.jl, and the
// samplers of Distributions.jl (Normal, Truncated{Normal}, InverseGamma, InverseGaussian, InverseWishart;
// call sites Draw.pl.jl:57,91,101,218,260,312,335,505,526,546,569,596).
// Contract (DESIGN.md "Random stream"): every word is a pure function of
//   key = { seed_lo, seed_hi ^ chain*0x9E3779B9 },  ctr = { unit, sweep, site, attempt }
// with site = domain<<28 | kind<<20 | index, so results are independent of launch geometry and of how
// persons are sharded over GPUs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace erirt {

constexpr uint32_t DOM_PERSON = 1u, DOM_ITEM = 2u, DOM_GLOBAL = 3u;
// person-domain kinds (unit = global person id)
constexpr uint32_t PK_NORMALS = 0;   // words 0,1 -> theta normal; words 2,3 -> zeta normal
constexpr uint32_t PK_NU = 1;        // LatentQr nu_i: words 0,1 normal, word 2 uniform
constexpr uint32_t PK_PG = 2;        // attempt 0 of cells (i,2*idx) [words 0,1] and (i,2*idx+1) [words 2,3]
constexpr uint32_t PK_PG_RETRY = 3;  // attempts >= 1 of cell (i,idx), attempt number in ctr.w
constexpr uint32_t PK_NU_CELL = 4;   // CrossQr nu_ij, idx = j/2: words 0,1 -> Box-Muller pair (cos: j even, sin: j odd), word 2 / 3 uniform
// item-domain kinds (unit = item)
constexpr uint32_t IK_B = 0, IK_A = 1, IK_LAMBDA = 2, IK_SIGMA2 = 3, IK_RHO = 4;
// global-domain kinds (unit = component)
constexpr uint32_t GK_BETA = 0, GK_SIGMAP = 1;
// data domain (erirt_generate_data, sweep = 0, unit = global person id): DK_Y idx = j/4, word j%4 -> uniform of Y_ij;
// DK_LOGT idx = j -> words 0,1 normal of the logT error (truncated normal: attempts in ctr.w), words 2,3 a spare normal;
// DK_AUX idx = j -> two Box-Muller pairs (the chi-square of the t5 errors)
constexpr uint32_t DOM_DATA = 4u;
constexpr uint32_t DK_Y = 0, DK_LOGT = 1, DK_AUX = 2;

__host__ __device__ constexpr uint32_t make_site(uint32_t dom, uint32_t kind, uint32_t idx = 0) {
  return (dom << 28) | (kind << 20) | idx;
}

struct PhiloxKey {
  uint32_t k0, k1;
};
__host__ __device__ inline PhiloxKey make_key(uint64_t seed, uint32_t chain) {
  PhiloxKey k;
  k.k0 = (uint32_t)(seed & 0xffffffffu);
  k.k1 = (uint32_t)(seed >> 32) ^ (chain * 0x9E3779B9u);
  return k;
}

// One Philox4x32-10 block.  Per round: 2 IMAD.WIDE + 2 LOP3; the key schedule is warp-uniform.
__device__ __forceinline__ uint4 philox(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site, uint32_t attempt) {
  uint32_t c0 = unit, c1 = sweep, c2 = site, c3 = attempt;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// The same block from a precomputed key schedule (round r uses rk[2r], rk[2r+1]).  The schedule lives in the kernel
// parameters, i.e. in the constant bank, so each round key is a direct operand of the round's LOP3 instead of an add.
struct PhiloxSched {
  uint32_t rk[20];
};
__host__ __device__ inline PhiloxSched make_sched(PhiloxKey key) {
  PhiloxSched s;
  for (int r = 0; r < 10; ++r) {
    s.rk[2 * r] = key.k0 + (uint32_t)r * 0x9E3779B9u;
    s.rk[2 * r + 1] = key.k1 + (uint32_t)r * 0xBB67AE85u;
  }
  return s;
}
__device__ __forceinline__ uint4 philox(const PhiloxSched& ks, uint32_t unit, uint32_t sweep, uint32_t site, uint32_t attempt) {
  uint32_t c0 = unit, c1 = sweep, c2 = site, c3 = attempt;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.rk[2 * r];
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.rk[2 * r + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}

// uniforms on (0,1): f64 keeps all 32 bits (exact); f32 rounds (w + 0.5) 2^-32 to 24 bits (may return 1.0f)
__device__ __forceinline__ double u01d(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }
__device__ __forceinline__ float u01f(uint32_t w) { return fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
template <typename R>
__device__ __forceinline__ R u01(uint32_t w);
template <>
__device__ __forceinline__ double u01<double>(uint32_t w) { return u01d(w); }
template <>
__device__ __forceinline__ float u01<float>(uint32_t w) { return u01f(w); }

// MUFU approximations (2^-22 relative error), one SASS instruction each
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// f32 uses the approximate reciprocal / rsqrt (well inside the 1e-5 contract), f64 the IEEE operations
__device__ __forceinline__ float rdiv(float a, float b) { return a * fast_rcp(b); }
__device__ __forceinline__ double rdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float sqrt_of(float x) { return x * fast_rsqrt(x); }  // sqrt(x), x > 0
__device__ __forceinline__ float rlog(float x) { return 0.6931471805599453f * fast_lg2(x); }
__device__ __forceinline__ double rlog(double x) { return log(x); }
__device__ __forceinline__ double sqrt_of(double x) { return sqrt(x); }

// Box-Muller cosine branch
__device__ __forceinline__ double normal2(uint32_t w0, uint32_t w1) {
  return sqrt(-2.0 * log(u01d(w0))) * cospi(2.0 * u01d(w1));
}
__device__ __forceinline__ float normal2f(uint32_t w0, uint32_t w1) {
  const float u = fminf(u01f(w0), 0.99999994f);
  const float t = -1.3862943611198906f * fast_lg2(u);  // -2 ln u > 0
  return t * fast_rsqrt(t) * __cosf(6.283185307179586f * u01f(w1));
}
template <typename R>
__device__ __forceinline__ R normal2r(uint32_t w0, uint32_t w1);
template <>
__device__ __forceinline__ double normal2r<double>(uint32_t w0, uint32_t w1) { return normal2(w0, w1); }
template <>
__device__ __forceinline__ float normal2r<float>(uint32_t w0, uint32_t w1) { return normal2f(w0, w1); }

// ---- item/global-site variates, always f64 (J + O(F) draws per sweep) ----
__device__ __noinline__ double site_normal(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site) {
  uint4 w = philox(key, unit, sweep, site, 0);
  return normal2(w.x, w.y);
}

// N(mu, sd^2) truncated to (0, inf): rejection from the parent when the standardised bound alpha = -mu/sd <= 0.5,
// Robert's (1995) translated-exponential proposal otherwise.
__device__ __noinline__ double site_tnorm_pos(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site, double mu, double sd,
                                               uint32_t att0 = 0) {
  double alpha = -mu / sd;
  if (!(alpha == alpha)) return alpha;  // NaN parameters: propagate instead of spinning
  if (alpha <= 0.5) {
    for (uint32_t att = att0;; ++att) {
      uint4 w = philox(key, unit, sweep, site, att);
      double z = normal2(w.x, w.y);
      if (z >= alpha || att > 10000u) return mu + sd * z;
    }
  }
  double lam = 0.5 * (alpha + sqrt(alpha * alpha + 4.0));
  for (uint32_t att = att0;; ++att) {
    uint4 w = philox(key, unit, sweep, site, att);
    double x = alpha - log(u01d(w.x)) / lam;
    double d = x - lam;
    if (u01d(w.y) <= exp(-0.5 * d * d) || att > 10000u) return mu + sd * x;
  }
}

// Gamma(shape, 1) by Marsaglia-Tsang (2000); shape >= 1 here (delta + N/2, (N+3)/2, ...)
__device__ __noinline__ double site_gamma(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site, double shape, uint32_t att0 = 0) {
  if (!(shape > 0.0)) return shape * 0.0 / 0.0;  // NaN / non-positive shape
  double boost = 1.0;
  if (shape < 1.0) {
    uint4 w = philox(key, unit, sweep, site, 0xffffffffu);
    boost = pow(u01d(w.x), 1.0 / shape);
    shape += 1.0;
  }
  double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t att = att0;; ++att) {
    uint4 w = philox(key, unit, sweep, site, att);
    double x = normal2(w.x, w.y);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = u01d(w.z);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v) || att > 10000u) return boost * d * v;
  }
}

// ---- the same variates from the raw material of attempt 0 (z0 = normal2(w.x, w.y), u0 = u01d(w.z) of the site's attempt-0
//      block), which the global kernel computes for all items at once on separate lanes before the parameters are known.
//      Decision-identical to the functions above: attempt 0 is judged from the raw values, later attempts fall back to them. ----
__device__ inline double tnorm_pos_from_raw(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site, double mu, double sd, double z0) {
  const double alpha = -mu / sd;
  if (!(alpha == alpha)) return alpha;
  if (alpha <= 0.5) return z0 >= alpha ? mu + sd * z0 : site_tnorm_pos(key, unit, sweep, site, mu, sd, 1u);
  return site_tnorm_pos(key, unit, sweep, site, mu, sd, 0u);  // Robert's exponential proposal reads the words differently (rare)
}
__device__ inline double gamma_from_raw(PhiloxKey key, uint32_t unit, uint32_t sweep, uint32_t site, double shape, double x0, double u0) {
  if (!(shape >= 1.0)) return site_gamma(key, unit, sweep, site, shape, 0u);  // boost path / NaN
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double v = 1.0 + c * x0;
  if (v > 0.0) {
    v = v * v * v;
    const double x2 = x0 * x0;
    // Marsaglia-Tsang's squeeze u < 1 - 0.0331 x^4 implies the exact test, so it never changes a decision; it skips two logs
    if (u0 < 1.0 - 0.0331 * x2 * x2 || log(u0) < 0.5 * x2 + d - d * v + d * log(v)) return d * v;
  }
  return site_gamma(key, unit, sweep, site, shape, 1u);
}

// Inverse Gaussian by Michael-Schucany-Haas (1976) from a normal z and a uniform u; the smaller root is
// written without cancellation: mu + mu/(2 lam)(w - sqrt(w(4 lam + w))) == 2 lam mu/(2 lam + w + sqrt(w(4 lam + w))), w = mu z^2.
template <typename R>
__device__ __forceinline__ R ig_msh(R mu, R lam, R z, R u) {
  R y = z * z;
  if (!isfinite(mu)) return rdiv(lam, y);
  R w = mu * y;
  R x1 = rdiv(R(2) * lam * mu, R(2) * lam + w + sqrt_of(w * (R(4) * lam + w) + R(1e-30)));
  return (u * (mu + x1) <= mu) ? x1 : rdiv(mu * mu, x1);
}

}  // namespace erirt

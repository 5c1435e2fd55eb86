/* oracle/rng.c -- Philox4x32-10 and the scalar variate generators.  TEST INFRASTRUCTURE ONLY.
 *
 * Replaces Julia's global RNG (implicit in every rand/randn of src/Draw.pl.jl) and the samplers of
 * Distributions.jl (un-vendored; algorithms restated from the literature):
 *   Normal            -> Box-Muller                                  (Draw.pl.jl:57,74,101,125,...)
 *   Truncated(Normal) -> rejection / Robert (1995) exponential tail  (Draw.pl.jl:91,218,228,248)
 *   Gamma             -> Marsaglia-Tsang (2000)                      (InverseGamma: Draw.pl.jl:260,270,286,546,569,596)
 *   InverseGaussian   -> Michael-Schucany-Haas (1976)                (Draw.pl.jl:312,335)
 */
#include "rng.h"
#include "oracle.h"

#define ORC_MAX_ATTEMPTS 100000u /* NaN inputs must terminate (the device loops have the same bound) */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, uint32_t attempt, uint32_t w[4]) {
  uint32_t ctr[4] = {unit, sweep, site, attempt};
  uint32_t k[2] = {key.k0, key.k1};
  orc_philox4x32_10(ctr, k, w);
}

double orc_site_normal(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site) {
  uint32_t w[4];
  orc_philox(key, unit, sweep, site, 0, w);
  return orc_normal2(w[0], w[1]);
}

/* N(mu, sd^2) truncated to (0, inf).  alpha = -mu/sd is the standardised lower bound. */
double orc_site_tnorm_pos(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, double mu, double sd) {
  double alpha = -mu / sd;
  uint32_t w[4];
  if (alpha <= 0.5) {
    for (uint32_t att = 0; att < ORC_MAX_ATTEMPTS; ++att) {
      orc_philox(key, unit, sweep, site, att, w);
      double z = orc_normal2(w[0], w[1]);
      if (z >= alpha) return mu + sd * z;
    }
    return NAN; /* NaN parameters (the reference would throw here) */
  }
  double lam = 0.5 * (alpha + sqrt(alpha * alpha + 4.0));
  for (uint32_t att = 0; att < ORC_MAX_ATTEMPTS; ++att) {
    orc_philox(key, unit, sweep, site, att, w);
    double x = alpha - log(orc_u01(w[0])) / lam;
    double d = x - lam;
    if (orc_u01(w[1]) <= exp(-0.5 * d * d)) return mu + sd * x;
  }
  return NAN;
}

/* Gamma(shape, 1), shape >= 1 (always true here: shape = delta + N/2 etc.) */
double orc_site_gamma(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, double shape) {
  double boost = 1.0;
  uint32_t w[4];
  if (shape < 1.0) { /* Gamma(a) = Gamma(a+1) * U^(1/a); uniform from attempt 0xffffffff */
    orc_philox(key, unit, sweep, site, 0xffffffffu, w);
    boost = pow(orc_u01(w[0]), 1.0 / shape);
    shape += 1.0;
  }
  double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t att = 0; att < ORC_MAX_ATTEMPTS; ++att) {
    orc_philox(key, unit, sweep, site, att, w);
    double x = orc_normal2(w[0], w[1]);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = orc_u01(w[2]);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
  }
  return NAN;
}

double orc_ig_msh(double mu, double lam, double z, double u) {
  double y = z * z;
  if (!isfinite(mu)) return lam / y; /* Levy limit */
  double w = mu * y;
  /* smaller root of the MSH quadratic, written without cancellation:
     mu + mu/(2 lam) (w - sqrt(w (4 lam + w)))  ==  2 lam mu / (2 lam + w + sqrt(w (4 lam + w))) */
  double x1 = 2.0 * lam * mu / (2.0 * lam + w + sqrt(w * (4.0 * lam + w)));
  return (u <= mu / (mu + x1)) ? x1 : mu * mu / x1;
}

void orc_variates(int kind, double p1, double p2, int64_t n, uint64_t seed, double* out) {
  orc_key key = orc_make_key(seed, 0);
  for (int64_t i = 0; i < n; ++i) {
    uint32_t unit = (uint32_t)(i & 0xfffff), sweep = (uint32_t)(i >> 20) + 1;
    switch (kind) {
      case 0: out[i] = p1 + p2 * orc_site_normal(key, unit, sweep, SITE(DOM_ITEM, IK_B, 0)); break;
      case 1: out[i] = orc_site_tnorm_pos(key, unit, sweep, SITE(DOM_ITEM, IK_A, 0), p1, p2); break;
      case 2: out[i] = orc_site_gamma(key, unit, sweep, SITE(DOM_ITEM, IK_SIGMA2, 0), p1); break;
      case 3: out[i] = p2 / orc_site_gamma(key, unit, sweep, SITE(DOM_ITEM, IK_SIGMA2, 0), p1); break;
      default: out[i] = NAN;
    }
  }
}

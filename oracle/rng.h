/* oracle/rng.h -- counter-based random stream contract of the Gibbs engine (oracle side).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Every random number is a pure function of (seed, chain, unit, sweep, site, attempt):
 *   key = { seed_lo, seed_hi ^ (chain * 0x9E3779B9) }
 *   ctr = { unit, sweep, site, attempt }             -> Philox4x32-10 -> 4 words
 *   site = domain<<28 | kind<<20 | index             (index < 2^20)
 * so results do not depend on launch geometry or on how persons are sharded over GPUs.
 */
#ifndef ERIRT_ORACLE_RNG_H
#define ERIRT_ORACLE_RNG_H
#include <stdint.h>
#include <math.h>

#define DOM_PERSON 1u
#define DOM_ITEM 2u
#define DOM_GLOBAL 3u
#define SITE(dom, kind, idx) (((dom) << 28) | ((uint32_t)(kind) << 20) | (uint32_t)(idx))

/* person-domain kinds (unit = global person id) */
#define PK_NORMALS 0  /* words 0,1 -> theta normal ; words 2,3 -> zeta normal */
#define PK_NU 1       /* LatentQr nu_i: words 0,1 normal, word 2 uniform */
#define PK_PG 2       /* attempt 0 of the PG draw of cells (i, 2*idx) [words 0,1] and (i, 2*idx+1) [words 2,3] */
#define PK_PG_RETRY 3 /* attempts >= 1 of cell (i, idx); attempt number in ctr.w */
#define PK_NU_CELL 4  /* CrossQr nu_ij, idx = j/2: words 0,1 -> Box-Muller pair (cos: j even, sin: j odd), word 2 / 3 uniform */
/* item-domain kinds (unit = item j) */
#define IK_B 0
#define IK_A 1
#define IK_LAMBDA 2
#define IK_SIGMA2 3
#define IK_RHO 4
/* global-domain kinds (unit = component) */
#define GK_BETA 0
#define GK_SIGMAP 1

typedef struct { uint32_t k0, k1; } orc_key;

static inline orc_key orc_make_key(uint64_t seed, uint32_t chain) {
  orc_key k;
  k.k0 = (uint32_t)(seed & 0xffffffffu);
  k.k1 = (uint32_t)(seed >> 32) ^ (chain * 0x9E3779B9u);
  return k;
}

void orc_philox(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, uint32_t attempt, uint32_t w[4]);

static inline double orc_u01(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }
static inline double orc_normal2(uint32_t w0, uint32_t w1) {
  return sqrt(-2.0 * log(orc_u01(w0))) * cos(6.283185307179586476925286766559 * orc_u01(w1));
}

/* variates at an item/global site; `attempt` loops internally */
double orc_site_normal(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site);
double orc_site_tnorm_pos(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, double mu, double sd);
double orc_site_gamma(orc_key key, uint32_t unit, uint32_t sweep, uint32_t site, double shape);
/* inverse Gaussian by Michael-Schucany-Haas from (normal z, uniform u) */
double orc_ig_msh(double mu, double lam, double z, double u);
double orc_pg1(orc_key key, uint32_t person, uint32_t sweep, int32_t j, double z, int32_t* attempts);
#endif

/* oracle/gen.c -- CPU restatement of the device-side data generator (erirt_generate_data).
 * TEST INFRASTRUCTURE ONLY (see oracle.h): the checker of the generator kernel, never the product path.
 *
 * Follows the N x J part of the reference's simulators, /root/reference/src/SimTools.jl:
 *   Y_ij ~ Bernoulli(logistic(a_j (theta_i - b_j)))              :131, :165, :241, :329, :363  (rand.(BernoulliLogit.(...)))
 *   logT_ij = lambda_j - zeta_i [- theta_i rho_j] + e_ij
 *     err 0: Truncated(Normal(lambda_j - zeta_i, sqrt(sigma2_j)), 0, Inf)   setDataRtIrtNull :136-139, setDataRtIrt :170-173
 *     err 1: Normal(0, 1)                                                   setDataRtIrtLatent :336
 *     err 2: Normal(0, 0.3)   err 3: TDist(5)   err 4: Gamma(0.5, 1) - 1    setDataRtIrtCross "norm"/"tail"/"skew" :236-247
 * on the counter-based stream of rng.h (data domain: unit = global person id, sweep = 0).
 */
#include <math.h>
#include <stdint.h>
#include "rng.h"

#define DOM_DATA 4u
#define DK_Y 0
#define DK_LOGT 1
#define DK_AUX 2

static void bm_pair(uint32_t w0, uint32_t w1, double* c, double* s) {
  const double r = sqrt(-2.0 * log(orc_u01(w0)));
  const double ang = 6.283185307179586476925286766559 * orc_u01(w1);
  *c = r * cos(ang);
  *s = r * sin(ang);
}

/* Y, logT: column-major n x J (logT may be NULL when has_rt == 0); item vectors of length J; sigma2 / rho may be NULL (1 / 0). */
void orc_generate_data(int64_t n, int32_t J, int64_t person_offset, uint64_t seed, int32_t has_rt, int32_t err, const double* theta,
                       const double* zeta, const double* a, const double* b, const double* lambda, const double* sigma2,
                       const double* rho, double* Y, double* logT) {
  const orc_key key = orc_make_key(seed, 0xDA7Au);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t gid = (uint32_t)(person_offset + i);
    const double th = theta[i], ze = has_rt ? zeta[i] : 0.0;
    uint32_t wy[4], wa[4], wb[4];
    for (int j = 0; j < J; ++j) {
      if ((j & 3) == 0) orc_philox(key, gid, 0u, SITE(DOM_DATA, DK_Y, j >> 2), 0u, wy);
      const double eta = a[j] * (th - b[j]);
      Y[i + n * (int64_t)j] = orc_u01(wy[j & 3]) < 1.0 / (1.0 + exp(-eta)) ? 1.0 : 0.0;
      if (!has_rt) continue;
      const double mu = lambda[j] - ze - th * (rho ? rho[j] : 0.0);
      const uint32_t site = SITE(DOM_DATA, DK_LOGT, j);
      double x;
      if (err == 0) {
        x = orc_site_tnorm_pos(key, gid, 0u, site, mu, sigma2 ? sqrt(sigma2[j]) : 1.0);
      } else {
        orc_philox(key, gid, 0u, site, 0u, wa);
        const double z = orc_normal2(wa[0], wa[1]);
        if (err == 1) x = mu + z;
        else if (err == 2) x = mu + 0.3 * z;
        else if (err == 4) x = mu + (0.5 * z * z - 1.0);
        else { /* t(5) = Z / sqrt(chi2_5 / 5) */
          double n1, n2, n3, n4;
          orc_philox(key, gid, 0u, SITE(DOM_DATA, DK_AUX, j), 0u, wb);
          bm_pair(wb[0], wb[1], &n1, &n2);
          bm_pair(wb[2], wb[3], &n3, &n4);
          const double n5 = orc_normal2(wa[2], wa[3]);
          x = mu + z / sqrt((n1 * n1 + n2 * n2 + n3 * n3 + n4 * n4 + n5 * n5) / 5.0);
        }
      }
      logT[i + n * (int64_t)j] = x;
    }
  }
}

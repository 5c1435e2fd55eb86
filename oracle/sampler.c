/* oracle/sampler.c -- the conditional draws of src/Draw.pl.jl and the seven sample! scans.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Dense, unexpanded float64 sums exactly as the Julia
 * source writes them; arrays are column-major like Julia's (M[i + N*j]).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "rng.h"

#define MAXD 64 /* largest small dense system: 2*(nFeat+1) or nFeat+2 */
#define OMP_FOR _Pragma("omp parallel for schedule(static) num_threads(x->nthreads) if (x->nthreads > 1)")

typedef struct {
  const orc_cfg* c;
  const double *Y, *logT, *X;
  orc_state* st;
  orc_key key;
  int N, J, F, nthreads;
  double k1, k2;       /* k1Rt, k2Rt: Draw.pl.jl:163-164 */
  double muLam, sdLam; /* mean(Data.logT), std(Data.logT): keyword defaults at Draw.pl.jl:215 */
  uint32_t sweep;
  double *tmpN, *tmpN2;
} ctx;

/* ---------- small dense linear algebra ---------- */
static int chol_lower(int n, const double* A, double* L) { /* column-major n x n */
  memset(L, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double d = A[j + n * j];
    for (int k = 0; k < j; ++k) d -= L[j + n * k] * L[j + n * k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    L[j + n * j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i + n * j];
      for (int k = 0; k < j; ++k) s -= L[i + n * k] * L[j + n * k];
      L[i + n * j] = s / d;
    }
  }
  return 0;
}
/* inverse of an SPD matrix through its Cholesky factor */
static int spd_inverse(int n, const double* A, double* Ainv) {
  double L[MAXD * MAXD], Li[MAXD * MAXD];
  if (chol_lower(n, A, L)) return -1;
  memset(Li, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) { /* Li = L^{-1}, lower */
    Li[j + n * j] = 1.0 / L[j + n * j];
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= L[i + n * k] * Li[k + n * j];
      Li[i + n * j] = s / L[i + n * i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = (i > j ? i : j); k < n; ++k) s += Li[k + n * i] * Li[k + n * j];
      Ainv[i + n * j] = s;
    }
  return 0;
}
static int spd_solve(int n, const double* A, const double* rhs, double* sol) {
  double L[MAXD * MAXD], y[MAXD];
  if (chol_lower(n, A, L)) return -1;
  for (int i = 0; i < n; ++i) {
    double s = rhs[i];
    for (int k = 0; k < i; ++k) s -= L[i + n * k] * y[k];
    y[i] = s / L[i + n * i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= L[k + n * i] * sol[k];
    sol[i] = s / L[i + n * i];
  }
  return 0;
}

/* design row x_i = [1, X_i, (theta_i)]; returns length */
static inline double xrow(const ctx* x, int i, int k, int with_theta) {
  if (k == 0) return 1.0;
  if (k <= x->F) return x->X[i + (size_t)x->N * (k - 1)];
  (void)with_theta;
  return x->st->theta[i];
}

/* ---------- IRT part ---------- */

/* drawRaPgRandomVariable, Draw.pl.jl:36-40 */
static void draw_omega(ctx* x) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  OMP_FOR
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < J; ++j) {
      double eta = s->a[j] * (s->theta[i] - s->b[j]);
      s->omega[i + (size_t)N * j] = orc_pg1(x->key, (uint32_t)i, x->sweep, j, eta, NULL);
    }
}

/* drawSubjAbility (Draw.pl.jl:49-62) when use_x, drawSubjAbilityNull (:67-80) otherwise.
 * prior_var = Para.Σp[1,1] (1.0 for GibbsMlIrt, whose Σp is never initialised: SURVEY Q7). */
static void draw_theta(ctx* x, int use_x, double prior_var) {
  const int N = x->N, J = x->J, F = x->F;
  orc_state* s = x->st;
  OMP_FOR
  for (int i = 0; i < N; ++i) {
    double mu0 = 0.0;
    if (use_x) {
      mu0 = s->beta[0];
      for (int k = 0; k < F; ++k) mu0 += x->X[i + (size_t)N * k] * s->beta[k + 1];
    }
    double sA = 0.0, sB = 0.0;
    for (int j = 0; j < J; ++j) {
      double w = s->omega[i + (size_t)N * j], a = s->a[j];
      double kap = x->Y[i + (size_t)N * j] - 0.5;
      sA += a * a * w;
      sB += a * (kap + a * s->b[j] * w);
    }
    double parV = 1.0 / (1.0 / prior_var + sA);
    double parM = parV * (mu0 / prior_var + sB);
    uint32_t w4[4];
    orc_philox(x->key, (uint32_t)i, x->sweep, SITE(DOM_PERSON, PK_NORMALS, 0), 0, w4);
    x->tmpN[i] = parM + sqrt(parV) * orc_normal2(w4[0], w4[1]);
  }
  memcpy(s->theta, x->tmpN, sizeof(double) * N);
}

/* drawItemDiscrimination, Draw.pl.jl:88-93 (mu_a0 = 1, sigma_a0 = 1) */
static void draw_a(ctx* x) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  OMP_FOR
  for (int j = 0; j < J; ++j) {
    double sV = 0.0, sM = 0.0;
    for (int i = 0; i < N; ++i) {
      double d = s->theta[i] - s->b[j];
      sV += d * d * s->omega[i + (size_t)N * j];
      sM += (x->Y[i + (size_t)N * j] - 0.5) * d;
    }
    double parV = 1.0 / (1.0 + sV);
    double parM = parV * (1.0 + sM);
    s->a[j] = orc_site_tnorm_pos(x->key, (uint32_t)j, x->sweep, SITE(DOM_ITEM, IK_A, 0), parM, sqrt(parV));
  }
}

/* drawItemDifficulty, Draw.pl.jl:98-105 (mu_b0 = 0, sigma_b0 = 1, clamp to [-4,4]) */
static void draw_b(ctx* x) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  OMP_FOR
  for (int j = 0; j < J; ++j) {
    double a = s->a[j], sW = 0.0, sM = 0.0;
    for (int i = 0; i < N; ++i) {
      double w = s->omega[i + (size_t)N * j];
      sW += a * a * w;
      sM += a * ((x->Y[i + (size_t)N * j] - 0.5) - s->theta[i] * a * w);
    }
    double parV = 1.0 / (1.0 + sW);
    double parM = parV * (0.0 - sM);
    double b = parM + sqrt(parV) * orc_site_normal(x->key, (uint32_t)j, x->sweep, SITE(DOM_ITEM, IK_B, 0));
    s->b[j] = b < -4.0 ? -4.0 : (b > 4.0 ? 4.0 : b);
  }
}

/* ---------- RT part: person speed ---------- */
enum { Z_NULL, Z_X, Z_LATENT, Z_LATENTQR, Z_CROSS, Z_CROSSQR };

/* drawSubjSpeedNull :119-127, drawSubjSpeed :132-141, ...Latent :147-156, ...LatentQr :161-174,
 * ...Cross :179-187, ...CrossQr :192-206 */
static void draw_zeta(ctx* x, int kind) {
  const int N = x->N, J = x->J, F = x->F;
  orc_state* s = x->st;
  const double S22 = s->Sigma[3];
  OMP_FOR
  for (int i = 0; i < N; ++i) {
    double mu0 = 0.0, var0 = 1.0;
    switch (kind) {
      case Z_NULL: mu0 = 0.0; var0 = 1.0; break; /* fixed 1.0, Draw.pl.jl:121 */
      case Z_X: {
        const double* b2 = s->beta + (F + 1);
        mu0 = b2[0];
        for (int k = 0; k < F; ++k) mu0 += x->X[i + (size_t)N * k] * b2[k + 1];
        var0 = S22;
      } break;
      case Z_LATENT:
      case Z_LATENTQR: {
        mu0 = s->beta[0];
        for (int k = 0; k < F; ++k) mu0 += x->X[i + (size_t)N * k] * s->beta[k + 1];
        mu0 += s->theta[i] * s->beta[F + 1];
        var0 = S22;
        if (kind == Z_LATENTQR) { mu0 += x->k1 * s->nu[i]; var0 = S22 * (x->k2 * s->nu[i]); }
      } break;
      case Z_CROSS:
      case Z_CROSSQR: mu0 = 0.0; var0 = S22; break;
    }
    double sP = 0.0, sM = 0.0;
    for (int j = 0; j < J; ++j) {
      double lt = x->logT[i + (size_t)N * j];
      if (kind == Z_CROSS) {
        sP += 1.0 / s->sigma2[j];
        sM += (s->lambda[j] - lt - s->theta[i] * s->rho[j]) / s->sigma2[j];
      } else if (kind == Z_CROSSQR) {
        double nu = s->nu[i + (size_t)N * j];
        double den = s->sigma2[j] * (x->k2 * nu);
        sP += 1.0 / den;
        sM += (s->lambda[j] - lt - s->theta[i] * s->rho[j] + x->k1 * nu) / den;
      } else {
        sP += 1.0 / s->sigma2[j];
        sM += (s->lambda[j] - lt) / s->sigma2[j];
      }
    }
    double parV = 1.0 / (1.0 / var0 + sP);
    double parM = parV * (mu0 / var0 + sM);
    uint32_t w4[4];
    orc_philox(x->key, (uint32_t)i, x->sweep, SITE(DOM_PERSON, PK_NORMALS, 0), 0, w4);
    x->tmpN[i] = parM + sqrt(parV) * orc_normal2(w4[2], w4[3]);
  }
  memcpy(s->zeta, x->tmpN, sizeof(double) * N);
}

/* ---------- RT part: item parameters ---------- */
enum { RT_PLAIN, RT_CROSS, RT_CROSSQR };

/* drawItemIntensity :215-220, ...Cross :225-231, ...CrossQr :239-251 */
static void draw_lambda(ctx* x, int kind) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  const double pm = x->muLam / (x->sdLam * x->sdLam), pv = 1.0 / (x->sdLam * x->sdLam);
  OMP_FOR
  for (int j = 0; j < J; ++j) {
    double prec = 0.0, sM = 0.0;
    for (int i = 0; i < N; ++i) {
      double lt = x->logT[i + (size_t)N * j];
      if (kind == RT_PLAIN) {
        sM += lt + s->zeta[i];
      } else if (kind == RT_CROSS) {
        sM += (lt + s->zeta[i] + s->theta[i] * s->rho[j]) / s->sigma2[j];
      } else {
        double nu = s->nu[i + (size_t)N * j], den = s->sigma2[j] * (x->k2 * nu);
        prec += 1.0 / den;
        sM += (lt + s->zeta[i] + s->theta[i] * s->rho[j] - x->k1 * nu) / den;
      }
    }
    if (kind == RT_PLAIN) { prec = N / s->sigma2[j]; sM = sM / s->sigma2[j]; }
    else if (kind == RT_CROSS) prec = N / s->sigma2[j];
    double parV = 1.0 / (pv + prec);
    double parM = parV * (pm + sM);
    s->lambda[j] = orc_site_tnorm_pos(x->key, (uint32_t)j, x->sweep, SITE(DOM_ITEM, IK_LAMBDA, 0), parM, sqrt(parV));
  }
}

/* drawItemTimeResidual :257-262, ...Cross :267-273, ...CrossQr :278-288 (delta_a = delta_b = 1e-3) */
static void draw_sigma2(ctx* x, int kind) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  OMP_FOR
  for (int j = 0; j < J; ++j) {
    double ss = 0.0, snu = 0.0;
    for (int i = 0; i < N; ++i) {
      double r = x->logT[i + (size_t)N * j] - s->lambda[j] + s->zeta[i];
      if (kind != RT_PLAIN) r += s->theta[i] * s->rho[j];
      if (kind == RT_CROSSQR) {
        double nu = s->nu[i + (size_t)N * j];
        r -= x->k1 * nu;
        ss += r * r / (2.0 * (x->k2 * nu));
        snu += nu;
      } else {
        ss += r * r;
      }
    }
    double parA, parB;
    if (kind == RT_CROSSQR) { parA = 1e-3 + N * 3 / 2.0; parB = 1e-3 + ss + snu; }
    else { parA = 1e-3 + N / 2.0; parB = 1e-3 + ss / 2.0; }
    s->sigma2[j] = parB / orc_site_gamma(x->key, (uint32_t)j, x->sweep, SITE(DOM_ITEM, IK_SIGMA2, 0), parA);
  }
}

/* ---------- structural: quantile weights ---------- */
static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* drawQrWeightsCrossQr, Draw.pl.jl:303-320 */
static void draw_nu_cell(ctx* x) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  OMP_FOR
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < J; ++j) {
      double sc = sqrt(s->sigma2[j] * x->k2);
      double parA = fabs(x->logT[i + (size_t)N * j] - s->lambda[j] + s->zeta[i] + s->theta[i] * s->rho[j]) / sc;
      double parB = sqrt(2.0 * x->k2 + x->k1 * x->k1) / sc;
      double mu = parB / parA;
      if (mu < 1e-10) mu = 1e-10;
      /* one Philox block serves the cell pair (j, j^1): Box-Muller cosine / sine branch, words 2 / 3 as the MSH uniform */
      uint32_t w4[4];
      orc_philox(x->key, (uint32_t)i, x->sweep, SITE(DOM_PERSON, PK_NU_CELL, j >> 1), 0, w4);
      double rad = sqrt(-2.0 * log(orc_u01(w4[0]))), ang = 6.283185307179586476925286766559 * orc_u01(w4[1]);
      double zn = (j & 1) ? rad * sin(ang) : rad * cos(ang);
      double ig = orc_ig_msh(mu, parB * parB, zn, orc_u01(w4[(j & 1) ? 3 : 2]));
      s->nu[i + (size_t)N * j] = clampd(1.0 / ig, 1e-10, 1e10);
    }
}

/* drawQrWeightsLatentQr, Draw.pl.jl:325-343 */
static void draw_nu_person(ctx* x) {
  const int N = x->N, F = x->F;
  orc_state* s = x->st;
  const double sc = sqrt(s->Sigma[3] * x->k2);
  const double parB = sqrt(2.0 * x->k2 + x->k1 * x->k1) / sc;
  OMP_FOR
  for (int i = 0; i < N; ++i) {
    double xb = s->beta[0];
    for (int k = 0; k < F; ++k) xb += x->X[i + (size_t)N * k] * s->beta[k + 1];
    xb += s->theta[i] * s->beta[F + 1];
    double parA = fabs(s->zeta[i] - xb) / sc;
    double mu = parB / parA;
    if (mu < 1e-10) mu = 1e-10;
    uint32_t w4[4];
    orc_philox(x->key, (uint32_t)i, x->sweep, SITE(DOM_PERSON, PK_NU, 0), 0, w4);
    double ig = orc_ig_msh(mu, parB * parB, orc_normal2(w4[0], w4[1]), orc_u01(w4[2]));
    s->nu[i] = clampd(1.0 / ig, 1e-10, 1e10);
  }
}

/* ---------- structural: regression coefficients ---------- */
/* G = x'x (d x d), g = x'y for design [1 X (theta)] */
static void gram(const ctx* x, int d, const double* y, const double* y2, double* G, double* g, double* g2) {
  const int N = x->N;
  memset(G, 0, sizeof(double) * d * d);
  memset(g, 0, sizeof(double) * d);
  if (g2) memset(g2, 0, sizeof(double) * d);
  double row[MAXD];
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < d; ++k) row[k] = xrow(x, i, k, 1);
    for (int r = 0; r < d; ++r) {
      for (int q = 0; q <= r; ++q) G[r + d * q] += row[r] * row[q];
      g[r] += row[r] * y[i];
      if (g2) g2[r] += row[r] * y2[i];
    }
  }
  for (int r = 0; r < d; ++r)
    for (int q = r + 1; q < d; ++q) G[r + d * q] = G[q + d * r];
}

/* getSubjCoefficientsMlIrt, Draw.pl.jl:351-357: beta = (x'x) \ x'theta, x = [1 X] */
static int get_beta_mlirt(ctx* x) {
  const int p = x->F + 1;
  double G[MAXD * MAXD], g[MAXD];
  gram(x, p, x->st->theta, NULL, G, g, NULL);
  if (spd_solve(p, G, g, x->st->beta)) return -2;
  if (!x->c->intercept) x->st->beta[0] = 0.0; /* GibbsRtIrt.pl.jl:225-227 */
  return 0;
}

static void beta_normals(ctx* x, int n, double* z) {
  for (int k = 0; k < n; ++k) z[k] = orc_site_normal(x->key, (uint32_t)k, x->sweep, SITE(DOM_GLOBAL, GK_BETA, 0));
}

/* drawSubjCoefficients, Draw.pl.jl:380-393 */
static int draw_beta_rtirt(ctx* x) {
  const int p = x->F + 1, d = 2 * p;
  orc_state* s = x->st;
  double G[MAXD * MAXD], gT[MAXD], gZ[MAXD];
  gram(x, p, s->theta, s->zeta, G, gT, gZ);
  const double S11 = s->Sigma[0], S21 = s->Sigma[1], S12 = s->Sigma[2], S22 = s->Sigma[3];
  (void)S21; /* inv(Symmetric(Σp)) reads the upper triangle */
  double det = S11 * S22 - S12 * S12;
  double iO[4] = {S22 / det, -S12 / det, -S12 / det, S11 / det};
  double M[MAXD * MAXD], parV[MAXD * MAXD], rhs[MAXD], parM[MAXD], L[MAXD * MAXD], z[MAXD];
  const double add = (x->c->compat & ORC_COMPAT_BETA_PRIOR_DIAG) ? 0.0 : 1.0; /* 1/σβ₀² .+ (every element), Q1 */
  for (int br = 0; br < 2; ++br)
    for (int bc = 0; bc < 2; ++bc)
      for (int r = 0; r < p; ++r)
        for (int q = 0; q < p; ++q)
          M[(br * p + r) + d * (bc * p + q)] = iO[br + 2 * bc] * G[r + p * q] + add;
  if (add == 0.0) for (int k = 0; k < d; ++k) M[k + d * k] += 1.0;
  if (spd_inverse(d, M, parV)) return -3;
  /* vec(x'η invΩ'): column c of (x'η invΩ') = sum_k (x'η)[:,k] invΩ[c,k] */
  for (int cidx = 0; cidx < 2; ++cidx)
    for (int r = 0; r < p; ++r) rhs[cidx * p + r] = gT[r] * iO[cidx + 2 * 0] + gZ[r] * iO[cidx + 2 * 1];
  for (int r = 0; r < d; ++r) {
    double acc = 0.0;
    for (int q = 0; q < d; ++q) acc += parV[r + d * q] * rhs[q];
    parM[r] = acc;
  }
  if (chol_lower(d, parV, L)) return -3;
  beta_normals(x, d, z);
  for (int r = 0; r < d; ++r) {
    double acc = parM[r];
    for (int q = 0; q <= r; ++q) acc += L[r + d * q] * z[q];
    s->beta[r] = acc;
  }
  if (!x->c->intercept) { s->beta[0] = 0.0; s->beta[p] = 0.0; } /* GibbsRtIrt.pl.jl:293-295 */
  return 0;
}

/* drawSubjCoefficientsLatent, Draw.pl.jl:399-416 */
static int draw_beta_latent(ctx* x) {
  const int d = x->F + 2;
  orc_state* s = x->st;
  double G[MAXD * MAXD], g[MAXD], M[MAXD * MAXD], parV[MAXD * MAXD], parM[MAXD], L[MAXD * MAXD], z[MAXD];
  gram(x, d, s->zeta, NULL, G, g, NULL);
  const double iO = 1.0 / s->Sigma[3];
  const double add = (x->c->compat & ORC_COMPAT_BETA_PRIOR_DIAG) ? 0.0 : 1.0;
  for (int k = 0; k < d * d; ++k) M[k] = iO * G[k] + add;
  if (add == 0.0) for (int k = 0; k < d; ++k) M[k + d * k] += 1.0;
  if (spd_inverse(d, M, parV)) return -3;
  for (int r = 0; r < d; ++r) {
    double acc = 0.0;
    for (int q = 0; q < d; ++q) acc += parV[r + d * q] * (g[q] * iO);
    parM[r] = acc;
  }
  if (chol_lower(d, parV, L)) return -3;
  beta_normals(x, d, z);
  for (int r = 0; r < d; ++r) {
    double acc = parM[r];
    for (int q = 0; q <= r; ++q) acc += L[r + d * q] * z[q];
    s->beta[r] = acc;
  }
  if (!x->c->intercept) s->beta[0] = 0.0; /* GibbsRtIrtLatent.pl.jl:184-186 */
  return 0;
}

/* getSubjCoefficientsLatentQr, Draw.pl.jl:446-458.  The Kronecker product of an N-vector with x'x
 * gives a tall least-squares system whose solution is (x'x)^{-1} x'(ζ - k1 ν). */
static int get_beta_latentqr(ctx* x) {
  const int d = x->F + 2, N = x->N;
  orc_state* s = x->st;
  double G[MAXD * MAXD], g[MAXD];
  for (int i = 0; i < N; ++i) x->tmpN2[i] = s->zeta[i] - x->k1 * s->nu[i];
  gram(x, d, x->tmpN2, NULL, G, g, NULL);
  if (spd_solve(d, G, g, s->beta)) return -2;
  if (!x->c->intercept) s->beta[0] = 0.0; /* GibbsRtIrtLatent.pl.jl:287-289 */
  return 0;
}

/* drawSubjCorrCross :463-469, drawSubjCorrCrossQr :474-489 (mu_rho = 0, sigma_rho = 1) */
static int draw_rho(ctx* x, int qr) {
  const int N = x->N, J = x->J;
  orc_state* s = x->st;
  int bad = 0;
  OMP_FOR
  for (int j = 0; j < J; ++j) {
    double sV = 0.0, sM = 0.0;
    for (int i = 0; i < N; ++i) {
      double th = s->theta[i], lt = x->logT[i + (size_t)N * j];
      if (qr) {
        double nu = s->nu[i + (size_t)N * j];
        if (!(nu > 0.0)) bad = 1; /* @assert all(ν .> 0), Draw.pl.jl:477 */
        double den = s->sigma2[j] * (x->k2 * nu);
        sV += th * th / den;
        sM += th * (s->lambda[j] - s->zeta[i] - lt + x->k1 * nu) / den;
      } else {
        sV += th * th / s->sigma2[j];
        sM += th * (s->lambda[j] - s->zeta[i] - lt) / s->sigma2[j];
      }
    }
    double parV = 1.0 / (1.0 + sV);
    double parM = parV * (0.0 + sM);
    s->rho[j] = parM + sqrt(parV) * orc_site_normal(x->key, (uint32_t)j, x->sweep, SITE(DOM_ITEM, IK_RHO, 0));
  }
  return bad ? -4 : 0;
}

/* ---------- structural: covariance ---------- */
static void cov2one_2x2(double* S) { /* Draw.pl.jl:507-511 */
  double r = S[2] / sqrt(S[0]) / sqrt(S[3]);
  double r21 = S[1] / sqrt(S[0]) / sqrt(S[3]);
  S[0] = 1.0; S[3] = 1.0; S[2] = r; S[1] = r21;
}

void orc_inv_wishart2_key(orc_key key, uint32_t sweep, double df, const double Psi[4], double out[4]) {
  /* rand(InverseWishart(df, Ψ)) = inv(rand(Wishart(df, inv(Ψ)))), Wishart by Bartlett */
  double det = Psi[0] * Psi[3] - Psi[1] * Psi[2];
  double S[4] = {Psi[3] / det, -Psi[1] / det, -Psi[2] / det, Psi[0] / det};
  double l11 = sqrt(S[0]), l21 = S[1] / l11, l22 = sqrt(S[3] - l21 * l21);
  double a11 = sqrt(2.0 * orc_site_gamma(key, 0, sweep, SITE(DOM_GLOBAL, GK_SIGMAP, 0), 0.5 * df));
  double a22 = sqrt(2.0 * orc_site_gamma(key, 1, sweep, SITE(DOM_GLOBAL, GK_SIGMAP, 0), 0.5 * (df - 1.0)));
  double a21 = orc_site_normal(key, 2, sweep, SITE(DOM_GLOBAL, GK_SIGMAP, 0));
  /* X = L A (lower), W = X X' */
  double x11 = l11 * a11, x21 = l21 * a11 + l22 * a21, x22 = l22 * a22;
  double w11 = x11 * x11, w21 = x21 * x11, w22 = x21 * x21 + x22 * x22;
  double dw = w11 * w22 - w21 * w21;
  out[0] = w22 / dw; out[1] = -w21 / dw; out[2] = -w21 / dw; out[3] = w11 / dw;
}
void orc_inv_wishart2(double df, const double Psi[4], uint64_t seed, uint32_t sweep, double out[4]) {
  orc_inv_wishart2_key(orc_make_key(seed, 0), sweep, df, Psi, out);
}

/* drawSubjCovariance :499-515 (use_x) / drawSubjCovarianceNull :522-535 */
static void draw_Sigma_iw(ctx* x, int use_x) {
  const int N = x->N, F = x->F, p = F + 1;
  orc_state* s = x->st;
  double e11 = 0, e12 = 0, e22 = 0;
  for (int i = 0; i < N; ++i) {
    double m1 = 0, m2 = 0;
    if (use_x) {
      m1 = s->beta[0]; m2 = s->beta[p];
      for (int k = 0; k < F; ++k) {
        m1 += x->X[i + (size_t)N * k] * s->beta[k + 1];
        m2 += x->X[i + (size_t)N * k] * s->beta[p + k + 1];
      }
    }
    double e1 = s->theta[i] - m1, e2 = s->zeta[i] - m2;
    e11 += e1 * e1; e12 += e1 * e2; e22 += e2 * e2;
  }
  double Psi[4] = {e11 + 1.0, e12, e12, e22 + 1.0};
  orc_inv_wishart2_key(x->key, x->sweep, (double)N + 3.0, Psi, s->Sigma);
  if (x->c->cov2one) cov2one_2x2(s->Sigma);
}

/* drawSubjCovarianceCross :542-557, ...Latent :563-579, ...LatentQr :585-606 */
enum { SG_CROSS, SG_LATENT, SG_LATENTQR };
static void draw_Sigma_ig(ctx* x, int kind) {
  const int N = x->N, F = x->F;
  orc_state* s = x->st;
  double parA = 1e-3 + N / 2.0, parB;
  if (kind == SG_CROSS) {
    double ss = 0.0;
    for (int i = 0; i < N; ++i) ss += s->zeta[i] * s->zeta[i];
    parB = 1e-3 + ss / 2.0;
  } else {
    double ss = 0.0, snu = 0.0, snu2 = 0.0, sw = 0.0;
    for (int i = 0; i < N; ++i) {
      double xb = s->beta[0];
      for (int k = 0; k < F; ++k) xb += x->X[i + (size_t)N * k] * s->beta[k + 1];
      xb += s->theta[i] * s->beta[F + 1];
      double r = s->zeta[i] - xb;
      if (kind == SG_LATENTQR) {
        double nu = s->nu[i];
        r -= x->k1 * nu;
        snu += nu; snu2 += nu * nu;
        sw += r * r / (2.0 * x->k2 * nu);
      }
      ss += r * r;
    }
    if (kind == SG_LATENT) parB = 1e-3 + ss / 2.0;
    else {
      parA = 1e-3 + N * 3 / 2.0;
      /* as written: sum(r.^2 / (2*k2e)) with vector/vector = r² (2k2ν)' / ((2k2ν)'(2k2ν)), an N x N matrix (Q2) */
      double quirk = ss * (2.0 * x->k2 * snu) / (4.0 * x->k2 * x->k2 * snu2);
      parB = 1e-3 + ((x->c->compat & ORC_COMPAT_LATENTQR_SCALE_ELEMENTWISE) ? sw : quirk) + snu;
    }
  }
  double sv = parB / orc_site_gamma(x->key, 0, x->sweep, SITE(DOM_GLOBAL, GK_SIGMAP, 0), parA);
  s->Sigma[0] = 1.0; s->Sigma[1] = 0.0; s->Sigma[2] = 0.0; s->Sigma[3] = sv;
  if (x->c->cov2one) cov2one_2x2(s->Sigma);
}

/* ---------- log-likelihoods ---------- */
static inline double log1pexp(double e) { return (e > 0 ? e : 0) + log1p(exp(-fabs(e))); }
#define LOG2PI 1.8378770664093454835606594728112

static double loglik_ctx(const ctx* x) {
  const int N = x->N, J = x->J, F = x->F, p = F + 1, model = x->c->model;
  const orc_state* s = x->st;
  double lb = 0.0, lt = 0.0, ls = 0.0;
  for (int j = 0; j < J; ++j)
    for (int i = 0; i < N; ++i) {
      double eta = s->a[j] * (s->theta[i] - s->b[j]);
      lb += x->Y[i + (size_t)N * j] * eta - log1pexp(eta);
    }
  if (model != ORC_MLIRT) {
    for (int j = 0; j < J; ++j)
      for (int i = 0; i < N; ++i) {
        double mu = s->lambda[j] - s->zeta[i], var = s->sigma2[j];
        if (model == ORC_CROSS || model == ORC_CROSSQR) mu -= s->theta[i] * s->rho[j];
        if (model == ORC_CROSSQR) {
          double nu = s->nu[i + (size_t)N * j];
          mu += x->k1 * nu; var *= x->k2 * nu;
        }
        double r = x->logT[i + (size_t)N * j] - mu;
        lt += -0.5 * (LOG2PI + log(var)) - 0.5 * r * r / var;
      }
  }
  const double S11 = s->Sigma ? s->Sigma[0] : 1, S21 = s->Sigma ? s->Sigma[1] : 0, S12 = s->Sigma ? s->Sigma[2] : 0,
               S22 = s->Sigma ? s->Sigma[3] : 1;
  (void)S21;
  for (int i = 0; i < N; ++i) {
    double xb1 = 0, xb2 = 0;
    switch (model) {
      case ORC_MLIRT: { /* Normal(xβ, 1) on θ, GibbsRtIrt.pl.jl:201 */
        xb1 = s->beta[0];
        for (int k = 0; k < F; ++k) xb1 += x->X[i + (size_t)N * k] * s->beta[k + 1];
        double r = s->theta[i] - xb1;
        ls += -0.5 * LOG2PI - 0.5 * r * r;
      } break;
      case ORC_RTIRT:
        xb1 = s->beta[0]; xb2 = s->beta[p];
        for (int k = 0; k < F; ++k) {
          xb1 += x->X[i + (size_t)N * k] * s->beta[k + 1];
          xb2 += x->X[i + (size_t)N * k] * s->beta[p + k + 1];
        }
        /* fallthrough */
      case ORC_NULL:
      case ORC_CROSS:
      case ORC_CROSSQR: { /* MvNormal(μη_i, Σp) on (θ_i, ζ_i), GibbsRtIrt.pl.jl:269 */
        double e1 = s->theta[i] - xb1, e2 = s->zeta[i] - xb2;
        double det = S11 * S22 - S12 * S12;
        double q = (S22 * e1 * e1 - 2.0 * S12 * e1 * e2 + S11 * e2 * e2) / det;
        ls += -LOG2PI - 0.5 * log(det) - 0.5 * q;
      } break;
      case ORC_LATENT:
      case ORC_LATENTQR: { /* Normal(xβ [+k1ν], sqrt(Σ22 [k2ν])) on ζ, GibbsRtIrtLatent.pl.jl:158,261 */
        double mu = s->beta[0];
        for (int k = 0; k < F; ++k) mu += x->X[i + (size_t)N * k] * s->beta[k + 1];
        mu += s->theta[i] * s->beta[F + 1];
        double var = S22;
        if (model == ORC_LATENTQR) { mu += x->k1 * s->nu[i]; var *= x->k2 * s->nu[i]; }
        double r = s->zeta[i] - mu;
        ls += -0.5 * (LOG2PI + log(var)) - 0.5 * r * r / var;
      } break;
    }
  }
  return lb + lt + ls;
}

/* ---------- scans ---------- */
int orc_beta_len(const orc_cfg* c) {
  switch (c->model) {
    case ORC_MLIRT: return c->nFeat + 1;
    case ORC_RTIRT: case ORC_NULL: return 2 * (c->nFeat + 1);
    case ORC_LATENT: case ORC_LATENTQR: return c->nFeat + 2;
    default: return 0;
  }
}
int orc_qr_width(const orc_cfg* c) {
  switch (c->model) {
    case ORC_MLIRT: return c->nFeat + 1;
    case ORC_RTIRT: case ORC_NULL: return 2 * (c->nFeat + 1) + 4;
    case ORC_CROSS: return c->nItem + 4;
    case ORC_CROSSQR: return c->nItem + 4 + c->nSubj * c->nItem;
    case ORC_LATENT: return c->nFeat + 2 + 4;
    case ORC_LATENTQR: return c->nFeat + 2 + 4 + c->nSubj;
  }
  return -1;
}

static int init_ctx(ctx* x, const orc_cfg* c, const double* Y, const double* logT, const double* X, orc_state* st) {
  memset(x, 0, sizeof(*x));
  x->c = c; x->Y = Y; x->logT = logT; x->X = X; x->st = st;
  x->N = c->nSubj; x->J = c->nItem; x->F = c->nFeat;
  x->nthreads = c->nthreads > 0 ? c->nthreads : 1;
  x->key = orc_make_key(c->seed, c->chain);
  if (2 * (x->F + 1) > MAXD) return -1;
  double q = c->qRt;
  x->k1 = (1.0 - 2.0 * q) / (q * (1.0 - q));
  x->k2 = 2.0 / (q * (1.0 - q));
  if (logT) {
    size_t n = (size_t)x->N * x->J;
    double m = 0.0;
    for (size_t k = 0; k < n; ++k) m += logT[k];
    m /= (double)n;
    double v = 0.0;
    for (size_t k = 0; k < n; ++k) v += (logT[k] - m) * (logT[k] - m);
    x->muLam = m;
    x->sdLam = sqrt(v / (double)(n - 1));
  }
  x->tmpN = (double*)malloc(sizeof(double) * x->N);
  x->tmpN2 = (double*)malloc(sizeof(double) * x->N);
  return (x->tmpN && x->tmpN2) ? 0 : -1;
}
static void free_ctx(ctx* x) { free(x->tmpN); free(x->tmpN2); }

double orc_loglik(const orc_cfg* c, const double* Y, const double* logT, const double* X, const orc_state* st) {
  ctx x;
  if (init_ctx(&x, c, Y, logT, X, (orc_state*)st)) return NAN;
  double v = loglik_ctx(&x);
  free_ctx(&x);
  return v;
}

static int one_sweep(ctx* x) {
  const orc_cfg* c = x->c;
  orc_state* s = x->st;
  int rc = 0;
  switch (c->model) {
    case ORC_MLIRT: /* GibbsRtIrt.pl.jl:224-238 : β, ω, a, b, θ  (a before b, Q8) */
      if ((rc = get_beta_mlirt(x))) return rc;
      draw_omega(x);
      draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_b(x);
      draw_theta(x, 1, 1.0);
      break;
    case ORC_RTIRT: /* GibbsRtIrt.pl.jl:292-313 */
      if ((rc = draw_beta_rtirt(x))) return rc;
      draw_Sigma_iw(x, 1);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 1, s->Sigma[0]);
      draw_lambda(x, RT_PLAIN); draw_sigma2(x, RT_PLAIN); draw_zeta(x, Z_X);
      break;
    case ORC_NULL: /* GibbsRtIrt.pl.jl:380-396 */
      memset(s->beta, 0, sizeof(double) * 2 * (x->F + 1));
      draw_Sigma_iw(x, 0);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 0, s->Sigma[0]);
      draw_lambda(x, RT_PLAIN); draw_sigma2(x, RT_PLAIN); draw_zeta(x, Z_NULL);
      break;
    case ORC_CROSS: /* GibbsRtIrtCross.pl.jl:190-205 */
      if ((rc = draw_rho(x, 0))) return rc;
      draw_Sigma_ig(x, SG_CROSS);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 0, s->Sigma[0]);
      draw_lambda(x, RT_CROSS); draw_sigma2(x, RT_CROSS); draw_zeta(x, Z_CROSS);
      break;
    case ORC_CROSSQR: /* GibbsRtIrtCross.pl.jl:278-294 */
      draw_nu_cell(x);
      if ((rc = draw_rho(x, 1))) return rc;
      draw_Sigma_ig(x, SG_CROSS);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 0, s->Sigma[0]);
      draw_lambda(x, RT_CROSSQR); draw_sigma2(x, RT_CROSSQR); draw_zeta(x, Z_CROSSQR);
      break;
    case ORC_LATENT: /* GibbsRtIrtLatent.pl.jl:182-203 */
      if ((rc = draw_beta_latent(x))) return rc;
      draw_Sigma_ig(x, SG_LATENT);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 0, s->Sigma[0]);
      draw_lambda(x, RT_PLAIN); draw_sigma2(x, RT_PLAIN); draw_zeta(x, Z_LATENT);
      break;
    case ORC_LATENTQR: /* GibbsRtIrtLatent.pl.jl:284-306 */
      draw_nu_person(x);
      if ((rc = get_beta_latentqr(x))) return rc;
      draw_Sigma_ig(x, SG_LATENTQR);
      draw_omega(x); draw_b(x); draw_a(x);
      if (c->onepl) for (int j = 0; j < x->J; ++j) s->a[j] = 1.0;
      draw_theta(x, 0, s->Sigma[0]);
      draw_lambda(x, RT_PLAIN); draw_sigma2(x, RT_PLAIN); draw_zeta(x, Z_LATENTQR);
      break;
    default: return -1;
  }
  return 0;
}

int orc_sample(const orc_cfg* c, const double* Y, const double* logT, const double* X, orc_state* st,
               int64_t first_sweep, int64_t n_sweeps, double* tr_ra, double* tr_rt, double* tr_qr, double* tr_ll,
               int qr_skip_nu) {
  ctx x;
  int rc = init_ctx(&x, c, Y, logT, X, st);
  if (rc) return rc;
  const int N = x.N, J = x.J, W = N + 2 * J, nb = orc_beta_len(c);
  int qw = orc_qr_width(c);
  int nu_len = (c->model == ORC_LATENTQR) ? N : (c->model == ORC_CROSSQR ? N * J : 0);
  if (qr_skip_nu) qw -= nu_len;
  for (int64_t t = 0; t < n_sweeps && rc == 0; ++t) {
    x.sweep = (uint32_t)(first_sweep + t);
    rc = one_sweep(&x);
    if (rc) break;
    if (tr_ra) {
      double* r = tr_ra + (size_t)t * W;
      memcpy(r, st->theta, sizeof(double) * N);
      memcpy(r + N, st->a, sizeof(double) * J);
      memcpy(r + N + J, st->b, sizeof(double) * J);
    }
    if (tr_rt && c->model != ORC_MLIRT) {
      double* r = tr_rt + (size_t)t * W;
      memcpy(r, st->zeta, sizeof(double) * N);
      memcpy(r + N, st->lambda, sizeof(double) * J);
      memcpy(r + N + J, st->sigma2, sizeof(double) * J);
    }
    if (tr_qr) {
      double* r = tr_qr + (size_t)t * qw;
      int o = 0;
      if (c->model == ORC_CROSS || c->model == ORC_CROSSQR) { memcpy(r, st->rho, sizeof(double) * J); o = J; }
      else { memcpy(r, st->beta, sizeof(double) * nb); o = nb; }
      if (c->model != ORC_MLIRT) { memcpy(r + o, st->Sigma, sizeof(double) * 4); o += 4; }
      if (nu_len && !qr_skip_nu) memcpy(r + o, st->nu, sizeof(double) * nu_len);
    }
    if (tr_ll) tr_ll[t] = loglik_ctx(&x);
  }
  free_ctx(&x);
  return rc;
}

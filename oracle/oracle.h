/*
 * oracle.h -- CPU restatement of the Gibbs sweep of ExtendedRtIrtModeling.jl (reference v0.2.6).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product
 * (extendedrtirtmodeling.jl_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference ships no golden vectors or seeded expected outputs
 * (test/test-basic-test.jl:2 is a failing placeholder) and Julia is not installed here, so this
 * restatement cannot be checked against the reference's own outputs.  It is pinned instead by
 *  (1) Random123 known-answer vectors for Philox4x32-10,
 *  (2) closed-form moments / KS tests of every variate generator against scipy.stats,
 *  (3) an independent dense numpy transcription of src/Draw.pl.jl (oracle/draw_np.py).
 *
 * Formulas follow /root/reference/src/Draw.pl.jl line by line (dense, unexpanded sums, float64);
 * scan orders follow the seven sample! methods.  The random variates live in third-party Julia
 * packages that are not vendored (PolyaGammaSamplers 0.1, Distributions 0.21-0.25, Project.toml:14-15,37-38);
 * their published algorithms are restated in rng.c / pg.c driven by a counter-based Philox stream, so
 * bit-level agreement is between this oracle and the CUDA kernels, distributional agreement with Julia.
 */
#ifndef ERIRT_ORACLE_H
#define ERIRT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* model ids (shared numbering with include/erirt_b200.h) */
enum {
  ORC_MLIRT = 0,     /* GibbsMlIrt          src/GibbsRtIrt.pl.jl:210-257      */
  ORC_RTIRT = 1,     /* GibbsRtIrt          src/GibbsRtIrt.pl.jl:278-346      */
  ORC_NULL = 2,      /* GibbsRtIrtNull      src/GibbsRtIrt.pl.jl:367-426      */
  ORC_CROSS = 3,     /* GibbsRtIrtCross     src/GibbsRtIrtCross.pl.jl:176-235 */
  ORC_CROSSQR = 4,   /* GibbsRtIrtCrossQr   src/GibbsRtIrtCross.pl.jl:265-325 */
  ORC_LATENT = 5,    /* GibbsRtIrtLatent    src/GibbsRtIrtLatent.pl.jl:168-233 */
  ORC_LATENTQR = 6   /* GibbsRtIrtLatentQr  src/GibbsRtIrtLatent.pl.jl:271-337 (README's GibbsRtIrtQuantile) */
};

/* compat flags: 0 = behave exactly as the reference source is written */
enum {
  ORC_COMPAT_BETA_PRIOR_DIAG = 1,   /* Q1: add 1/sigma_beta^2 to the diagonal only (Draw.pl.jl:386,410 add it to every element) */
  ORC_COMPAT_LATENTQR_SCALE_ELEMENTWISE = 2 /* Q2: sum_i r_i^2/(2 k2 nu_i) instead of the vec/vec matrix division of Draw.pl.jl:594 */
};

typedef struct {
  int32_t model;
  int32_t nSubj, nItem, nFeat;
  double qRt;           /* Cond.qRt (qRa is never read by the reference) */
  int32_t intercept;    /* sample! kwarg, default false */
  int32_t onepl;        /* itemtype == "1pl" */
  int32_t cov2one;      /* sample! kwarg (default true except Latent/LatentQr) */
  int32_t compat;
  uint64_t seed;
  uint32_t chain;
  int32_t nthreads;     /* OpenMP threads for the N x J loops (1 = serial, the reference's execution model) */
} orc_cfg;

/* Mutable sampler state == InputPara (src/Base.pl.jl:100-115).  All float64.
 * omega: N x J column-major; nu: N (LatentQr) or N x J column-major (CrossQr);
 * beta: p (MlIrt), p x 2 column-major (RtIrt/Null), F+2 (Latent*); Sigma: 2x2 column-major. */
typedef struct {
  double *omega, *theta, *a, *b, *zeta, *lambda, *sigma2, *nu, *beta, *rho, *Sigma;
} orc_state;

/* Sizes of one trace row per model (ra = rt = N + 2J). */
int orc_qr_width(const orc_cfg* c);      /* full width incl. the nu block of the Qr models */
int orc_beta_len(const orc_cfg* c);

/* Run sweeps first_sweep .. first_sweep+n_sweeps-1 (1-based RNG sweep index) on `st` in place.
 * Y, logT: N x J column-major (logT may be NULL for MlIrt); X: N x F column-major.
 * Trace outputs are row-major [sweep][width]; any may be NULL.  tr_qr rows have orc_qr_width()
 * entries unless qr_skip_nu != 0, in which case the trailing nu block is dropped.
 * Returns 0, or a negative error code. */
int orc_sample(const orc_cfg* c, const double* Y, const double* logT, const double* X, orc_state* st,
               int64_t first_sweep, int64_t n_sweeps,
               double* tr_ra, double* tr_rt, double* tr_qr, double* tr_ll, int qr_skip_nu);

/* Log-likelihood of a given state (getLogLikelihood*), used for DIC's D-hat. */
double orc_loglik(const orc_cfg* c, const double* Y, const double* logT, const double* X, const orc_state* st);

/* ---- building blocks exposed for unit tests ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* PG(1,z) for a rows x cols grid of z (row-major), cell (i,j) uses the same counters as the sampler. */
void orc_pg_grid(const double* z, int64_t rows, int32_t cols, int64_t row0, uint64_t seed, uint32_t chain,
                 uint32_t sweep, double* out, int32_t* attempts);
/* 1/IG draws for the person-level nu site: nu_i = clamp(1/IG(mu_i, lam),1e-10,1e10) */
/* gen.c: restatement of erirt_generate_data (the N x J part of setData*, src/SimTools.jl:117-368) */
void orc_generate_data(int64_t n, int32_t J, int64_t person_offset, uint64_t seed, int32_t has_rt, int32_t err, const double* theta,
                       const double* zeta, const double* a, const double* b, const double* lambda, const double* sigma2,
                       const double* rho, double* Y, double* logT);
void orc_nu_person(const double* mu, double lam, int64_t n, int64_t row0, uint64_t seed, uint32_t chain,
                   uint32_t sweep, double* out);
double orc_inv_normal_tail(double y);    /* Phic^{-1}(y) */
/* item/global-site variates (for distribution tests): kind 0 normal, 1 trunc-normal(mu,sd;0,inf),
 * 2 gamma(shape), 3 inverse-gamma(shape, scale) ; unit = index, sweep = draw number */
void orc_variates(int kind, double p1, double p2, int64_t n, uint64_t seed, double* out);
void orc_inv_wishart2(double df, const double Psi[4], uint64_t seed, uint32_t sweep, double out[4]);

#ifdef __cplusplus
}
#endif
#endif

"""Shared test helpers: seeded synthetic data per model and oracle/engine runners."""
import numpy as np

MODELS = ["MlIrt", "RtIrt", "RtIrtNull", "RtIrtLatent", "RtIrtLatentQr", "RtIrtCross", "RtIrtCrossQr"]
ALL_MODELS = MODELS


def make_problem(model, N, J, F, seed=0, q=0.85):
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=N)
    a = rng.uniform(0.7, 1.4, J)
    b = rng.normal(0, 0.5, J)
    X = rng.normal(size=(N, F)) if F > 0 else np.zeros((N, 0))
    Y = (rng.uniform(size=(N, J)) < 1 / (1 + np.exp(-a * (theta[:, None] - b)))).astype(float)
    zeta = 0.3 * theta + rng.normal(0, 0.5, N)
    logT = 3 + rng.normal(0, 0.5, (N, J)) - zeta[:, None]
    nb = {"MlIrt": F + 1, "RtIrt": 2 * (F + 1), "RtIrtNull": 2 * (F + 1), "RtIrtLatent": F + 2, "RtIrtLatentQr": F + 2}.get(model, 0)
    init = dict(theta=rng.normal(size=N), zeta=rng.normal(size=N), a=np.ones(J), b=np.zeros(J), lambda_=np.zeros(J),
                sigma2=np.ones(J), beta=rng.normal(size=max(nb, 1)), rho=rng.normal(size=J), Sigma=np.eye(2).ravel())
    if model == "RtIrtNull":
        init["beta"] = np.zeros(nb)
    return dict(model=model, N=N, J=J, F=F, q=q, Y=Y, logT=logT, X=X, init=init, nb=nb)


def run_oracle(O, pb, n_sweeps, seed=99, chain=0, **opts):
    cfg = O.make_cfg(pb["model"], pb["N"], pb["J"], pb["F"], qRt=pb["q"], seed=seed, chain=chain, **opts)
    logT = None if pb["model"] == "MlIrt" else pb["logT"]
    return O.sample(cfg, pb["Y"], logT, pb["X"], pb["init"], n_sweeps)


def run_engine(E, pb, n_sweeps, seed=99, chain=0, dtype="f64", intercept=False, itemtype="2pl", cov2one=None, compat=0,
               use_graph=False, n_chain=1, person_trace=True, shard=None):
    model = pb["model"]
    if cov2one is None:
        cov2one = model not in ("RtIrtLatent", "RtIrtLatentQr")
    eng = E.Engine(model, pb["N"], pb["J"], pb["F"], n_iter=max(n_sweeps // n_chain, 1), n_chain=n_chain, n_burnin=0,
                   q_rt=pb["q"], intercept=intercept, itemtype=itemtype, cov2one=cov2one, dtype=dtype, seed=seed,
                   chain=chain, compat=compat, person_trace=person_trace, use_graph=use_graph)
    eng.set_data(pb["Y"], None if model == "MlIrt" else pb["logT"], pb["X"] if pb["F"] > 0 else None)
    init = pb["init"]
    st = dict(theta=init["theta"], a=init["a"], b=init["b"])
    if model != "MlIrt":
        st.update(zeta=init["zeta"], lambda_=init["lambda_"], sigma2=init["sigma2"], Sigma=init["Sigma"])
    if pb["nb"]:
        st["beta"] = init["beta"][: pb["nb"]]
    if "Cross" in model:
        st["rho"] = init["rho"]
    eng.set_state(**st)
    eng.sample(n_sweeps)
    return eng


def relerr(x, y, atol=0.0):
    x = np.asarray(x, float).ravel(order="F")
    y = np.asarray(y, float).ravel(order="F")
    return np.abs(x - y) / (np.abs(y) + atol + 1e-300)


def pg_cdf(x, z, n_grid=600_001, terms=12):
    """CDF of PG(1, z) = J*(1, |z|/2) / 4 from the alternating-series density of J* (Devroye 2009), integrated numerically
    (trapezoid, h = 1e-5: error << 1e-4, enough for a KS test on 1e6 draws)."""
    c = abs(z) / 2
    grid = np.linspace(1e-5, 6, n_grid)
    n = np.arange(terms)[:, None]
    xs = grid[None, :]
    with np.errstate(under="ignore"):
        small = np.pi * (n + 0.5) * (2 / (np.pi * xs)) ** 1.5 * np.exp(-2 * (n + 0.5) ** 2 / xs)
        large = np.pi * (n + 0.5) * np.exp(-((n + 0.5) ** 2) * np.pi ** 2 * xs / 2)
        a = np.where(xs <= 0.64, small, large)
        f = np.exp(np.logaddexp(c, -c) - np.log(2.0) - c * c * grid / 2) * ((-1.0) ** n * a).sum(axis=0)
    cdf = np.concatenate([[0], np.cumsum((f[1:] + f[:-1]) / 2 * np.diff(grid))])
    cdf /= max(cdf[-1], 1.0 - 1e-9) if cdf[-1] > 1 else 1.0
    return np.interp(4 * np.asarray(x), grid, cdf)


def ks_uniformity(w, z):
    """KS distance of draws w from PG(1, z) and the p-value (asymptotic Kolmogorov law)."""
    from scipy import stats
    u = np.sort(pg_cdf(w, z))
    n = u.size
    d = max(np.max(np.arange(1, n + 1) / n - u), np.max(u - np.arange(0, n) / n))
    return d, float(stats.kstwobign.sf(d * np.sqrt(n)))


def posterior_agreement(runs, z=4.5):
    """`runs`: dict label -> (draws, params) arrays of INDEPENDENT chains of the same posterior (burn-in removed).
    For every pair of runs and every non-constant parameter:
      |mean_1 - mean_2| <= z sqrt(MCSE_1^2 + MCSE_2^2),   MCSE = SD / sqrt(ESS)            (z ~ 4 by Bonferroni)
      |log(sd_1 / sd_2)| <= z sqrt(1/(2 ESS'_1) + 1/(2 ESS'_2)),   ESS' = min(ESS of x, ESS of (x - mean)^2)
    Returns the worst standardised deviations (mean, sd) and the list of violations."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from erirt_b200.diagnostics import ess_rhat
    stats_ = {}
    for lab, arr in runs.items():
        rows = []
        for c in range(arr.shape[1]):
            x = arr[:, c]
            if np.ptp(x) == 0:
                rows.append(None)
                continue
            e1 = ess_rhat(x)[0]
            e2 = ess_rhat((x - x.mean()) ** 2)[0]
            rows.append((x.mean(), x.std(ddof=1), e1, min(e1, e2)))
        stats_[lab] = rows
    labs = list(runs)
    worst_m = worst_s = 0.0
    bad = []
    for i in range(len(labs)):
        for j in range(i + 1, len(labs)):
            A, B = stats_[labs[i]], stats_[labs[j]]
            for c, (ra, rb) in enumerate(zip(A, B)):
                if ra is None or rb is None:
                    assert ra is None and rb is None and runs[labs[i]][0, c] == runs[labs[j]][0, c], ("constant column differs", c)
                    continue
                zm = abs(ra[0] - rb[0]) / np.sqrt(ra[1] ** 2 / ra[2] + rb[1] ** 2 / rb[2])
                zs = abs(np.log(ra[1] / rb[1])) / np.sqrt(0.5 / ra[3] + 0.5 / rb[3])
                worst_m, worst_s = max(worst_m, zm), max(worst_s, zs)
                if zm > z or zs > z:
                    bad.append((labs[i], labs[j], c, float(zm), float(zs)))
    return worst_m, worst_s, bad

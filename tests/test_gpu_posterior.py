"""Posterior parity (BASELINE.json north_star: "posterior means and SDs on the README simulation must agree with the
reference within Monte Carlo standard error") and the law of the device's PG(1, z) draws.  Run with -m gpu.

Independent chains (different seeds) of the CPU oracle, the GPU sampler in f64 and in f32 must agree, parameter by parameter,
on the posterior mean within z MCSEs and on the posterior SD within z standard errors of log SD (helpers.posterior_agreement;
z = 4.5: about 400 comparisons per test).  The reference's interleaved "chains" are one long chain, so nChain = 1."""
import numpy as np
import pytest

from helpers import ks_uniformity, posterior_agreement
from test_oracle import PG_KS_Z

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]


@pytest.fixture(scope="module")
def E():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import erirt_b200
    erirt_b200._lib.load()
    return erirt_b200


@pytest.mark.parametrize("z", PG_KS_Z)
def test_pg_f32_distribution_ks_1e6(E, z):
    """The f32 fast path (packed attempt 0, constant-bound squeeze, replay / retry queues; pg_fast.cuh) draws from PG(1, z):
    KS on 1e6 draws per z across the method switch |z| = 3.125 and the attempt-0 limit |z| = 16 (src/Draw.pl.jl:36-40)."""
    n = 1_000_000
    w = E.k_pg(np.full((n // 100, 100), z), seed=202, sweep=5, dtype="f32").ravel()
    d, p = ks_uniformity(w, z)
    assert p > 1e-3, (z, d, p)
    m = 0.25 if z == 0 else np.tanh(z / 2) / (2 * z)
    v = 1 / 24 if z == 0 else (np.sinh(z) - z) / (4 * z ** 3 * np.cosh(z / 2) ** 2)
    assert abs(w.mean() - m) < 4.5 * np.sqrt(v / n) + 2e-7


def _engine_chain(E, model, Y, logT, X, N, J, F, ns, dtype, seed, init, q_rt=0.5, cov2one=True):
    eng = E.Engine(model, N, J, F, n_iter=ns, n_chain=1, n_burnin=ns // 2, q_rt=q_rt, cov2one=cov2one, dtype=dtype, seed=seed,
                   person_trace=False, use_graph=True)
    eng.set_data(Y, logT, X if F > 0 else None)
    eng.set_state(**init)
    eng.sample(ns)
    cols = [eng.get_trace("ra", N, 2 * J)[ns // 2:, :, 0]]
    if model != "MlIrt":
        cols.append(eng.get_trace("rt", N, 2 * J)[ns // 2:, :, 0])
    cols.append(eng.get_trace("qr", 0, eng.trace_width("qr") - (N if model == "RtIrtLatentQr" else 0))[ns // 2:, :, 0])
    ll = eng.get_trace("logLike")[:, 0, 0]
    assert np.all(np.isfinite(ll))
    eng.close()
    return np.concatenate(cols, axis=1)


def _oracle_chain(O, model, Y, logT, X, N, J, F, ns, seed, init, q_rt=0.5, cov2one=None):
    import os
    cfg = O.make_cfg(model, N, J, F, qRt=q_rt, seed=seed, cov2one=cov2one, nthreads=min(8, os.cpu_count() or 1))
    full = dict(a=np.ones(J), b=np.zeros(J), lambda_=np.zeros(J), sigma2=np.ones(J), Sigma=np.eye(2).ravel())
    full.update(init)
    r = O.sample(cfg, Y, logT, X, full, ns, person_trace=True, qr_skip_nu=True, want_ll=False)  # small N: the full-width trace is cheap
    cols = [r["ra"][ns // 2:, N:]]
    if model != "MlIrt":
        cols.append(r["rt"][ns // 2:, N:])
    cols.append(r["qr"][ns // 2:])
    return np.concatenate(cols, axis=1)


def test_readme_simulation_posterior_means_and_sds(E, oracle):
    """README.md:61-77 = BASELINE configs[0]: GibbsMlIrt, setCond(nSubj=1000, nItem=15), nIter=5000 (nBurnin = 2500, src/Base.pl.jl:60),
    single chain.  Oracle vs GPU f64 vs GPU f32, three independent chains: a, b (30 columns) and beta (intercept fixed at 0)."""
    N, J, F, ns = 1000, 15, 3, 5000
    Cond = E.setCond(nSubj=N, nItem=J, nFeat=F, nIter=ns, nChain=1)
    tp = E.setTrueParaMlIrt(Cond, rng=1234)
    Data = E.setDataMlIrt(Cond, tp, rng=1234)
    rng = np.random.default_rng(5)
    init = dict(theta=rng.standard_normal(N), beta=rng.standard_normal(F + 1))
    runs = {"oracle": _oracle_chain(oracle, "MlIrt", Data.Y, None, Data.X, N, J, F, ns, 11, init),
            "gpu_f64": _engine_chain(E, "MlIrt", Data.Y, None, Data.X, N, J, F, ns, "f64", 12, init),
            "gpu_f32": _engine_chain(E, "MlIrt", Data.Y, None, Data.X, N, J, F, ns, "f32", 13, init)}
    wm, ws, bad = posterior_agreement(runs)
    print(f"[posterior] README MlIrt 1000x15, 5000 sweeps: worst |dmean|/se {wm:.2f}, worst |dlog sd|/se {ws:.2f}", flush=True)
    assert not bad, bad
    # and the chain recovers the generating parameters (the author's own acceptance criterion, src/SimTools.jl:42-45)
    a_hat = runs["gpu_f32"][:, :J].mean(axis=0)
    b_hat = runs["gpu_f32"][:, J:2 * J].mean(axis=0)
    assert np.sqrt(np.mean((a_hat - tp.a) ** 2)) < 0.15 and np.sqrt(np.mean((b_hat - tp.b) ** 2)) < 0.15


def test_quantile_model_posterior_means_and_sds(E, oracle):
    """BASELINE configs[2]: GibbsRtIrtQuantile (LatentQr), q = 0.85, TIMSS-shaped 631 x 14 with 10 covariates, 4000 sweeps (second
    half used): a, b, lambda, sigma2, beta, Sigma_p22 of the benchmarked sampler in f32 and f64 against the oracle."""
    N, J, F, ns, q = 631, 14, 10, 4000, 0.85
    Cond = E.setCond(nSubj=N, nItem=J, nFeat=F, nIter=ns, nChain=1, qRt=q, qRa=q)
    tp = E.setTrueParaRtIrtLatent(Cond, rng=77)
    Data = E.setDataRtIrtLatent(Cond, tp, type="skew", rng=77)
    rng = np.random.default_rng(6)
    init = dict(theta=rng.standard_normal(N), zeta=rng.standard_normal(N), beta=rng.standard_normal(F + 2))
    runs = {"oracle": _oracle_chain(oracle, "RtIrtLatentQr", Data.Y, Data.logT, Data.X, N, J, F, ns, 21, init, q_rt=q),
            "gpu_f64": _engine_chain(E, "RtIrtLatentQr", Data.Y, Data.logT, Data.X, N, J, F, ns, "f64", 22, init, q_rt=q, cov2one=False),
            "gpu_f32": _engine_chain(E, "RtIrtLatentQr", Data.Y, Data.logT, Data.X, N, J, F, ns, "f32", 23, init, q_rt=q, cov2one=False)}
    wm, ws, bad = posterior_agreement(runs)
    print(f"[posterior] LatentQr q=0.85 631x14 F=10, 4000 sweeps: worst |dmean|/se {wm:.2f}, worst |dlog sd|/se {ws:.2f}", flush=True)
    assert not bad, bad


@pytest.mark.parametrize("model", ["RtIrt", "RtIrtNull", "RtIrtCrossQr"])
def test_other_samplers_posterior_means_and_sds(E, oracle, model):
    """The same criterion for the joint model with latent regression (BASELINE configs[1], reduced to 600 x 8), the measurement-only
    model (configs[3]) and the cell-level quantile sampler: oracle vs GPU f32, 2400 sweeps."""
    from helpers import make_problem
    N, J, F, ns = 600, 8, (2 if model == "RtIrt" else 0), 2400
    pb = make_problem(model, N, J, F, seed=19)
    i = pb["init"]
    init = dict(theta=i["theta"], zeta=i["zeta"])
    if model == "RtIrt":
        init["beta"] = i["beta"][:pb["nb"]]
    if "Cross" in model:
        init["rho"] = i["rho"]
    runs = {"oracle": _oracle_chain(oracle, model, pb["Y"], pb["logT"], pb["X"], N, J, F, ns, 31, init, q_rt=pb["q"]),
            "gpu_f32": _engine_chain(E, model, pb["Y"], pb["logT"], pb["X"], N, J, F, ns, "f32", 32, init, q_rt=pb["q"])}
    wm, ws, bad = posterior_agreement(runs)
    print(f"[posterior] {model} 600x8, 2400 sweeps: worst |dmean|/se {wm:.2f}, worst |dlog sd|/se {ws:.2f}", flush=True)
    assert not bad, bad

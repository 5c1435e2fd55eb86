"""CPU tests of the oracle (test infrastructure): it is pinned by Random123 known-answer vectors, closed-form
moments of every variate generator, and an independent numpy transcription of src/Draw.pl.jl.
(The reference ships no golden vectors -- parity unpinned, see oracle/oracle.h.)"""
import numpy as np
import pytest
from scipy import stats

from helpers import ALL_MODELS

# Random123 kat_vectors, philox4x32 with 10 rounds: (ctr, key, expected)
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers(oracle):
    from oracle import draw_np
    for ctr, key, exp in PHILOX_KAT:
        assert oracle.philox(ctr, key) == exp
        assert list(draw_np.philox4x32_10(ctr, key)) == exp


def test_inverse_normal_tail(oracle):
    for y in [0.10564977366685535, 1e-2, 1e-5, 1e-9, 2.5e-11]:
        assert abs(oracle.inv_normal_tail(y) - stats.norm.isf(y)) < 1e-12 * stats.norm.isf(y)


def _coeff_table(name):
    """(A, B, coefficients) of a polynomial table of csrc/pg_coeffs.h (the generated header the device code includes)."""
    import os, re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "extendedrtirtmodeling.jl_b200", "csrc", "pg_coeffs.h")).read()
    A = float(re.search(rf"#define {name}_A (\S+)", src).group(1))
    B = float(re.search(rf"#define {name}_B (\S+)", src).group(1))
    body = src[src.index(f"#define {name}_COEFFS"):]
    body = body[:body.index("}")]
    return A, B, [float(v) for v in re.findall(r"(-?\d\.\d+(?:e-?\d+)?),", body)]


def test_device_inverse_tail_tables_match_the_oracle(oracle):
    """The committed polynomial tables of the device's Phic^-1 (f32: degree 8, 1e-7; f64: degree 24, rounding level; tools/gen_coeffs.py)
    evaluated as the kernels evaluate them -- Horner in f32, Estrin in f64 (pg.cuh: xq_poly, xqd_poly) -- against the oracle's Newton root over
    the whole argument range y = u P0, u in [2^-33, 1]."""
    P0 = 0.10564977366685535
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.random(400), 2.0 ** -rng.uniform(0, 33, 400), [1.0 - 2.0 ** -33, 2.0 ** -33]])
    zt = np.array([oracle.inv_normal_tail(P0 * v) for v in u])
    A, B, c = _coeff_table("ERIRT_XQD")
    assert len(c) == 25
    r = 1.0 / np.sqrt(-2.0 * np.log(P0 * u))
    t = A * r + B
    q = [c[2 * i] + c[2 * i + 1] * t for i in range(12)] + [c[24] + 0 * t]
    t2, t4, t8, t16 = t * t, t ** 4, t ** 8, t ** 16
    rr = [q[2 * i] + q[2 * i + 1] * t2 for i in range(6)] + [q[12]]
    ss = [rr[0] + rr[1] * t4, rr[2] + rr[3] * t4, rr[4] + rr[5] * t4, rr[6]]
    p = (ss[0] + ss[1] * t8) + (ss[2] + ss[3] * t8) * t16
    assert np.max(np.abs(1.0 / (r * p) - zt) / zt) < 2e-15
    A, B, c = _coeff_table("ERIRT_XQ")
    assert len(c) == 9
    r32 = r.astype(np.float32)
    x = np.float32(A) * r32 + np.float32(B)
    p32 = np.full_like(x, np.float32(c[8]))
    for k in range(7, -1, -1):
        p32 = p32 * x + np.float32(c[k])
    assert np.max(np.abs(1.0 / (r32 * p32) - zt) / zt) < 1e-6


@pytest.mark.parametrize("z", [0.0, 0.3, 1.0, 2.0, 3.0, 3.2, 5.0, 12.0])
def test_pg_moments(oracle, z):
    # E = tanh(z/2)/(2z), Var = (sinh z - z)/(4 z^3 cosh^2(z/2))  
    n = 150_000
    w = oracle.pg_grid(np.full((n // 50, 50), z), seed=11, sweep=2).ravel()
    m = 0.25 if z == 0 else np.tanh(z / 2) / (2 * z)
    v = 1 / 24 if z == 0 else (np.sinh(z) - z) / (4 * z ** 3 * np.cosh(z / 2) ** 2)
    assert abs(w.mean() - m) < 4.5 * np.sqrt(v / n)
    assert abs(w.var() - v) < 0.05 * v
    assert w.min() > 0


def _pg_cdf(x, z, terms=200):
    # J*(1,c) density by its alternating series, integrated numerically -> CDF of PG(1,z) = J*/4
    c = abs(z) / 2
    grid = np.linspace(1e-4, 6, 60001)
    n = np.arange(terms)[:, None]
    xs = grid[None, :]
    small = np.pi * (n + 0.5) * (2 / (np.pi * xs)) ** 1.5 * np.exp(-2 * (n + 0.5) ** 2 / xs)
    large = np.pi * (n + 0.5) * np.exp(-((n + 0.5) ** 2) * np.pi ** 2 * xs / 2)
    a = np.where(xs <= 0.64, small, large)
    f = np.cosh(c) * np.exp(-c * c * grid / 2) * ((-1.0) ** n * a).sum(axis=0)
    cdf = np.concatenate([[0], np.cumsum((f[1:] + f[:-1]) / 2 * np.diff(grid))])
    return np.interp(4 * np.asarray(x), grid, cdf)


@pytest.mark.parametrize("z", [0.0, 1.5, 3.5])
def test_pg_distribution_ks(oracle, z):
    w = oracle.pg_grid(np.full((400, 50), z), seed=3, sweep=1).ravel()
    d, p = stats.kstest(w, lambda x: _pg_cdf(x, z))
    assert p > 1e-3, (d, p)


PG_KS_Z = [0.0, 1.5, 3.12, 3.13, 6.0, 12.0, 16.0, 16.5, 20.0]  # both methods, the switch |z| = 3.125, the attempt-0 limit |z| = 16


@pytest.mark.parametrize("z", PG_KS_Z)
def test_pg_distribution_ks_1e6(oracle, z):
    """The mixed-envelope sampler (attempt 0 always Method A for |z| <= 16, retries by regime; oracle/pg.c) is this repository's own
    arrangement of the Polson-Scott-Windle sampler, so its law is checked on 1e6 draws per z: KS against PG(1, z) and the mean."""
    from helpers import ks_uniformity
    n = 1_000_000
    w = oracle.pg_grid(np.full((n // 100, 100), z), seed=101, sweep=3).ravel()
    d, p = ks_uniformity(w, z)
    assert p > 1e-3, (z, d, p)
    m = 0.25 if z == 0 else np.tanh(z / 2) / (2 * z)
    v = 1 / 24 if z == 0 else (np.sinh(z) - z) / (4 * z ** 3 * np.cosh(z / 2) ** 2)
    assert abs(w.mean() - m) < 4.5 * np.sqrt(v / n)


def test_scalar_variates(oracle):
    n = 200_000
    x = oracle.variates("normal", 1.5, 2.0, n)
    assert abs(x.mean() - 1.5) < 0.02 and abs(x.std() - 2.0) < 0.02
    assert stats.kstest(x, stats.norm(1.5, 2.0).cdf).pvalue > 1e-3
    for mu, sd in [(1.0, 0.2), (-0.5, 1.0), (-3.0, 1.0)]:
        t = oracle.variates("tnorm", mu, sd, n)
        a = (0 - mu) / sd
        assert t.min() > 0
        assert stats.kstest(t, stats.truncnorm(a, np.inf, loc=mu, scale=sd).cdf).pvalue > 1e-3
    for shape in [1.0, 2.5, 500.001]:
        g = oracle.variates("gamma", shape, 0.0, n)
        assert stats.kstest(g, stats.gamma(shape).cdf).pvalue > 1e-3
    ig = oracle.variates("invgamma", 501.0, 300.0, n)
    assert abs(ig.mean() - 300.0 / 500.0) < 1e-3  # mean = scale/(shape-1)


def test_inverse_gaussian_nu(oracle):
    n = 200_000
    mu, lam = 3.0, 7.0
    nu = oracle.nu_person(np.full(n, mu), lam, seed=4, sweep=1)
    x = 1.0 / nu
    assert abs(x.mean() - mu) < 4 * np.sqrt(mu ** 3 / lam / n)
    assert stats.kstest(x, stats.invgauss(mu / lam, scale=lam).cdf).pvalue > 1e-3


def test_inverse_wishart_mean(oracle):
    Psi = np.array([[3.0, 0.8], [0.8, 2.0]])
    df = 12.0
    acc = np.zeros((2, 2))
    n = 20000
    for s in range(1, n + 1):
        acc += oracle.inv_wishart2(df, Psi, 77, s)
    assert np.allclose(acc / n, Psi / (df - 3), rtol=0.03)  # E = Psi/(df - p - 1), p = 2


@pytest.mark.parametrize("model", ALL_MODELS)
@pytest.mark.parametrize("alt", [False, True])
def test_c_oracle_matches_numpy_transcription(oracle, model, alt):
    """Two independently written restatements of Draw.pl.jl (C loops vs dense numpy broadcasting) agree."""
    from oracle import draw_np as D
    rng = np.random.default_rng(7)
    N, J, F, q = 30, 6, 2, 0.85
    opts = dict(intercept=True, itemtype="1pl", cov2one=False) if alt else {}
    theta = rng.normal(size=N)
    a, b = rng.uniform(0.7, 1.4, J), rng.normal(0, 0.5, J)
    X = rng.normal(size=(N, F))
    Y = (rng.uniform(size=(N, J)) < 1 / (1 + np.exp(-a * (theta[:, None] - b)))).astype(float)
    logT = 3 + rng.normal(0, 0.5, (N, J))
    cfg = oracle.make_cfg(model, N, J, F, qRt=q, seed=99, **opts)
    nb = oracle.lib().orc_beta_len(cfg)
    init = dict(theta=rng.normal(size=N), zeta=rng.normal(size=N), a=np.ones(J), b=np.zeros(J), lambda_=np.zeros(J),
                sigma2=np.ones(J), beta=rng.normal(size=max(nb, 1)), rho=rng.normal(size=J), Sigma=np.eye(2).ravel())
    nsw = 2
    res = oracle.sample(cfg, Y, logT if model != "MlIrt" else None, X, init, nsw)
    st = D.Stream(99, 0)
    C = dict(N=N, J=J, F=F, qRt=q)
    Dd = dict(Y=Y, κ=Y - 0.5, logT=logT, X=X)
    P = {"θ": init["theta"].copy(), "ζ": init["zeta"].copy(), "a": init["a"].copy(), "b": init["b"].copy(),
         "λ": init["lambda_"].copy(), "σ²t": init["sigma2"].copy(), "ρ": init["rho"].copy(), "Σp": np.eye(2), "ν": np.ones(N)}
    P["β"] = init["beta"][:2 * (F + 1)].reshape(F + 1, 2, order="F").copy() if model == "RtIrt" else init["beta"][:max(nb, 1)].copy()
    cov2one = opts.get("cov2one", model not in ("RtIrtLatent", "RtIrtLatentQr"))
    lls = []
    for s in range(1, nsw + 1):
        st.sweep = s
        P = D.sweep(model, st, C, Dd, P, intercept=opts.get("intercept", False), onepl=alt, cov2one=cov2one)
        lls.append(D.loglik(model, C, Dd, P))

    def close(x, y, tol=1e-9):
        x, y = np.asarray(x, float).ravel(order="F"), np.asarray(y, float).ravel(order="F")
        assert np.allclose(x, y, rtol=tol, atol=tol), np.abs(x - y).max()

    close(res["theta"], P["θ"]); close(res["a"], P["a"]); close(res["b"], P["b"]); close(res["omega"], P["ω"])
    if model != "MlIrt":
        close(res["zeta"], P["ζ"]); close(res["lambda"], P["λ"]); close(res["sigma2"], P["σ²t"]); close(res["Sigma"], P["Σp"])
    if "Cross" in model:
        close(res["rho"], P["ρ"])
    elif model != "RtIrtNull":
        close(res["beta"][: np.asarray(P["β"]).size], P["β"])
    if model.endswith("Qr"):
        close(res["nu"], P["ν"])
    close(res["ll"], lls, 1e-10)


def test_oracle_recovers_item_parameters(oracle):
    """Author's own acceptance criterion (README flow): parameter recovery on a simulated data set."""
    from helpers import make_problem, run_oracle
    pb = make_problem("RtIrtNull", 1500, 8, 0, seed=5)
    res = run_oracle(oracle, pb, 300)
    ra = res["ra"][150:]
    a_hat = ra[:, pb["N"]:pb["N"] + pb["J"]].mean(0)
    b_hat = ra[:, pb["N"] + pb["J"]:].mean(0)
    rng = np.random.default_rng(5)
    rng.normal(size=pb["N"])
    a_true = rng.uniform(0.7, 1.4, pb["J"])
    b_true = rng.normal(0, 0.5, pb["J"])
    assert np.sqrt(np.mean((a_hat - a_true) ** 2)) < 0.15
    assert np.sqrt(np.mean((b_hat - b_true) ** 2)) < 0.15


def test_oracle_reproduces_committed_fixture(oracle):
    """tests/golden/oracle_v2.npz (made by tests/golden/make_golden.py) pins the oracle's own stream layout and formulas; it is
    not reference-pinned (the Julia package cannot run here and ships no vectors)."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = np.load(mg.GOLDEN)
    got = mg.compute()
    assert set(got.keys()) == set(want.files)
    for k in want.files:
        w, g = np.asarray(want[k], dtype=np.float64), np.asarray(got[k], dtype=np.float64)
        assert w.shape == g.shape, k
        assert np.allclose(g, w, rtol=1e-11, atol=1e-12), (k, float(np.max(np.abs(g - w))))


@pytest.mark.parametrize("error,mean,var", [("tnorm", 0.0, 0.3), ("unit", 0.0, 1.0), ("norm", 0.0, 0.09), ("tail", 0.0, 5.0 / 3.0), ("skew", -0.5, 0.5)])
def test_generator_error_laws(oracle, error, mean, var):
    """oracle/gen.c (restatement of erirt_generate_data): Bernoulli-logit responses and the five error laws of the reference's
    simulators (src/SimTools.jl:117-368): N(0, sigma2_j) truncated, N(0,1), N(0,0.3), t(5), Gamma(0.5,1)-1."""
    rng = np.random.default_rng(1)
    N, J = 100_000, 6
    th, ze = rng.normal(size=N), 0.3 * rng.normal(size=N)
    a, b, lam = rng.uniform(0.7, 1.4, J), rng.normal(0, 0.5, J), np.full(J, 3.0)
    rho = rng.normal(0, 0.2, J) if error in ("norm", "tail", "skew") else None
    Y, T = oracle.generate_data(N, J, th, a, b, ze, lam, np.full(J, 0.3), rho, error=error, seed=11)
    p = 1 / (1 + np.exp(-a * (th[:, None] - b)))
    assert set(np.unique(Y)) == {0.0, 1.0} and abs(Y.mean() - p.mean()) < 4 * np.sqrt(0.25 / Y.size)
    e = (T - (lam[None, :] - ze[:, None] - (np.outer(th, rho) if rho is not None else 0.0))).ravel()
    assert abs(e.mean() - mean) < 5 * np.sqrt(var / e.size)
    assert abs(e.var() - var) < (0.15 if error == "tail" else 0.03) * var + 1e-3
    if error == "skew":
        assert e.min() >= -1.0 and abs(np.mean(((e - e.mean()) / e.std()) ** 3) - np.sqrt(8.0)) < 0.15
    if error == "tnorm":
        assert T.min() > 0.0
    # counters are global person ids: a shard generates exactly its rows of the one data set
    Y2, T2 = oracle.generate_data(1000, J, th[5000:6000], a, b, ze[5000:6000], lam, np.full(J, 0.3), rho, error=error, seed=11, person_offset=5000)
    assert np.array_equal(Y2, Y[5000:6000]) and np.array_equal(T2, T[5000:6000])
    # MlIrt: responses only
    Y3, T3 = oracle.generate_data(1000, J, th[:1000], a, b, seed=11)
    assert T3 is None and np.array_equal(Y3, Y[:1000])

"""Oracle parity of the BENCHMARKED configuration at benchmark size (run with -m gpu on the B200 box).

bench.py times GibbsRtIrtQuantile (= the LatentQr sampler, src/GibbsRtIrtLatent.pl.jl:271-337) at nSubj = 1M x nItem = 100.
At that size a CTA of the f32 person kernel owns >= 16 tiles, so the cross-tile register accumulators of the item
statistics and their staged fold into the f64 accumulators run -- code the small parity tests never reach.  These tests
compare one and three sweeps at 1M x 100, 500k x 100 and 1M x 13 (other tiles-per-CTA counts and TPP) in f32 AND f64
with the CPU oracle (OpenMP over persons; formulas Draw.pl.jl:36-62, :215-262, :325-343) from the initial state of
bench.py, and force the rarely taken branches of the f32 sampler (work-queue overflow, rows with |z| > 16).

Tolerances (BASELINE.json north_star): 1e-5 relative in f32, 1e-9 in f64 for item / structural columns and logLike;
person columns (theta, zeta, nu) within 1e-5 for >= 99.5 % of the persons in f32 (a PG accept/reject decision taken
the other way by f32 rounding moves that person's theta; the oracle decides in f64).  "Relative" is |x - y| / (|y| + 0.1):
b, beta and the person parameters are locations of O(1) spread that cross zero, so a purely relative error is undefined for
them (measured at 1M x 100 in f32: absolute error of a, b <= 1e-6 = 4e-4 posterior SDs, a handful of flipped PG cells per item)."""
import os

import numpy as np
import pytest

from helpers import relerr

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

SEED = 1234
Q_RT = 0.85
F = 3


@pytest.fixture(scope="module")
def E():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import erirt_b200
    erirt_b200._lib.load()
    return erirt_b200


def bench_like_problem(N, J, a0=None):
    """Data of setDataRtIrtLatent(type="skew") (src/SimTools.jl:300-343) and the initial state of bench.py:init_state."""
    rng = np.random.default_rng(SEED + 100 + J)
    a = rng.uniform(0.7, 1.4, J)
    b = rng.normal(0, 0.5, J)
    lam = rng.normal(4.0, 0.2, J)
    beta = rng.normal(0, 1, F + 1)
    theta = rng.standard_normal(N)
    X = rng.standard_normal((N, F))
    zeta = X @ beta[:F] + beta[F] * theta + (0.5 * rng.standard_normal(N) ** 2 - 1.0)
    Y = np.empty((N, J), dtype=np.float64, order="F")
    logT = np.empty((N, J), dtype=np.float64, order="F")
    for j in range(J):  # column by column: no N x J temporaries beyond the two outputs
        Y[:, j] = rng.random(N) < 1.0 / (1.0 + np.exp(-a[j] * (theta - b[j])))
        logT[:, j] = lam[j] - zeta + rng.standard_normal(N)
    r2 = np.random.default_rng(SEED + 7)
    init = dict(theta=r2.standard_normal(N), zeta=r2.standard_normal(N), beta=r2.standard_normal(F + 2),
                a=np.ones(J) if a0 is None else np.asarray(a0, float), b=np.zeros(J), lambda_=np.zeros(J), sigma2=np.ones(J),
                Sigma=np.eye(2).ravel())
    return dict(model="RtIrtLatentQr", N=N, J=J, F=F, q=Q_RT, Y=Y, logT=logT, X=X, init=init, nb=F + 2)


def oracle_run(O, pb, ns):
    cfg = O.make_cfg("RtIrtLatentQr", pb["N"], pb["J"], F, qRt=Q_RT, seed=SEED, nthreads=os.cpu_count() or 1)
    return O.sample(cfg, pb["Y"], pb["logT"], pb["X"], pb["init"], ns, person_trace=True)


def engine_run(E, pb, ns, dtype):
    eng = E.Engine("RtIrtQuantile", pb["N"], pb["J"], F, n_iter=max(ns, 1), n_chain=1, n_burnin=0, q_rt=Q_RT, cov2one=False, dtype=dtype,
                   seed=SEED, person_trace=True, use_graph=True)
    eng.set_data(pb["Y"], pb["logT"], pb["X"])
    i = pb["init"]
    eng.set_state(theta=i["theta"], zeta=i["zeta"], beta=i["beta"], a=i["a"], b=i["b"], lambda_=i["lambda_"], sigma2=i["sigma2"], Sigma=i["Sigma"])
    eng.sample(ns)
    return eng


def compare(eng, ref, pb, ns, tol_item, tol_person, frac_ok, label):
    """Item / structural columns and logLike: max relative error per sweep below tol_item[sweep]; person columns: share within tol_person."""
    N, J = pb["N"], pb["J"]
    report = {}
    for name in ("ra", "rt"):
        got = eng.get_trace(name, N, 2 * J)[:ns, :, 0]
        e = relerr(got, ref[name][:ns, N:], atol=0.1).reshape(got.shape, order="F")
        report[name] = e.max(axis=1)
        report[name + "_abs"] = np.abs(got - ref[name][:ns, N:]).max(axis=1)
    qw = F + 2 + 4
    got = eng.get_trace("qr", 0, qw)[:ns, :, 0]
    report["qr"] = relerr(got, ref["qr"][:ns, :qw], atol=0.1).reshape(got.shape, order="F").max(axis=1)
    ll = eng.get_trace("logLike")[:ns, 0, 0]
    report["ll"] = relerr(ll, ref["ll"][:ns])
    person = {}
    for name, col0, width in (("ra", 0, N), ("rt", 0, N), ("qr", qw, N)):
        got = eng.get_trace(name, col0, width)[:ns, :, 0]
        e = relerr(got, ref[name][:ns, col0:col0 + width], atol=1e-1).reshape(got.shape, order="F")
        person[name] = (e < tol_person).mean(axis=1)
    print(f"[fullsize] {label}: item max relerr per sweep " + ", ".join(f"{k} {np.array2string(v, precision=2)}" for k, v in report.items())
          + " | persons within tol " + ", ".join(f"{k} {np.array2string(v, precision=5)}" for k, v in person.items()), flush=True)
    for k, v in report.items():
        if k.endswith("_abs"):
            continue
        for s in range(ns):
            assert v[s] < tol_item[min(s, len(tol_item) - 1)], (label, k, "sweep", s + 1, v)
    for k, v in person.items():
        assert v.min() >= frac_ok, (label, k, v)


# (N, J, sweeps): 1M x 100 = 15 625 tiles of 64 persons on 444 CTAs (35 tiles per CTA: two 16-tile folds + the final one);
# 500k x 100 = 17-18 tiles per CTA (one fold + remainder); 1M x 13: TPP = 1, 128-person tiles, 13-14 tiles per CTA (final fold only)
SIZES = [(1_000_000, 100, 3), (500_000, 100, 3), (1_000_000, 13, 3)]


@pytest.mark.parametrize("N,J,ns", SIZES)
def test_benchmark_size_parity_f32_and_f64(E, oracle, N, J, ns):
    pb = bench_like_problem(N, J)
    ref = oracle_run(oracle, pb, ns)
    # f64: the generic kernel; 1e-9 over three sweeps (1e-12 per conditional, summation order over 1e6 persons differs)
    eng = engine_run(E, pb, ns, "f64")
    compare(eng, ref, pb, ns, [1e-9], 1e-9, 1.0, f"f64 {N}x{J}")
    eng.close()
    # f32: person_sweep_fast_kernel, the benchmarked kernel.  Sweep 1 is a pure kernel-vs-formula comparison (1e-5); in sweeps 2-3
    # the f32 chain carries its own rounding and the handful of PG decisions f32 takes the other way forward, so the item columns
    # are held to 1e-4 there (they are sums over 1e6 persons; a single flipped cell moves them by ~1e-6 relative)
    eng = engine_run(E, pb, ns, "f32")
    compare(eng, ref, pb, ns, [1e-5, 1e-4, 1e-4], 1e-5, 0.995, f"f32 {N}x{J}")
    st = eng.stats()
    assert 0.0 < st["pg_deferred_frac"] < 0.2
    eng.close()


@pytest.mark.parametrize("a_hi,a_rest,note", [(7.0, 1.0, "rows with |z| > 16 (the 'wide row' path: no attempt 0 for the row)"),
                                              (7.0, 4.0, "work-queue overflow (both ends) + wide rows: most cells are in the Method-B regime")])
def test_forced_queue_overflow_and_wide_rows_f32(E, oracle, a_hi, a_rest, note):
    """person_fast.cuh: `wide` rows (some |z| = |a (theta - b)| > 16) and the QCAP overflow branch of the push are not reached with
    sane parameters.  Start the chain from a = 7 for three items (|theta| > 2.3 gives |z| > 16, ~2 % of the rows) and optionally a = 4
    for the rest (|z| > 3.125 for most cells: Method-B regime, attempt 0 rarely accepted, more deferred cells than a warp's queues hold), one sweep
    so that the state every draw conditions on is exactly the oracle's."""
    N, J = 20_000, 100
    a0 = np.full(J, a_rest)
    a0[[3, 50, 97]] = a_hi
    pb = bench_like_problem(N, J, a0=a0)
    ref = oracle_run(oracle, pb, 1)
    # omega_1 as drawn by the prologue from the initial state, cell by cell (the oracle's state after sweep 1 still holds omega_1:
    # omega is the first draw of a sweep, Draw.pl.jl:36-40); a cell whose accept/reject decision f32 takes the other way differs
    eng0 = engine_run(E, pb, 0, "f32")
    om = eng0.get_state("omega")
    assert np.all(np.isfinite(om)) and np.all(om > 0)
    e = relerr(om, ref["omega"]).reshape(om.shape, order="F")
    wide = np.abs(pb["init"]["a"][None, :] * (pb["init"]["theta"][:, None] - pb["init"]["b"][None, :])).max(axis=1) > 16.0
    assert wide.sum() > 50, "no wide rows in this problem"
    assert (e < 2e-5).mean() > 0.998, ("omega_1", (e < 2e-5).mean())
    assert (e[wide] < 2e-5).mean() > 0.995, ("omega_1 of the wide rows", (e[wide] < 2e-5).mean())
    st0 = eng0.stats()
    eng0.close()
    eng = engine_run(E, pb, 1, "f32")
    compare(eng, ref, pb, 1, [1e-5], 1e-4, 0.99, f"f32 forced a_hi={a_hi} a_rest={a_rest}")
    st = eng.stats()
    print(f"[fullsize] forced branches ({note}): deferred fraction {st['pg_deferred_frac']:.3f} (prologue {st0['pg_deferred_frac']:.3f})", flush=True)
    if a_rest > 3:  # the prologue drew omega_1 with a = 4: a warp's queues hold 192 + 64 of its 1600 cells
        assert st0["pg_deferred_frac"] > (192 + 64) / 1600.0, "the queues did not overflow"
    eng.close()

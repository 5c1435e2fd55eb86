"""CPU tests of the host-side mirror of the reference API (structs, constructors, generators, diagnostics)."""
import numpy as np
import pytest

import erirt_b200 as E
from erirt_b200 import diagnostics, distributed


def test_setcond_forces_burnin_to_half():
    # src/Base.pl.jl:59-61: nBurnin = round(Int, nIter/2) regardless of the keyword
    c = E.setCond(nSubj=100, nItem=5, nIter=3000, nBurnin=10)
    assert c.nBurnin == 1500
    assert E.setCond(nIter=5).nBurnin == 2  # round half to even, like Julia
    d = E.setCond()
    assert (d.nSubj, d.nItem, d.nFeat, d.nIter, d.nChain, d.nThin, d.nRep, d.qRa, d.qRt) == (2000, 15, 3, 5000, 4, 1, 10, 0.5, 0.5)


def test_input_data_derives_kappa_and_logt():
    Y = np.array([[1, 0], [0, 1.0]])
    T = np.exp(np.array([[1.0, 2.0], [3.0, 4.0]]))
    d = E.InputData(Y=Y, T=T)
    assert np.allclose(d.kappa, Y - 0.5) and np.allclose(d.logT, [[1, 2], [3, 4]])


def test_input_para_greek_aliases():
    p = E.InputPara(theta=[1, 2], Sigma_p=np.eye(2))
    assert np.all(p.θ == [1, 2]) and np.all(p.Σp == np.eye(2))
    p.λ = [3.0]
    assert p.lambda_[0] == 3.0
    with pytest.raises(AttributeError):
        p.nope = 1


@pytest.mark.parametrize("cls,beta_shape", [(E.GibbsMlIrt, (4,)), (E.GibbsRtIrt, (4, 2)), (E.GibbsRtIrtNull, (0,)),
                                            (E.GibbsRtIrtLatent, (5,)), (E.GibbsRtIrtQuantile, (5,)), (E.GibbsRtIrtCross, (0,))])
def test_constructors_follow_set_initial_values(cls, beta_shape):
    c = E.setCond(nSubj=20, nItem=4, nFeat=3)
    m = cls(c, rng=1)
    assert m.Para.theta.shape == (20,) and np.all(m.Para.a == 1) and np.all(m.Para.b == 0)
    assert m.Para.beta.shape == beta_shape
    if cls is not E.GibbsMlIrt:
        assert np.all(m.Para.lambda_ == 0) and np.all(m.Para.sigma2t == 1) and np.all(m.Para.Sigma_p == np.eye(2))
    if cls is E.GibbsRtIrtCross:
        assert m.Para.rho.shape == (4,)


def test_sample_rejects_bad_itemtype():
    c = E.setCond(nSubj=20, nItem=4)
    m = E.GibbsMlIrt(c, rng=1)
    with pytest.raises(ValueError, match="Invalid input: the item type must be '1pl' or '2pl'."):
        E.sample(m, itemtype="3pl")


def test_generators_shapes_and_ranges():
    c = E.setCond(nSubj=200, nItem=6, nFeat=3)
    tp = E.setTrueParaRtIrt(c, rng=0)
    d = E.setDataRtIrt(c, tp, rng=0)
    assert d.Y.shape == (200, 6) and set(np.unique(d.Y)) <= {0.0, 1.0} and d.X.shape == (200, 3)
    assert np.all(d.logT > 0) and tp.theta.shape == (200,) and np.all(tp.a > 0)
    tl = E.setTrueParaRtIrtLatent(c, rng=0)
    dl = E.setDataRtIrtLatent(c, tl, type="skew", rng=0)
    assert tl.beta.shape == (4,) and dl.logT.shape == (200, 6) and abs(tl.beta[-1]) < 1
    tm = E.setTrueParaMlIrt(c, rng=0)
    dm = E.setDataMlIrt(c, tm, rng=0)
    assert dm.T is None and set(np.unique(dm.X[:, 0])) <= {0.0, 1.0}
    tc = E.setTrueParaRtIrtCross(c, rng=0)
    dc = E.setDataRtIrtCross(c, tc, type="tail", rng=0)
    assert dc.X is None and tc.rho.shape == (6,)
    assert E.getRmse([1, 2], [1, 4]) == pytest.approx(np.sqrt(2)) and E.getBias([1, 2], [1, 4]) == -1


def test_ess_rhat_on_known_chains():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2000, 4))
    e, r = diagnostics.ess_rhat(x)
    assert 6000 < e < 10000 and abs(r - 1) < 0.01
    ar = np.zeros((4000, 2))
    eps = rng.normal(size=ar.shape)
    for t in range(1, 4000):
        ar[t] = 0.9 * ar[t - 1] + eps[t]
    e, r = diagnostics.ess_rhat(ar)
    assert 200 < e < 900  # theory: n (1-rho)/(1+rho) = 8000/19 = 421
    shifted = x.copy()
    shifted[:, 0] += 3
    assert diagnostics.ess_rhat(shifted)[1] > 1.2
    assert np.isnan(diagnostics.ess_rhat(np.ones((100, 2)))[0])  # constant columns are skipped, SimTools.jl:430-433


def test_shard_bounds_partition_persons():
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (631, 2)]:
        parts = [distributed.shard_bounds(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (o0, c0), (o1, _) in zip(parts, parts[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert distributed.chain_assignment(8, 4, 1) == [1, 5]


def test_batched_ess_matches_scalar_estimator():
    rng = np.random.default_rng(3)
    n, P, m = 400, 6, 3
    x = rng.standard_normal((n, P, m))
    for p in range(1, P):
        for t in range(1, n):
            x[t, p] = 0.15 * p * x[t - 1, p] + x[t, p]
    x[:, 2, 1] += 2.0            # a chain that sits elsewhere: R-hat > 1.1
    x[:, 4] = np.round(x[:, 4], 1)  # ties
    x[:, 5] = 1.0                # constant column -> NaN, skipped by checkConvergence (SimTools.jl:430-433)
    e, r = diagnostics.ess_rhat_batched(x)
    for p in range(P):
        e0, r0 = diagnostics.ess_rhat(x[:, p, :])
        assert (np.isnan(e0) and np.isnan(e[p])) or abs(e0 - e[p]) < 1e-9 * e0
        assert (np.isnan(r0) and np.isnan(r[p])) or abs(r0 - r[p]) < 1e-12


def test_simtools_metrics_and_convergence_table():
    import erirt_b200 as E
    true = np.array([1.0, 1.2, 0.8, 1.1])
    rng = np.random.default_rng(1)
    obj = {"True": {"a": true}}
    for r in range(1, 6):
        obj[r] = {"a": true + 0.01 * rng.standard_normal(4) + 0.02}
    m = E.getMetrics(obj, par="a")
    assert abs(m["Bias"] - 0.02) < 0.01 and m["Rmse"] < 0.04 and m["Corr"] > 0.99
    m2 = E.getMetrics2(obj, par="a")
    assert abs(m2["relativeBias"] - 0.02) < 0.015

    class Fake:
        pass
    M = Fake()
    M.Cond = E.setCond(nSubj=5, nItem=3, nIter=1000, nChain=2)
    M.Post = E.OutputPost()
    M.Post.ra = rng.standard_normal((1000, 11, 2))
    M.Post.ra[:, :5] = np.nan            # person columns absent (person_trace=False)
    M.Post.rt = rng.standard_normal((1000, 11, 2))
    M.Post.qr = np.ones((1000, 4, 2))   # constant columns are not counted
    conv = E.checkConvergence(M)
    assert conv["essN"].endswith("/ 17") and conv["ess"] == 100.0 and conv["rhat"] == 100.0


def test_csv_ingest_timss_shaped(tmp_path):
    """A TIMSS-2019-shaped table (item scores, *_S response times in seconds, covariates, ids, a missing cell) -> InputData."""
    import erirt_b200 as E
    rng = np.random.default_rng(2)
    n, J = 40, 5
    items = [f"ME62{100 + j}" for j in range(J)]
    hdr = ["IDSTUD"] + items + [i + "_S" for i in items] + ["BSBG01", "BSBGHER_Z"]
    Y = (rng.random((n, J)) < 0.7).astype(int)
    T = np.exp(rng.normal(3.7, 0.6, (n, J)))
    X = np.column_stack([rng.integers(1, 3, n), rng.standard_normal(n)])
    lines = [",".join(hdr)]
    for i in range(n):
        lines.append(",".join([str(1000 + i)] + [str(v) for v in Y[i]] + [f"{v:.3f}" for v in T[i]] + [str(X[i, 0]), f"{X[i, 1]:.6f}"]))
    bad = lines[7].split(",")
    bad[8] = ""  # a missing response time
    lines[7] = ",".join(bad)
    p = tmp_path / "toy.csv"
    p.write_text("\n".join(lines) + "\n")
    D = E.readCsvData(str(p), y_cols=items, t_cols=[i + "_S" for i in items], x_cols=["BSBG01", "BSBGHER_Z"])
    assert D.Y.shape == (n - 1, J) and D.T.shape == (n - 1, J) and D.X.shape == (n - 1, 2) and D.dropped_rows == [8]
    keep = [i for i in range(n) if i != 6]
    assert np.array_equal(D.Y, Y[keep]) and np.allclose(D.logT, np.log(np.round(T[keep], 3)))
    with pytest.raises(ValueError):
        E.readCsvData(str(p), y_cols=items, t_cols=[i + "_S" for i in items], drop_missing=False)


def test_set_data_on_device_draws_only_the_person_level_part():
    """setDataOnDevice (the N x J part of setData* is left to erirt_generate_data): person-level draws per model as in
    src/SimTools.jl:117-368, the error law handed to the device, nothing of size N x J on the host."""
    C = E.setCond(nSubj=4000, nItem=7, nFeat=3, nIter=10, nChain=1)
    cases = [("MlIrt", E.setTrueParaMlIrt, "norm", "unit", True, False), ("RtIrtNull", E.setTrueParaRtIrt, "norm", "tnorm", False, True),
             ("RtIrt", E.setTrueParaRtIrt, "norm", "tnorm", True, True), ("RtIrtCross", E.setTrueParaRtIrtCross, "tail", "tail", False, True),
             ("RtIrtLatent", E.setTrueParaRtIrtLatent, "skew", "unit", True, True)]
    for model, mk, typ, err, has_x, has_zeta in cases:
        tp = mk(C, rng=3)
        D = E.setDataOnDevice(C, tp, model, type=typ, rng=3)
        assert isinstance(D, E.DeviceData) and D.error == err and D.Y is None and D.logT is None
        assert tp.theta.shape == (4000,) and (np.size(tp.zeta) == 4000) == has_zeta
        assert (D.X is not None and D.X.shape == (4000, 3)) == has_x
        assert abs(tp.theta.std() - (np.sqrt(1 + np.sum(tp.beta[:, 0] ** 2)) if model == "RtIrt" else tp.theta.std())) < 0.1
    # same seed -> same data-domain seed: the DIC evaluation regenerates the data set the chain was sampled on
    tp1, tp2 = E.setTrueParaRtIrt(C, rng=3), E.setTrueParaRtIrt(C, rng=3)
    assert E.setDataOnDevice(C, tp1, "RtIrt", rng=9).seed == E.setDataOnDevice(C, tp2, "RtIrt", rng=9).seed
    with pytest.raises(ValueError):
        E.setDataOnDevice(C, E.setTrueParaRtIrtCross(C, rng=3), "RtIrtCross", type="cauchy", rng=3)
    with pytest.raises(ValueError):
        E.setDataOnDevice(C, tp1, "NoSuchModel", rng=3)


def test_beta_intercept_rows_are_dropped_per_column():
    """comparePara / getMetrics on β of GibbsRtIrt: vec(β) is column-major over (nFeat+1) x 2 and the reference drops row 1 of each
    column (β[2:end, :], src/SimTools.jl:505); nFeat = 3 so that a wrong 'last len(true) entries' rule would misalign."""
    from erirt_b200 import simtools
    F = 3
    true = np.arange(1.0, 2 * F + 1).reshape(F, 2, order="F")          # slopes only, F x 2
    est = np.vstack([[100.0, 200.0], true + 0.5])                     # (F+1) x 2 with an intercept row
    obj = {"True": {"β": true.ravel(order="F")}, "Run1": {"β": est.ravel(order="F")}, "Run2": {"β": est.ravel(order="F")}}
    m = simtools.getMetrics(obj, par="β")
    assert abs(m["Bias"] - 0.5) < 1e-12 and abs(m["Rmse"] - 0.5) < 1e-12

    class M:
        pass
    mc = M()
    mc.truePara, mc.Post = E.InputPara(), M()
    mc.Post.mean = E.InputPara()
    mc.truePara.beta, mc.Post.mean.beta = true, est.ravel(order="F")
    import io
    tab = simtools.comparePara(mc, par="β", file=io.StringIO())
    assert np.allclose(tab[:, 2], 0.5)

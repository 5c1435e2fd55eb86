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

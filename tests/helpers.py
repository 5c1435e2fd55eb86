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

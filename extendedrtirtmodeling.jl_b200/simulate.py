"""Host-side mirrors of the reference's simulation helpers (src/SimTools.jl:74-368): true-parameter
generators and data generators.  Same distributions, numpy RNG (the reference uses Julia's global RNG, so
streams differ; only the distributions are part of the contract).  They define the synthetic inputs of the
benchmarks."""
import numpy as np

from .structs import InputData, InputPara


def _rng(rng):
    return rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)


def _tnorm_pos(rng, mu, sd, size, upper=np.inf, lower=0.0):
    out = np.empty(size)
    flat = out.reshape(-1)
    mu_b = np.broadcast_to(mu, out.shape).reshape(-1)
    sd_b = np.broadcast_to(sd, out.shape).reshape(-1)
    todo = np.arange(flat.size)
    rounds = 0
    while todo.size and rounds < 8:
        v = rng.normal(mu_b[todo], sd_b[todo])
        ok = (v > lower) & (v < upper)
        flat[todo[ok]] = v[ok]
        todo = todo[~ok]
        rounds += 1
    if todo.size:
        # far tails (mean many SDs outside the interval: plain rejection would never end; Distributions.jl's Truncated{Normal}
        # switches to a tail sampler there too): exact inverse-CDF draw in the standardised interval
        from scipy import stats
        a = (lower - mu_b[todo]) / sd_b[todo]
        b = (upper - mu_b[todo]) / sd_b[todo]
        flat[todo] = stats.truncnorm.rvs(a, b, loc=mu_b[todo], scale=sd_b[todo], random_state=rng)
    return out


def setTrueParaMlIrt(Cond, rng=None):
    """src/SimTools.jl:100-112"""
    rng = _rng(rng)
    p = InputPara()
    p.a = _tnorm_pos(rng, 1.0, 0.2, Cond.nItem)
    p.b = rng.normal(0.0, 0.5, Cond.nItem)
    p.beta = rng.normal(0.0, 1.0, (Cond.nFeat, 1))
    return p


def setTrueParaRtIrt(Cond, trueStdRa=1.0, trueStdRt=1.0, trueCorr=0.0, rng=None):
    """src/SimTools.jl:74-95"""
    rng = _rng(rng)
    p = InputPara()
    p.a = _tnorm_pos(rng, 1.0, 0.2, Cond.nItem)
    p.b = rng.normal(0.0, 0.5, Cond.nItem)
    p.lambda_ = _tnorm_pos(rng, 4.0, 0.2, Cond.nItem)
    p.sigma2t = rng.lognormal(np.log(0.3), 0.2, Cond.nItem)
    sd = np.diag([trueStdRa, trueStdRt])
    p.Sigma_p = sd @ np.array([[1.0, trueCorr], [trueCorr, 1.0]]) @ sd
    p.beta = rng.normal(0.0, 1.0, (Cond.nFeat, 2))
    return p


def setTrueParaRtIrtCross(Cond, trueStdRa=1.0, trueStdRt=1.0, rng=None):
    """src/SimTools.jl:186-208"""
    rng = _rng(rng)
    p = InputPara()
    p.a = _tnorm_pos(rng, 1.0, 0.2, Cond.nItem)
    p.b = rng.normal(0.0, 0.5, Cond.nItem)
    p.lambda_ = _tnorm_pos(rng, 3.0, 0.5, Cond.nItem)
    p.sigma2t = rng.lognormal(np.log(0.3), 0.2, Cond.nItem)
    p.Sigma_p = np.diag([trueStdRa ** 2, trueStdRt ** 2])
    p.rho = rng.normal(0.0, 0.2, Cond.nItem)
    return p


def setTrueParaRtIrtLatent(Cond, trueStdRa=1.0, trueStdRt=1.0, rng=None):
    """src/SimTools.jl:260-290"""
    rng = _rng(rng)
    p = InputPara()
    p.a = _tnorm_pos(rng, 1.0, 0.2, Cond.nItem)
    p.b = rng.normal(0.0, 0.5, Cond.nItem)
    p.lambda_ = _tnorm_pos(rng, 3.0, 0.2, Cond.nItem)
    p.Sigma_p = np.diag([trueStdRa ** 2, trueStdRt ** 2])
    rho = _tnorm_pos(rng, 0.0, 0.5, 1, upper=1.0, lower=-1.0)[0]
    p.beta = np.concatenate([rng.normal(0.0, 0.5, Cond.nFeat), [rho]])
    return p


def _bernoulli_logit(rng, eta):
    return (rng.random(eta.shape) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)


def setDataMlIrt(Cond, truePara, rng=None):
    """src/SimTools.jl:349-368"""
    rng = _rng(rng)
    N, F = Cond.nSubj, Cond.nFeat
    X = np.empty((N, F))
    X[:, 0] = rng.random(N) < 0.5
    X[:, 1:] = rng.normal(0.0, 1.0, (N, F - 1))
    truePara.theta = rng.normal((X @ truePara.beta).ravel(), 1.0)
    Y = _bernoulli_logit(rng, truePara.a[None, :] * (truePara.theta[:, None] - truePara.b[None, :]))
    return InputData(Y=Y, X=X)


def _mvn2(rng, Sigma, n):
    return rng.multivariate_normal(np.zeros(2), Sigma, n)


def setDataRtIrtNull(Cond, truePara, rng=None):
    """src/SimTools.jl:117-144"""
    rng = _rng(rng)
    noise = _mvn2(rng, truePara.Sigma_p, Cond.nSubj)
    truePara.theta, truePara.zeta = noise[:, 0].copy(), noise[:, 1].copy()
    Y = _bernoulli_logit(rng, truePara.a[None, :] * (truePara.theta[:, None] - truePara.b[None, :]))
    mu = truePara.lambda_[None, :] - truePara.zeta[:, None]
    logT = _tnorm_pos(rng, mu, np.sqrt(truePara.sigma2t)[None, :], mu.shape)
    return InputData(Y=Y, T=np.exp(logT))


def setDataRtIrt(Cond, truePara, rng=None):
    """src/SimTools.jl:149-178"""
    rng = _rng(rng)
    X = rng.normal(0.0, 1.0, (Cond.nSubj, Cond.nFeat))
    subj = X @ truePara.beta + _mvn2(rng, truePara.Sigma_p, Cond.nSubj)
    truePara.theta, truePara.zeta = subj[:, 0].copy(), subj[:, 1].copy()
    Y = _bernoulli_logit(rng, truePara.a[None, :] * (truePara.theta[:, None] - truePara.b[None, :]))
    mu = truePara.lambda_[None, :] - truePara.zeta[:, None]
    logT = _tnorm_pos(rng, mu, np.sqrt(truePara.sigma2t)[None, :], mu.shape)
    return InputData(Y=Y, X=X, T=np.exp(logT))


def _errors(rng, type, size):
    if type == "norm":
        return rng.normal(0.0, 0.3, size)
    if type == "tail":
        return rng.standard_t(5, size)
    if type == "skew":
        return rng.gamma(0.5, 1.0, size) - 1.0
    raise ValueError("type must be 'norm', 'tail' or 'skew'")


def setDataRtIrtCross(Cond, truePara, type="norm", rng=None):
    """src/SimTools.jl:220-255"""
    rng = _rng(rng)
    noise = _mvn2(rng, truePara.Sigma_p, Cond.nSubj)
    truePara.theta, truePara.zeta = noise[:, 0].copy(), noise[:, 1].copy()
    Y = _bernoulli_logit(rng, truePara.a[None, :] * (truePara.theta[:, None] - truePara.b[None, :]))
    mu = truePara.lambda_[None, :] - truePara.zeta[:, None] - np.outer(truePara.theta, truePara.rho)
    logT = mu + _errors(rng, type, mu.shape)
    return InputData(Y=Y, T=np.exp(logT))


def setDataRtIrtLatent(Cond, truePara, type="norm", rng=None, dtype=np.float64):
    """src/SimTools.jl:300-343"""
    rng = _rng(rng)
    N = Cond.nSubj
    truePara.theta = rng.standard_normal(N)
    X = rng.normal(0.0, 1.0, (N, Cond.nFeat))
    x = np.column_stack([X, truePara.theta])
    truePara.zeta = x @ truePara.beta + _errors(rng, type, N)
    Y = _bernoulli_logit(rng, truePara.a[None, :] * (truePara.theta[:, None] - truePara.b[None, :]))
    logT = truePara.lambda_[None, :] - truePara.zeta[:, None] + rng.standard_normal((N, Cond.nItem))
    return InputData(Y=Y, T=np.exp(logT), X=X)


class DeviceData:
    """Stand-in for InputData whose N x J part (Y, logT) is generated ON THE DEVICE by erirt_generate_data when `sample` creates
    the engine: only the person-level part of setData* (theta, zeta, X: O(N) values) is drawn on the host.  SURVEY.md 8f-1."""

    def __init__(self, truePara, X=None, error="unit", seed=1234):
        self.truePara, self.X, self.error, self.seed = truePara, X, error, int(seed)
        self.Y = self.T = self.logT = self.kappa = None  # never materialised on the host (engine.get_data() copies them back on request)

    def generate_into(self, eng, has_rt, lo=0, hi=None):
        """Fill the engine (a whole chain, or the shard [lo, hi) of one) from the person-level draws held here."""
        p = self.truePara
        sl = slice(lo, hi)
        eng.generate_data(p.theta[sl], p.a, p.b, p.zeta[sl] if has_rt else None, p.lambda_ if has_rt else None,
                          p.sigma2t if (has_rt and np.size(p.sigma2t)) else None, p.rho if np.size(p.rho) else None,
                          None if self.X is None else self.X[sl], error=self.error, seed=self.seed)


def setDataOnDevice(Cond, truePara, model, type="norm", rng=None, seed=None):
    """setData* of src/SimTools.jl:117-368 with the N x J part left to the device.  model: "MlIrt", "RtIrtNull", "RtIrt", "RtIrtCross",
    "RtIrtLatent" (also for the *Qr samplers).  The person-level draws follow the reference line by line; returns a DeviceData."""
    rng = _rng(rng)
    N, F = Cond.nSubj, Cond.nFeat
    seed = int(rng.integers(1, 2 ** 62)) if seed is None else seed
    if model == "MlIrt":            # :349-368
        X = np.empty((N, F))
        X[:, 0] = rng.random(N) < 0.5
        X[:, 1:] = rng.normal(0.0, 1.0, (N, F - 1))
        truePara.theta = rng.normal((X @ truePara.beta).ravel(), 1.0)
        return DeviceData(truePara, X, "unit", seed)
    if model == "RtIrtNull":        # :117-144
        noise = _mvn2(rng, truePara.Sigma_p, N)
        truePara.theta, truePara.zeta = noise[:, 0].copy(), noise[:, 1].copy()
        return DeviceData(truePara, None, "tnorm", seed)
    if model == "RtIrt":            # :149-178
        X = rng.normal(0.0, 1.0, (N, F))
        subj = X @ truePara.beta + _mvn2(rng, truePara.Sigma_p, N)
        truePara.theta, truePara.zeta = subj[:, 0].copy(), subj[:, 1].copy()
        return DeviceData(truePara, X, "tnorm", seed)
    if model in ("RtIrtCross", "RtIrtCrossQr"):   # :220-255, cell-level errors of the given type
        if type not in ("norm", "tail", "skew"):
            raise ValueError("type must be 'norm', 'tail' or 'skew'")
        noise = _mvn2(rng, truePara.Sigma_p, N)
        truePara.theta, truePara.zeta = noise[:, 0].copy(), noise[:, 1].copy()
        return DeviceData(truePara, None, type, seed)
    if model in ("RtIrtLatent", "RtIrtLatentQr", "RtIrtQuantile"):  # :300-343, person-level errors of the given type, N(0,1) cells
        truePara.theta = rng.standard_normal(N)
        X = rng.normal(0.0, 1.0, (N, F))
        truePara.zeta = np.column_stack([X, truePara.theta]) @ truePara.beta + _errors(rng, type, N)
        return DeviceData(truePara, X, "unit", seed)
    raise ValueError(f"unknown model {model!r}")


def getRmse(a, b):
    """src/SimTools.jl:42"""
    return float(np.sqrt(np.mean((np.asarray(a) - np.asarray(b)) ** 2)))


def getBias(a, b):
    """src/SimTools.jl:43"""
    return float(np.mean(np.asarray(a) - np.asarray(b)))

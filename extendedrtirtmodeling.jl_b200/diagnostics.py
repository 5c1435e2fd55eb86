"""Convergence diagnostics used by the reference's reporting layer (precis / checkConvergence,
src/Base.pl.jl:152-167, src/SimTools.jl:419-443): rank-normalised split-chain bulk ESS and R-hat
(Vehtari et al. 2021, the estimator behind recent MCMCChains.ess_rhat; the reference does not pin the
version, so this is stated rather than matched bit for bit)."""
import numpy as np
from scipy import stats as _st


def _split(x):
    n = x.shape[0] // 2
    return np.concatenate([x[:n], x[-n:]], axis=1) if n > 0 else x


def _rank_normalise(x):
    r = _st.rankdata(x.ravel(), method="average").reshape(x.shape)
    return _st.norm.ppf((r - 0.375) / (x.size + 0.25))


def _autocov(x):
    n = x.shape[0]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=0, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=0)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=0)[:n].real
    return ac / n


def _ess_rhat_raw(x):
    """x: (draws, chains)"""
    n, m = x.shape
    if n < 4 or not np.all(np.isfinite(x)) or np.ptp(x) == 0:
        return np.nan, np.nan
    ac = _autocov(x)
    chain_var = ac[0] * n / (n - 1.0)
    W = chain_var.mean()
    B = n * x.mean(axis=0).var(ddof=1) if m > 1 else 0.0
    var_plus = W * (n - 1.0) / n + B / n
    rhat = np.sqrt(var_plus / W) if W > 0 else np.nan
    rho = 1.0 - (W - ac.mean(axis=1)) / var_plus
    rho[0] = 1.0
    # Geyer's initial monotone sequence
    tau, t, prev = -1.0, 0, np.inf
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau += 2.0 * pair
        prev = pair
        t += 2
    tau = max(tau, 1.0 / np.log10(n * m))
    return n * m / tau, rhat


def ess_rhat(x):
    """Bulk ESS and R-hat of draws x with shape (draws, chains)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    z = _rank_normalise(_split(x)) if np.ptp(x) > 0 else _split(x)
    return _ess_rhat_raw(z)


def summarize(arr, names=None):
    """arr: (draws, params, chains) -> list of dict(mean, std, ess, rhat, q025, q975), like getPrecisTable."""
    out = []
    for c in range(arr.shape[1]):
        x = arr[:, c, :]
        e, r = ess_rhat(x)
        out.append(dict(name=names[c] if names else str(c), mean=float(x.mean()), std=float(x.std(ddof=1)), ess=e, rhat=r,
                        q025=float(np.quantile(x, 0.025)), q975=float(np.quantile(x, 0.975))))
    return out

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


def ess_rhat_batched(arr, device=None):
    """Bulk ESS and R-hat of many parameters at once: arr (draws, params, chains) -> (ess[params], rhat[params]).
    Same estimator as ess_rhat (rank-normalised split chains, FFT autocovariance, Geyer's initial monotone sequence), written
    with torch tensor operations so that the whole table of checkConvergence (src/SimTools.jl:419-443: every column of
    Post.ra / Post.rt / Post.qr) is one batched computation; device="cuda" keeps it on the GPU next to the traces
    (SURVEY 8f-2), device=None runs the identical code on the CPU."""
    import torch
    x = torch.as_tensor(np.asarray(arr, dtype=np.float64), device=device)
    n0, P, m0 = x.shape
    h = n0 // 2
    if h > 0:
        x = torch.cat([x[:h], x[n0 - h:]], dim=2)  # split chains
    n, _, m = x.shape
    const = (x.amax(dim=(0, 2)) - x.amin(dim=(0, 2))) == 0
    bad = const | ~torch.isfinite(x).all(dim=0).all(dim=1)
    xs = torch.nan_to_num(x)
    # rank normalisation per parameter over all draws and chains (average ranks for ties)
    flat = xs.permute(1, 0, 2).reshape(P, n * m)
    order = flat.argsort(dim=1, stable=True)
    ranks = torch.empty_like(flat)
    ar = torch.arange(1, n * m + 1, dtype=torch.float64, device=x.device).expand(P, -1)
    ranks.scatter_(1, order, ar)
    srt = flat.gather(1, order)
    # average the ranks of equal values: group boundaries in sorted order
    newgrp = torch.ones_like(srt, dtype=torch.bool)
    newgrp[:, 1:] = srt[:, 1:] != srt[:, :-1]
    gid = newgrp.cumsum(dim=1) - 1
    gsum = torch.zeros_like(srt).scatter_add_(1, gid, ar)
    gcnt = torch.zeros_like(srt).scatter_add_(1, gid, torch.ones_like(srt))
    avg_sorted = (gsum / gcnt.clamp(min=1)).gather(1, gid)
    ranks.scatter_(1, order, avg_sorted)
    z = torch.special.ndtri((ranks - 0.375) / (n * m + 0.25)).reshape(P, n, m).permute(1, 0, 2)  # (n, P, m)
    # autocovariance by FFT
    L = 1 << int(np.ceil(np.log2(2 * n)))
    zc = z - z.mean(dim=0, keepdim=True)
    f = torch.fft.rfft(zc, n=L, dim=0)
    ac = torch.fft.irfft(f * f.conj(), n=L, dim=0)[:n] / n  # (n, P, m)
    chain_var = ac[0] * n / (n - 1.0)
    W = chain_var.mean(dim=1)
    B = n * z.mean(dim=0).var(dim=1, unbiased=True) if m > 1 else torch.zeros_like(W)
    var_plus = W * (n - 1.0) / n + B / n
    rhat = torch.sqrt(var_plus / W)
    rho = 1.0 - (W[None, :] - ac.mean(dim=2)) / var_plus[None, :]
    rho[0] = 1.0
    # Geyer: pair sums, cut at the first negative pair, running minimum
    npairs = n // 2
    pairs = rho[0:2 * npairs:2] + rho[1:2 * npairs:2]  # (npairs, P)
    alive = (pairs >= 0).to(torch.float64).cumprod(dim=0)
    mono = torch.cummin(torch.where(alive > 0, pairs, torch.full_like(pairs, float("inf"))), dim=0).values
    tau = -1.0 + 2.0 * (torch.where(alive > 0, mono, torch.zeros_like(mono))).sum(dim=0)
    tau = torch.clamp(tau, min=1.0 / np.log10(n * m))
    ess = n * m / tau
    nan = torch.full_like(ess, float("nan"))
    ess = torch.where(bad | (n < 4), nan, ess)
    rhat = torch.where(bad | (n < 4), nan, rhat)
    return ess.cpu().numpy(), rhat.cpu().numpy()


def ess_rhat_device(arr, skip=0, device=0):
    """Bulk ESS and R-hat of arr (draws, params, chains) by the CUDA kernel of the engine (erirt_ess_rhat, csrc/diagnostics.cuh:
    one CTA per parameter: bitonic sort for the ranks, direct autocovariances, Geyer's sequence), first `skip` draws discarded.
    Same estimator as ess_rhat above, which is its checker in the GPU tests.  Raises when the CUDA library or a GPU is missing."""
    import ctypes as C
    from . import _lib
    x = np.asarray(arr, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    n, P, m = x.shape
    xf = np.ascontiguousarray(x.transpose(2, 1, 0))  # [chain][param][draw] == Julia's column-major [nIter, P, nChain]
    ess, rhat = np.empty(P), np.empty(P)
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.load().erirt_ess_rhat(xf.ctypes.data_as(dp), n, P, m, int(skip), int(device), ess.ctypes.data_as(dp), rhat.ctypes.data_as(dp)))
    return ess, rhat

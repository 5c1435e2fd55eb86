"""Config / data / state containers mirroring src/Base.pl.jl:45-149 of the reference."""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class SimConditions:
    """struct SimConditions, src/Base.pl.jl:45-56"""
    nSubj: int
    nItem: int
    nFeat: int
    nIter: int
    nChain: int
    nBurnin: int
    nThin: int
    nRep: int
    qRa: float
    qRt: float


def setCond(nSubj=2000, nItem=15, nFeat=3, nIter=5000, nChain=4, nBurnin=None, nThin=1, nRep=10, qRa=0.5, qRt=0.5):
    """setCond, src/Base.pl.jl:59-62: nBurnin is ALWAYS round(Int, nIter/2), whatever is passed (quirk Q13)."""
    nBurnin = int(round(nIter / 2))
    return SimConditions(nSubj, nItem, nFeat, nIter, nChain, nBurnin, nThin, nRep, float(qRa), float(qRt))


class InputData:
    """struct InputData, src/Base.pl.jl:67-78: κ = Y .- 0.5, logT = log.(T)."""

    def __init__(self, Y=None, T=None, X=None):
        self.Y = None if Y is None else np.asarray(Y, dtype=np.float64)
        self.kappa = None if Y is None else self.Y - 0.5
        self.T = None if T is None else np.asarray(T, dtype=np.float64)
        self.logT = None if T is None else np.log(self.T)
        self.X = None if X is None else np.asarray(X, dtype=np.float64)


class InputData4R:
    """struct InputData4R, src/Base.pl.jl:86-95 (fields passed through untouched)."""

    def __init__(self, Y=None, kappa=None, T=None, logT=None, X=None):
        self.Y, self.kappa, self.T, self.logT, self.X = Y, kappa, T, logT, X


class InputPara:
    """mutable struct InputPara, src/Base.pl.jl:100-115.  ASCII field names with the reference's Greek
    names as aliases (ω θ ζ λ ν β ρ Σp; σ²t is not a valid Python identifier -> sigma2t)."""
    _FIELDS = ("omega", "theta", "a", "b", "zeta", "lambda_", "sigma2t", "nu", "beta", "rho", "Sigma_p")
    _ALIASES = {"ω": "omega", "θ": "theta", "ζ": "zeta", "λ": "lambda_", "ν": "nu", "β": "beta", "ρ": "rho",
                "Σp": "Sigma_p", "σ2t": "sigma2t"}

    def __init__(self, **kw):
        for f in self._FIELDS:
            object.__setattr__(self, f, np.zeros(0))
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        alias = type(self)._ALIASES.get(name)
        if alias is None:
            raise AttributeError(name)
        return object.__getattribute__(self, alias)

    def __setattr__(self, name, value):
        name = type(self)._ALIASES.get(name, name)
        if name not in self._FIELDS:
            raise AttributeError(f"InputPara has no field {name}")
        object.__setattr__(self, name, np.asarray(value, dtype=np.float64))


class OutputDic:
    """mutable struct OutputDic, src/Base.pl.jl:130-136"""

    def __init__(self, pD=None, DIC=None):
        self.pD, self.DIC = pD, DIC


class OutputPost:
    """OutputPost* structs (src/GibbsRtIrt.pl.jl:35-71 etc.): trace arrays in the Julia layout
    [nIter, P, nChain] plus `mean::InputPara`."""

    def __init__(self):
        self.ra = self.rt = self.qr = self.logLike = None
        self.mean = None
        self.sd = None          # extension: posterior SDs of the person parameters (running moments)
        self.person_cols = True  # False when ra/rt hold only the item columns (large nSubj)

"""Second, independent restatement of /root/reference/src/Draw.pl.jl in dense numpy broadcasting.

TEST INFRASTRUCTURE ONLY.  Each function transcribes one Julia function expression by expression
(`Para.a' .* (Para.θ .- Para.b')` -> `a[None, :] * (θ[:, None] - b[None, :])`), uses the same
counter-based random stream as oracle/rng.h (re-implemented here in pure Python integers), and is
used by tests/test_oracle_pin.py to pin the C oracle: two restatements written independently must
agree to rounding on every block of every model.  Small sizes only (pure-Python RNG loops).
"""
import math

import numpy as np

M32 = 0xFFFFFFFF
DOM_PERSON, DOM_ITEM, DOM_GLOBAL = 1, 2, 3
PK_NORMALS, PK_NU, PK_PG, PK_PG_RETRY, PK_NU_CELL = 0, 1, 2, 3, 4
IK_B, IK_A, IK_LAMBDA, IK_SIGMA2, IK_RHO = 0, 1, 2, 3, 4
GK_BETA, GK_SIGMAP = 0, 1
PG_T = 0.64
PG_P0 = 0.10564977366685535
PG_Q0 = 4 * PG_P0


def site(dom, kind, idx=0):
    return (dom << 28) | (kind << 20) | idx


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Stream:
    def __init__(self, seed, chain=0):
        self.key = (seed & M32, ((seed >> 32) ^ ((chain * 0x9E3779B9) & M32)) & M32)
        self.sweep = 1

    def words(self, unit, s, attempt=0):
        return philox4x32_10((unit, self.sweep, s, attempt), self.key)

    @staticmethod
    def u01(w):
        return (w + 0.5) / 4294967296.0

    @classmethod
    def normal2(cls, w0, w1):
        return math.sqrt(-2.0 * math.log(cls.u01(w0))) * math.cos(2.0 * math.pi * cls.u01(w1))

    def normal(self, unit, s):
        w = self.words(unit, s)
        return self.normal2(w[0], w[1])

    def tnorm_pos(self, unit, s, mu, sd):
        alpha = -mu / sd
        att = 0
        if alpha <= 0.5:
            while True:
                w = self.words(unit, s, att)
                z = self.normal2(w[0], w[1])
                if z >= alpha:
                    return mu + sd * z
                att += 1
        lam = 0.5 * (alpha + math.sqrt(alpha * alpha + 4.0))
        while True:
            w = self.words(unit, s, att)
            x = alpha - math.log(self.u01(w[0])) / lam
            if self.u01(w[1]) <= math.exp(-0.5 * (x - lam) ** 2):
                return mu + sd * x
            att += 1

    def gamma(self, unit, s, shape):
        d = shape - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        att = 0
        while True:
            w = self.words(unit, s, att)
            att += 1
            x = self.normal2(w[0], w[1])
            v = 1.0 + c * x
            if v <= 0:
                continue
            v = v ** 3
            if math.log(self.u01(w[2])) < 0.5 * x * x + d - d * v + d * math.log(v):
                return d * v


def ig_msh(mu, lam, z, u):
    y = z * z
    w = mu * y
    x1 = 2.0 * lam * mu / (2.0 * lam + w + math.sqrt(w * (4.0 * lam + w)))
    return x1 if u <= mu / (mu + x1) else mu * mu / x1


def inv_normal_tail(y):
    from scipy import stats
    return float(stats.norm.isf(y))


def _series(U, x, right):
    e = (-0.5 * math.pi ** 2 * x) if right else (-2.0 / x)
    S = 1.0
    for n in range(1, 400):
        rn = (2 * n + 1) * math.exp(e * (n * n + n))
        if n & 1:
            S -= rn
            if U <= S:
                return True
        else:
            S += rn
            if U > S:
                return False
    return False


def _attempt_a(c, wa, wb):
    um, up = Stream.u01(wa), Stream.u01(wb)
    K = math.pi ** 2 / 8 + 0.5 * c * c
    Rm1 = (2 * PG_Q0 / math.pi) * K * math.exp(K * PG_T)
    v = um * (1 + Rm1)
    if v < 1:
        x = PG_T - math.log(up) / K
        return x if _series(v, x, True) else None
    Z = inv_normal_tail(up * PG_P0)
    x = 1 / (Z * Z)
    ua = (v - 1) / Rm1
    tilt = math.exp(-0.5 * c * c * x)
    if ua >= tilt:
        return None
    return x if _series(ua / tilt, x, False) else None


def _attempt_b(c, w):
    um = Stream.u01(w[0])
    K = math.pi ** 2 / 8 + 0.5 * c * c
    p = (math.pi / (2 * K)) * math.exp(-K * PG_T)
    ql = 2 * math.exp(-c)
    Pr = p / (p + ql)
    if um < Pr:
        x = PG_T - math.log(Stream.u01(w[1])) / K
        return x if _series(um / Pr, x, True) else None
    ua = (um - Pr) / (1 - Pr)
    x = ig_msh(1 / c, 1.0, Stream.normal2(w[1], w[2]), Stream.u01(w[3]))
    if x >= PG_T:
        return None
    return x if _series(ua, x, False) else None


def pg1(st, person, j, z):
    """Stream layout of oracle/pg.c: attempt 0 = Method A from the pair block when |z| <= 16; retry block r >= 1 holds
    Method-A attempts 2r-1, 2r (c <= 1/t) or Method-B attempt r (c > 1/t)."""
    c = 0.5 * abs(z)
    x = None
    if abs(z) <= 16.0:
        w = st.words(person, site(DOM_PERSON, PK_PG, j >> 1), 0)
        x = _attempt_a(c, w[(j & 1) * 2], w[(j & 1) * 2 + 1])
    r = 1
    while x is None:
        w = st.words(person, site(DOM_PERSON, PK_PG_RETRY, j), r)
        if c <= 1 / PG_T:
            x = _attempt_a(c, w[0], w[1])
            if x is None:
                x = _attempt_a(c, w[2], w[3])
        else:
            x = _attempt_b(c, w)
        r += 1
    return 0.25 * x


# ----------------------------------------------------------------------------------------------
# Draw.pl.jl, transcribed.  P is a dict with the InputPara fields; D has Y, κ, logT, X; C has N,J,F,qRt.
# ----------------------------------------------------------------------------------------------
def _col(v):
    return np.asarray(v, dtype=float).reshape(-1, 1)


def _row(v):
    return np.asarray(v, dtype=float).reshape(1, -1)


def _person_normals(st, N, which):
    out = np.empty(N)
    for i in range(N):
        w = st.words(i, site(DOM_PERSON, PK_NORMALS))
        out[i] = Stream.normal2(w[0], w[1]) if which == 0 else Stream.normal2(w[2], w[3])
    return out


def _k12(C):
    q = C["qRt"]
    return (1 - 2 * q) / (q * (1 - q)), 2 / (q * (1 - q))


def drawRaPgRandomVariable(st, P):  # Draw.pl.jl:36-40
    η = _row(P["a"]) * (_col(P["θ"]) - _row(P["b"]))
    ω = np.empty_like(η)
    for i in range(η.shape[0]):
        for j in range(η.shape[1]):
            ω[i, j] = pg1(st, i, j, η[i, j])
    return ω


def drawSubjAbility(st, C, D, P, null=False, prior_var=None):  # :49-62 / :67-80
    N = C["N"]
    x = np.hstack([np.ones((N, 1)), D["X"]]) if not null else None
    θμ0 = (x @ np.asarray(P["β"]).reshape(C["F"] + 1, -1, order="F")[:, 0]).reshape(-1, 1) if not null else 0.0
    θσ02 = P["Σp"][0, 0] if prior_var is None else prior_var
    a, b, ω = _row(P["a"]), _row(P["b"]), P["ω"]
    parV = 1 / (1 / θσ02 + np.sum(a ** 2 * ω, axis=1, keepdims=True))
    parM = parV * (θμ0 / θσ02 + np.sum(a * (D["κ"] + a * b * ω), axis=1, keepdims=True))
    return (parM + np.sqrt(parV) * _col(_person_normals(st, N, 0))).ravel()


def drawItemDiscrimination(st, D, P):  # :88-93
    θ, b = _col(P["θ"]), _row(P["b"])
    parV = 1 / (1 / 1 ** 2 + np.sum((θ - b) ** 2 * P["ω"], axis=0))
    parM = parV * (1.0 / 1 ** 2 + np.sum(D["κ"] * (θ - b), axis=0))
    return np.array([st.tnorm_pos(j, site(DOM_ITEM, IK_A), parM[j], math.sqrt(parV[j])) for j in range(parV.size)])


def drawItemDifficulty(st, D, P):  # :98-105
    a, θ = _row(P["a"]), _col(P["θ"])
    parV = 1 / (1 / 1 ** 2 + np.sum(a ** 2 * P["ω"], axis=0))
    parM = parV * (0.0 / 1 ** 2 - np.sum(a * (D["κ"] - (θ @ a) * P["ω"]), axis=0))
    b = np.array([parM[j] + math.sqrt(parV[j]) * st.normal(j, site(DOM_ITEM, IK_B)) for j in range(parV.size)])
    return np.clip(b, -4, 4)


def drawSubjSpeed(st, C, D, P, kind):  # :119-206
    N, F = C["N"], C["F"]
    k1, k2 = _k12(C)
    θ, λ, σ2 = _col(P["θ"]), _row(P["λ"]), _row(P["σ²t"])
    logT = D["logT"]
    if kind == "null":
        ζμ0, ζσ02 = 0.0, 1.0
    elif kind == "x":
        x = np.hstack([np.ones((N, 1)), D["X"]])
        ζμ0 = (x @ np.asarray(P["β"]).reshape(F + 1, 2, order="F")[:, 1]).reshape(-1, 1)
        ζσ02 = P["Σp"][1, 1]
    elif kind in ("latent", "latentqr"):
        x = np.hstack([np.ones((N, 1)), D["X"], θ])
        ζμ0 = (x @ np.asarray(P["β"]).ravel()).reshape(-1, 1)
        ζσ02 = P["Σp"][1, 1]
        if kind == "latentqr":
            ν = _col(P["ν"])
            ζμ0 = ζμ0 + k1 * ν
            ζσ02 = P["Σp"][1, 1] * (k2 * ν)
    else:
        ζμ0, ζσ02 = 0.0, P["Σp"][1, 1]
    if kind == "cross":
        parV = 1 / (1 / ζσ02 + np.sum(1.0 / σ2 * np.ones_like(logT), axis=1, keepdims=True))
        parM = parV * (ζμ0 / ζσ02 + np.sum((λ - logT - θ @ _row(P["ρ"])) / σ2, axis=1, keepdims=True))
    elif kind == "crossqr":
        k1e, k2e = k1 * P["ν"], k2 * P["ν"]
        parV = 1 / (1 / ζσ02 + np.sum(1.0 / (σ2 * k2e), axis=1, keepdims=True))
        parM = parV * (ζμ0 / ζσ02 + np.sum((λ - logT - θ @ _row(P["ρ"]) + k1e) / (σ2 * k2e), axis=1, keepdims=True))
    else:
        parV = 1 / (1 / ζσ02 + np.sum(1.0 / σ2))
        parM = parV * (ζμ0 / ζσ02 + np.sum((λ - logT) / σ2, axis=1, keepdims=True))
    parV = np.broadcast_to(parV, (N, 1))
    return (parM + np.sqrt(parV) * _col(_person_normals(st, N, 1))).ravel()


def drawItemIntensity(st, C, D, P, kind):  # :215-251
    N = C["N"]
    k1, k2 = _k12(C)
    logT, ζ, θ, σ2 = D["logT"], _col(P["ζ"]), _col(P["θ"]), _row(P["σ²t"])
    μλ, σλ = np.mean(logT), np.std(logT, ddof=1)
    if kind == "plain":
        parV = 1 / (1 / σλ ** 2 + N / σ2)
        parM = parV * (μλ / σλ ** 2 + np.sum(logT + ζ, axis=0, keepdims=True) / σ2)
    elif kind == "cross":
        parV = 1 / (1 / σλ ** 2 + N / σ2)
        parM = parV * (μλ / σλ ** 2 + np.sum((logT + ζ + θ @ _row(P["ρ"])) / σ2, axis=0, keepdims=True))
    else:
        k1e, k2e = k1 * P["ν"], k2 * P["ν"]
        parV = 1 / (1 / σλ ** 2 + np.sum(1.0 / (σ2 * k2e), axis=0, keepdims=True))
        parM = parV * (μλ / σλ ** 2 + np.sum((logT + ζ + θ @ _row(P["ρ"]) - k1e) / (σ2 * k2e), axis=0, keepdims=True))
    parV, parM = parV.ravel(), parM.ravel()
    return np.array([st.tnorm_pos(j, site(DOM_ITEM, IK_LAMBDA), parM[j], math.sqrt(parV[j])) for j in range(parV.size)])


def drawItemTimeResidual(st, C, D, P, kind):  # :257-288
    N = C["N"]
    k1, k2 = _k12(C)
    logT, ζ, θ, λ = D["logT"], _col(P["ζ"]), _col(P["θ"]), _row(P["λ"])
    if kind == "plain":
        parA = 1e-3 + N / 2
        parB = 1e-3 + np.sum((logT - λ + ζ) ** 2, axis=0) / 2
    elif kind == "cross":
        parA = 1e-3 + N / 2
        parB = 1e-3 + np.sum((logT - λ + ζ + θ @ _row(P["ρ"])) ** 2, axis=0) / 2
    else:
        k1e, k2e = k1 * P["ν"], k2 * P["ν"]
        parA = 1e-3 + N * 3 / 2
        parB = 1e-3 + np.sum((logT - λ + ζ + θ @ _row(P["ρ"]) - k1e) ** 2 / (2 * k2e), axis=0) + np.sum(P["ν"], axis=0)
    return np.array([parB[j] / st.gamma(j, site(DOM_ITEM, IK_SIGMA2), parA) for j in range(parB.size)])


def drawQrWeightsCrossQr(st, C, D, P):  # :303-320
    k1, k2 = _k12(C)
    N, J = C["N"], C["J"]
    σ2 = _row(P["σ²t"])
    parA = np.abs(D["logT"] - _row(P["λ"]) + _col(P["ζ"]) + _col(P["θ"]) @ _row(P["ρ"])) / np.sqrt(σ2 * k2)
    parB = np.sqrt(2 * k2 + k1 ** 2) / np.sqrt(σ2 * k2)
    with np.errstate(divide="ignore"):
        μ = np.clip(parB / parA, 1e-10, np.inf)
    parB = np.broadcast_to(parB, (N, J))
    ν = np.empty((N, J))
    for i in range(N):
        for j in range(J):
            w = st.words(i, site(DOM_PERSON, PK_NU_CELL, j >> 1))
            rad = math.sqrt(-2.0 * math.log(Stream.u01(w[0])))
            ang = 2.0 * math.pi * Stream.u01(w[1])
            zn = rad * math.sin(ang) if j & 1 else rad * math.cos(ang)
            ν[i, j] = 1 / ig_msh(μ[i, j], parB[i, j] ** 2, zn, Stream.u01(w[3 if j & 1 else 2]))
    return np.clip(ν, 1e-10, 1e10)


def drawQrWeightsLatentQr(st, C, D, P):  # :325-343
    k1, k2 = _k12(C)
    N = C["N"]
    x = np.hstack([np.ones((N, 1)), D["X"], _col(P["θ"])])
    parA = np.abs(np.asarray(P["ζ"]) - x @ np.asarray(P["β"]).ravel()) / math.sqrt(P["Σp"][1, 1] * k2)
    parB = math.sqrt(2 * k2 + k1 ** 2) / math.sqrt(P["Σp"][1, 1] * k2)
    μ = np.clip(parB / parA, 1e-10, np.inf)
    ν = np.empty(N)
    for i in range(N):
        w = st.words(i, site(DOM_PERSON, PK_NU))
        ν[i] = 1 / ig_msh(μ[i], parB ** 2, Stream.normal2(w[0], w[1]), Stream.u01(w[2]))
    return np.clip(ν, 1e-10, 1e10)


def getSubjCoefficientsMlIrt(C, D, P):  # :351-357
    x = np.hstack([np.ones((C["N"], 1)), D["X"]])
    return np.linalg.solve(x.T @ x, x.T @ np.asarray(P["θ"]))


def _beta_normals(st, n):
    return np.array([st.normal(k, site(DOM_GLOBAL, GK_BETA)) for k in range(n)])


def drawSubjCoefficients(st, C, D, P):  # :380-393
    N, F = C["N"], C["F"]
    η = np.column_stack([P["θ"], P["ζ"]])
    x = np.hstack([np.ones((N, 1)), D["X"]])
    invΩ = np.linalg.inv(P["Σp"])
    parV = np.linalg.inv(1 / 1 ** 2 + np.kron(invΩ, x.T @ x))
    parM = parV @ (0.0 / 1 ** 2 + (x.T @ η @ invΩ.T).ravel(order="F"))
    β = parM + np.linalg.cholesky((parV + parV.T) / 2) @ _beta_normals(st, 2 * (F + 1))
    return β.reshape(F + 1, 2, order="F")


def drawSubjCoefficientsLatent(st, C, D, P):  # :399-416
    N, F = C["N"], C["F"]
    x = np.hstack([np.ones((N, 1)), D["X"], _col(P["θ"])])
    invΩ = 1 / P["Σp"][1, 1]
    parV = np.linalg.inv(1 / 1 ** 2 + invΩ * (x.T @ x))
    parM = parV @ (0.0 + (x.T @ np.asarray(P["ζ"])) * invΩ)
    return parM + np.linalg.cholesky((parV + parV.T) / 2) @ _beta_normals(st, F + 2)


def getSubjCoefficientsLatentQr(C, D, P):  # :446-458 -- literally: (k2eΣp ⊗ x'x) \ vec(x'(ζ - k1e) k2eΣp')
    k1, k2 = _k12(C)
    N = C["N"]
    x = np.hstack([np.ones((N, 1)), D["X"], _col(P["θ"])])
    k1e, k2e = k1 * np.asarray(P["ν"]), k2 * np.asarray(P["ν"])
    k2eΣp = _col(1 / (P["Σp"][1, 1] * k2e))
    A = np.kron(k2eΣp, x.T @ x)
    rhs = (_col(x.T @ (np.asarray(P["ζ"]) - k1e)) @ k2eΣp.T).ravel(order="F")
    return np.linalg.lstsq(A, rhs, rcond=None)[0]


def drawSubjCorr(st, C, D, P, qr):  # :463-489
    k1, k2 = _k12(C)
    θ, ζ, λ, σ2 = _col(P["θ"]), _col(P["ζ"]), _row(P["λ"]), _row(P["σ²t"])
    if qr:
        k1e, k2e = k1 * P["ν"], k2 * P["ν"]
        parV = 1 / (1 / 1 ** 2 + np.sum(θ ** 2 / (σ2 * k2e), axis=0))
        parM = parV * (0.0 + np.sum((θ * (λ - ζ - D["logT"] + k1e)) / (σ2 * k2e), axis=0))
    else:
        parV = 1 / (1 / 1 ** 2 + np.sum(θ ** 2 / σ2, axis=0))
        parM = parV * (0.0 + np.sum((θ * (λ - ζ - D["logT"])) / σ2, axis=0))
    return np.array([parM[j] + math.sqrt(parV[j]) * st.normal(j, site(DOM_ITEM, IK_RHO)) for j in range(parV.size)])


def _cov2one(s):
    s = s.copy()
    d = np.diag([s[0, 0] ** -0.5, 1.0])
    s = d @ s @ d
    d = np.diag([1.0, s[1, 1] ** -0.5])
    s = d @ s @ d
    s[0, 0] = s[1, 1] = 1.0
    return s


def _inv_wishart(st, df, Psi):
    S = np.linalg.inv(Psi)
    L = np.linalg.cholesky(S)
    A = np.zeros((2, 2))
    A[0, 0] = math.sqrt(2 * st.gamma(0, site(DOM_GLOBAL, GK_SIGMAP), df / 2))
    A[1, 1] = math.sqrt(2 * st.gamma(1, site(DOM_GLOBAL, GK_SIGMAP), (df - 1) / 2))
    A[1, 0] = st.normal(2, site(DOM_GLOBAL, GK_SIGMAP))
    Xm = L @ A
    return np.linalg.inv(Xm @ Xm.T)


def drawSubjCovariance(st, C, D, P, cov2one, null=False):  # :499-535
    N, F = C["N"], C["F"]
    η = np.column_stack([P["θ"], P["ζ"]])
    if null:
        e = η
    else:
        x = np.hstack([np.ones((N, 1)), D["X"]])
        e = η - x @ np.asarray(P["β"]).reshape(F + 1, 2, order="F")
    s = _inv_wishart(st, N + 3, e.T @ e + np.eye(2))
    return _cov2one(s) if cov2one else s


def drawSubjCovarianceIG(st, C, D, P, cov2one, kind, compat=0):  # :542-606
    N = C["N"]
    k1, k2 = _k12(C)
    ζ = np.asarray(P["ζ"])
    if kind == "cross":
        parA = 1e-3 + N / 2
        parB = 1e-3 + np.sum(ζ ** 2) / 2
    else:
        x = np.hstack([np.ones((N, 1)), D["X"], _col(P["θ"])])
        r = ζ - x @ np.asarray(P["β"]).ravel()
        if kind == "latent":
            parA = 1e-3 + N / 2
            parB = 1e-3 + np.sum(r ** 2) / 2
        else:
            ν = np.asarray(P["ν"])
            k1e, k2e = k1 * ν, k2 * ν
            parA = 1e-3 + N * 3 / 2
            if compat & 2:
                parB = 1e-3 + np.sum((r - k1e) ** 2 / (2 * k2e)) + np.sum(ν)
            else:
                # Julia: vector / vector = A * pinv(B) -> N x N matrix  A B' / (B'B)
                A, B = _col((r - k1e) ** 2), _col(2 * k2e)
                parB = 1e-3 + np.sum(A @ B.T / float(B.T @ B)) + np.sum(ν)
    sv = parB / st.gamma(0, site(DOM_GLOBAL, GK_SIGMAP), parA)
    Σp = np.array([[1.0, 0.0], [0.0, sv]])
    return _cov2one(Σp) if cov2one else Σp


# ---- log-likelihoods (GibbsRtIrt.pl.jl:195-204,262-272,351-361; Cross :158-169,240-258; Latent :151-161,243-264)
def loglik(model, C, D, P):
    from scipy import stats
    N, F = C["N"], C["F"]
    k1, k2 = _k12(C)
    θ, a, b = _col(P["θ"]), _row(P["a"]), _row(P["b"])
    pr = a * (θ - b)
    out = np.sum(D["Y"] * pr - np.logaddexp(0, pr))
    if model == "MlIrt":
        x = np.hstack([np.ones((N, 1)), D["X"]])
        return out + np.sum(stats.norm.logpdf(np.asarray(P["θ"]), x @ np.asarray(P["β"]).ravel(), 1.0))
    ζ, λ, σ2 = _col(P["ζ"]), _row(P["λ"]), _row(P["σ²t"])
    μt = λ - ζ
    sd = np.sqrt(σ2)
    if model in ("RtIrtCross", "RtIrtCrossQr"):
        μt = μt - θ @ _row(P["ρ"])
    if model == "RtIrtCrossQr":
        μt = μt + k1 * P["ν"]
        sd = np.sqrt(σ2 * (k2 * P["ν"]))
    out += np.sum(stats.norm.logpdf(D["logT"], μt, sd))
    if model in ("RtIrt", "RtIrtNull", "RtIrtCross", "RtIrtCrossQr"):
        η = np.column_stack([P["θ"], P["ζ"]])
        if model == "RtIrt":
            x = np.hstack([np.ones((N, 1)), D["X"]])
            μη = x @ np.asarray(P["β"]).reshape(F + 1, 2, order="F")
        else:
            μη = np.zeros((N, 2))
        Σ = np.asarray(P["Σp"])
        Σs = np.array([[Σ[0, 0], Σ[0, 1]], [Σ[0, 1], Σ[1, 1]]])
        out += sum(stats.multivariate_normal.logpdf(η[i], μη[i], Σs) for i in range(N))
    else:
        x = np.hstack([np.ones((N, 1)), D["X"], θ])
        μ = x @ np.asarray(P["β"]).ravel()
        sdζ = math.sqrt(P["Σp"][1, 1])
        if model == "RtIrtLatentQr":
            ν = np.asarray(P["ν"])
            μ = μ + k1 * ν
            sdζ = np.sqrt(P["Σp"][1, 1] * k2 * ν)
        out += np.sum(stats.norm.logpdf(np.asarray(P["ζ"]), μ, sdζ))
    return float(out)


# ---- the seven scans (sample! bodies) ----
def sweep(model, st, C, D, P, intercept=False, onepl=False, cov2one=True, compat=0):
    J = C["J"]

    def irt(order_ab=False, null_theta=True, prior_var=None):
        P["ω"] = drawRaPgRandomVariable(st, P)
        if order_ab:
            P["a"] = drawItemDiscrimination(st, D, P)
            if onepl:
                P["a"] = np.ones(J)
            P["b"] = drawItemDifficulty(st, D, P)
        else:
            P["b"] = drawItemDifficulty(st, D, P)
            P["a"] = drawItemDiscrimination(st, D, P)
            if onepl:
                P["a"] = np.ones(J)
        P["θ"] = drawSubjAbility(st, C, D, P, null=null_theta, prior_var=prior_var)

    if model == "MlIrt":
        P["β"] = getSubjCoefficientsMlIrt(C, D, P)
        if not intercept:
            P["β"][0] = 0.0
        irt(order_ab=True, null_theta=False, prior_var=1.0)
    elif model == "RtIrt":
        P["β"] = drawSubjCoefficients(st, C, D, P)
        if not intercept:
            P["β"][0, :] = 0.0
        P["Σp"] = drawSubjCovariance(st, C, D, P, cov2one)
        irt(null_theta=False)
        P["λ"] = drawItemIntensity(st, C, D, P, "plain")
        P["σ²t"] = drawItemTimeResidual(st, C, D, P, "plain")
        P["ζ"] = drawSubjSpeed(st, C, D, P, "x")
    elif model == "RtIrtNull":
        P["β"] = np.zeros((C["F"] + 1, 2))
        P["Σp"] = drawSubjCovariance(st, C, D, P, cov2one, null=True)
        irt()
        P["λ"] = drawItemIntensity(st, C, D, P, "plain")
        P["σ²t"] = drawItemTimeResidual(st, C, D, P, "plain")
        P["ζ"] = drawSubjSpeed(st, C, D, P, "null")
    elif model in ("RtIrtCross", "RtIrtCrossQr"):
        qr = model == "RtIrtCrossQr"
        if qr:
            P["ν"] = drawQrWeightsCrossQr(st, C, D, P)
        P["ρ"] = drawSubjCorr(st, C, D, P, qr)
        P["Σp"] = drawSubjCovarianceIG(st, C, D, P, cov2one, "cross")
        irt()
        kind = "crossqr" if qr else "cross"
        P["λ"] = drawItemIntensity(st, C, D, P, kind)
        P["σ²t"] = drawItemTimeResidual(st, C, D, P, kind)
        P["ζ"] = drawSubjSpeed(st, C, D, P, kind)
    elif model in ("RtIrtLatent", "RtIrtLatentQr"):
        qr = model == "RtIrtLatentQr"
        if qr:
            P["ν"] = drawQrWeightsLatentQr(st, C, D, P)
            P["β"] = getSubjCoefficientsLatentQr(C, D, P)
        else:
            P["β"] = drawSubjCoefficientsLatent(st, C, D, P)
        if not intercept:
            P["β"][0] = 0.0
        P["Σp"] = drawSubjCovarianceIG(st, C, D, P, cov2one, "latentqr" if qr else "latent", compat)
        irt()
        P["λ"] = drawItemIntensity(st, C, D, P, "plain")
        P["σ²t"] = drawItemTimeResidual(st, C, D, P, "plain")
        P["ζ"] = drawSubjSpeed(st, C, D, P, "latentqr" if qr else "latent")
    else:
        raise ValueError(model)
    return P

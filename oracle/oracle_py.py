"""ctypes binding of the CPU oracle (oracle/_build/liberirt_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liberirt_oracle.so")

MODELS = {"MlIrt": 0, "RtIrt": 1, "RtIrtNull": 2, "RtIrtCross": 3, "RtIrtCrossQr": 4, "RtIrtLatent": 5,
          "RtIrtLatentQr": 6, "RtIrtQuantile": 6}
COMPAT_BETA_PRIOR_DIAG = 1
COMPAT_LATENTQR_SCALE_ELEMENTWISE = 2


class Cfg(C.Structure):
    _fields_ = [("model", C.c_int32), ("nSubj", C.c_int32), ("nItem", C.c_int32), ("nFeat", C.c_int32),
                ("qRt", C.c_double), ("intercept", C.c_int32), ("onepl", C.c_int32), ("cov2one", C.c_int32),
                ("compat", C.c_int32), ("seed", C.c_uint64), ("chain", C.c_uint32), ("nthreads", C.c_int32)]


class State(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in
                ("omega", "theta", "a", "b", "zeta", "lambda_", "sigma2", "nu", "beta", "rho", "Sigma")]


def build(force=False):
    """Compile the oracle if the shared object is missing (gcc is in the image)."""
    srcs = [os.path.join(_HERE, f) for f in ("rng.c", "pg.c", "sampler.c", "gen.c", "oracle.h", "rng.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_sample.restype = C.c_int
        L.orc_sample.argtypes = [C.POINTER(Cfg), dp, dp, dp, C.POINTER(State), C.c_int64, C.c_int64, dp, dp, dp, dp,
                                 C.c_int]
        L.orc_loglik.restype = C.c_double
        L.orc_loglik.argtypes = [C.POINTER(Cfg), dp, dp, dp, C.POINTER(State)]
        L.orc_qr_width.restype = C.c_int
        L.orc_qr_width.argtypes = [C.POINTER(Cfg)]
        L.orc_beta_len.restype = C.c_int
        L.orc_beta_len.argtypes = [C.POINTER(Cfg)]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_pg_grid.argtypes = [dp, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, dp,
                                  C.POINTER(C.c_int32)]
        L.orc_nu_person.argtypes = [dp, C.c_double, C.c_int64, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, dp]
        L.orc_inv_normal_tail.restype = C.c_double
        L.orc_inv_normal_tail.argtypes = [C.c_double]
        L.orc_variates.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int64, C.c_uint64, dp]
        L.orc_inv_wishart2.argtypes = [C.c_double, dp, C.c_uint64, C.c_uint32, dp]
        L.orc_generate_data.argtypes = [C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_int32, C.c_int32] + [dp] * 9
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def pg_grid(z, seed=1234, chain=0, sweep=1, row0=0, return_attempts=False):
    z = np.ascontiguousarray(z, dtype=np.float64)
    rows, cols = z.shape
    out = np.empty_like(z)
    att = np.empty(z.shape, dtype=np.int32)
    lib().orc_pg_grid(_dp(z), rows, cols, row0, seed, chain, sweep, _dp(out), att.ctypes.data_as(C.POINTER(C.c_int32)))
    return (out, att) if return_attempts else out


def nu_person(mu, lam, seed=1234, chain=0, sweep=1, row0=0):
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    out = np.empty_like(mu)
    lib().orc_nu_person(_dp(mu), lam, mu.size, row0, seed, chain, sweep, _dp(out))
    return out


def variates(kind, p1, p2, n, seed=1234):
    out = np.empty(n, dtype=np.float64)
    lib().orc_variates({"normal": 0, "tnorm": 1, "gamma": 2, "invgamma": 3}[kind], p1, p2, n, seed, _dp(out))
    return out


def inv_wishart2(df, Psi, seed, sweep):
    Psi = np.asfortranarray(Psi, dtype=np.float64)
    out = np.empty((2, 2), dtype=np.float64, order="F")
    lib().orc_inv_wishart2(df, _dp(Psi), seed, sweep, _dp(out))
    return out


def inv_normal_tail(y):
    return lib().orc_inv_normal_tail(float(y))


def make_cfg(model, N, J, F, qRt=0.5, intercept=False, itemtype="2pl", cov2one=None, compat=0, seed=1234, chain=0,
             nthreads=1):
    mid = MODELS[model] if isinstance(model, str) else int(model)
    if cov2one is None:
        cov2one = mid not in (5, 6)  # defaults: GibbsRtIrtLatent.pl.jl:168,271 use cov2one=false
    return Cfg(mid, N, J, F, qRt, int(intercept), int(itemtype == "1pl"), int(cov2one), compat, seed, chain, nthreads)


class OracleState:
    """Owns float64 arrays for every InputPara field; `.c` is the ctypes view."""

    FIELDS = ("omega", "theta", "a", "b", "zeta", "lambda_", "sigma2", "nu", "beta", "rho", "Sigma")

    def __init__(self, cfg, init):
        N, J, F, m = cfg.nSubj, cfg.nItem, cfg.nFeat, cfg.model
        nb = lib().orc_beta_len(C.byref(cfg))
        shapes = {"omega": N * J, "theta": N, "a": J, "b": J, "zeta": N, "lambda_": J, "sigma2": J,
                  "nu": N * J if m == 4 else N, "beta": max(nb, 2 * (F + 1)), "rho": J, "Sigma": 4}
        self.arr = {}
        for k, n in shapes.items():
            v = np.zeros(n, dtype=np.float64)
            src = init.get(k.rstrip("_"), init.get(k)) if init else None
            if src is not None:
                src = np.asarray(src, dtype=np.float64).ravel(order="F")
                v[: src.size] = src
            elif k in ("a", "sigma2"):
                v[:] = 1.0
            elif k == "Sigma":
                v[:] = [1, 0, 0, 1]
            elif k == "nu":
                v[:] = 1.0
            self.arr[k] = v
        self.c = State(*[_dp(self.arr[k]) for k in self.FIELDS])


def sample(cfg, Y, logT, X, init, n_sweeps, first_sweep=1, person_trace=True, qr_skip_nu=False, want_ll=True):
    """Run the oracle.  Y/logT: (N,J), X: (N,F).  Returns dict with final state and traces
    ra/rt/qr/ll as (n_sweeps, width) arrays (row per sweep)."""
    L = lib()
    N, J = cfg.nSubj, cfg.nItem
    Yf = np.asfortranarray(Y, dtype=np.float64)
    Tf = np.asfortranarray(logT, dtype=np.float64) if logT is not None else None
    Xf = np.asfortranarray(X, dtype=np.float64) if X is not None and cfg.nFeat > 0 else None
    st = OracleState(cfg, init)
    W = N + 2 * J
    qw = L.orc_qr_width(C.byref(cfg))
    if qr_skip_nu:
        qw -= {6: N, 4: N * J}.get(cfg.model, 0)
    ra = np.empty((n_sweeps, W)) if person_trace else None
    rt = np.empty((n_sweeps, W)) if (person_trace and cfg.model != 0) else None
    qr = np.empty((n_sweeps, qw))
    ll = np.empty(n_sweeps) if want_ll else None
    rc = L.orc_sample(C.byref(cfg), _dp(Yf), _dp(Tf), _dp(Xf), C.byref(st.c), first_sweep, n_sweeps, _dp(ra), _dp(rt),
                      _dp(qr), _dp(ll), int(qr_skip_nu))
    if rc != 0:
        raise RuntimeError(f"oracle failed rc={rc}")
    out = {k.rstrip("_"): v for k, v in st.arr.items()}
    out["omega"] = out["omega"].reshape((N, J), order="F")
    if cfg.model == 4:
        out["nu"] = out["nu"].reshape((N, J), order="F")
    out.update(ra=ra, rt=rt, qr=qr, ll=ll)
    return out


def loglik(cfg, Y, logT, X, state):
    Yf = np.asfortranarray(Y, dtype=np.float64)
    Tf = np.asfortranarray(logT, dtype=np.float64) if logT is not None else None
    Xf = np.asfortranarray(X, dtype=np.float64) if X is not None and cfg.nFeat > 0 else None
    st = OracleState(cfg, state)
    return lib().orc_loglik(C.byref(cfg), _dp(Yf), _dp(Tf), _dp(Xf), C.byref(st.c))


ERROR_TYPES = {"tnorm": 0, "unit": 1, "norm": 2, "tail": 3, "skew": 4}


def generate_data(N, J, theta, a, b, zeta=None, lambda_=None, sigma2=None, rho=None, error="unit", seed=1234, person_offset=0):
    """CPU restatement of erirt_generate_data (oracle/gen.c): returns (Y, logT) column-major float64; logT is None without zeta."""
    def vec(v, n):
        if v is None:
            return None
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
        assert v.size == n
        return v
    has_rt = zeta is not None
    th, ze = vec(theta, N), vec(zeta, N)
    av, bv, lv, sv, rv = (vec(v, J) for v in (a, b, lambda_, sigma2, rho))
    Y = np.empty((N, J), dtype=np.float64, order="F")
    T = np.empty((N, J), dtype=np.float64, order="F") if has_rt else None
    p = lambda v: _dp(v) if v is not None else None  # noqa: E731
    lib().orc_generate_data(N, J, person_offset, seed, int(has_rt), ERROR_TYPES[error], p(th), p(ze), p(av), p(bv), p(lv), p(sv), p(rv), p(Y), p(T))
    return Y, T

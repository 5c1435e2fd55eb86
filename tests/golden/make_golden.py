#!/usr/bin/env python
"""Regenerate tests/golden/oracle_v2.npz: outputs of the CPU oracle on small seeded problems.

The reference (Julia) cannot run in this container or on the GPU box and ships no golden vectors, so these fixtures are NOT
reference-pinned ("parity unpinned").  They pin the *stream layout and formulas of the oracle itself*: any later
change to the Philox counter layout, the PG attempt order or a conditional shows up as a diff here, on the CPU suite (oracle
vs fixture) and on the GPU suite (f64 engine vs fixture).   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import MODELS, make_problem, run_oracle  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

GOLDEN = os.path.join(HERE, "oracle_v2.npz")
N, J, F, SWEEPS, SEED = 96, 9, 2, 3, 41


def z_grid():
    rng = np.random.default_rng(17)
    z = rng.normal(0, 1.8, (40, 6))
    z.ravel()[:10] = [0.0, 3.124, 3.126, -4.0, 8.0, 1e-4, -25.0, 0.64, 15.99, 16.01]
    return z


def compute():
    out = {}
    z = z_grid()
    out["pg_z"] = z
    om, att = O.pg_grid(z, seed=5, chain=2, sweep=3, row0=10, return_attempts=True)
    out["pg_omega"], out["pg_attempts"] = om, att
    for model in MODELS:
        pb = make_problem(model, N, J, F, seed=SEED)
        ref = run_oracle(O, pb, SWEEPS)
        for k in ("ra", "rt", "qr", "ll"):
            if k in ref and ref[k] is not None:
                arr = np.asarray(ref[k])
                if model == "RtIrtCrossQr" and k == "qr":
                    arr = arr[:, : J + 4]  # the N*J block of cell-level nu is not traced by the engine
                out[f"{model}_{k}"] = arr
    # device-side data generator (erirt_generate_data / oracle/gen.c): responses and log-times of the five error laws, a shard offset
    th, ze, a, b, lam, s2, rho = gen_inputs()
    for err in ("tnorm", "unit", "norm", "tail", "skew"):
        Y, T = O.generate_data(GEN_N, GEN_J, th, a, b, ze, lam, s2, rho if err in ("norm", "tail", "skew") else None, error=err, seed=GEN_SEED,
                               person_offset=GEN_OFFSET)
        out[f"gen_{err}_Y"], out[f"gen_{err}_logT"] = Y, T
    return out


GEN_N, GEN_J, GEN_SEED, GEN_OFFSET = 48, 7, 2024, 1000


def gen_inputs():
    rng = np.random.default_rng(23)
    return (rng.normal(size=GEN_N), 0.4 * rng.normal(size=GEN_N), rng.uniform(0.7, 1.4, GEN_J), rng.normal(0, 0.5, GEN_J),
            rng.uniform(2.5, 3.5, GEN_J), rng.uniform(0.2, 0.4, GEN_J), rng.normal(0, 0.2, GEN_J))


if __name__ == "__main__":
    O.build()
    np.savez_compressed(GOLDEN, **compute())
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")

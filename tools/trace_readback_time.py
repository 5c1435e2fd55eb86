"""Read-back time of the full Post.ra / Post.rt / Post.qr arrays at BASELINE configs[1] (GibbsRtIrt 10k x 30, nChain=3) for a given nIter."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import erirt_b200 as E
from helpers import make_problem
n_iter = int(os.environ.get("NITER", "1000"))
pb = make_problem("RtIrt", 10_000, 30, 3, seed=1)
eng = E.Engine("RtIrt", 10_000, 30, 3, n_iter=n_iter, n_chain=3, dtype="f32", person_trace=True)
eng.set_data(pb["Y"].astype(bool), pb["logT"], pb["X"])
i = pb["init"]
eng.set_state(theta=i["theta"], zeta=i["zeta"], beta=i["beta"][:8])
t0 = time.perf_counter(); eng.sample(3 * n_iter); t1 = time.perf_counter()
ra = eng.get_trace("ra"); rt = eng.get_trace("rt"); qr = eng.get_trace("qr"); ll = eng.get_trace("logLike"); t2 = time.perf_counter()
assert np.all(np.isfinite(ra)) and ra.shape == (n_iter, 10_060, 3)
print(f"nIter={n_iter} x 3 chains: sample {1e3*(t1-t0):.1f} ms ({3*n_iter/(t1-t0):.0f} sweeps/s), read-back of {(ra.nbytes+rt.nbytes+qr.nbytes)/1e6:.0f} MB {1e3*(t2-t1):.1f} ms")

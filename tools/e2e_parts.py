"""End-to-end parts of one sample! through the C ABI from pinned host buffers (C5), repeated in one process.  ERIRT_POOL_KEEP_MB=0 disables the pool retention."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import erirt_b200 as E, bench
from erirt_b200._lib import check
tp = bench.true_params(); dev = torch.device("cuda", 0)
dY, dT, dX, off, n = bench.gen_shard_torch(tp, 0, 1, dev)
hY = torch.empty(dY.shape, dtype=torch.uint8, pin_memory=True).copy_(dY)
hT = torch.empty(dT.shape, dtype=torch.float64, pin_memory=True).copy_(dT)
hX = torch.empty(dX.shape, dtype=torch.float64, pin_memory=True).copy_(dX)
del dY, dT, dX; torch.cuda.empty_cache(); torch.cuda.synchronize()
th, ze, be = bench.init_state(off, n)
hM = torch.empty((3, 2, n), dtype=torch.float64, pin_memory=True).numpy()
for rep in range(int(os.environ.get("REPS", "4"))):
    K = 200
    t = [time.perf_counter()]
    eng = E.Engine("RtIrtQuantile", n, bench.N_ITEM, bench.N_FEAT, n_iter=K + 31, n_chain=1, n_burnin=0, q_rt=bench.Q_RT, cov2one=False, dtype="f32", seed=1, person_trace=False, use_graph=True)
    t.append(time.perf_counter())
    if os.environ.get("GEN"):  # data generated on the device instead of uploaded
        eng.generate_data(th, tp.a, tp.b, ze, tp.lambda_, None, None, hX.numpy().T, error="unit", seed=5)
    else:
        check(eng.lib.erirt_set_data_y8(eng.h, hY.data_ptr(), n, hT.data_ptr(), n, hX.data_ptr(), n))
    t.append(time.perf_counter())
    eng.set_state(theta=th, zeta=ze, beta=be); t.append(time.perf_counter())
    eng.sample(K); t.append(time.perf_counter())
    a = eng.get_trace("ra", n, 200); b = eng.get_trace("rt", n, 200); c = eng.get_trace("qr", 0, 9); d = eng.get_trace("logLike"); t.append(time.perf_counter())
    m = [eng.get_moments(f, out=(hM[i, 0], hM[i, 1])) for i, f in enumerate(("theta", "zeta", "nu"))]; t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    names = ["create", "set_data", "set_state", "sample", "get_trace", "get_moments", "close"]
    print(" ".join(f"{nm}={1e3*(t[i+1]-t[i]):.1f}ms" for i, nm in enumerate(names)), "total", f"{1e3*(t[-2]-t[0]):.1f}ms")

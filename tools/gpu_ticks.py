"""Per-phase clock64() attribution of the person kernel (diagnostic build diag_tick.so, see profiles/)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ERIRT_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diag_tick.so")
import numpy as np, torch
import erirt_b200 as E
import bench
names = ["issue+person part1+logT sums", "mbar wait (TMA)", "row sums", "barrier 1", "person part 2 + u rows", "barrier 2", "PG main pass",
         "queue push", "barrier 3", "drain", "barrier 4", "statistics pass", "flush + Gram", "fence+barrier 5+store issue", "-", "-"]
tp = bench.true_params()
dev = torch.device("cuda", 0)
dY, dT, dX, off, n = bench.gen_shard_torch(tp, 0, 1, dev)
th, ze, be = bench.init_state(off, n)
eng = E.Engine("RtIrtQuantile", n, bench.N_ITEM, bench.N_FEAT, n_iter=40, n_chain=1, n_burnin=0, q_rt=bench.Q_RT, cov2one=False,
               dtype="f32", seed=1, person_trace=False, use_graph=True)
eng.set_data_device(dY.data_ptr(), n, dT.data_ptr(), n, dX.data_ptr(), n)
eng.set_state(theta=th, zeta=ze, beta=be)
eng.sample(5)
L = eng.lib
buf = (ctypes.c_ulonglong * 16)()
L.erirt_diag_ticks(buf, 1)
K = 20
eng.sample(K)
L.erirt_diag_ticks(buf, 0)
t = np.array(list(buf), dtype=np.float64)
st = eng.stats()
print("ms per sweep", st["last_sample_ms"] / K)
warps = int(os.environ.get("ERIRT_TICK_CTAS", "444")) * 4
print("per-warp mean cycles per sweep and share:")
tot = t.sum()
for nm, v in zip(names, t):
    print(f"  {nm:28s} {v / warps / K:12.0f} cyc  {100 * v / tot:5.1f} %")
print("sum per warp per sweep", tot / warps / K, "cycles =", tot / warps / K / 1.965e6, "ms")
gb = (ctypes.c_longlong * 16)()
L.erirt_diag_gticks(gb)
print("global_draw_kernel, cycles since kernel start (last sweep): staging %d, loglik %d, structural lane %d, item lane 32 %d, barrier %d, end %d, raw variates %d | structural (Latent*): XX built %d, solve done %d, ss %d, scale %d, gamma %d"
      % tuple(gb[i] for i in range(12)))

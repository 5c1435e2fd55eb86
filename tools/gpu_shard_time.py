"""ms per sweep of the production library on one GPU for the benchmark problem (C5) and for the shard one of eight GPUs holds (C5s8, no
exchange: the local part of a sharded sweep).  usage: python tools/gpu_shard_time.py [C5s8 C5 ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import erirt_b200 as E
import bench

dev = torch.device("cuda", 0)
tp = bench.true_params()
for which in (sys.argv[1:] or ["C5s8", "C5"]):
    world = 8 if which == "C5s8" else 1
    dY, dT, dX, off, n = bench.gen_shard_torch(tp, 0, world, dev)
    th, ze, be = bench.init_state(off, n)
    eng = E.Engine("RtIrtQuantile", n, bench.N_ITEM, bench.N_FEAT, n_iter=1200, n_chain=1, n_burnin=0, q_rt=bench.Q_RT, cov2one=False,
                   dtype="f32", seed=1, person_trace=False, use_graph=True)
    eng.set_data_device(dY.data_ptr(), n, dT.data_ptr(), n, dX.data_ptr(), n)
    eng.set_state(theta=th, zeta=ze, beta=be)
    eng.sample(40)
    res = []
    for rep in range(3):
        K = 320 if world == 8 else 100
        eng.sample(K)
        res.append(eng.stats()["last_sample_ms"] / K * 1e3)
    print(which, "us per sweep", " ".join(f"{r:.2f}" for r in res), flush=True)
    eng.close()
    del dY, dT, dX

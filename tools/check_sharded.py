"""Sharding invariance on real GPUs: a chain whose persons are sharded over WORLD_SIZE GPUs (item statistics exchanged every
sweep, by the fused peer-memory all-reduce or by ncclAllReduce) must reproduce the single-GPU chain up to f64 summation
order, because the Philox counters use global person ids.   torchrun --nproc-per-node 2 tools/check_sharded.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import erirt_b200 as E
    from erirt_b200 import distributed as D
    from helpers import make_problem
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for model, dtype, tol, mode in (("RtIrtLatentQr", "f64", 1e-9, "peer"), ("RtIrtLatentQr", "f64", 1e-9, "nccl"), ("RtIrt", "f64", 1e-9, "peer"),
                                    ("RtIrtNull", "f32", 2e-4, "peer"), ("MlIrt", "f64", 1e-9, "peer"), ("RtIrtCross", "f64", 1e-9, "peer"),
                                    ("RtIrtCrossQr", "f64", 1e-7, "nccl")):
        # CrossQr amplifies rounding differences by ~30x per sweep (weights 1/nu_ij, nu clamped to [1e-10, 1e10], Draw.pl.jl:318;
        # the single-GPU engine itself is 7e-12 from the oracle after 4 sweeps), so it is compared over 3 sweeps
        N, J, F, ns = 5003, 21, 3, (3 if model == "RtIrtCrossQr" else 10)
        pb = make_problem(model, N, J, F, seed=31)
        shard, cnt = D.make_shard(N, nccl=(mode == "nccl"))  # peer mode runs without any NCCL communicator
        off = shard[3]
        cov2one = model not in ("RtIrtLatent", "RtIrtLatentQr")

        def run(n, offset, sh, dev):
            eng = E.Engine(model, n, J, F, n_iter=ns, n_chain=1, n_burnin=0, q_rt=pb["q"], cov2one=cov2one, dtype=dtype, seed=5,
                           person_trace=True, device=dev, use_graph=True, n_subj_total=N, subj_offset=offset)
            if sh is not None:
                eng.comm_init(sh[0], sh[1], sh[2])
                if mode == "peer":
                    D.attach_peers(eng)  # fused one-shot exchange over peer memory instead of ncclAllReduce
            sl = slice(offset, offset + n)
            eng.set_data(pb["Y"][sl], None if model == "MlIrt" else pb["logT"][sl], pb["X"][sl])
            i = pb["init"]
            st = dict(theta=i["theta"][sl], a=i["a"], b=i["b"])
            if model != "MlIrt":
                st.update(zeta=i["zeta"][sl], lambda_=i["lambda_"], sigma2=i["sigma2"], Sigma=i["Sigma"])
            if "Cross" in model:
                st["rho"] = i["rho"]
            if pb["nb"]:
                st["beta"] = i["beta"][: pb["nb"]]
            eng.set_state(**st)
            eng.sample(ns)
            return eng

        sharded = run(cnt, off, shard, local)
        items_s = sharded.get_trace("ra", cnt, 2 * J)[:, :, 0]
        theta_s = D.gather_person_vector(sharded.get_trace("ra", 0, cnt)[ns - 1, :, 0], N)
        ll_s = sharded.get_trace("logLike")[:, 0, 0]
        if rank == 0:
            whole = run(N, 0, None, local)
            items_w = whole.get_trace("ra", N, 2 * J)[:, :, 0]
            theta_w = whole.get_trace("ra", 0, N)[ns - 1, :, 0]
            ll_w = whole.get_trace("logLike")[:, 0, 0]
            e1 = np.max(np.abs(items_s - items_w) / (np.abs(items_w) + 1e-3))
            e2 = np.quantile(np.abs(theta_s - theta_w) / (np.abs(theta_w) + 1e-1), 0.995)
            e3 = np.max(np.abs(ll_s - ll_w) / np.abs(ll_w))
            good = e1 < tol and e2 < tol and e3 < tol
            ok &= bool(good)
            print(f"{model} {dtype} world={world} exchange={mode}: items {e1:.2e} theta(q99.5) {e2:.2e} loglik {e3:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        D.close_sharded(sharded)
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDING_OK" if ok else "SHARDING_FAIL")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck), one tool per gpurun call, and for the repo's own
substitute where compute-sanitizer is closed (profiles/r02_sanitizer_unavailable.txt):

    ERIRT_CHECKS_BUILD=1 python tools/make_tick_build.py          # diag_checks.so: device-side index asserts (-DERIRT_CHECKS)
    ERIRT_B200_LIB=$PWD/diag_checks.so ERIRT_GUARDS=1 python tools/sanitize_run.py    # + 256-byte guard zones after every device buffer


    compute-sanitizer --tool racecheck python tools/sanitize_run.py            # single GPU
    compute-sanitizer --tool memcheck --target-processes all \
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sanitize_run.py   # peer exchange

Cases: the f32 fast kernel (TPP 1, 2, 4), the generic kernel (f64, Cross family), and -- with the persistent grid capped by
ERIRT_MAX_GRID -- a CTA that owns > 16 tiles, so that the cross-tile register accumulators are folded through the staging area that
aliases the logT tile (the situation of the 1M x 100 benchmark).  Under torchrun the persons of every case are sharded over the
ranks and the statistics go through the fused peer-memory exchange of global_draw_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import erirt_b200 as E  # noqa: E402
from helpers import make_problem  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch
    import torch.distributed as dist
    from erirt_b200 import distributed as D
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# (model, N, J, F, dtype, max_grid)
cases = [("RtIrtLatentQr", 300, 13, 2, "f32", 0), ("RtIrtLatentQr", 200, 100, 3, "f32", 0), ("RtIrt", 150, 150, 1, "f32", 0), ("MlIrt", 130, 21, 0, "f32", 0),
         ("RtIrtLatentQr", 200, 21, 2, "f64", 0), ("RtIrtCrossQr", 150, 13, 0, "f32", 0), ("RtIrtCross", 129, 7, 0, "f64", 0),
         ("RtIrtNull", 1000, 40, 0, "f32", 0), ("RtIrtLatent", 257, 33, 4, "f32", 0), ("RtIrtCrossQr", 131, 9, 0, "f64", 0),
         ("RtIrt", 3000, 30, 3, "f32", 0), ("MlIrt", 1000, 15, 3, "f64", 0),
         ("RtIrtLatentQr", 64 * 2 * 18 * world + 5, 100, 3, "f32", 2)]  # 18 tiles per CTA on every rank: one 16-tile fold + the final one
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for model, N, J, F, dt, max_grid in cases:
    if max_grid:
        os.environ["ERIRT_MAX_GRID"] = str(max_grid)
    else:
        os.environ.pop("ERIRT_MAX_GRID", None)
    pb = make_problem(model, N, J, F, seed=3)
    off, n = (0, N) if world == 1 else D.shard_bounds(N, world, rank)
    cov2one = model not in ("RtIrtLatent", "RtIrtLatentQr")
    eng = E.Engine(model, n, J, F, n_iter=3, n_chain=1, n_burnin=0, q_rt=pb["q"], cov2one=cov2one, dtype=dt, seed=5, person_trace=True,
                   device=local, use_graph=False, n_subj_total=N, subj_offset=off)
    if world > 1:
        eng.comm_init(rank, world, None)
        D.attach_peers(eng)
    sl = slice(off, off + n)
    eng.set_data(pb["Y"][sl], None if model == "MlIrt" else pb["logT"][sl], pb["X"][sl] if F > 0 else None)
    i = pb["init"]
    st = dict(theta=i["theta"][sl], a=i["a"], b=i["b"])
    if model != "MlIrt":
        st.update(zeta=i["zeta"][sl], lambda_=i["lambda_"], sigma2=i["sigma2"], Sigma=i["Sigma"])
    if pb["nb"]:
        st["beta"] = i["beta"][: pb["nb"]]
    if "Cross" in model:
        st["rho"] = i["rho"]
    eng.set_state(**st)
    eng.sample(3)
    a = eng.get_trace("ra", n, 2 * J)[:3, :, 0]
    assert np.all(np.isfinite(a)), (model, dt)
    eng.get_moments("theta")
    eng.loglik_current() if hasattr(eng, "loglik_current") else None
    guards = eng.check_guards()
    assert guards in (0, -1), f"{guards} guard bytes overwritten ({model} {dt})"
    if world > 1:
        D.close_sharded(eng)
    else:
        eng.close()
    if rank == 0:
        print("ok", model, N, J, F, dt, "max_grid", max_grid, "world", world, "guard bytes overwritten:", guards, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

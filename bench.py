#!/usr/bin/env python
"""Benchmark of the Gibbs sweep path (BASELINE.json metric: "Gibbs sweeps/sec & ESS/sec, nSubj=1M nItem=100").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one Gibbs sweep of GibbsRtIrtQuantile (the LatentQr sampler, q = 0.85, F = 3 covariates) over the
synthetic BASELINE config 5: nSubj = 1,000,000 x nItem = 100, data from setDataRtIrtLatent(type="skew")
(src/SimTools.jl:300-343), f32 storage / per-cell arithmetic, f64 statistics, log-likelihood and running
person moments ON (the reference computes logLike and stores the trace every sweep).  With N > 1 the persons of
the one chain are sharded over the N GPUs of the box and the item statistics are all-reduced every sweep (one-shot
exchange over NVLink peer memory fused into the global draw kernel; ERIRT_EXCHANGE=nccl selects ncclAllReduce), so
the total work is fixed: "scaling": "strong".

Keys (see DESIGN.md "Measurement"):
  value       sweeps/s, data resident in HBM, CUDA events on the library's stream, max over ranks
  e2e         sweeps/s through the public C-ABI calls with HOST (pinned) buffers: erirt_create + erirt_set_data
              (H2D + ingest) + erirt_set_state + K sweeps + erirt_get_trace/erirt_get_moments (D2H), all timed
  roofline    person-sweep kernel: algorithmic HBM bytes per launch / its mean CUDA-event duration, against the
              measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a C restatement of the reference; Julia is not installed) on one host core, on a
              bounded sample of the persons, extrapolated linearly in nSubj
--impl reference times that same CPU restatement with all host threads (OpenMP over persons/items).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SUBJ, N_ITEM, N_FEAT, Q_RT = 1_000_000, 100, 3, 0.85
N_BLOCKS = 8  # data is generated in 8 seeded person blocks so that every GPU count sees the same data set
SEED = 1234
METRIC = "gibbs_sweeps_per_sec"
UNIT = "sweeps/s"
WORKLOAD = "GibbsRtIrtQuantile(LatentQr) synthetic nSubj=1M nItem=100 nFeat=3 qRt=0.85 (BASELINE config 5)"


def true_params():
    import erirt_b200 as E
    Cond = E.setCond(nSubj=N_SUBJ, nItem=N_ITEM, nFeat=N_FEAT, qRt=Q_RT, qRa=Q_RT)
    return E.setTrueParaRtIrtLatent(Cond, rng=SEED)


def gen_block_torch(tp, block, n, device):
    """setDataRtIrtLatent(type="skew") for one block of persons, on the device, Julia column-major float64:
    returns Y^T, logT^T as (J, n) and X^T as (F, n) contiguous tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED * 1000 + block)
    f64 = dict(dtype=torch.float64, device=device)
    a = torch.as_tensor(tp.a, **f64)[:, None]
    b = torch.as_tensor(tp.b, **f64)[:, None]
    lam = torch.as_tensor(tp.lambda_, **f64)[:, None]
    beta = torch.as_tensor(tp.beta, **f64)
    theta = torch.randn(n, generator=g, **f64)
    X = torch.randn(N_FEAT, n, generator=g, **f64)
    # Gamma(0.5, 1) - 1 errors: Gamma(1/2) = Z^2/2
    err = 0.5 * torch.randn(n, generator=g, **f64) ** 2 - 1.0
    zeta = (beta[:N_FEAT, None] * X).sum(0) + beta[N_FEAT] * theta + err
    eta = a * (theta[None, :] - b)
    Y = (torch.rand(N_ITEM, n, generator=g, **f64) < torch.sigmoid(eta)).to(torch.float64)
    logT = lam - zeta[None, :] + torch.randn(N_ITEM, n, generator=g, **f64)
    return Y.contiguous(), logT.contiguous(), X.contiguous()


def gen_shard_torch(tp, rank, world, device):
    import torch
    per = N_SUBJ // N_BLOCKS
    blocks = range(rank * N_BLOCKS // world, (rank + 1) * N_BLOCKS // world)
    parts = [gen_block_torch(tp, bk, per, device) for bk in blocks]
    Y = torch.cat([p[0] for p in parts], dim=1).contiguous()
    T = torch.cat([p[1] for p in parts], dim=1).contiguous()
    X = torch.cat([p[2] for p in parts], dim=1).contiguous()
    return Y, T, X, blocks[0] * per, per * len(blocks)


def init_state(offset, n):
    rng = np.random.default_rng(SEED + 7)
    theta = rng.standard_normal(N_SUBJ)[offset:offset + n]
    zeta = rng.standard_normal(N_SUBJ)[offset:offset + n]
    beta = rng.standard_normal(N_FEAT + 2)
    return theta, zeta, beta


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every few milliseconds
    from a thread (ctypes releases the GIL while the library runs), so that even a 15 ms timed region gets samples;
    nvidia-smi -lms 100 is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons = None, False, [], [], set()

    def _nvml_loop(self, nv, handle):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(handle))
                for nm, bit in bits.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            if self.stop_flag:
                break
            time.sleep(0.02)  # 50 Hz: NVML calls take driver locks, a tighter loop was seen to perturb 8-GPU runs

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))]
            self.stop_flag = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)) if self.mx else None,
                           reasons=sorted(self.reasons), samples=len(self.sm), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic():
    """dram bytes per launch of the person kernel from the last committed ncu --set full capture (or None)."""
    p = os.path.join(ROOT, "profiles", "person_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def ess_per_sweep(eng, first, n, qw):
    """min / median bulk ESS per sweep over the traced item + structural parameters of sweeps [first, first+n)."""
    from erirt_b200.diagnostics import ess_rhat
    N, J = eng.N, eng.J
    cols = [eng.get_trace("ra", N, 2 * J)[first:first + n, :, 0], eng.get_trace("rt", N, 2 * J)[first:first + n, :, 0],
            eng.get_trace("qr", 0, qw)[first:first + n, :, 0]]
    ess = []
    for arr in cols:
        for c in range(arr.shape[1]):
            x = arr[:, c]
            if np.ptp(x) > 0:
                ess.append(ess_rhat(x)[0])
    ess = np.asarray(ess)
    return float(np.nanmin(ess) / n), float(np.nanmedian(ess) / n)


def cpu_oracle_rate(tp, n_sample, n_sweeps, nthreads, warm=1):
    """sweeps/s of the CPU restatement on the first n_sample persons of block 0 (host-generated, same
    distributions), extrapolated linearly to N_SUBJ persons."""
    import erirt_b200 as E
    from oracle import oracle_py as O
    Cond = E.setCond(nSubj=n_sample, nItem=N_ITEM, nFeat=N_FEAT, qRt=Q_RT)
    import copy
    tpc = copy.deepcopy(tp)
    D = E.setDataRtIrtLatent(Cond, tpc, type="skew", rng=SEED)
    cfg = O.make_cfg("RtIrtLatentQr", n_sample, N_ITEM, N_FEAT, qRt=Q_RT, seed=SEED, nthreads=nthreads)
    theta, zeta, beta = init_state(0, n_sample)
    init = dict(theta=theta, zeta=zeta, beta=beta, a=np.ones(N_ITEM), b=np.zeros(N_ITEM), lambda_=np.zeros(N_ITEM),
                sigma2=np.ones(N_ITEM), Sigma=np.eye(2).ravel())
    if warm:
        O.sample(cfg, D.Y, D.logT, D.X, init, warm, person_trace=False, qr_skip_nu=True)
    t0 = time.perf_counter()
    O.sample(cfg, D.Y, D.logT, D.X, init, n_sweeps, person_trace=False, qr_skip_nu=True)
    dt = time.perf_counter() - t0
    return (n_sweeps / dt) * (n_sample / N_SUBJ), dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Julia package cannot
    run here (no julia binary), so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    O.build()
    tp = true_params()
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    # calibrate so that the whole run stays within ~2 minutes
    rate, _ = cpu_oracle_rate(tp, 2000, 2, cores, warm=1)
    per_sweep_full = 1.0 / rate
    budget = 100.0 / max(total, 1)
    n_sample = int(min(N_SUBJ, max(2000, N_SUBJ * budget / per_sweep_full)))
    n_sample -= n_sample % 8
    value, dt = cpu_oracle_rate(tp, n_sample, args.steps, cores, warm=args.warmup)
    sample = (f"oracle (C restatement of Draw.pl.jl, not Julia) with {cores} OpenMP threads on the first {n_sample} persons x {N_ITEM} items, "
              f"{args.steps} sweeps in {dt:.1f} s, sweeps/s scaled by {n_sample}/{N_SUBJ} (cost is linear in persons)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="erirt_b200", choices=["erirt_b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-y", choices=["u8", "f64"], default="u8", help="host type of Y in the end-to-end pass (Julia Matrix{Bool} or Float64)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--short", action="store_true", help="profiling run: timed sweeps only, no e2e / cpu / roofline passes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import erirt_b200 as E
    from erirt_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert N_BLOCKS % world == 0, "GPU count must divide 8"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        shard, _ = D.make_shard(N_SUBJ, nccl=False)

    tp = true_params()
    dY, dT, dX, offset, n_local = gen_shard_torch(tp, rank, world, dev)
    theta0, zeta0, beta0 = init_state(offset, n_local)
    K, W = args.steps, args.warmup
    n_iter = K + W + 8

    def make_engine(time_kernels=False, use_graph=True):
        eng = E.Engine("RtIrtQuantile", n_local, N_ITEM, N_FEAT, n_iter=n_iter, n_chain=1, n_burnin=0, q_rt=Q_RT,
                       cov2one=False, dtype=args.dtype, seed=SEED, person_trace=False, device=local, use_graph=use_graph,
                       n_subj_total=N_SUBJ, subj_offset=offset, time_kernels=time_kernels)
        if world > 1:
            peer = os.environ.get("ERIRT_EXCHANGE", "peer") == "peer"
            sh, _ = D.make_shard(N_SUBJ, nccl=not peer)  # NCCL mode: a fresh id per communicator
            eng.comm_init(sh[0], sh[1], sh[2])
            if peer:
                D.attach_peers(eng)  # one-shot exchange over NVLink peer memory fused into the global draw kernel; no NCCL at all
        return eng

    def close_engine(e):
        if world > 1:
            D.close_sharded(e)  # unmap the peers' exchange buffers on every rank before anybody frees its own
        else:
            e.close()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput ("value") ----------------
    eng = make_engine()
    eng.set_data_device(dY.data_ptr(), n_local, dT.data_ptr(), n_local, dX.data_ptr(), n_local)
    eng.set_state(theta=theta0, zeta=zeta0, beta=beta0)
    eng.sample(W)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    eng.sample(K)
    barrier()
    clk = clocks.stop()
    ms = max_over_ranks(eng.stats()["last_sample_ms"])
    st = eng.stats()
    value = K / (ms / 1000.0)
    qw = N_FEAT + 2 + 4
    ess_min = ess_med = None
    if rank == 0 and not args.short:
        try:
            ess_min, ess_med = ess_per_sweep(eng, W, K, qw)
        except Exception:
            pass
    close_engine(eng)

    # ---------------- roofline of the person kernel (plain launches bracketed by CUDA events) ----------------
    roofline = None
    if not args.short:
        engk = make_engine(time_kernels=True, use_graph=False)
        engk.set_data_device(dY.data_ptr(), n_local, dT.data_ptr(), n_local, dX.data_ptr(), n_local)
        engk.set_state(theta=theta0, zeta=zeta0, beta=beta0)
        engk.sample(min(W, 10))
        barrier()
        kk = min(K, 100)
        engk.sample(kk)
        sk = engk.stats()
        pk_ms = max_over_ranks(sk["person_kernel_ms"])
        peak, which = measured_peak()
        bytes_launch = sk["bytes_per_sweep"]
        achieved = bytes_launch / (pk_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "person_sweep_fast_kernel" if args.dtype == "f32" else "person_sweep_kernel", "achieved": achieved, "peak": peak, "peak_source": which,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(), "algorithmic_bytes_per_launch": bytes_launch,
                    "kernel_ms": pk_ms, "launches_timed": kk, "pg_deferred_frac": sk["pg_deferred_frac"]}
        close_engine(engk)

    # ---------------- end to end through the C ABI with host buffers ("e2e") ----------------
    e2e = None
    if not args.no_e2e and not args.short:
        # the host holds what the reference's InputData holds: Y as Julia's Matrix{Bool} (one byte per response, `rand.(BernoulliLogit…)`
        # src/SimTools.jl:165; --e2e-y f64 passes a Float64 Y instead), logT and X as Float64; all column-major, pinned
        y8 = args.e2e_y == "u8"
        hY = torch.empty(dY.shape, dtype=torch.uint8 if y8 else torch.float64, pin_memory=True).copy_(dY)
        hT = torch.empty(dT.shape, dtype=torch.float64, pin_memory=True).copy_(dT)
        hX = torch.empty(dX.shape, dtype=torch.float64, pin_memory=True).copy_(dX)
        hM = torch.empty((3, 2, n_local), dtype=torch.float64, pin_memory=True)  # Post.mean / SD of theta, zeta, nu land here
        hMn = hM.numpy()
        del dY, dT, dX
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        enge = make_engine()
        t1 = time.perf_counter()
        from erirt_b200._lib import check
        set_data = enge.lib.erirt_set_data_y8 if y8 else enge.lib.erirt_set_data
        check(set_data(enge.h, hY.data_ptr(), n_local, hT.data_ptr(), n_local, hX.data_ptr(), n_local))
        enge.set_state(theta=theta0, zeta=zeta0, beta=beta0)
        t2 = time.perf_counter()
        enge.sample(K)
        t3 = time.perf_counter()
        tr_a = enge.get_trace("ra", n_local, 2 * N_ITEM)
        tr_t = enge.get_trace("rt", n_local, 2 * N_ITEM)
        tr_q = enge.get_trace("qr", 0, qw)
        tr_l = enge.get_trace("logLike")
        moms = [enge.get_moments(f, out=(hMn[k, 0], hMn[k, 1])) for k, f in enumerate(("theta", "zeta", "nu"))]
        t4 = time.perf_counter()
        close_engine(enge)  # erirt_destroy is part of the call a user makes (the Julia shim destroys the handle inside sample!)
        t5 = time.perf_counter()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = hY.numel() * hY.element_size() + (hT.numel() + hX.numel()) * 8 + (2 * n_local + N_FEAT + 2) * 8
        d2h = K * (tr_a.shape[1] + tr_t.shape[1] + tr_q.shape[1] + 1) * 8 + 3 * 2 * n_local * 8
        assert np.all(np.isfinite(tr_l[:K])) and np.all(np.isfinite(moms[0][0]))
        e2e = {"value": K / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world / K), "d2h_bytes_per_step": int(d2h * world / K),
               "seconds": dt, "parts_ms_rank0": {"create": (t1 - t0) * 1e3, "set_data_and_state": (t2 - t1) * 1e3, "sample": (t3 - t2) * 1e3, "read_back": (t4 - t3) * 1e3, "destroy": (t5 - t4) * 1e3},
               "note": "erirt_create + " + ("erirt_set_data_y8 (pinned host: Y as Matrix{Bool} bytes, logT/X f64" if y8 else "erirt_set_data (pinned host f64")
                                     + "; chunked H2D + ingest) + erirt_set_state + K sweeps + "
                                     "erirt_get_trace/erirt_get_moments (D2H into pinned buffers) + erirt_destroy; bytes are totals of the call divided by K"}

    # ---------------- the same call with the N x J data generated on the device (informational: no N x J upload) ----------------
    e2e_gen = None
    if not args.no_e2e and not args.short:
        rng = np.random.default_rng(SEED + 11)
        th_all = rng.standard_normal(N_SUBJ)
        X_all = rng.standard_normal((N_SUBJ, N_FEAT))
        ze_all = np.column_stack([X_all, th_all]) @ tp.beta + (0.5 * rng.standard_normal(N_SUBJ) ** 2 - 1.0)  # setDataRtIrtLatent(type="skew")
        sl = slice(offset, offset + n_local)
        th_g, ze_g, X_g = th_all[sl].copy(), ze_all[sl].copy(), np.asfortranarray(X_all[sl])
        barrier()
        t0 = time.perf_counter()
        engg = make_engine()
        engg.generate_data(th_g, tp.a, tp.b, ze_g, tp.lambda_, None, None, X_g, error="unit", seed=SEED)
        engg.set_state(theta=theta0, zeta=zeta0, beta=beta0)
        engg.sample(K)
        tr_g = [engg.get_trace("ra", n_local, 2 * N_ITEM), engg.get_trace("rt", n_local, 2 * N_ITEM), engg.get_trace("qr", 0, qw), engg.get_trace("logLike")]
        moms_g = [engg.get_moments(f, out=(hMn[k, 0], hMn[k, 1])) for k, f in enumerate(("theta", "zeta", "nu"))]
        close_engine(engg)
        barrier()
        dtg = max_over_ranks(time.perf_counter() - t0)
        assert np.all(np.isfinite(tr_g[3][:K])) and np.all(np.isfinite(moms_g[0][0]))
        e2e_gen = {"value": K / dtg, "unit": UNIT, "seconds": dtg, "h2d_bytes_per_step": int((2 + N_FEAT) * n_local * 8 * world / K),
                   "note": "erirt_create + erirt_generate_data (person-level theta, zeta, X from the host; responses and log-times generated "
                           "on the device) + erirt_set_state + K sweeps + read-back + erirt_destroy"}

    # ---------------- CPU baseline beside it (rank 0, single GPU run only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.short:
        try:
            n_sample = 20000
            v, dtc = cpu_oracle_rate(tp, n_sample, 8, 1, warm=1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle (C restatement of Draw.pl.jl; the Julia reference cannot run here) on 1 core of {os.cpu_count()}, "
                             f"first {n_sample} persons x {N_ITEM} items, 8 sweeps in {dtc:.1f} s, scaled by {n_sample}/{N_SUBJ}"}
        except Exception as ex:  # the bench line must still be printed
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": (f"persons sharded over {world} GPUs, item statistics all-reduced per sweep by "
                                           + ("a one-shot exchange over NVLink peer memory fused into the global draw kernel"
                                              if os.environ.get("ERIRT_EXCHANGE", "peer") == "peer" else "ncclAllReduce")) if world > 1 else "single GPU",
                           "l2": "inputs (1.3 GB/sweep) larger than the 126 MB L2", "cuda_graph": True,
                           "loglik_and_moments": "on"},
                "clocks": clk, "gpu_launches": 2 * K, "e2e": e2e, "e2e_device_generated_data": e2e_gen, "roofline": roofline, "cpu_baseline": cpu,
                "ess_per_sweep": {"min": ess_min, "median": ess_med, "estimator": "rank-normalised bulk ESS over the K timed sweeps (a, b, lambda, sigma2, beta, Sigma)"},
                "ess_per_sec": {"min": None if ess_min is None else ess_min * value, "median": None if ess_med is None else ess_med * value},
                "bytes_per_sweep": st["bytes_per_sweep"] * world}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

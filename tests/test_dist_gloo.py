"""world_size-2 gloo test of the multi-process host logic (rendezvous on 127.0.0.1): broadcasting the NCCL id
bytes, shard bounds, and gathering person-level results.  The NCCL all-reduce itself needs GPUs (-m gpu / bench)."""
import os

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from erirt_b200 import distributed as D
    payload = bytes(range(128)) if rank == 0 else b"\0" * 128
    got = D.broadcast_bytes(payload, 0)
    n_total = 11
    off, cnt = D.shard_bounds(n_total, world, rank)
    local = np.arange(off, off + cnt, dtype=np.float64) * 2
    full = D.gather_person_vector(local, n_total)
    # the transport of the 64-byte CUDA IPC handles of the fused peer exchange (attach_peers): rank order, equal lengths
    handles = D.allgather_bytes(bytes([rank + 1]) * 64)
    # runSimulation deals replication r to rank (r-1) mod world; chains c to GPU c mod world
    reps = [r for r in range(1, 8) if (r - 1) % world == rank]
    # independent chains (BASELINE configs[3]): chain c runs on rank c mod world, traces gathered in chain order on every rank
    ran = []

    def run_chain(c):
        ran.append(c)
        return dict(ra=np.full((4, 3, 1), float(c)), logLike=np.full((4, 1, 1), 10.0 * c))
    chains = D.run_independent_chains(run_chain, 5)
    chains_ok = (ran == D.chain_assignment(5, world, rank) and chains["ra"].shape == (4, 3, 5) and chains["logLike"].shape == (4, 1, 5)
                 and chains["ra"][0, 0, :].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0] and chains["logLike"][1, 0, :].tolist() == [0.0, 10.0, 20.0, 30.0, 40.0])
    ret[rank] = (chains_ok and got == bytes(range(128)) and handles == b"\x01" * 64 + b"\x02" * 64 and D.make_shard(n_total, nccl=False)[0][2] is None,
                 full.tolist(), (off, cnt), reps, D.chain_assignment(5, world, rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29531, ret), nprocs=world, join=True)
    for r in range(world):
        ok, full = ret[r][0], ret[r][1]
        assert ok
        assert full == [2.0 * i for i in range(11)]
    assert ret[0][2] == (0, 6) and ret[1][2] == (6, 5)
    assert ret[0][3] == [1, 3, 5, 7] and ret[1][3] == [2, 4, 6]
    assert ret[0][4] == [0, 2, 4] and ret[1][4] == [1, 3]

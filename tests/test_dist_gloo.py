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
    ret[rank] = (got == bytes(range(128)), full.tolist(), (off, cnt))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_plumbing():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29531, ret), nprocs=world, join=True)
    for r in range(world):
        ok, full, _ = ret[r]
        assert ok
        assert full == [2.0 * i for i in range(11)]
    assert ret[0][2] == (0, 6) and ret[1][2] == (6, 5)

"""Multi-GPU plumbing, one process per GPU (torchrun): person sharding of one chain and independent chains.

torch.distributed is used only as the bootstrap (rendezvous, broadcasting the 128-byte NCCL id, all-gathering the
CUDA IPC handles of the exchange buffers, gathering results).  The per-sweep exchange of item statistics runs inside
the library: a one-shot all-reduce over NVLink peer memory fused into the global draw kernel (attach_peers), or an
ncclAllReduce on the library's stream when no peer buffers are attached (csrc/erirt_b200.cu, launch_global).  SURVEY.md 8e."""
import numpy as np


def shard_bounds(n_total, world, rank):
    """Contiguous person blocks, sizes differing by at most one: returns (offset, count)."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def chain_assignment(n_chains, world, rank):
    """Independent chains c -> GPU c mod world."""
    return [c for c in range(n_chains) if c % world == rank]


def broadcast_bytes(payload, src=0, group=None, nbytes=128):
    """Broadcast a fixed-size byte string from `src` with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def make_shard(n_total, group=None, nccl=True):
    """Returns the `shard` tuple accepted by api.sample / Engine.comm_init for this rank:
    (rank, world, nccl_unique_id, subj_offset, n_subj_total) plus the local person count.  nccl=False: no NCCL
    communicator (unique id None); the caller attaches the peer exchange (attach_peers)."""
    import torch.distributed as dist
    from .engine import nccl_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = None
    if nccl:
        uid = nccl_unique_id() if rank == 0 else b"\0" * 128
        uid = broadcast_bytes(uid, 0, group)
    offset, count = shard_bounds(n_total, world, rank)
    return (rank, world, uid, offset, n_total), count


def allgather_bytes(payload, group=None):
    """All-gather equal-length byte strings; returns their concatenation in rank order."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    outs = [torch.zeros_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, mine, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)


def attach_peers(engine, group=None):
    """Switch a person-sharded engine (after comm_init) to the fused peer-memory exchange: export this GPU's buffer,
    all-gather the IPC handles, map the peers.  One process per GPU on one node (NVLink / NVSwitch peers)."""
    engine.peer_attach(allgather_bytes(engine.peer_export(), group))


def close_sharded(engine, group=None):
    """Tear a person-sharded engine down in the order the peer mappings need: unmap on every rank, barrier, free."""
    import torch.distributed as dist
    engine.peer_detach()
    dist.barrier(group=group)
    engine.close()


def gather_person_vector(local, n_total, group=None):
    """All-gather the shards of a person-level vector (theta/zeta/nu means) back into one array of n_total."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    sizes = [shard_bounds(n_total, world, r)[1] for r in range(world)]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.float64, device=dev)
    buf[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=dev)
    outs = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])


def run_independent_chains(run_chain, n_chains, group=None):
    """Independent chains, SURVEY 8e / BASELINE configs[3]: chain c runs on rank c mod world (one GPU per rank, chains of a rank one
    after the other), no collective while sampling; the traces are gathered at the end.  `run_chain(c)` returns a dict
    name -> array whose LAST axis is the chain axis of length 1 (e.g. Post.ra[:, item columns, :] of a one-chain run); every rank
    gets dict name -> array with the chains concatenated in chain order along that axis."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = {c: run_chain(c) for c in chain_assignment(n_chains, world, rank)}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
        mine = {c: r for part in parts for c, r in part.items()}
    names = list(mine[0].keys())
    return {nm: np.concatenate([np.asarray(mine[c][nm]) for c in range(n_chains)], axis=-1) for nm in names}

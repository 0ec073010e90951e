"""Multi-GPU use of the library from Python: one process per GPU, one context per process.

The sharding itself lives in the CUDA library (csrc/comm.cu, csrc/transcript.h): after `join`,
`Participant.verify_distribution_shares` / `distribute_secret` on the group are collective calls --
every rank passes the whole box, the library takes participants rank, rank + N, ... (every rank gets
the same mix of short and long addition chains, hence equal work), runs the kernels on its slice,
exchanges the fixed-size transcript rows with ONE ncclAllGather on its own stream and hashes them in
`publickeys` order (participant.rs:238-245, 438-447).  This module only carries the NCCL unique id
from rank 0 to the other ranks, over whatever `torch.distributed` group the processes already share
(gloo or nccl): plumbing, not data path.
"""
from __future__ import annotations

from .lib import COMM_ID_BYTES, comm_unique_id


def shard_indices(rank: int, world: int, n_total: int):
    """0-based participant indices the library assigns to `rank` (positions are index + 1)."""
    return list(range(rank, n_total, world))


def rows_per_rank(n_total: int, world: int) -> int:
    return -(-n_total // world)


def gather_layout(rows_by_participant, world: int, row_bytes: int) -> bytes:
    """Arrange per-participant rows the way the all-gather delivers them: [rank][local row][row_bytes],
    zero padded to rows_per_rank (test helper for the host half of the sharded transcript)."""
    n_total = len(rows_by_participant)
    rpr = rows_per_rank(n_total, world)
    out = bytearray(world * rpr * row_bytes)
    for i, row in enumerate(rows_by_participant):
        off = ((i % world) * rpr + i // world) * row_bytes
        out[off:off + len(row)] = row
    return bytes(out)


def join(ctx, rank: int, world: int, dist=None):
    """Create the library-side communicator for `ctx` (a lib.Context).  `dist` is an initialised
    torch.distributed module (any backend); rank 0's NCCL unique id is broadcast through it."""
    if world == 1:
        return
    import torch
    if dist is None:
        import torch.distributed as dist
    uid = comm_unique_id() if rank == 0 else bytes(COMM_ID_BYTES)
    t = torch.frombuffer(bytearray(uid), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    ctx.comm_init(bytes(t.cpu().numpy().tobytes()), world, rank)

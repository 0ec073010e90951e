"""Participant sharding for the multi-GPU path (one process per GPU).

Every share is independent, so the participants of one DistributionSharesBox are dealt round
robin to the ranks (rank r owns indices r, r+N, r+2N, ...: every rank gets the same mix of small
and large positions, hence equal work).  Each rank produces fixed-width rows (X, a1, a2) for its
indices; one all-gather combines them and `interleave` restores `publickeys` order for the single
running SHA-256 transcript (participant.rs:238-245, 438-447 hash in that order)."""
from __future__ import annotations


def shard_indices(rank: int, world: int, n_total: int):
    """0-based participant indices owned by `rank` (positions are index + 1)."""
    return list(range(rank, n_total, world))


def interleave(gathered, world: int, kinds: int, n: int, width: int):
    """gathered: array-like of shape [world, kinds, n, width] (numpy or torch) as produced by
    all_gather_into_tensor over per-rank [kinds, n, width] blocks.  Returns `kinds` byte strings of
    n*world rows in participant order (row j of rank r is participant j*world + r)."""
    import numpy as np
    g = np.asarray(gathered).reshape(world, kinds, n, width)
    return [g[:, k].transpose(1, 0, 2).reshape(-1).tobytes() for k in range(kinds)]

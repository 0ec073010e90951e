"""The N>1 host logic on CPU: two gloo ranks each produce the (X, a1, a2) rows of their round-robin
shard (computed by the oracle here -- no GPU), all-gather them, re-interleave, and rank 0's transcript
must give exactly the single-process challenge (participant.rs:438-454 hashes in publickeys order)."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpvss_rs_b200 import synth
    from mpvss_rs_b200.sharding import interleave, shard_indices
    from oracle import pvss
    from oracle.groups import Secp256k1Group
    g = Secp256k1Group()
    n_total, t = 10, 4
    sks = synth.private_keys(5, n_total, g.name, g.order())
    pks = [g.generate_public_key(s) for s in sks]
    box = pvss.distribute_secret(g, 42, pks, t, synth.coefficients(5, t, g.order()),
                                 synth.witnesses(5, n_total, g.order()))
    mine = shard_indices(rank, world, n_total)
    n = len(mine)
    rows = np.zeros((3, n, 33), dtype=np.uint8)
    for j, i in enumerate(mine):                       # this rank's participants only
        pk = pks[i]
        kb = g.element_to_bytes(pk)
        x = pvss.x_horner_schedule(g, box.commitments, i + 1)
        a1, a2 = pvss.verifier_commitments(g, g.subgroup_generator(), x, pk, box.shares[kb], box.responses[kb],
                                           box.challenge)
        for k, e in enumerate((x, a1, a2)):
            rows[k, j] = np.frombuffer(g.element_to_bytes(e), dtype=np.uint8)
    local = torch.from_numpy(rows)
    gathered = torch.empty((world * 3, n, 33), dtype=torch.uint8)   # concatenation along dim 0
    dist.all_gather_into_tensor(gathered, local)
    if rank == 0:
        xs, a1s, a2s = interleave(gathered.numpy(), world, 3, n, 33)
        h = hashlib.sha256()
        for i, pk in enumerate(pks):
            for blob in (xs, None, a1s, a2s):
                e = g.element_to_bytes(box.shares[g.element_to_bytes(pk)]) if blob is None else blob[i * 33:(i + 1) * 33]
                h.update(pvss.framed(e))
        result.put(g.hash_to_scalar(h.digest()) == box.challenge and pvss.verify_distribution_shares(g, box))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_verification_gloo():
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert result.get(timeout=5) is True


def test_shard_indices_cover_everything():
    from mpvss_rs_b200.sharding import shard_indices
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(r, world, 4096 * world))
        assert seen == list(range(4096 * world))
        assert len({len(shard_indices(r, world, 4096 * world)) for r in range(world)}) == 1

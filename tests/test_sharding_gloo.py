"""The N>1 host logic on CPU (gloo, world_size 2): each rank produces the framed transcript rows of its
round-robin shard (computed by the oracle here -- no GPU), one all-gather combines them exactly as
ncclAllGather does inside the library ([rank][local row][4 frames]), and the LIBRARY's host routine
(csrc/transcript.h through mpvss_transcript_digest) must hash them in `publickeys` order to the
single-process transcript digest (participant.rs:438-454).  Covers both groups' frame formats, a box
size that does not divide by the world size, and ModpGroup values shorter than 256 bytes."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame(body, eb):
    return len(body).to_bytes(8, "big") + body + bytes(eb - len(body))


def _worker(rank, world, port, group_name, result):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpvss_rs_b200 import lib, synth
    from mpvss_rs_b200.sharding import rows_per_rank, shard_indices
    from oracle import pvss
    from oracle.groups import ModpGroup, Secp256k1Group
    g = ModpGroup() if group_name == "modp" else Secp256k1Group()
    eb = 256 if group_name == "modp" else 33
    n_total, t = 7, 3
    bound = g.q if group_name == "modp" else g.order()
    sks = synth.private_keys(5, n_total, g.name, g.order(), bound)
    pks = [g.generate_public_key(s) for s in sks]
    box = pvss.distribute_secret(g, 42, pks, t, synth.coefficients(5, t, g.order()), synth.witnesses(5, n_total, bound))
    trace = {}
    assert pvss.verify_distribution_shares(g, box, trace=trace)
    if group_name == "modp":              # a short element exercises the minimal-length frames
        trace["a1"][3] = 0x1234
        trace["X"][5] = 0
    rpr, row = rows_per_rank(n_total, world), 4 * (8 + eb)
    local = np.zeros((rpr, row), dtype=np.uint8)
    for j, i in enumerate(shard_indices(rank, world, n_total)):    # this rank's participants only
        y = box.shares[g.element_to_bytes(pks[i])]
        fr = b"".join(_frame(g.element_to_bytes(e), eb) for e in (trace["X"][i], y, trace["a1"][i], trace["a2"][i]))
        local[j] = np.frombuffer(fr, dtype=np.uint8)
    gathered = torch.empty((world * rpr, row), dtype=torch.uint8)  # [rank][local row][row bytes]
    dist.all_gather_into_tensor(gathered, torch.from_numpy(local))
    got = lib.transcript_digest(group_name, gathered.numpy().tobytes(), n_total, world)
    h = hashlib.sha256()
    for i, pk in enumerate(pks):
        for e in (trace["X"][i], box.shares[g.element_to_bytes(pk)], trace["a1"][i], trace["a2"][i]):
            h.update(pvss.framed(g.element_to_bytes(e)))
    result.put((rank, got == h.digest()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("group_name", ["modp", "secp256k1"])
def test_two_rank_sharded_transcript_gloo(group_name):
    from mpvss_rs_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libmpvss_b200.so not built")
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = 29600 + (os.getpid() + len(group_name)) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, group_name, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(result.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]


def test_gather_layout_matches_single_rank():
    """The same rows through world sizes 1, 2, 3, 8 (including more ranks than participants) hash alike."""
    from mpvss_rs_b200 import lib
    from mpvss_rs_b200.sharding import gather_layout
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libmpvss_b200.so not built")
    rng = np.random.default_rng(3)
    for group_name, eb in (("modp", 256), ("ristretto255", 32)):
        rows = []
        for i in range(5):
            fr = b""
            for k in range(4):
                ln = eb if (group_name != "modp" or (i + k) % 3) else int(rng.integers(1, eb))
                fr += _frame(bytes(rng.integers(1, 256, ln, dtype=np.uint8)), eb)
            rows.append(fr)
        ref = lib.transcript_digest(group_name, b"".join(rows), 5, 1)
        h = hashlib.sha256()
        for fr in rows:
            for k in range(4):
                f = fr[k * (8 + eb):(k + 1) * (8 + eb)]
                h.update(f[:8 + int.from_bytes(f[:8], "big")])
        assert ref == h.digest()
        for world in (2, 3, 8):
            assert lib.transcript_digest(group_name, gather_layout(rows, world, 4 * (8 + eb)), 5, world) == ref


def test_shard_indices_cover_everything():
    from mpvss_rs_b200.sharding import shard_indices
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(r, world, 4096 * world))
        assert seen == list(range(4096 * world))
        assert len({len(shard_indices(r, world, 4096 * world)) for r in range(world)}) == 1

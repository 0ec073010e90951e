"""BASELINE.json configs 3 and 4 at full size through size-independent properties (the oracle
cannot run them in reasonable time): dealer X_i == verifier X_i for every i, verification accepts
the box and rejects a flipped response, extracted shares carry valid proofs, and t shares (a
contiguous and a scattered subset) reconstruct the secret."""
import copy

import pytest

import mpvss_rs_b200 as m
from mpvss_rs_b200 import synth
from oracle import pvss
from oracle.groups import GROUPS

pytestmark = pytest.mark.gpu
SECRET = pvss.string_to_secret("Hello MPVSS Example.")


def _run(gname, n, t, reconstruct_sets):
    og = GROUPS[gname]()
    g = m.Group(gname)
    sks = synth.private_keys(77, n, gname, og.order())
    co = synth.coefficients(77, t, og.order())
    ws = synth.witnesses(77, n, og.order())
    dealer = m.Participant(g)
    pks = g.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    tr = {}
    assert dealer.verify_distribution_shares(box, trace=tr) is True
    # dealer-side X_i = p_i * G against the verifier's Horner over the commitments, spot-checked
    # against the oracle's scalar evaluation
    idx = [0, 1, n // 3, n - 1]
    ps = [pvss.poly_eval_mod(co, i + 1, og.order()) for i in idx]
    assert [tr["X"][i] for i in idx] == g.fixed_base_exp(ps)
    assert [box.shares[pks[i]] for i in idx] == g.batch_exp([pks[i] for i in idx], ps)
    bad = copy.copy(box)
    bad.responses = dict(box.responses)
    bad.responses[pks[n // 2]] = (bad.responses[pks[n // 2]] + 1) % og.order()
    assert dealer.verify_distribution_shares(bad) is False
    for sel in reconstruct_sets:
        sbs = dealer.extract_secret_shares(box, [sks[i] for i in sel], [ws[i] for i in sel])
        assert all(dealer.verify_shares(sbs, box, [pks[i] for i in sel]))
        assert dealer.reconstruct(sbs, box) == SECRET
        assert dealer.reconstruct(sbs[:-1], box) is None      # t - 1 shares: None (participant.rs:1458)


def test_config3_secp256k1_n4096_t2731():
    n, t = 4096, 2731
    scattered = sorted(set(range(0, n, 3)) | set(range(1, n, 3)))[:t]
    assert len(scattered) == t
    _run("secp256k1", n, t, [list(range(t)), scattered])


def test_config4_ristretto255_n16384_t10923():
    n, t = 16384, 10923
    scattered = sorted(set(range(0, n, 3)) | set(range(1, n, 3)))[:t]
    assert len(scattered) == t
    _run("ristretto255", n, t, [list(range(t)), scattered])


# ---- the configurations that carry the reported numbers ---------------------------------------
def _cpu_rows(box_bytes, sample, schedule):
    """X_i of the sampled participants from the oracle's C restatement (OpenSSL), both schedules."""
    import bench
    return bench.cpu_reference_step(box_bytes, sample, 4, schedule)[1]


def test_headline_modp_n4096_t2731(modp_group):
    """BASELINE metric configuration: dealer X_i = g^P(i) equals the verifier's chain-Horner X_i for
    every participant, a flipped response is rejected, and 16 sampled X_i agree with the CPU restatement
    running the reference's own schedule (t full exponentiations, participant.rs:423-434) and the
    Horner schedule."""
    import bench
    n, t = 4096, 2731
    box = bench.build_box(modp_group, n, t, 0x6D70767373)
    c = modp_group.codec
    dealer = m.Participant(modp_group)
    pbox = bench.participant_box(modp_group, box)
    tr = {}
    assert dealer.verify_distribution_shares(pbox, trace=tr) is True
    assert c.enc_elems(tr["X"]) == box["x_dealer"]
    sample = bench.spread_sample(n, 16)
    for schedule in (0, 1):
        xs = _cpu_rows(box, sample, schedule)
        assert xs == [box["x_dealer"][i * 256:(i + 1) * 256] for i in sample], schedule
    bad = copy.copy(pbox)
    bad.responses = dict(pbox.responses)
    k = c.key(pbox.publickeys[n // 2])
    bad.responses[k] = (bad.responses[k] + 1) % c.order
    assert dealer.verify_distribution_shares(bad) is False


def test_config5_slice_modp_t43691(modp_group):
    """A slice of BASELINE config 5 (ModpGroup n = 65536, t = 43691): 64 positions spread over 1..65536
    evaluated against the full 43691 commitments; X_i must equal g^P(i) from the scalar kernel + fixed-base
    table, i.e. the dealer's value, and 4 of them the CPU restatement's Horner schedule."""
    og = GROUPS["modp"]()
    t = 43691
    co = synth.coefficients(11, t, og.order())
    positions = sorted({1, 2, 65535, 65536} | {1 + (k * 65536) // 60 for k in range(60)})
    comm = modp_group.fixed_base_exp(co, generator=1)
    xs = modp_group.poly_eval_exp(comm, positions)
    ps = modp_group.scalar_poly_eval(co, positions)
    assert ps[:3] == [pvss.poly_eval_mod(co, p, og.order()) for p in positions[:3]]
    assert xs == modp_group.fixed_base_exp(ps, generator=1)


def test_scalar_poly_eval_all_groups():
    for gname in ("modp", "secp256k1", "ristretto255"):
        og = GROUPS[gname]()
        g = m.Group(gname)
        co = synth.coefficients(3, 9, og.order())
        pos = [1, 2, 3, 77, 4096, 65536, (1 << 31) - 1]
        assert g.scalar_poly_eval(co, pos) == [pvss.poly_eval_mod(co, p, og.order()) for p in pos]


@pytest.mark.parametrize("gname,n,t", [("modp", 700, 5), ("secp256k1", 300, 7), ("ristretto255", 300, 7)])
def test_device_side_transcripts_match_the_host_hash(gname, n, t):
    """SURVEY 8 f1.  (a) 'device_hash': the whole-box transcript as one SHA-256 chain on the device gives the digest
    and verdict of the host pass (n = 700 ModpGroup rows contain frames shorter than 256 bytes with probability
    1 - (255/256)^2800); (b) the per-share transcripts of extract_secret_share / verify_share are always hashed on
    the device: a changed challenge, response or share is rejected share by share."""
    og = GROUPS[gname]()
    g = m.Group(gname)
    sks = synth.private_keys(9, n, gname, og.order(), g.codec.key_bound)
    co = synth.coefficients(9, t, og.order())
    ws = synth.witnesses(9, n, og.order())
    d = m.Participant(g)
    pks = g.fixed_base_exp(sks)
    box = d.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    host, dev = {}, {}
    assert d.verify_distribution_shares(box, trace=host) is True
    bad = copy.copy(box)
    bad.responses = dict(box.responses)
    k3 = g.codec.key(pks[3])
    bad.responses[k3] = (bad.responses[k3] + 1) % og.order()
    g.ctx.set_int("device_hash", 1)
    try:
        assert d.verify_distribution_shares(box, trace=dev) is True
        assert dev["digest"] == host["digest"]
        assert d.verify_distribution_shares(bad) is False
    finally:
        g.ctx.set_int("device_hash", 0)
    # per-share proofs (bit-exact against the oracle in test_gpu_*::test_full_round_bit_exact): accepted as dealt,
    # rejected share by share when the challenge, the response or the share is changed
    k = 40
    sbs = d.extract_secret_shares(box, sks[:k], ws[:k])
    assert all(d.verify_shares(sbs, box, pks[:k]))
    sbs[1].challenge = (sbs[1].challenge + 1) % og.order()
    sbs[2].response = (sbs[2].response + 1) % og.order()
    sbs[4].share = sbs[5].share
    ok = d.verify_shares(sbs, box, pks[:k])
    assert ok == [i not in (1, 2, 4) for i in range(k)]


def test_modp_input_validation_flag(modp_group):
    """'validate' tunable: elements outside (0, q) or outside the order-g subgroup make the
    box verify as false; the reference (and the default here) checks nothing (modp.rs:154-156)."""
    og = GROUPS["modp"]()
    n, t = 6, 3
    sks = synth.private_keys(9, n, "modp", og.order(), og.q)
    pks = modp_group.fixed_base_exp(sks)
    dealer = m.Participant(modp_group)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=synth.coefficients(9, t, og.order()),
                                   witnesses=synth.witnesses(9, n, og.q))
    try:
        modp_group.ctx.set_int("validate", 1)
        assert dealer.verify_distribution_shares(box) is True
        # G = 2 generates the same order-g subgroup (q = 7 mod 8), so public keys pass; q - 1 has order 2
        k = modp_group.codec.key(pks[2])
        for evil in (og.q - 1, 0, og.q + 5):
            bad = copy.copy(box)
            bad.shares = dict(box.shares)
            bad.shares[k] = evil
            assert dealer.verify_distribution_shares(bad) is False
    finally:
        modp_group.ctx.set_int("validate", 0)


def _two_gpu_worker(rank, world, port, q):
    import os
    import sys
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for gname, n, t in (("modp", 37, 9), ("secp256k1", 41, 7), ("ristretto255", 8, 8)):
        og = GROUPS[gname]()
        bound = og.q if gname == "modp" else og.order()
        g1 = m.Group(gname, device=rank)            # single-GPU reference on the same device
        gN = m.Group(gname, device=rank)
        gN.join(rank, world, dist)
        sks = synth.private_keys(21, n, gname, og.order(), bound)
        co, ws = synth.coefficients(21, t, og.order()), synth.witnesses(21, n, bound)
        pks = g1.fixed_base_exp(sks)
        ref = m.Participant(g1).distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
        box = m.Participant(gN).distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)   # collective
        same_box = (box.commitments, box.shares, box.challenge, box.responses, box.U) == \
                   (ref.commitments, ref.shares, ref.challenge, ref.responses, ref.U)
        t1, tN = {}, {}
        ok1 = m.Participant(g1).verify_distribution_shares(ref, trace=t1)
        okN = m.Participant(gN).verify_distribution_shares(ref, trace=tN)                     # collective
        bad = copy.copy(ref)
        bad.responses = dict(ref.responses)
        k = g1.codec.key(pks[n - 1])
        bad.responses[k] = (bad.responses[k] + 1) % og.order()
        okbad = m.Participant(gN).verify_distribution_shares(bad)
        out[gname] = (same_box, ok1, okN, t1["digest"] == tN["digest"], okbad)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_calls_match_single_gpu():
    """Library-internal sharding (comm.cu): with a 2-rank NCCL communicator, mpvss_distribute and
    mpvss_verify_distribution give every rank the box / transcript digest / verdict of the 1-GPU path."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    for rank in (0, 1):
        for gname, v in res[rank].items():
            assert v == (True, True, True, True, False), (rank, gname, v)


def test_two_gpu_threads_in_one_process():
    """The Rust-shaped use: ONE process, one context per GPU, one thread per
    context; the collective verify returns the single-GPU digest and verdict on both threads."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mpvss_rs_b200.lib import comm_unique_id
    og = GROUPS["modp"]()
    n, t = 23, 8
    g1 = m.Group("modp", device=0)
    sks = synth.private_keys(31, n, "modp", og.order(), og.q)
    pks = g1.fixed_base_exp(sks)
    box = m.Participant(g1).distribute_secret(SECRET, pks, t, coeffs=synth.coefficients(31, t, og.order()),
                                              witnesses=synth.witnesses(31, n, og.q))
    ref = {}
    assert m.Participant(g1).verify_distribution_shares(box, trace=ref) is True
    uid, groups, out = comm_unique_id(), [m.Group("modp", device=r) for r in range(2)], [None, None]

    def worker(r):
        groups[r].ctx.comm_init(uid, 2, r)
        tr = {}
        out[r] = (m.Participant(groups[r]).verify_distribution_shares(box, trace=tr), tr["digest"])
    th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join(300)
    assert out[0] == out[1] == (True, ref["digest"])

"""GPU parity tests for ModpGroup: the CUDA path (through the C ABI) against the CPU
oracle on the same injected inputs.  Bit-exact (integer work): every commitment,
encrypted share, X_i, a1/a2, challenge, response, decrypted share, U and
reconstructed secret must be equal.  Mirrors the reference's own tests
(src/participant.rs:593-743, src/mpvss.rs:151-287, tests/mpvss_tests.rs:11-87)."""
import random

import pytest

from mpvss_rs_b200 import synth
from oracle import pvss
from oracle.groups import ModpGroup as OModp

pytestmark = pytest.mark.gpu

OG = OModp()
Q = OG.q
SECRET = pvss.string_to_secret("Hello MPVSS Example.")


def _setup(n, t, seed):
    sks = synth.private_keys(seed, n, "modp", OG.order(), Q)
    co = synth.coefficients(seed, t, OG.order())
    ws = synth.witnesses(seed, n, Q)
    return sks, co, ws


def test_batch_mul_and_exp(modp_group):
    rng = random.Random(7)
    n = 37
    a = [rng.randrange(Q) for _ in range(n)]
    b = [rng.randrange(Q) for _ in range(n)]
    a[0], b[0] = Q - 1, Q - 1
    a[1] = 0
    b[2] = 1
    assert modp_group.batch_mul(a, b) == [x * y % Q for x, y in zip(a, b)]
    e = [rng.randrange(Q - 1) for _ in range(n)]
    e[0], e[1], e[2], e[3] = 0, 1, Q - 2, 2
    a[1] = 5
    assert modp_group.batch_exp(a, e) == [pow(x, y, Q) for x, y in zip(a, e)]
    assert modp_group.batch_exp(4, e) == [pow(4, y, Q) for y in e]
    assert modp_group.fixed_base_exp(e) == [pow(2, y, Q) for y in e]
    assert modp_group.fixed_base_exp(e, 1) == [pow(4, y, Q) for y in e]
    # short exponents take the short-window path
    small = [rng.getrandbits(64) for _ in range(5)]
    assert modp_group.batch_exp(a[:5], small) == [pow(x, y, Q) for x, y in zip(a, small)]
    assert modp_group.ctx.last_kernel_launches >= 1


@pytest.mark.parametrize("tpi", [4, 8, 16])
def test_poly_eval_exp_matches_reference_schedule(modp_group, tpi):
    modp_group.ctx.set_int("modp_tpi", tpi)
    try:
        rng = random.Random(tpi)
        t = 6
        comm = [pow(4, rng.randrange(Q - 1), Q) for _ in range(t)]
        positions = [1, 2, 3, 4, 7, 8, 15, 16, 31, 1000, 4095, 65536, 5]
        got = modp_group.poly_eval_exp(comm, positions)
        want = [pvss.x_reference_schedule(OG, comm, p) for p in positions]
        assert got == want
        # t = 1: X_i = C_0
        assert modp_group.poly_eval_exp(comm[:1], [1, 9]) == [comm[0], comm[0]]
    finally:
        modp_group.ctx.set_int("modp_tpi", 8)


def test_dleq_commitments(modp_group):
    rng = random.Random(3)
    n = 9
    h1 = [rng.randrange(1, Q) for _ in range(n)]
    g2 = [rng.randrange(1, Q) for _ in range(n)]
    h2 = [rng.randrange(1, Q) for _ in range(n)]
    r = [rng.randrange(Q - 1) for _ in range(n)]
    c = rng.getrandbits(256)
    a1, a2 = modp_group.dleq_verify_commit(4, h1, g2, h2, r, c)
    assert a1 == [pow(4, r[i], Q) * pow(h1[i], c, Q) % Q for i in range(n)]
    assert a2 == [pow(g2[i], r[i], Q) * pow(h2[i], c, Q) % Q for i in range(n)]
    cs = [rng.getrandbits(256) for _ in range(n)]
    a1, a2 = modp_group.dleq_verify_commit(2, h1, g2, h2, r, cs)
    assert a1 == [pow(2, r[i], Q) * pow(h1[i], cs[i], Q) % Q for i in range(n)]
    assert a2 == [pow(g2[i], r[i], Q) * pow(h2[i], cs[i], Q) % Q for i in range(n)]
    w = [rng.randrange(Q) for _ in range(n)]
    p1, p2 = modp_group.dleq_prove_commit(4, g2, w)
    assert p1 == [pow(4, x, Q) for x in w]
    assert p2 == [pow(g2[i], w[i], Q) for i in range(n)]
    assert modp_group.multi_exp(g2, r) == __import__("functools").reduce(
        lambda acc, i: acc * pow(g2[i], r[i], Q) % Q, range(n), 1)


@pytest.mark.parametrize("n,t,subset", [(3, 3, [0, 1, 2]), (4, 3, [0, 1, 3]), (5, 3, [0, 2, 4]), (7, 2, [0, 2])])
def test_full_round_bit_exact(modp_group, n, t, subset):
    """examples/mpvss_all.rs (n=3,t=3 = BASELINE config 1), examples/mpvss_sub.rs (4 of 3,
    {1,2,4}), participant.rs:832-903 (3-of-5 from {1,3,5}), participant.rs:703-743 (t=2, {1,3})."""
    import mpvss_rs_b200 as m
    sks, co, ws = _setup(n, t, 100 + n)
    dealer = m.Participant(modp_group)
    pks = modp_group.fixed_base_exp(sks)
    assert pks == [OG.generate_public_key(s) for s in sks]
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    obox = pvss.distribute_secret(OG, SECRET, pks, t, co, ws)
    assert box.commitments == obox.commitments
    assert box.shares == obox.shares and box.positions == obox.positions
    assert box.challenge == obox.challenge
    assert box.responses == obox.responses
    assert box.U == obox.U and box.publickeys == obox.publickeys
    tr, otr = {}, {}
    assert dealer.verify_distribution_shares(box, trace=tr) is True
    assert pvss.verify_distribution_shares(OG, obox, trace=otr) is True
    assert tr["X"] == otr["X"] == obox.trace["X"]
    assert tr["a1"] == otr["a1"] == obox.trace["a1"] and tr["a2"] == otr["a2"] == obox.trace["a2"]
    # extraction + share proofs
    w2 = synth.witnesses(200 + n, n, Q, "extract")
    sbs = dealer.extract_secret_shares(box, sks, w2)
    for i in range(n):
        osb = pvss.extract_secret_share(OG, obox, sks[i], w2[i])
        assert (sbs[i].publickey, sbs[i].share, sbs[i].challenge, sbs[i].response) == \
               (osb.publickey, osb.share, osb.challenge, osb.response)
    assert dealer.verify_shares(sbs, box, pks) == [True] * n
    # reconstruction from a subset
    chosen = [sbs[i] for i in subset]
    tr2, otr2 = {}, {}
    got = dealer.reconstruct(chosen, box, trace=tr2)
    want = pvss.reconstruct(OG, [pvss.ShareBox(s.publickey, s.share, s.challenge, s.response) for s in chosen],
                            obox, trace=otr2)
    assert tr2["G_s"] == otr2["G_s"] == obox.trace["G_s"]
    assert got == want == SECRET
    assert m.string_from_secret(got) == "Hello MPVSS Example."
    # too few shares -> None (participant.rs:469)
    if t > 1:
        assert dealer.reconstruct(chosen[: t - 1], box) is None


def test_tampering_is_rejected(modp_group):
    import copy
    import mpvss_rs_b200 as m
    n, t = 5, 3
    sks, co, ws = _setup(n, t, 55)
    dealer = m.Participant(modp_group)
    pks = modp_group.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    assert dealer.verify_distribution_shares(box)
    k = next(iter(box.responses))
    bad = copy.deepcopy(box)
    bad.responses[k] = (bad.responses[k] + 1) % (Q - 1)
    assert dealer.verify_distribution_shares(bad) is False
    bad = copy.deepcopy(box)
    bad.shares[k] = bad.shares[k] * 4 % Q
    assert dealer.verify_distribution_shares(bad) is False
    bad = copy.deepcopy(box)
    bad.commitments[1] = bad.commitments[1] * 4 % Q
    assert dealer.verify_distribution_shares(bad) is False
    bad = copy.deepcopy(box)
    del bad.responses[k]                     # participant.rs:415-420
    assert dealer.verify_distribution_shares(bad) is False
    sb = dealer.extract_secret_share(box, sks[0], ws[1])
    assert dealer.verify_share(sb, box, pks[0]) is True
    sb.share = sb.share * 2 % Q
    assert dealer.verify_share(sb, box, pks[0]) is False
    assert dealer.verify_share(sb, box, 12345) is False       # unknown key -> false (participant.rs:371-375)
    # a private key without inverse mod q-1 yields None (participant.rs:314)
    even_sk = sks[0] + 1
    evbox = dealer.distribute_secret(SECRET, [modp_group.generate_public_key(even_sk)] + pks[1:], t,
                                     coeffs=co, witnesses=ws)
    assert dealer.extract_secret_share(evbox, even_sk, ws[0]) is None
    # ... and inside a batch it only spoils its own instance: the inverses of the batch come from ONE inversion
    # (product tree on the device), so the non-invertible key must not leak into its neighbours
    n = len(sks)
    got = dealer.extract_secret_shares(evbox, [even_sk] + sks[1:], ws)
    assert got[0] is None and all(g is not None for g in got[1:])
    want = dealer.extract_secret_shares(box, sks, ws)
    for i in range(1, n):            # same keys, witnesses and encrypted shares as in the untouched box
        assert evbox.shares[modp_group.codec.key(pks[i])] == box.shares[modp_group.codec.key(pks[i])]
        assert (got[i].publickey, got[i].share, got[i].challenge, got[i].response) == \
               (want[i].publickey, want[i].share, want[i].challenge, want[i].response)
    assert all(dealer.verify_shares(got[1:], evbox, pks[1:]))


def test_medium_box_against_oracle(modp_group):
    """n=48, t=32: every X_i / a1 / a2 and the digest against the oracle (Horner oracle schedule,
    itself asserted equal to the reference schedule in the CPU suite)."""
    import mpvss_rs_b200 as m
    n, t = 48, 32
    sks, co, ws = _setup(n, t, 9)
    dealer = m.Participant(modp_group)
    pks = modp_group.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    obox = pvss.distribute_secret(OG, SECRET, pks, t, co, ws, x_schedule=pvss.x_horner_schedule)
    assert box.commitments == obox.commitments and box.challenge == obox.challenge
    assert box.responses == obox.responses and box.shares == obox.shares and box.U == obox.U
    tr, otr = {}, {}
    assert dealer.verify_distribution_shares(box, trace=tr)
    assert pvss.verify_distribution_shares(OG, obox, x_schedule=pvss.x_horner_schedule, trace=otr)
    assert tr["X"] == otr["X"] and tr["a1"] == otr["a1"] and tr["a2"] == otr["a2"]


@pytest.mark.parametrize("n,t", [(1024, 683)])
def test_config2_properties(modp_group, n, t):
    """BASELINE config 2 (MODP n=1024 t=683) through size-independent properties: the dealer's
    X_i = g^P(i) (one fixed-base exponentiation) must equal the verifier's Horner product over the
    commitments for every i, the box must verify, a single flipped response must not, and t shares
    must reconstruct the secret."""
    import copy
    import mpvss_rs_b200 as m
    sks, co, ws = _setup(n, t, 2)
    dealer = m.Participant(modp_group)
    pks = modp_group.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    tr = {}
    assert dealer.verify_distribution_shares(box, trace=tr)
    ps = [pvss.poly_eval_mod(co, i + 1, Q - 1) for i in (0, 1, n // 2, n - 1)]
    for idx, p in zip((0, 1, n // 2, n - 1), ps):
        assert tr["X"][idx] == pow(4, p, Q)
        assert box.shares[dealer.group.codec.key(pks[idx])] == pow(pks[idx], p, Q)
    assert tr["X"] == modp_group.fixed_base_exp([pvss.poly_eval_mod(co, i + 1, Q - 1) for i in range(n)], 1)
    bad = copy.deepcopy(box)
    k = dealer.group.codec.key(pks[n - 1])
    bad.responses[k] ^= 1
    assert dealer.verify_distribution_shares(bad) is False
    sbs = dealer.extract_secret_shares(box, sks[:t], ws[:t])
    assert all(dealer.verify_shares(sbs, box, pks[:t]))
    assert dealer.reconstruct(sbs, box) == SECRET


def test_dleq_and_pvss_wrappers(modp_group):
    """dleq.rs:380-403 (r = w - alpha*c mod q-1) and a prove/verify round trip through the DLEQ mirror;
    mpvss.rs:151-287 through the PVSS mirror."""
    import hashlib
    import mpvss_rs_b200 as m
    d = m.DLEQ(modp_group)
    w, alpha, c = 81647, 163027, 127997
    d.init(4, pow(4, alpha, Q), 2, pow(2, alpha, Q), alpha, w)
    d.c = c
    assert d.get_r() == (w - alpha * c) % (Q - 1)
    assert d.get_a1() == pow(4, w, Q) and d.get_a2() == pow(2, w, Q)
    hasher = hashlib.sha256()
    d.update_hash(hasher)
    d.c = m.hash_to_scalar(modp_group, hasher.digest())
    d.r = d.get_r()
    assert d.verify() is True
    d.c += 1
    assert d.verify() is False
    n, t = 4, 3
    sks, co, ws = _setup(n, t, 321)
    pks = modp_group.fixed_base_exp(sks)
    box = m.Participant(modp_group).distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    assert m.PVSS(modp_group).verify_distribution_shares(box) is True


def test_overlap_modes_and_lane_counts_agree(modp_group):
    """Every placement of the X-independent a2 launch (before the Horner launch, beside it, as persistent
    one-warp CTAs in the idle warp slots) and every lane count per value yields the same X, a1, a2."""
    import mpvss_rs_b200 as m
    n, t = 1024, 683
    sks, co, ws = _setup(n, t, 5)
    dealer = m.Participant(modp_group)
    pks = modp_group.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    ref = None
    try:
        for key, values in (("modp_overlap", (0, 2, 3)), ("modp_tpi", (4, 16, 8)), ("modp_chunks", (2, 3, 8, 0)),
                            ("modp_wpc", (4, 0))):
            for v in values:
                modp_group.ctx.set_int(key, v)
                tr = {}
                assert dealer.verify_distribution_shares(box, trace=tr) is True, (key, v)
                cur = (tr["X"], tr["a1"], tr["a2"], tr["digest"])
                ref = ref or cur
                assert cur == ref, (key, v)
    finally:
        modp_group.ctx.set_int("modp_overlap", 3)
        modp_group.ctx.set_int("modp_tpi", 8)
        modp_group.ctx.set_int("modp_chunks", 0)
        modp_group.ctx.set_int("modp_wpc", 0)


def test_chunked_horner_small_boxes(modp_group):
    """Few positions: the polynomial is evaluated in K contiguous chunks and recombined as
    X = prod_k H_k^(pos^(k B) mod (q-1)); same X as the single chain, the dealer's g^P(i) and the oracle, for
    even and odd positions (the CRT parity fix of the chunk exponents) and chunk counts that do not divide t."""
    n, t = 9, 70
    co = synth.coefficients(44, t, OG.order())
    comm = modp_group.fixed_base_exp(co, generator=1)
    pos = [1, 2, 3, 4, 7, 100, 4095, 4096, 65537]
    want = modp_group.fixed_base_exp([pvss.poly_eval_mod(co, p, OG.order()) for p in pos], generator=1)
    try:
        for k in (1, 2, 3, 5, 0):
            modp_group.ctx.set_int("modp_chunks", k)
            assert modp_group.poly_eval_exp(comm, pos) == want, k
    finally:
        modp_group.ctx.set_int("modp_chunks", 0)
    assert want[0] == pvss.x_reference_schedule(OG, comm, 1) and want[5] == pvss.x_horner_schedule(OG, comm, 100)


def test_bucket_multi_exp_equals_direct(modp_group):
    """multi_exp / reconstruct through the bucket method (Pippenger, 8-bit windows) give the same element as
    one exponentiation per base + product tree, and both equal the integer result."""
    import mpvss_rs_b200 as m
    rng = random.Random(77)
    k = 300
    bases = [pow(2, rng.randrange(Q - 1), Q) for _ in range(k)]
    exps = [rng.randrange(Q - 1) for _ in range(k)]
    exps[0], exps[1] = 0, 1
    want = 1
    for b, e in zip(bases, exps):
        want = want * pow(b, e, Q) % Q
    try:
        modp_group.ctx.set_int("modp_msm", 0)
        assert modp_group.multi_exp(bases, exps) == want
        modp_group.ctx.set_int("modp_msm", 1)
        assert modp_group.multi_exp(bases, exps) == want
        # short exponents (fewer windows) and a single base
        assert modp_group.multi_exp(bases[:5], [3, 0, 255, 256, 65535]) == \
            pow(bases[0], 3, Q) * pow(bases[2], 255, Q) * pow(bases[3], 256, Q) * pow(bases[4], 65535, Q) % Q
        assert modp_group.multi_exp(bases[:1], exps[5:6]) == pow(bases[0], exps[5], Q)
        # reconstruct from a scattered subset, G^s against the oracle
        n, t = 40, 27
        sks, co, ws = _setup(n, t, 15)
        dealer = m.Participant(modp_group)
        pks = modp_group.fixed_base_exp(sks)
        box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
        sel = sorted(rng.sample(range(n), t))
        sbs = dealer.extract_secret_shares(box, [sks[i] for i in sel], [ws[i] for i in sel])
        tr = {}
        assert dealer.reconstruct(sbs, box, trace=tr) == SECRET
        assert tr["G_s"] == pow(2, co[0] % (Q - 1), Q)
    finally:
        modp_group.ctx.set_int("modp_msm", 2)

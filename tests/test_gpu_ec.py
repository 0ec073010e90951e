"""GPU parity tests for Secp256k1Group and Ristretto255Group: the CUDA path (through the C ABI)
against the CPU oracle on the same injected inputs, bit-exact on every encoded element and
scalar.  Mirrors the reference's tests (src/participant.rs:752-903, examples/mpvss_*_secp256k1.rs,
examples/mpvss_*_ristretto255.rs)."""
import copy
import random

import pytest

import mpvss_rs_b200 as m
from mpvss_rs_b200 import synth
from oracle import pvss
from oracle.groups import Ristretto255Group, Secp256k1Group

pytestmark = pytest.mark.gpu
SECRET = pvss.string_to_secret("Hello MPVSS Example.")
ORACLES = {"secp256k1": Secp256k1Group, "ristretto255": Ristretto255Group}


@pytest.fixture(scope="module", params=["secp256k1", "ristretto255"])
def pair(request):
    return m.Group(request.param), ORACLES[request.param]()


def _enc(og, es):
    return [og.element_to_bytes(e) for e in es]


def test_batch_ops(pair):
    g, og = pair
    order = og.order()
    rng = random.Random(5)
    n = 33
    ks = [rng.randrange(1, order) for _ in range(n)]
    pts = [og.exp(og.generator(), k) for k in ks]
    e = [rng.randrange(order) for _ in range(n)]
    e[0], e[1], e[2] = 0, 1, order - 1
    assert g.fixed_base_exp(ks) == _enc(og, pts)
    assert g.batch_exp(_enc(og, pts), e) == _enc(og, [og.exp(p, x) for p, x in zip(pts, e)])
    assert g.batch_exp(_enc(og, pts)[3], e) == _enc(og, [og.exp(pts[3], x) for x in e])
    rev = pts[::-1]
    rev[0] = pts[-1]                                           # P + P
    rev[1] = og.element_inverse(pts[-2])                       # P + (-P) = identity
    assert g.batch_mul(_enc(og, pts[::-1]), _enc(og, rev)) == _enc(og, [og.mul(a, b) for a, b in zip(pts[::-1], rev)])
    acc = og.identity()
    for p, x in zip(pts, e):
        acc = og.mul(acc, og.exp(p, x))
    assert g.multi_exp(_enc(og, pts), e) == og.element_to_bytes(acc)
    # DLEQ commitments, shared and per-item challenges
    h1, g2, h2 = pts[:8], pts[8:16], pts[16:24]
    r = e[:8]
    c = rng.randrange(order)
    a1, a2 = g.dleq_verify_commit(og.element_to_bytes(og.generator()), _enc(og, h1), _enc(og, g2), _enc(og, h2), r, c)
    for i in range(8):
        w1, w2 = pvss.verifier_commitments(og, og.generator(), h1[i], g2[i], h2[i], r[i], c)
        assert (a1[i], a2[i]) == (og.element_to_bytes(w1), og.element_to_bytes(w2))
    cs = [rng.randrange(order) for _ in range(8)]
    a1, a2 = g.dleq_verify_commit(og.element_to_bytes(og.generator()), _enc(og, h1), _enc(og, g2), _enc(og, h2), r, cs)
    for i in range(8):
        w1, w2 = pvss.verifier_commitments(og, og.generator(), h1[i], g2[i], h2[i], r[i], cs[i])
        assert (a1[i], a2[i]) == (og.element_to_bytes(w1), og.element_to_bytes(w2))
    # invalid encodings are reported, never computed on (reference: bytes_to_element -> None)
    bad = _enc(og, pts[:4])
    bad[2] = b"\xff" * len(bad[2])
    with pytest.raises(m.MpvssError) as ei:
        g.batch_exp(bad, e[:4])
    assert ei.value.status == -3


def test_poly_eval_exp(pair):
    g, og = pair
    rng = random.Random(8)
    for t in (1, 5, 40):
        comm = [og.exp(og.generator(), rng.randrange(1, og.order())) for _ in range(t)]
        positions = [1, 2, 3, 4, 7, 16, 255, 4096, 65536, 5]
        got = g.poly_eval_exp(_enc(og, comm), positions)
        assert got == _enc(og, [pvss.x_reference_schedule(og, comm, p) for p in positions])


@pytest.mark.parametrize("n,t,subset", [(3, 3, [0, 1, 2]), (4, 3, [0, 1, 3]), (5, 3, [0, 2, 4])])
def test_full_round_bit_exact(pair, n, t, subset):
    g, og = pair
    name = og.name
    sks = synth.private_keys(300 + n, n, name, og.order())
    co = synth.coefficients(300 + n, t, og.order())
    ws = synth.witnesses(300 + n, n, og.order())
    dealer = m.Participant(g)
    opks = [og.generate_public_key(s) for s in sks]
    pks = g.fixed_base_exp(sks)
    assert pks == _enc(og, opks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    obox = pvss.distribute_secret(og, SECRET, opks, t, co, ws)
    assert box.commitments == _enc(og, obox.commitments)
    assert box.positions == obox.positions
    assert box.shares == {k: og.element_to_bytes(v) for k, v in obox.shares.items()}
    assert box.challenge == obox.challenge and box.responses == obox.responses and box.U == obox.U
    tr, otr = {}, {}
    assert dealer.verify_distribution_shares(box, trace=tr) is True
    assert pvss.verify_distribution_shares(og, obox, trace=otr) is True
    for k in ("X", "a1", "a2"):
        assert tr[k] == _enc(og, otr[k])
    w2 = synth.witnesses(400 + n, n, og.order(), "extract")
    sbs = dealer.extract_secret_shares(box, sks, w2)
    osbs = [pvss.extract_secret_share(og, obox, sks[i], w2[i]) for i in range(n)]
    for sb, osb in zip(sbs, osbs):
        assert (sb.publickey, sb.share, sb.challenge, sb.response) == \
               (og.element_to_bytes(osb.publickey), og.element_to_bytes(osb.share), osb.challenge, osb.response)
    assert dealer.verify_shares(sbs, box, pks) == [True] * n
    tr2, otr2 = {}, {}
    got = dealer.reconstruct([sbs[i] for i in subset], box, trace=tr2)
    want = pvss.reconstruct(og, [osbs[i] for i in subset], obox, trace=otr2)
    assert tr2["G_s"] == og.element_to_bytes(otr2["G_s"])
    assert got == want == SECRET
    assert dealer.reconstruct([sbs[i] for i in subset][: t - 1], box) is None
    # tampering
    k0 = next(iter(box.responses))
    bad = copy.deepcopy(box)
    bad.responses[k0] = (bad.responses[k0] + 1) % og.order()
    assert dealer.verify_distribution_shares(bad) is False
    bad = copy.deepcopy(box)
    bad.shares[k0] = pks[1] if bad.shares[k0] != pks[1] else pks[2]
    assert dealer.verify_distribution_shares(bad) is False
    sb = copy.deepcopy(sbs[0])
    sb.response = (sb.response + 1) % og.order()
    assert dealer.verify_share(sb, box, pks[0]) is False


def test_medium_box_properties(pair):
    """n=300, t=200: dealer X_i (one fixed-base multiplication) == verifier X_i (chunked Horner over
    the commitments) for every i; the box verifies; t shares reconstruct the secret; spot checks
    against the oracle."""
    g, og = pair
    n, t = 300, 200
    sks = synth.private_keys(9, n, og.name, og.order())
    co = synth.coefficients(9, t, og.order())
    ws = synth.witnesses(9, n, og.order())
    dealer = m.Participant(g)
    pks = g.fixed_base_exp(sks)
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    tr = {}
    assert dealer.verify_distribution_shares(box, trace=tr)
    ps = [pvss.poly_eval_mod(co, i + 1, og.order()) for i in range(n)]
    assert tr["X"] == g.fixed_base_exp(ps)
    for i in (0, 17, n - 1):
        assert tr["X"][i] == og.element_to_bytes(og.exp(og.generator(), ps[i]))
        assert box.shares[pks[i]] == og.element_to_bytes(og.exp(og.bytes_to_element(pks[i]), ps[i]))
    sbs = dealer.extract_secret_shares(box, sks[5:5 + t], ws[5:5 + t])
    assert all(dealer.verify_shares(sbs, box, pks[5:5 + t]))
    assert dealer.reconstruct(sbs, box) == SECRET


def test_dleq_wrapper_matches_reference_semantics(pair):
    """src/dleq.rs tests restated (dleq.rs:357-441): a1/a2, r = w - alpha*c, prove -> verify round trip."""
    import hashlib
    g, og = pair
    rng = random.Random(77)
    order = og.order()
    alpha, w = rng.randrange(1, order), rng.randrange(1, order)
    g1 = og.generator()
    g2 = og.exp(og.generator(), rng.randrange(1, order))
    h1, h2 = og.exp(g1, alpha), og.exp(g2, alpha)
    E = og.element_to_bytes
    d = m.DLEQ(g)
    d.init(E(g1), E(h1), E(g2), E(h2), alpha, w)
    assert d.get_a1() == E(og.exp(g1, w)) and d.get_a2() == E(og.exp(g2, w))
    hasher = hashlib.sha256()
    d.update_hash(hasher)
    d.c = m.hash_to_scalar(g, hasher.digest())
    oh = hashlib.sha256()
    pvss.append_transcript(og, h1, h2, og.exp(g1, w), og.exp(g2, w), oh)
    assert d.c == pvss.challenge_from(og, oh)
    d.r = d.get_r()
    assert d.r == pvss.prover_response(og, w, alpha, d.c)
    assert d.verify() is True
    d.r = (d.r + 1) % order
    assert d.verify() is False

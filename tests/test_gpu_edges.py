"""Edge cases through the C ABI: degenerate boxes (n = t = 1, zero coefficients -> identity
commitments), argument errors mapped to status codes instead of the reference's panics/None,
non-canonical scalars, unknown public keys (the reference's HashMap misses)."""
import ctypes

import pytest

import mpvss_rs_b200 as m
from mpvss_rs_b200 import lib as L
from mpvss_rs_b200 import synth
from mpvss_rs_b200.lib import buf, ptr
from oracle import pvss
from oracle.groups import GROUPS

pytestmark = pytest.mark.gpu
SECRET = pvss.string_to_secret("edge")


@pytest.fixture(scope="module", params=["modp", "secp256k1", "ristretto255"])
def pair(request):
    return m.Group(request.param), GROUPS[request.param]()


def _host(og, e):
    """oracle element -> host-mirror representation"""
    return e if og.name == "modp" else og.element_to_bytes(e)


def test_single_participant_box(pair):
    g, og = pair
    kb = getattr(og, "q", og.order())
    sk = synth.private_keys(1, 1, og.name, og.order(), getattr(og, "q", None))
    co, ws = synth.coefficients(1, 1, og.order()), synth.witnesses(1, 1, kb)
    dealer = m.Participant(g)
    pks = g.fixed_base_exp(sk)
    box = dealer.distribute_secret(SECRET, pks, 1, coeffs=co, witnesses=ws)
    obox = pvss.distribute_secret(og, SECRET, [og.generate_public_key(sk[0])], 1, co, ws)
    assert box.commitments == [_host(og, c) for c in obox.commitments]
    assert box.challenge == obox.challenge and box.U == obox.U
    assert dealer.verify_distribution_shares(box)
    sb = dealer.extract_secret_share(box, sk[0], ws[0])
    assert dealer.verify_share(sb, box, pks[0])
    assert dealer.reconstruct([sb], box) == SECRET


def test_zero_coefficients_give_identity_commitments(pair):
    """a_1 = a_2 = 0: C_1, C_2 are the identity element; X_i = C_0 for every i (the accumulator of the
    reference's product starts at the identity, participant.rs:207)."""
    g, og = pair
    n, t = 4, 3
    kb = getattr(og, "q", og.order())
    sks = synth.private_keys(3, n, og.name, og.order(), getattr(og, "q", None))
    co = [synth.coefficients(3, 1, og.order())[0], 0, 0]
    ws = synth.witnesses(3, n, kb)
    dealer = m.Participant(g)
    pks = g.fixed_base_exp(sks)
    opks = [og.generate_public_key(s) for s in sks]
    box = dealer.distribute_secret(SECRET, pks, t, coeffs=co, witnesses=ws)
    obox = pvss.distribute_secret(og, SECRET, opks, t, co, ws)
    assert box.commitments == [_host(og, c) for c in obox.commitments]
    assert box.commitments[1] == _host(og, og.identity())
    assert box.challenge == obox.challenge and box.responses == obox.responses
    tr, otr = {}, {}
    assert dealer.verify_distribution_shares(box, trace=tr) and pvss.verify_distribution_shares(og, obox, trace=otr)
    assert tr["X"] == [_host(og, x) for x in otr["X"]] == [box.commitments[0]] * n
    sbs = dealer.extract_secret_shares(box, sks, ws)
    assert dealer.reconstruct(sbs[:t], box) == SECRET


def test_argument_errors_are_status_codes(pair):
    g, og = pair
    c, ctx = g.codec, g.ctx
    out = buf(size=4 * c.eb)
    one = buf(c.enc_scalar(1))
    # null pointers / zero counts -> MPVSS_ERR_ARG, never a crash
    assert ctx.lib.mpvss_batch_exp(ctx.h, None, c.eb, ptr(one), 1, ptr(out)) == L.ERR_ARG
    assert ctx.lib.mpvss_batch_exp(ctx.h, ptr(out), c.eb, ptr(one), 0, ptr(out)) == L.ERR_ARG
    assert ctx.lib.mpvss_fixed_base_exp(ctx.h, 7, ptr(one), 1, ptr(out)) == L.ERR_ARG
    assert b"bad arguments" in ctx.lib.mpvss_last_error(ctx.h) or b"generator" in ctx.lib.mpvss_last_error(ctx.h)
    # threshold > n: the reference asserts (participant.rs:166); here a status
    n, t = 2, 3
    sks = synth.private_keys(5, n, og.name, og.order(), getattr(og, "q", None))
    pks = g.fixed_base_exp(sks)
    bufs = [buf(size=8 * max(c.eb, c.sb)) for _ in range(5)]
    st = ctx.lib.mpvss_distribute(ctx.h, n, t, ptr(buf(b"x")), 1, ptr(buf(c.enc_scalars([1, 2, 3]))),
                                  ptr(buf(c.enc_scalars([4, 5]))), ptr(buf(c.enc_elems(pks))), ptr(bufs[0]),
                                  ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]), ptr(bufs[4]), None)
    assert st == L.ERR_ARG
    # positions must be >= 1
    pos = (ctypes.c_int64 * 1)(0)
    assert ctx.lib.mpvss_poly_eval_exp(ctx.h, ptr(buf(c.enc_elems(pks[:1]))), 1, pos, 1, ptr(out)) == L.ERR_ARG
    assert ctx.lib.mpvss_ctx_set_int(ctx.h, b"no_such_knob", 1) == L.ERR_ARG
    # run before stage
    fresh = m.Group(og.name)
    ok = ctypes.c_int(0)
    assert fresh.ctx.lib.mpvss_verify_distribution_run(fresh.ctx.h, ctypes.byref(ok), None, None, None, None) == L.ERR_ARG


def test_unknown_keys_follow_the_reference_maps(pair):
    """extract/verify_share with a key that is not in the box: None / false (participant.rs:310, 371-375)."""
    g, og = pair
    n, t = 3, 2
    kb = getattr(og, "q", og.order())
    sks = synth.private_keys(8, n + 1, og.name, og.order(), getattr(og, "q", None))
    pks = g.fixed_base_exp(sks)
    dealer = m.Participant(g)
    box = dealer.distribute_secret(SECRET, pks[:n], t, coeffs=synth.coefficients(8, t, og.order()),
                                   witnesses=synth.witnesses(8, n, kb))
    assert dealer.extract_secret_share(box, sks[n], 12345) is None
    sbs = dealer.extract_secret_shares(box, sks, [11, 12, 13, 14])
    assert sbs[n] is None and all(s is not None for s in sbs[:n])
    assert dealer.verify_share(sbs[0], box, pks[n]) is False
    stranger = m.ShareBox(pks[n], sbs[0].share, sbs[0].challenge, sbs[0].response)
    assert dealer.reconstruct([sbs[0], stranger], box) is None          # participant.rs:480


def test_noncanonical_scalars():
    """secp256k1: Scalar::from_repr rejects values >= n (participant.rs:1143 unwrap) -> MPVSS_ERR_ENCODING;
    ristretto255: from_bytes_mod_order reduces (ristretto255.rs:104)."""
    from oracle.groups import ED_L, SECP_N, Ristretto255Group
    gs = m.Group("secp256k1")
    c = gs.codec
    out = buf(size=c.eb)
    bad = buf(SECP_N.to_bytes(32, "big"))
    gen = gs.fixed_base_exp([1])[0]
    assert gs.ctx.lib.mpvss_batch_exp(gs.ctx.h, ptr(buf(gen)), c.eb, ptr(bad), 1, ptr(out)) == L.ERR_ENCODING
    gr = m.Group("ristretto255")
    og = Ristretto255Group()
    big = buf((ED_L + 5).to_bytes(32, "little"))
    out = buf(size=32)
    gen = gr.fixed_base_exp([1])[0]
    assert gr.ctx.lib.mpvss_batch_exp(gr.ctx.h, ptr(buf(gen)), 32, ptr(big), 1, ptr(out)) == L.OK
    assert bytes(out) == og.element_to_bytes(og.exp(og.generator(), 5))


@pytest.mark.parametrize("gname", ["secp256k1", "ristretto255"])
def test_malformed_box_contents_verify_as_false(gname):
    """Untrusted box contents that do not decode -- an invalid point, a non-canonical response, a position out
    of range -- make verify_distribution_shares / verify_share return False like the reference
    (bytes_to_element -> None, participant.rs:415-420), never raise."""
    import copy
    from oracle.groups import GROUPS, SECP_N
    og = GROUPS[gname]()
    g = m.Group(gname)
    n, t = 5, 3
    sks = synth.private_keys(12, n, gname, og.order())
    pks = g.fixed_base_exp(sks)
    d = m.Participant(g)
    box = d.distribute_secret(SECRET, pks, t, coeffs=synth.coefficients(12, t, og.order()),
                              witnesses=synth.witnesses(12, n, og.order()))
    assert d.verify_distribution_shares(box) is True
    bad_point = (b"\x05" + b"\x11" * 32) if gname == "secp256k1" else b"\xff" * 32
    for where in ("share", "commitment", "position"):
        evil = copy.copy(box)
        if where == "share":
            evil.shares = dict(box.shares)
            evil.shares[pks[1]] = bad_point
        elif where == "commitment":
            evil.commitments = [box.commitments[0], bad_point] + box.commitments[2:]
        else:
            evil.positions = dict(box.positions)
            evil.positions[pks[2]] = 1 << 40
        assert d.verify_distribution_shares(evil) is False, where
    sbs = d.extract_secret_shares(box, sks, synth.witnesses(13, n, og.order()))
    assert all(d.verify_shares(sbs, box, pks))
    sbs[1] = m.ShareBox(sbs[1].publickey, bad_point, sbs[1].challenge, sbs[1].response)
    assert d.verify_shares(sbs, box, pks) == [True, False, True, True, True]
    if gname == "secp256k1":                      # a response >= n cannot be a k256 Scalar
        c = g.codec
        ok = (ctypes.c_int * 1)()
        st = g.ctx.lib.mpvss_verify_shares(g.ctx.h, 1, ptr(buf(pks[0])), ptr(buf(sbs[0].share)), ptr(buf(box.shares[pks[0]])),
                                           ptr(buf(c.enc_scalar(sbs[0].challenge))), ptr(buf(SECP_N.to_bytes(32, "big"))), ok)
        assert st == L.OK and ok[0] == 0

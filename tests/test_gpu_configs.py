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

"""Wire format of the boxes (mpvss_rs_b200/wire.py): CPU round trips with a codec-only stand-in for the
group handle, and on the GPU verification straight from the serialised bytes."""
import types

import pytest

from mpvss_rs_b200 import wire
from mpvss_rs_b200.participant import DistributionSharesBox, ShareBox, _Codec
from mpvss_rs_b200 import synth
from oracle import pvss
from oracle.groups import GROUPS


def _oracle_box(name, n=4, t=3):
    og = GROUPS[name]()
    c = _Codec(name)
    sks = synth.private_keys(31, n, name, og.order(), getattr(og, "q", None))
    opks = [og.generate_public_key(s) for s in sks]
    obox = pvss.distribute_secret(og, 777, opks, t, synth.coefficients(31, t, og.order()),
                                  synth.witnesses(31, n, getattr(og, "q", og.order())))
    host = (lambda e: e) if name == "modp" else og.element_to_bytes
    box = DistributionSharesBox()
    box.commitments = [host(x) for x in obox.commitments]
    box.publickeys = [host(p) for p in opks]
    for p in opks:
        k = og.element_to_bytes(p)
        box.positions[k], box.shares[k], box.responses[k] = obox.positions[k], host(obox.shares[k]), obox.responses[k]
    box.challenge, box.U = obox.challenge, obox.U
    return c, box, og, sks


@pytest.mark.parametrize("name", ["modp", "secp256k1", "ristretto255"])
def test_box_round_trip(name):
    c, box, og, sks = _oracle_box(name)
    g = types.SimpleNamespace(codec=c)
    blob = wire.box_to_bytes(g, box)
    assert len(blob) == 20 + 3 * c.eb + 4 * 8 + 4 * c.eb * 2 + 4 * c.sb + c.sb + c.eb
    back = wire.box_from_bytes(g, blob)
    assert back == box
    with pytest.raises(ValueError):
        wire.box_from_bytes(g, blob[:-1])
    with pytest.raises(ValueError):
        wire.box_from_bytes(types.SimpleNamespace(codec=_Codec("modp" if name != "modp" else "secp256k1")), blob)
    sb = ShareBox(box.publickeys[0], box.publickeys[1], 12345, 67890)
    assert wire.sharebox_from_bytes(g, wire.sharebox_to_bytes(g, sb)) == sb


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["modp", "secp256k1", "ristretto255"])
def test_verify_from_bytes_on_gpu(name):
    import mpvss_rs_b200 as m
    c, box, og, sks = _oracle_box(name, n=6, t=4)
    g = m.Group(name)
    blob = wire.box_to_bytes(g, box)
    assert wire.verify_distribution_bytes(g, blob) is True
    assert m.Participant(g).verify_distribution_shares(wire.box_from_bytes(g, blob)) is True
    tampered = bytearray(blob)
    tampered[-c.eb - 1] ^= 1                      # last byte of the challenge
    assert wire.verify_distribution_bytes(g, bytes(tampered)) is False

"""Tiny full round for each group, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpvss_rs_b200 as m
from mpvss_rs_b200 import synth

for name in ("modp", "secp256k1", "ristretto255"):
    g = m.Group(name)
    c = g.codec
    n, t = 6, 4
    sks = synth.private_keys(1, n, name, c.order, c.key_bound)
    pks = g.fixed_base_exp(sks)
    d = m.Participant(g)
    box = d.distribute_secret(424242, pks, t, coeffs=synth.coefficients(1, t, c.order),
                              witnesses=synth.witnesses(1, n, c.key_bound))
    assert d.verify_distribution_shares(box)
    sbs = d.extract_secret_shares(box, sks, synth.witnesses(2, n, c.key_bound))
    assert all(d.verify_shares(sbs, box, pks))
    assert d.reconstruct(sbs[:t], box) == 424242
    g.ctx.set_int("device_hash", 1)     # the whole-box transcript through the device-side SHA-256 as well
    assert d.verify_distribution_shares(box)
    g.ctx.set_int("device_hash", 0)
    if name == "modp":
        for tpi in (4, 16, 8):
            g.ctx.set_int("modp_tpi", tpi)
            assert d.verify_distribution_shares(box)
    print(name, "ok")

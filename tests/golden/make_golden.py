"""Regenerates tests/golden/pvss_golden.json from the CPU oracle.

The reference cannot run here (no Rust toolchain), so the vectors come from the oracle
(oracle/groups.py, oracle/pvss.py), which is itself pinned by the reference's scalar KATs and by
RFC 3526 / SEC1 / RFC 9496 vectors (tests/test_oracle.py).  The file freezes complete protocol
transcripts for the three groups so that neither the oracle nor the CUDA path can drift silently.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mpvss_rs_b200 import synth  # noqa: E402
from oracle import pvss  # noqa: E402
from oracle.groups import GROUPS  # noqa: E402

CASES = [("modp", 4, 3, [0, 1, 3]), ("secp256k1", 4, 3, [0, 1, 3]), ("ristretto255", 4, 3, [0, 1, 3]),
         ("modp", 3, 3, [0, 1, 2]), ("secp256k1", 5, 3, [0, 2, 4]), ("ristretto255", 5, 2, [1, 4])]
SECRET = "Hello MPVSS Example."


def build(gname, n, t, subset, seed):
    g = GROUPS[gname]()
    E = lambda e: g.element_to_bytes(e).hex()
    sks = synth.private_keys(seed, n, gname, g.order(), getattr(g, "q", None))
    pks = [g.generate_public_key(s) for s in sks]
    co = synth.coefficients(seed, t, g.order())
    ws = synth.witnesses(seed, n, getattr(g, "q", g.order()))
    w2 = synth.witnesses(seed, n, getattr(g, "q", g.order()), "extract")
    secret = pvss.string_to_secret(SECRET)
    box = pvss.distribute_secret(g, secret, pks, t, co, ws)
    tr = {}
    assert pvss.verify_distribution_shares(g, box, trace=tr)
    sbs = [pvss.extract_secret_share(g, box, sks[i], w2[i]) for i in range(n)]
    assert all(pvss.verify_share(g, sbs[i], box, pks[i]) for i in range(n))
    rt = {}
    assert pvss.reconstruct(g, [sbs[i] for i in subset], box, trace=rt) == secret
    key = lambda pk: g.element_to_bytes(pk)
    return {
        "group": gname, "n": n, "t": t, "seed": seed, "subset": subset, "secret": SECRET,
        "private_keys": [hex(x) for x in sks], "coefficients": [hex(x) for x in co],
        "witnesses": [hex(x) for x in ws], "extract_witnesses": [hex(x) for x in w2],
        "publickeys": [E(p) for p in pks], "commitments": [E(c) for c in box.commitments],
        "shares": [E(box.shares[key(p)]) for p in pks],
        "responses": [hex(box.responses[key(p)]) for p in pks],
        "challenge": hex(box.challenge), "U": hex(box.U),
        "X": [E(x) for x in tr["X"]], "a1": [E(x) for x in tr["a1"]], "a2": [E(x) for x in tr["a2"]],
        "share_boxes": [{"share": E(s.share), "challenge": hex(s.challenge), "response": hex(s.response)} for s in sbs],
        "G_s": E(rt["G_s"]),
    }


if __name__ == "__main__":
    out = [build(g, n, t, sub, 1000 + i) for i, (g, n, t, sub) in enumerate(CASES)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pvss_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "cases")

"""Pins the CPU oracle: the reference's scalar known-answer tests, the RFC 3526 prime, SEC1 and
RFC 9496 vectors, OpenSSL cross-checks, and the reference's own protocol round-trip tests
restated with injected randomness (src/participant.rs:593-903, tests/mpvss_tests.rs:11-87)."""
import hashlib
import random

import pytest

from mpvss_rs_b200 import synth
from oracle import pvss
from oracle.groups import (ED_L, SECP_N, ModpGroup, Ristretto255Group, Secp256k1Group, ext_gcd,
                           lagrange_coefficient, mod_inverse, ristretto_decode, ristretto_eq)


# ---- reference KATs ----------------------------------------------------------------
def test_polynomial_kats():
    co = [3, 2, 2, 4]                                         # polynomial.rs:75-108
    assert [pvss.poly_get_value(co, x) for x in range(4)] == [3, 11, 47, 135]
    co = [105211, 1548877, 892134, 3490857, 324, 14234735]    # polynomial.rs:111-125
    assert pvss.poly_get_value(co, 278) % 15486967 == 4115179
    assert pvss.poly_eval_mod(co, 278, 15486967) == 4115179


def test_util_kats_verbatim():
    """The reference's own vectors, value for value (src/util.rs:84-138, 164-190)."""
    assert ext_gcd(26, 3) == (1, -1, 9)                       # util.rs:84-91: (g, x, y)
    assert 26 * -1 + 3 * 9 == 1
    assert mod_inverse(3, 26) == 9                            # util.rs:93-113
    assert mod_inverse(4, 32) is None
    values = [0, 1, 2, 3, 4, 5, 6]                            # util.rs:115-138
    assert lagrange_coefficient(9, values) == (0, 1)
    assert lagrange_coefficient(1, values) == (720, 120)
    assert lagrange_coefficient(2, values) == (360, -24)
    assert lagrange_coefficient(3, values) == (240, 12)
    assert lagrange_coefficient(3, [1, 3, 4]) == (4, -2)
    assert 1337 ^ 42 == 1299                                  # util.rs:155-161 (the U mask is this XOR)
    h = hashlib.sha256()                                      # util.rs:164-190: SHA-256 over decimal strings
    h.update(b"43589072349864890574839")
    h.update(b"14735247304952934566")
    assert h.hexdigest() == "e25e5b7edf4ea66e5238393fb4f183e0fc1593c69a522f9255a51bd0bc2b7ba7"
    assert int(h.hexdigest(), 16) == 102389418883295205726805934198606438410316463205994911160958467170744727731111


def test_util_kats_family():
    g, x, y = ext_gcd(240, 46)
    assert g == 2 and 240 * x + 46 * y == 2
    assert mod_inverse(4, 26) is None
    assert lagrange_coefficient(1, [1, 2, 3]) == (6, 2)       # prod j / prod (j - i)
    assert lagrange_coefficient(2, [1, 2, 3]) == (3, -1)
    assert lagrange_coefficient(3, [1, 2, 3]) == (2, 2)
    assert lagrange_coefficient(4, [1, 2, 3]) == (0, 1)


def test_dleq_init_kat_verbatim():
    """dleq.rs:356-377: a1 = g1^w, a2 = g2^w mod q for the reference's fixed small values."""
    g = ModpGroup()
    g1, g2, w = 8443, 1299721, 81647
    a1, a2 = pvss.prover_commitments(g, g1, g2, w) if hasattr(pvss, "prover_commitments") else (g.exp(g1, w), g.exp(g2, w))
    assert (a1, a2) == (pow(g1, w, g.q), pow(g2, w, g.q))


def test_ristretto_scalar_conversion_kats_verbatim():
    """ristretto255.rs:260-275, 697-717: BigInt -> Scalar -> BigInt round trips (big-endian in, LE inside)."""
    g = Ristretto255Group()
    for v in (0x0102030405060708, 123456789, 1 << 200):
        assert g.scalar_from_int(v) == v
        assert int.from_bytes(g.scalar_to_bytes(g.scalar_from_int(v)), "little") == v
    assert len(g.scalar_to_bytes(g.scalar_from_int(42))) == 32          # ristretto255.rs:690-696
    a, b, c = ED_L - 3, ED_L // 2 + 7, 11                                 # ristretto255.rs:340-376: sums agree
    assert (g.scalar_from_int(a) + g.scalar_from_int(b) + g.scalar_from_int(c)) % ED_L == g.scalar_from_int((a + b + c) % ED_L)


def test_dleq_response_kat():
    g = ModpGroup()                                           # dleq.rs:380-403
    w, alpha, c = 81647, 163027, 127997
    assert pvss.prover_response(g, w, alpha, c) == (w - alpha * c) % (g.q - 1)


def test_modp_group_constants():
    g = ModpGroup()
    # RFC 3526 section 3: 2^2048 - 2^1984 - 1 + 2^64 * ([2^1918 pi] + 124476)
    assert g.q.bit_length() == 2048 and g.q % (1 << 64) == (1 << 64) - 1 and g.q >> 1984 == (1 << 64) - 1
    assert hashlib.sha256(g.q.to_bytes(256, "big")).hexdigest() == hashlib.sha256(
        bytes.fromhex(__import__("oracle.groups", fromlist=["x"]).RFC3526_2048_HEX)).hexdigest()
    assert pow(2, g.g, g.q) == 1 and pow(4, g.g, g.q) == 1    # both generators have order g
    assert g.exp(g.generator(), 0) == 1 and g.mul(3, g.identity()) == 3   # modp.rs:233-257
    assert 0 <= g.hash_to_scalar(b"x") < g.g
    assert g.element_to_bytes(0) == b"\x00" and g.element_to_bytes(256) == b"\x01\x00"


def test_ristretto_order_constant():
    assert ED_L == int("1000000000000000000000000000000014def9dea2f79cd65812631a5cf5d3ed", 16)  # ristretto255.rs:378-401


# ---- external standards ---------------------------------------------------------------
def test_secp256k1_vectors():
    g = Secp256k1Group()
    enc = lambda k: g.element_to_bytes(g.exp(g.generator(), k)).hex()
    assert enc(1) == "0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"
    assert enc(2) == "02c6047f9441ed7d6d3045406e95c07cd85c778e4b8cef3ca7abac09b95c709ee5"
    assert enc(3) == "02f9308a019258c31049344f85f89d5229b531c845836f99b08601f113bce036f9"
    assert enc(SECP_N - 1) == "0379be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"
    p2 = g.exp(g.generator(), 2)
    assert g.mul(g.generator(), g.generator()) == p2          # secp256k1.rs:213-222
    assert g.mul(p2, g.element_inverse(p2)) is None
    assert g.bytes_to_element(g.element_to_bytes(p2)) == p2
    assert g.bytes_to_element(b"\x02" + b"\xff" * 32) is None


def test_secp256k1_against_openssl():
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.hazmat.primitives import serialization
    g = Secp256k1Group()
    rng = random.Random(1)
    for _ in range(8):
        k = rng.randrange(1, SECP_N)
        pub = ec.derive_private_key(k, ec.SECP256K1()).public_key()
        comp = pub.public_bytes(serialization.Encoding.X962, serialization.PublicFormat.CompressedPoint)
        assert g.element_to_bytes(g.exp(g.generator(), k)) == comp


RFC9496_MULTIPLES = [  # RFC 9496 appendix A.1, multiples 0..15 of the generator
    "0000000000000000000000000000000000000000000000000000000000000000",
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
    "da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57",
    "e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e",
    "f64746d3c92b13050ed8d80236a7f0007c3b3f962f5ba793d19a601ebb1df403",
    "44f53520926ec81fbd5a387845beb7df85a96a24ece18738bdcfa6a7822a176d",
    "903293d8f2287ebe10e2374dc1a53e0bc887e592699f02d077d5263cdd55601c",
    "02622ace8f7303a31cafc63f8fc48fdc16e1c8c8d234b2f0d6685282a9076031",
    "20706fd788b2720a1ed2a5dad4952b01f413bcf0e7564de8cdc816689e2db95f",
    "bce83f8ba5dd2fa572864c24ba1810f9522bc6004afe95877ac73241cafdab42",
    "e4549ee16b9aa03099ca208c67adafcafa4c3f3e4e5303de6026e3ca8ff84460",
    "aa52e000df2e16f55fb1032fc33bc42742dad6bd5a8fc0be0167436c5948501f",
    "46376b80f409b29dc2b5f6f0c52591990896e5716f41477cd30085ab7f10301e",
    "e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e",
]


def test_ristretto255_vectors():
    g = Ristretto255Group()
    for k, want in enumerate(RFC9496_MULTIPLES):
        p = g.exp(g.generator(), k)
        assert g.element_to_bytes(p).hex() == want
        q = ristretto_decode(bytes.fromhex(want))
        assert q is not None and ristretto_eq(p, q)
    # RFC 9496 A.2: a few encodings that must be rejected
    for bad in ["00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
                "0100000000000000000000000000000000000000000000000000000000000000",
                "edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f"]:
        assert ristretto_decode(bytes.fromhex(bad)) is None
    a = [5, 7, 11]                                            # ristretto255.rs:523-549
    cs = [g.exp(g.generator(), x) for x in a]
    assert ristretto_eq(g.mul(g.mul(cs[0], cs[1]), cs[2]), g.exp(g.generator(), sum(a)))
    assert g.scalar_from_int(ED_L + 5) == 5                   # ristretto255.rs:78-105


def _ed_compress(P):
    """standard Ed25519 encoding of an extended point: y with the sign of x in bit 255"""
    from oracle.groups import ED_P
    X, Y, Z, _ = P
    zi = pow(Z, -1, ED_P)
    x, y = X * zi % ED_P, Y * zi % ED_P
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def test_ristretto_edwards_arithmetic_against_libsodium():
    """The oracle's Edwards arithmetic under ristretto255 at FULL-WIDTH scalars against an independent
    implementation: libsodium's crypto_scalarmult_ed25519_noclamp / crypto_core_ed25519_add through PyNaCl
    (the bundled libsodium exports no ristretto255 API; the encodings are pinned by RFC 9496 above)."""
    nb = pytest.importorskip("nacl.bindings")
    g = Ristretto255Group()
    rng = random.Random(2024)
    B = g.generator()
    Benc = _ed_compress(B)
    assert Benc.hex() == "5866666666666666666666666666666666666666666666666666666666666666"  # Ed25519 basepoint
    pts = []
    for _ in range(12):
        k = rng.randrange(1, ED_L)
        P = g.exp(B, k)
        want = nb.crypto_scalarmult_ed25519_noclamp(k.to_bytes(32, "little"), Benc)
        assert _ed_compress(P) == want
        pts.append((k, P, want))
    for (k1, P1, e1), (k2, P2, e2) in zip(pts, pts[1:]):
        assert _ed_compress(g.mul(P1, P2)) == nb.crypto_core_ed25519_add(e1, e2)
        # variable-base: k2 * (k1 * B)
        assert _ed_compress(g.exp(P1, k2)) == nb.crypto_scalarmult_ed25519_noclamp(k2.to_bytes(32, "little"), e1)
    # the ristretto encoding of a point and of its libsodium twin decode to the same ristretto element
    k, P, e = pts[0]
    assert ristretto_eq(ristretto_decode(g.element_to_bytes(P)), P)


def test_secp256k1_dleq_commitments_against_openssl():
    """a1 = r*G + c*X, a2 = r*y + c*Y and the Horner X_i of the oracle against OpenSSL EC_POINT arithmetic
    (oracle/cpu_baseline.c), both schedules, at full-width scalars."""
    from oracle import cpu_baseline as cb
    g = Secp256k1Group()
    n, t = 5, 4
    sks = synth.private_keys(8, n, "secp256k1", g.order())
    pks = [g.generate_public_key(s) for s in sks]
    box = pvss.distribute_secret(g, 99, pks, t, synth.coefficients(8, t, g.order()), synth.witnesses(8, n, g.order()))
    tr = {}
    assert pvss.verify_distribution_shares(g, box, trace=tr)
    enc = g.element_to_bytes
    keys = [enc(pk) for pk in pks]
    for schedule in (0, 1):
        xs, a1, a2 = cb.secp_verify([enc(c) for c in box.commitments], list(range(1, n + 1)), keys,
                                    [enc(box.shares[k]) for k in keys], [box.responses[k] for k in keys],
                                    box.challenge, nthreads=2, schedule=schedule)
        assert xs == [enc(x) for x in tr["X"]] and a1 == [enc(x) for x in tr["a1"]] and a2 == [enc(x) for x in tr["a2"]]
    # the k256 identity encoding question (33 zero bytes) never arises inside a valid transcript
    assert all(e != bytes(33) for e in xs + a1 + a2)


def test_modp_transcript_against_openssl():
    """X_i, a1, a2 of the ModpGroup oracle against OpenSSL BN_mod_exp (both schedules)."""
    from oracle import cpu_baseline as cb
    g = ModpGroup()
    n, t = 4, 3
    sks = synth.private_keys(6, n, "modp", g.order(), g.q)
    pks = [g.generate_public_key(s) for s in sks]
    box = pvss.distribute_secret(g, 99, pks, t, synth.coefficients(6, t, g.order()), synth.witnesses(6, n, g.q))
    tr = {}
    assert pvss.verify_distribution_shares(g, box, trace=tr)
    keys = [g.element_to_bytes(pk) for pk in pks]
    for schedule in (0, 1):
        xs, a1, a2 = cb.modp_verify(g.q, box.commitments, list(range(1, n + 1)), pks, [box.shares[k] for k in keys],
                                    [box.responses[k] for k in keys], box.challenge, nthreads=2, schedule=schedule)
        assert (xs, a1, a2) == (tr["X"], tr["a1"], tr["a2"])


# ---- protocol round trips (reference tests restated with injected randomness) -------------
SECRET = pvss.string_to_secret("Hello MPVSS Example.")


@pytest.mark.parametrize("gname", ["modp", "secp256k1", "ristretto255"])
@pytest.mark.parametrize("n,t,subset", [(3, 3, [0, 1, 2]), (4, 3, [0, 1, 3]), (5, 3, [0, 2, 4])])
def test_round_trip(gname, n, t, subset):
    g = {"modp": ModpGroup, "secp256k1": Secp256k1Group, "ristretto255": Ristretto255Group}[gname]()
    sks = synth.private_keys(n, n, gname, g.order(), getattr(g, "q", None))
    pks = [g.generate_public_key(s) for s in sks]
    co = synth.coefficients(n, t, g.order())
    ws = synth.witnesses(n, n, getattr(g, "q", g.order()))
    box = pvss.distribute_secret(g, SECRET, pks, t, co, ws)
    assert pvss.verify_distribution_shares(g, box)
    # the GPU's Horner schedule gives the same X_i as the reference's product of t exponentiations
    for i in range(n):
        a = pvss.x_reference_schedule(g, box.commitments, i + 1)
        b = pvss.x_horner_schedule(g, box.commitments, i + 1)
        assert g.element_to_bytes(a) == g.element_to_bytes(b) == g.element_to_bytes(
            g.exp(g.subgroup_generator(), box.trace["p"][i]))
    sbs = [pvss.extract_secret_share(g, box, sks[i], ws[i]) for i in range(n)]
    assert all(pvss.verify_share(g, sbs[i], box, pks[i]) for i in range(n))
    assert pvss.reconstruct(g, [sbs[i] for i in subset], box) == SECRET
    assert pvss.reconstruct(g, sbs[: t - 1], box) is None
    # tampering
    k = g.element_to_bytes(pks[0])
    box.responses[k] = (box.responses[k] + 1) % g.order()
    assert not pvss.verify_distribution_shares(g, box)


def test_modp_threshold2_positions_1_3():
    g = ModpGroup()                                           # participant.rs:703-743 regression
    n, t = 3, 2
    sks = synth.private_keys(77, n, "modp", g.order(), g.q)
    pks = [g.generate_public_key(s) for s in sks]
    box = pvss.distribute_secret(g, SECRET, pks, t, synth.coefficients(77, t, g.order()), synth.witnesses(77, n, g.q))
    sbs = [pvss.extract_secret_share(g, box, sks[i], 12345 + i) for i in (0, 2)]
    assert pvss.reconstruct(g, sbs, box) == SECRET

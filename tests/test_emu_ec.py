"""CPU execution of the elliptic-curve *kernel bodies* (fp256.cuh, secp.cuh, rist.cuh,
ec_kernels.cuh -- the source nvcc compiles) through tests/emu, bit-exact against the oracle."""
import ctypes
import random

import numpy as np
import pytest

import emu_util as eu
from oracle import pvss
from oracle.groups import ED_L, SECP_N, Ristretto255Group, Secp256k1Group

U8 = lambda b: (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


@pytest.fixture(scope="module")
def L():
    return eu.build_ec()


@pytest.mark.parametrize("m", [eu.SECP_P, eu.SECP_N, eu.ED_P, eu.ED_L])
def test_fp256_field_ops(L, m):
    rng = random.Random(m & 0xFFFF)
    M = eu.modulus_words(m)
    n = 64
    A = [rng.randrange(m) for _ in range(n)]
    B = [rng.randrange(m) for _ in range(n)]
    A[:4] = [m - 1, 0, 1, m - 2]
    B[:4] = [m - 1, 5, 1, m - 1]
    a = np.concatenate([eu.to_limbs(x, 8) for x in A])
    b = np.concatenate([eu.to_limbs(x, 8) for x in B])
    o = [np.zeros(8 * n, dtype=np.uint32) for _ in range(4)]
    L.emu_fp_ops(eu.P(M), eu.P(a), eu.P(b), n, eu.P(o[0]), eu.P(o[1]), eu.P(o[2]), eu.P(o[3]))
    Ri = pow(1 << 256, -1, m)
    for i in range(n):
        assert eu.from_limbs(o[0][8 * i:8 * i + 8]) == A[i] * B[i] * Ri % m
        assert eu.from_limbs(o[1][8 * i:8 * i + 8]) == (A[i] + B[i]) % m
        assert eu.from_limbs(o[2][8 * i:8 * i + 8]) == (A[i] - B[i]) % m
        assert eu.from_limbs(o[3][8 * i:8 * i + 8]) == (pow(A[i], -1, m) if A[i] else 0)


@pytest.mark.parametrize("which,m", [(0, eu.SECP_P), (1, eu.ED_P)])
def test_special_form_field(L, which, m):
    """fpspecial.cuh: plain-representation product / square with the 2^256 = c fold, fully reduced."""
    rng = random.Random(which + 17)
    M = eu.modulus_words(m)
    edge = [0, 1, 2, m - 1, m - 2, (1 << 255) - 1 if m > (1 << 255) else m - 19, (1 << 128) - 1,
            ((1 << 256) - 1) % m, m // 2, m // 2 + 1, 0xFFFFFFFF, 0xFFFFFFFF00000000 % m]
    A = edge + [rng.randrange(m) for _ in range(200)]
    B = edge[::-1] + [rng.randrange(m) for _ in range(200)]
    for k in range(40):      # structured operands: runs of all-ones / zero limbs
        limbs = [rng.choice([0, 0xFFFFFFFF, 0xFFFFFFFE, 1, rng.getrandbits(32)]) for _ in range(8)]
        A.append(sum(l << (32 * i) for i, l in enumerate(limbs)) % m)
        B.append(sum(l << (32 * (7 - i)) for i, l in enumerate(limbs)) % m)
    # products congruent to a small value v: after the first fold the 256-bit remainder sits just
    # below 2^256, so the second fold wraps (secp: the carry handed to the conditional subtraction;
    # 25519: the extra bit-255 wrap) -- unreachable with random operands
    c = (1 << 256) % m
    for v in [c - 1, c, c + 1, 2 * c, 2 * c + 1, 977 * c, (1 << 31) * c, (1 << 32) * c - 1, 19, 18, 20, 37, 38, 39,
              (1 << 64) + c, 1, 0]:
        for _ in range(6):
            x = rng.randrange(1 << 200, m)
            A.append(x)
            B.append(v * pow(x, -1, m) % m)
    n = len(A)
    a = np.concatenate([eu.to_limbs(x, 8) for x in A])
    b = np.concatenate([eu.to_limbs(x, 8) for x in B])
    om, osq = np.zeros(8 * n, dtype=np.uint32), np.zeros(8 * n, dtype=np.uint32)
    L.emu_fpsp_ops(which, eu.P(M), eu.P(a), eu.P(b), n, eu.P(om), eu.P(osq))
    for i in range(n):
        assert eu.from_limbs(om[8 * i:8 * i + 8]) == A[i] * B[i] % m, (which, i)
        assert eu.from_limbs(osq[8 * i:8 * i + 8]) == A[i] * A[i] % m, (which, i)


CURVES = {
    "secp": (Secp256k1Group, eu.secp_consts, SECP_N, 33),
    "rist": (Ristretto255Group, eu.rist_consts, ED_L, 32),
}


@pytest.mark.parametrize("cv", ["secp", "rist"])
def test_point_kernels(L, cv):
    Gc, consts, order, EB = CURVES[cv]
    G = Gc()
    C = consts()
    assert getattr(L, f"emu_{cv}_sizeof_consts")() == C.size * 4
    exp2 = getattr(L, f"emu_{cv}_exp2")
    rng = random.Random(11)
    n = 7
    pts = [G.exp(G.generator(), rng.randrange(1, order)) for _ in range(n)]
    pts2 = [G.exp(G.generator(), rng.randrange(1, order)) for _ in range(n)]
    e1 = [rng.randrange(order) for _ in range(n)]
    e2 = [rng.randrange(order) for _ in range(n)]
    e1[0], e2[1], e1[2], e2[2] = 0, 0, 1, order - 1
    pts2[3], e2[3] = pts[3], order - e1[3]      # sums to the identity
    pts2[4], e2[4] = pts[4], e1[4]              # equal addends (doubling inside add)
    pts[5] = G.identity()                       # identity as a base
    enc = lambda ps: U8(b"".join(G.element_to_bytes(p) for p in ps))
    b1, b2 = enc(pts), enc(pts2)
    E1 = np.concatenate([eu.to_limbs(x, 8) for x in e1])
    E2 = np.concatenate([eu.to_limbs(x, 8) for x in e2])
    out = (ctypes.c_uint8 * (EB * n))()
    st = np.zeros(n, dtype=np.uint32)
    exp2(eu.P(C), b1, EB, eu.P(E1), 8, b2, EB, eu.P(E2), 8, n, out, eu.P(st))
    for i in range(n):
        want = G.element_to_bytes(G.mul(G.exp(pts[i], e1[i]), G.exp(pts2[i], e2[i])))
        assert bytes(out)[EB * i:EB * i + EB] == want, i
    assert not st.any()
    exp2(eu.P(C), b2, 0, eu.P(E1), 8, None, 0, None, 0, n, out, eu.P(st))   # one shared base
    for i in range(n):
        assert bytes(out)[EB * i:EB * i + EB] == G.element_to_bytes(G.exp(pts2[0], e1[i]))
    bad = bytearray(bytes(b2))
    bad[0] ^= 5                                                          # invalid prefix / odd s
    bad[EB + 1:2 * EB] = b"\xff" * (EB - 1)                              # x >= p / s >= p
    exp2(eu.P(C), U8(bytes(bad)), EB, eu.P(E1), 8, None, 0, None, 0, n, out, eu.P(st))
    assert list(st[:3]) == [1, 1, 0]
    getattr(L, f"emu_{cv}_add")(eu.P(C), b1, b2, n, out, eu.P(st))
    for i in range(n):
        assert bytes(out)[EB * i:EB * i + EB] == G.element_to_bytes(G.mul(pts[i], pts2[i]))


@pytest.mark.parametrize("cv", ["secp", "rist"])
def test_fixed_base_table(L, cv):
    """Generator table T[w][d-1] = (d * 16^w) G (affine), e * G from 64 lookups, and the DLEQ form
    e1 * G + e2 * B through it (Group::exp with a generator: secp256k1.rs:91-100 / ristretto255.rs:161-170)."""
    Gc, consts, order, EB = CURVES[cv]
    G = Gc()
    C = consts()
    rng = random.Random(23)
    n = 8
    e1 = [rng.randrange(order) for _ in range(n)]
    e2 = [rng.randrange(order) for _ in range(n)]
    e1[0], e1[1], e1[2], e1[3] = 0, 1, order - 1, 0xF << 252 if cv == "secp" else (1 << 252)
    e2[4] = 0
    pts2 = [G.exp(G.generator(), rng.randrange(1, order)) for _ in range(n)]
    pts2[5], e2[5] = G.generator(), order - e1[5]                  # e1 G + e2 G = identity
    pts2[6], e2[6] = G.generator(), e1[6]                          # equal addends
    gen = U8(G.element_to_bytes(G.generator()))
    b2 = U8(b"".join(G.element_to_bytes(p) for p in pts2))
    tbl = np.zeros(64 * 15 * 16, dtype=np.uint32)
    E1 = np.concatenate([eu.to_limbs(x % order, 8) for x in e1])
    E2 = np.concatenate([eu.to_limbs(x, 8) for x in e2])
    of, oe = (ctypes.c_uint8 * (EB * n))(), (ctypes.c_uint8 * (EB * n))()
    st = np.zeros(n, dtype=np.uint32)
    getattr(L, f"emu_{cv}_comb")(eu.P(C), gen, eu.P(tbl), eu.P(E1), b2, eu.P(E2), n, of, oe, eu.P(st))
    for i in range(n):
        k = e1[i] % order
        assert bytes(of)[EB * i:EB * i + EB] == G.element_to_bytes(G.exp(G.generator(), k)), i
        want = G.mul(G.exp(G.generator(), k), G.exp(pts2[i], e2[i]))
        assert bytes(oe)[EB * i:EB * i + EB] == G.element_to_bytes(want), i
    assert not st.any()


@pytest.mark.parametrize("cv", ["secp", "rist"])
@pytest.mark.parametrize("K", [1, 2, 7])
def test_chunked_horner_equals_reference_schedule(L, cv, K):
    Gc, consts, order, EB = CURVES[cv]
    G = Gc()
    C = consts()
    rng = random.Random(K)
    t = 7
    comm = [G.exp(G.generator(), rng.randrange(1, order)) for _ in range(t)]
    cb = U8(b"".join(G.element_to_bytes(p) for p in comm))
    positions = [1, 2, 3, 4, 5, 17, 255, 4096, 65536]
    pos = np.array(positions, dtype=np.uint32)
    n = len(positions)
    out = (ctypes.c_uint8 * (EB * n))()
    st = np.zeros(t, dtype=np.uint32)
    getattr(L, f"emu_{cv}_poly_eval_exp")(eu.P(C), cb, t, eu.P(pos), n, K, out, eu.P(st))
    for i in range(n):
        want = G.element_to_bytes(pvss.x_reference_schedule(G, comm, positions[i]))
        assert bytes(out)[EB * i:EB * i + EB] == want, (K, i)


@pytest.mark.parametrize("order", [SECP_N, ED_L])
def test_scalar_kernels(L, order):
    rng = random.Random(3)
    MN = eu.modulus_words(order)
    t = 9
    positions = [1, 2, 3, 1000, 65536]
    pos = np.array(positions, dtype=np.uint32)
    n = len(positions)
    co = [rng.randrange(order) for _ in range(t)]
    o = np.zeros(8 * n, dtype=np.uint32)
    L.emu_ec_poly(eu.P(MN), eu.P(np.concatenate([eu.to_limbs(x, 8) for x in co])), t, eu.P(pos), n, eu.P(o))
    for i in range(n):
        assert eu.from_limbs(o[8 * i:8 * i + 8]) == pvss.poly_get_value(co, positions[i]) % order
    # Lagrange coefficients: factors are folded eight at a time (plain 256-bit product, then one field product):
    # cover fewer than 8, exactly 8 + 1, several folds with a remainder, and factors near 2^31 (248-bit products)
    big = (1 << 31) - 1
    for vals in ([1, 3, 5, 6, 40], list(range(1, 10)), [big - 7 * j for j in range(18)] + [1, 2],
                 [rng.randrange(1, 1 << 31) for _ in range(41)], list(range(100, 133))):
        assert len(set(vals)) == len(vals)
        k = len(vals)
        o = np.zeros(8 * k, dtype=np.uint32)
        L.emu_ec_lagrange(eu.P(MN), eu.P(np.array(vals, dtype=np.uint32)), k, eu.P(o))
        for i, xi in enumerate(vals):
            num = den = 1
            for xj in vals:
                if xj != xi:
                    num = num * xj % order
                    den = den * (xj - xi) % order
            assert eu.from_limbs(o[8 * i:8 * i + 8]) == num * pow(den, -1, order) % order
    dup = [4, 9, 4, 11]           # duplicate position: zero denominator, lambda = 0 (ristretto255.rs:1983-1989)
    o = np.zeros(8 * 4, dtype=np.uint32)
    L.emu_ec_lagrange(eu.P(MN), eu.P(np.array(dup, dtype=np.uint32)), 4, eu.P(o))
    assert eu.from_limbs(o[0:8]) == 0 and eu.from_limbs(o[16:24]) == 0
    xs = [rng.randrange(order) for _ in range(5)] + [0]
    o = np.zeros(8 * 6, dtype=np.uint32)
    st = np.zeros(6, dtype=np.uint32)
    L.emu_ec_inv(eu.P(MN), eu.P(np.concatenate([eu.to_limbs(x, 8) for x in xs])), 6, eu.P(o), eu.P(st))
    for i, x in enumerate(xs):
        assert eu.from_limbs(o[8 * i:8 * i + 8]) == (pow(x, -1, order) if x else 0)
    assert list(st) == [0, 0, 0, 0, 0, 1]

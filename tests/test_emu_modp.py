"""CPU execution of the *CUDA kernel bodies* for ModpGroup through the lane-per-thread
emulator (tests/emu): the same modp_arith.cuh / modp_kernels.cuh source that nvcc compiles,
checked bit-exactly against Python integers / the oracle."""
import random

import numpy as np
import pytest

import emu_util as eu
from oracle import pvss
from oracle.groups import ModpGroup

G = ModpGroup()
Q = G.q
R = 1 << 2048


@pytest.fixture(scope="module")
def lib():
    return eu.build()


def _struct(rng):
    if rng.random() < 0.3:
        return rng.randrange(R)
    limbs = [rng.choice([0, 0xFFFFFFFF, 1, 0xFFFFFFFE, rng.getrandbits(32)]) for _ in range(64)]
    return sum(l << (32 * i) for i, l in enumerate(limbs))


@pytest.mark.parametrize("tpi", [4, 8, 16])
@pytest.mark.parametrize("modulus", ["q", "g"])
def test_mont_mul_edge_patterns(lib, tpi, modulus):
    m = Q if modulus == "q" else G.g
    C = eu.consts_block(m)
    rng = random.Random(tpi * 3 + len(modulus))
    n = 48
    A = [_struct(rng) for _ in range(n)]
    B = [_struct(rng) for _ in range(n)]
    A[:5] = [R - 1, 0, m - 1, m, m + 5]
    B[:5] = [R - 1, 7, m - 1, m, 1]
    a = np.concatenate([eu.to_limbs(x) for x in A])
    b = np.concatenate([eu.to_limbs(x) for x in B])
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_mul(tpi, eu.P(C), eu.P(a), 64, eu.P(b), 64, n, 0, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == A[i] * B[i] % m, (tpi, i)
    # mode 3: the same product through the split-accumulator loop (mont_mul_il)
    assert lib.emu_modp_mul(tpi, eu.P(C), eu.P(a), 64, eu.P(b), 64, n, 3, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == A[i] * B[i] % m, (tpi, i)
    # mode 1: Montgomery form a * 2^2048 mod m, canonical
    assert lib.emu_modp_mul(tpi, eu.P(C), eu.P(a), 64, None, 0, n, 1, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == A[i] * R % m, (tpi, i)


@pytest.mark.parametrize("tpi", [4, 8, 16])
@pytest.mark.parametrize("modulus", ["q", "g"])
def test_mont_sqr_edge_patterns(lib, tpi, modulus):
    """Dedicated squaring (half-square accumulation + doubling): a*a/R mod m for any a < 2^2048,
    odd and even a, saturated limbs, values at and above the modulus."""
    m = Q if modulus == "q" else G.g
    C = eu.consts_block(m)
    rng = random.Random(tpi * 7 + len(modulus))
    n = 96
    A = [_struct(rng) for _ in range(n)]
    A[:12] = [R - 1, 0, 1, 2, m - 1, m, m + 5, R - 2, (1 << 2047), (1 << 2047) + 1, (1 << 32) - 1, R - (1 << 31)]
    for k in range(8):                       # a single saturated limb in every lane block, odd and even
        A[12 + k] = 0xFFFFFFFF << (32 * (8 * k + (k % 8)))
        A[20 + k] = (0xFFFFFFFF << (32 * (8 * k))) | 1
    a = np.concatenate([eu.to_limbs(x) for x in A])
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_mul(tpi, eu.P(C), eu.P(a), 64, None, 0, n, 2, eu.P(out)) == 0
    rinv = pow(R, -1, m)
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == A[i] * A[i] * rinv % m, (tpi, i)


@pytest.mark.parametrize("modulus", ["q", "g"])
def test_split_sqr_edge_patterns(lib, modulus):
    """Split squaring (thread-local block products, column sums through shared memory, reduction-only
    digit loop; TPI = 8): a*a/R mod m on saturated / sparse / out-of-range operands, and three squarings
    in a row (scratch reuse)."""
    m = Q if modulus == "q" else G.g
    C = eu.consts_block(m)
    rng = random.Random(99 + len(modulus))
    n = 96
    A = [_struct(rng) for _ in range(n)]
    A[:12] = [R - 1, 0, 1, 2, m - 1, m, m + 5, R - 2, (1 << 2047), (1 << 2047) + 1, (1 << 32) - 1, R - (1 << 31)]
    for k in range(8):
        A[12 + k] = ((1 << 256) - 1) << (256 * k)          # one saturated block
        A[20 + k] = (((1 << 256) - 1) << (256 * k)) | ((1 << 256) - 1) << (256 * ((k + 4) % 8))
    a = np.concatenate([eu.to_limbs(x) for x in A])
    out = np.zeros(64 * n, dtype=np.uint32)
    rinv = pow(R, -1, m)
    assert lib.emu_modp_sqr_split(eu.P(C), eu.P(a), n, 1, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == A[i] * A[i] * rinv % m, i
    assert lib.emu_modp_sqr_split(eu.P(C), eu.P(a), n, 3, eu.P(out)) == 0
    for i in range(n):
        x = A[i]
        for _ in range(3):
            x = x * x * rinv % m
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == x, i


@pytest.mark.parametrize("tpi", [4, 8, 16])
def test_horner_kernel_equals_reference_schedule(lib, tpi):
    C = eu.consts_block(Q)
    rng = random.Random(tpi)
    t = 5
    comm = [pow(4, rng.randrange(Q - 1), Q) for _ in range(t)]
    cm = np.concatenate([eu.to_limbs(c * R % Q) for c in comm])
    positions = [1, 2, 3, 5, 11, 14, 15, 64, 255][: 9 if tpi != 4 else 9]
    nd = 4
    pos = np.array(positions, dtype=np.uint32)
    n = len(positions)
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_horner(tpi, eu.P(C), eu.P(cm), t, eu.P(pos), n, nd, eu.P(out), None) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pvss.x_reference_schedule(G, comm, positions[i]), (tpi, i)
    # block-uniform skipping of window multiplications whose digit is zero for the whole warp
    gpw = 32 // tpi
    positions = [64 + k for k in range(gpw)] + [80 + 16 * 0 + k for k in range(gpw)]   # digits (1,0,0,x), (1,1,0,x)
    pos = np.array(positions, dtype=np.uint32)
    n = len(positions)
    skip = np.array([0b0110 if gpw <= 4 else 0b0100, 0b0010 if gpw <= 4 else 0b0000], dtype=np.uint32)
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_horner(tpi, eu.P(C), eu.P(cm), t, eu.P(pos), n, 4, eu.P(out), eu.P(skip)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pvss.x_reference_schedule(G, comm, positions[i]), (tpi, i)


@pytest.mark.parametrize("tpi", [8, 16])
def test_exp2_kernel(lib, tpi):
    C = eu.consts_block(Q)
    rng = random.Random(40 + tpi)
    n = 6
    b1 = [rng.randrange(Q) for _ in range(n)]
    e1 = [rng.getrandbits(96) for _ in range(n)]
    b2 = [rng.randrange(Q) for _ in range(n)]
    e2 = [rng.getrandbits(40) for _ in range(n)]
    e1[0], e2[1], b1[2] = 0, 0, 1
    B1 = np.concatenate([eu.to_limbs(x) for x in b1])
    E1 = np.concatenate([eu.to_limbs(x) for x in e1])
    B2 = np.concatenate([eu.to_limbs(x) for x in b2])
    E2 = np.concatenate([eu.to_limbs(x, 8) for x in e2])
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_exp2(tpi, eu.P(C), eu.P(B1), 64, eu.P(E1), 64, 24, eu.P(B2), 64, eu.P(E2), 8, 10, n,
                             eu.P(out), None) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pow(b1[i], e1[i], Q) * pow(b2[i], e2[i], Q) % Q
    # single exponentiation with one shared base (stride 0)
    assert lib.emu_modp_exp2(tpi, eu.P(C), eu.P(B1), 0, eu.P(E1), 64, 24, None, 0, None, 0, 0, n, eu.P(out), None) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pow(b1[0], e1[i], Q)


def test_fixed_base_comb(lib):
    """Fixed-base table (8-bit comb) and the exponentiation that consumes it: g^e = prod_w T[w][byte_w(e)]."""
    C = eu.consts_block(Q)
    rng = random.Random(12)
    rows = 6                                        # exponents below 2^48 keep the emulated build short
    tbl = np.zeros(rows * 256 * 64, dtype=np.uint32)
    assert lib.emu_modp_comb_build(8, eu.P(C), eu.P(eu.to_limbs(4)), eu.P(tbl), rows) == 0
    for w, d in [(0, 0), (0, 1), (0, 255), (1, 1), (3, 200), (5, 255)]:
        got = eu.from_limbs(tbl[(w * 256 + d) * 64:(w * 256 + d) * 64 + 64])
        assert got == pow(4, d << (8 * w), Q) * R % Q, (w, d)
    n = 5
    e1 = [rng.getrandbits(48) for _ in range(n)]
    e1[0], e1[1], e1[2] = 0, 1, (1 << 48) - 1
    b2 = [rng.randrange(Q) for _ in range(n)]
    e2 = [rng.getrandbits(32) for _ in range(n)]
    E1 = np.concatenate([eu.to_limbs(x) for x in e1])
    B2 = np.concatenate([eu.to_limbs(x) for x in b2])
    E2 = np.concatenate([eu.to_limbs(x, 8) for x in e2])
    out = np.zeros(64 * n, dtype=np.uint32)
    for tpi in (8, 16):
        assert lib.emu_modp_exp2(tpi, eu.P(C), eu.P(E1), 0, eu.P(E1), 64, 12, eu.P(B2), 64, eu.P(E2), 8, 8, n,
                                 eu.P(out), eu.P(tbl)) == 0
        for i in range(n):
            assert eu.from_limbs(out[64 * i:64 * i + 64]) == pow(4, e1[i], Q) * pow(b2[i], e2[i], Q) % Q, (tpi, i)


def test_scalar_polynomial_kernel(lib):
    """P(i) mod (q-1) kernel against Polynomial::get_value(i) % order (polynomial.rs:50-58)."""
    rng = random.Random(21)
    order = Q - 1
    t = 9
    co = [rng.randrange(order) for _ in range(t)]
    co[0], co[1] = order - 1, (1 << 2048) - 1          # also a coefficient that is not reduced
    positions = [1, 2, 3, 255, 4096, 65536, (1 << 31) - 1]
    n = len(positions)
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_poly(eu.P(np.concatenate([eu.to_limbs(c) for c in co])), t, eu.P(eu.to_limbs(order)),
                             eu.P(np.array(positions, dtype=np.uint32)), n, eu.P(out)) == 0
    for i, x in enumerate(positions):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pvss.poly_get_value(co, x) % order, i


def test_lagrange_kernel(lib):
    """num_i, den_i mod (q-1) and the sign, against util.rs:47-64 (oracle lagrange_coefficient)."""
    from oracle.groups import lagrange_coefficient
    order = Q - 1
    values = [1, 3, 4, 9, 200, 4096, 65536, 7]
    k = len(values)
    num, den = np.zeros(64 * k, dtype=np.uint32), np.zeros(64 * k, dtype=np.uint32)
    neg = np.zeros(k, dtype=np.uint32)
    assert lib.emu_modp_lagrange(eu.P(eu.to_limbs(order)), eu.P(np.array(values, dtype=np.uint32)), k, eu.P(num),
                                 eu.P(den), eu.P(neg)) == 0
    for i, x in enumerate(values):
        n_, d_ = lagrange_coefficient(x, values)
        assert eu.from_limbs(num[64 * i:64 * i + 64]) == abs(n_) % order
        assert eu.from_limbs(den[64 * i:64 * i + 64]) == abs(d_) % order
        assert bool(neg[i]) == (n_ * d_ < 0)

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


def _chain_ops(hl, p, limit=1 << 12):
    """ops of one Horner step for position p from the product's host code (modp_chain.h)"""
    import ctypes
    ops = (ctypes.c_uint16 * 64)()
    sq, ml = ctypes.c_uint32(0), ctypes.c_uint32(0)
    n = hl.hc_chain_ops(p, limit, ops, ctypes.byref(sq), ctypes.byref(ml))
    assert n > 0, p
    return list(ops[:n]), sq.value, ml.value


def _run_ops(ops, x, cj, mul=lambda a, b: a * b % Q):
    """execute an op list the way modp::horner_body does, on Python integers"""
    slot, acc = [None] * 6, x
    for op in ops:
        sv, a, b = (op >> 8) & 15, (op >> 4) & 15, op & 15
        if sv:
            slot[sv - 1] = acc
        if a:
            acc = slot[a - 1]
        acc = mul(acc, 1 if b == 6 else cj if b == 7 else slot[b])
    return acc


def test_addition_chain_ops_compute_x_to_the_p_times_c():
    """Host-side chain builder (power tree below the limit, sliding windows above it): every op list
    evaluates acc^p * C_j within the kernel's 6 slots and 48 ops; products are counted."""
    import host_util
    hl = host_util.build()
    rng = random.Random(5)
    x, cj = rng.randrange(2, Q), rng.randrange(2, Q)
    ps = list(range(1, 300)) + [4095, 4096, 4097, 65535, 65536, 131071, 131072, 131073, (1 << 31) - 1, 0x6DB6DB6D,
                                0x55555555, 0x7FFFFFFE] + [rng.randrange(1, 1 << 17) for _ in range(300)]
    total_len = 0
    for p in ps:
        for limit in (1 << 12, 1 << 17):
            ops, sq, ml = _chain_ops(hl, p, limit)
            assert len(ops) <= 48 and sq + ml == len(ops) and ops[-1] == 7
            assert _run_ops(ops, x, cj) == pow(x, p, Q) * cj % Q, (p, limit)
            if limit == 1 << 17 and p <= 4096:
                total_len += len(ops) - 1
    # the power tree is at least as short as the fixed 2-bit windows it replaces (2 + 3 (d - 1) products)
    for p in (1, 2, 3, 4, 15, 255, 1000, 4095):
        d = 1
        while p >> (2 * d):
            d += 1
        assert len(_chain_ops(hl, p, 1 << 17)[0]) - 1 <= 2 + 3 * (d - 1)


@pytest.mark.parametrize("tpi", [4, 8, 16])
def test_horner_kernel_equals_reference_schedule(lib, tpi):
    """X_i = prod_j C_j^(i^j) by the chain kernel == the reference's t full exponentiations
    (participant.rs:423-434), every lane group of a warp running a different chain of equal length."""
    import host_util
    hl = host_util.build()
    C = eu.consts_block(Q)
    rng = random.Random(tpi)
    t = 4
    comm = [pow(4, rng.randrange(Q - 1), Q) for _ in range(t)]
    cm = np.concatenate([eu.to_limbs(c * R % Q) for c in comm])
    gpw = 32 // tpi
    positions = [1, 2, 3, 5, 11, 14, 15, 64, 255, 4096, 70000, 200001][:max(gpw, 4) + 2]
    lists = [_chain_ops(hl, p)[0] for p in positions]
    nops = max(len(l) for l in lists)
    ops = np.full(48 * len(positions), 6, dtype=np.uint16)          # padding: products by one
    for i, l in enumerate(lists):
        ops[48 * i:48 * i + len(l) - 1] = l[:-1]
        ops[48 * i + nops - 1] = 7                                  # the C_j product closes the step
    n = len(positions)
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_horner(tpi, eu.P(C), eu.P(cm), t, ops.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint16)),
                               n, nops, eu.P(out), t - 1, t - 1) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == pvss.x_reference_schedule(G, comm, positions[i]), (tpi, i)
    # a chunk of the polynomial: coefficients 1..2 only, H = C_1 * C_2^pos (what the chunked launch computes)
    assert lib.emu_modp_horner(tpi, eu.P(C), eu.P(cm), t, ops.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint16)),
                               n, nops, eu.P(out), 2, 1) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == comm[1] * pow(comm[2], positions[i], Q) % Q, (tpi, i)


def test_frame_kernel_minimal_big_endian():
    """Transcript rows from the device: len_u64_be || minimal big-endian bytes (dleq.rs:58-61,
    modp.rs:150-152), including short values and zero."""
    lib = eu.build()
    rng = random.Random(9)
    vals = [[rng.randrange(Q) for _ in range(4)] for _ in range(5)]
    vals[1] = [0, 1, 255, 256]
    vals[2] = [(1 << 2040) - 1, 1 << 2040, (1 << 2047), Q - 1]
    vals[3] = [rng.getrandbits(b) for b in (2039, 1000, 33, 8)]
    n = len(vals)
    cols = [np.concatenate([eu.to_limbs(v[k]) for v in vals]) for k in range(4)]
    out = np.zeros(n * 4 * 264, dtype=np.uint8)
    import ctypes
    assert lib.emu_modp_frames(eu.P(cols[0]), eu.P(cols[1]), eu.P(cols[2]), eu.P(cols[3]), n,
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))) == 0
    raw = out.tobytes()
    for j in range(n):
        for k in range(4):
            fr = raw[(4 * j + k) * 264:(4 * j + k + 1) * 264]
            body = G.element_to_bytes(vals[j][k])
            assert fr[:8] == len(body).to_bytes(8, "big") and fr[8:8 + len(body)] == body
            assert not any(fr[8 + len(body):])


def test_response_kernel(lib):
    """r = (w - (alpha * c mod (q-1))) mod (q-1) with the reference's scalar_sub (modp.rs:180-192)."""
    rng = random.Random(17)
    order = Q - 1
    n = 12
    alpha = [rng.randrange(order) for _ in range(n)]
    w = [rng.randrange(Q) for _ in range(n)]
    c = [rng.getrandbits(256) for _ in range(n)]
    alpha[0], w[0], c[0] = order - 1, 0, (1 << 256) - 1
    alpha[1], w[1] = (1 << 2048) - 1, (1 << 2048) - 1          # unreduced inputs
    alpha[2], c[2] = 0, 5
    w[3] = order
    c[4] = 0
    A = np.concatenate([eu.to_limbs(x) for x in alpha])
    W = np.concatenate([eu.to_limbs(x) for x in w])
    Cc = np.concatenate([eu.to_limbs(x) for x in c])
    out = np.zeros(64 * n, dtype=np.uint32)
    assert lib.emu_modp_resp(eu.P(eu.to_limbs(order)), eu.P(A), 64, eu.P(W), eu.P(Cc), 64, n, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == G.scalar_sub(w[i], G.scalar_mul(alpha[i], c[i])) % order, i
    # one shared challenge (distribute_secret)
    assert lib.emu_modp_resp(eu.P(eu.to_limbs(order)), eu.P(A), 64, eu.P(W), eu.P(Cc), 0, n, eu.P(out)) == 0
    for i in range(n):
        assert eu.from_limbs(out[64 * i:64 * i + 64]) == (w[i] - alpha[i] * c[0]) % order, i


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


@pytest.mark.parametrize("parts", [1, 3, 8])
def test_lagrange_kernel(lib, parts):
    """num_i, den_i mod (q-1) and the sign, against util.rs:47-64 (oracle lagrange_coefficient); the products are
    cut into `parts` ranges of j whose partial products (rows part * k + i) multiply to the full one, the partial
    signs XOR to the sign.  parts = 8 with k = 11 leaves empty ranges (product 1)."""
    from oracle.groups import lagrange_coefficient
    _lagrange_case(lib, parts, [1, 3, 4, 9, 200, 4096, 65536, 7, 12345, 2, 77], lagrange_coefficient)
    # positions near 2^31: the oracle's restatement of util.rs scans 1..max and cannot be asked here
    _lagrange_case(lib, parts, [(1 << 31) - 1, 5, (1 << 30) + 3, 17, (1 << 31) - 9, 2, 1, 99, 100], None)


def _lagrange_case(lib, parts, values, lagrange_coefficient):
    order = Q - 1
    k = len(values)
    num, den = np.zeros(64 * k * parts, dtype=np.uint32), np.zeros(64 * k * parts, dtype=np.uint32)
    neg = np.zeros(k * parts, dtype=np.uint32)
    assert lib.emu_modp_lagrange(eu.P(eu.to_limbs(order)), eu.P(np.array(values, dtype=np.uint32)), k, eu.P(num),
                                 eu.P(den), eu.P(neg), parts) == 0
    for i, x in enumerate(values):
        pn = pd = 1
        sign = 0
        for p in range(parts):
            r = p * k + i
            pn = pn * eu.from_limbs(num[64 * r:64 * r + 64]) % order
            pd = pd * eu.from_limbs(den[64 * r:64 * r + 64]) % order
            sign ^= int(neg[r])
        # the oracle's pair is reduced by its gcd; compare the quotient instead of the two products
        full_n = full_d = 1
        for y in values:
            if y != x:
                full_n, full_d = full_n * y, full_d * abs(y - x)
        assert pn == full_n % order and pd == full_d % order
        assert bool(sign) == (sum(1 for y in values if y < x) % 2 == 1)
        if lagrange_coefficient is not None:
            n_, d_ = lagrange_coefficient(x, values)
            assert full_n * abs(d_) == full_d * abs(n_)       # same rational number as util.rs returns
            assert bool(sign) == (n_ * d_ < 0)


def test_bucket_multi_exponentiation(lib):
    """Pippenger buckets (8-bit windows): prod_i S_i^(e_i) for 3-byte exponents, empty and crowded buckets,
    zero digits and a zero exponent (the reconstruct fold, participant.rs:490-509)."""
    C = eu.consts_block(Q)
    rng = random.Random(31)
    k, windows = 21, 3
    bases = [rng.randrange(2, Q) for _ in range(k)]
    exps = [rng.getrandbits(24) for _ in range(k)]
    exps[0], exps[1], exps[2], exps[3] = 0, 0xFF00FF, 0x000001, 0xFFFFFF
    for i in range(4, 10):
        exps[i] = (exps[i] & 0xFFFF00) | 0x2A                    # a crowded bucket in window 0
    idx = np.zeros(windows * k, dtype=np.uint32)
    start = np.zeros(windows * 257, dtype=np.uint32)
    for w in range(windows):
        order = sorted(range(k), key=lambda i: ((exps[i] >> (8 * w)) & 0xFF, i))
        idx[w * k:(w + 1) * k] = order
        digs = [(exps[i] >> (8 * w)) & 0xFF for i in order]
        for d in range(257):
            start[w * 257 + d] = sum(1 for x in digs if x < d)
    bm = np.concatenate([eu.to_limbs(b * R % Q) for b in bases])
    buckets = np.zeros(windows * 256 * 64, dtype=np.uint32)
    wprod = np.zeros(windows * 64, dtype=np.uint32)
    out = np.zeros(64, dtype=np.uint32)
    assert lib.emu_modp_msm(eu.P(C), eu.P(bm), eu.P(idx), eu.P(start), windows, k, eu.P(buckets), eu.P(wprod),
                            eu.P(out)) == 0
    want = 1
    for b, e in zip(bases, exps):
        want = want * pow(b, e, Q) % Q
    assert eu.from_limbs(out) == want

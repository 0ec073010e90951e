"""CPU oracle: the three `Group` back-ends of the reference, restated in Python.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker.

PARITY STATUS: *partly pinned*.  The reference (AlexiaChen/mpvss-rs) cannot be
compiled here (no Rust toolchain, crates not vendored, no Cargo.lock) and its own
tests hold no golden group element / transcript / challenge.  The arithmetic
lives in third-party crates that are absent from /root/reference:
``num-bigint 0.2`` (Cargo.toml:15), ``k256 0.13`` (Cargo.toml:24),
``curve25519-dalek 4`` (Cargo.toml:27), ``sha2 0.10`` (Cargo.toml:21).  This file
restates their *published* algorithms (modular exponentiation, SEC1 secp256k1,
RFC 9496 ristretto255, FIPS 180-4) and is pinned by: the reference's scalar-side
known-answer tests (polynomial.rs:111-125, util.rs:84-138, dleq.rs:380-403,
ristretto255.rs:378-401), the RFC 3526 group-14 prime (modp.rs:47-58), the SEC1
generator multiples and the RFC 9496 generator multiples.  Protocol-level
transcripts stay "parity unpinned" against the Rust binary itself.

Every method cites the reference line it follows (paths relative to
/root/reference/src).
"""
from __future__ import annotations

import hashlib

# ---------------------------------------------------------------------------
# ModpGroup  (groups/modp.rs)
# ---------------------------------------------------------------------------

RFC3526_2048_HEX = (  # groups/modp.rs:47-56
    "ffffffffffffffffc90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74"
    "020bbea63b139b22514a08798e3404ddef9519b3cd3a431b302b0a6df25f1437"
    "4fe1356d6d51c245e485b576625e7ec6f44c42e9a637ed6b0bff5cb6f406b7ed"
    "ee386bfb5a899fa5ae9f24117c4b1fe649286651ece45b3dc2007cb8a163bf05"
    "98da48361c55d39a69163fa8fd24cf5f83655d23dca3ad961c62f356208552bb"
    "9ed529077096966d670c354e4abc9804f1746c08ca18217c32905e462e36ce3b"
    "e39e772c180e86039b2783a2ec07a28fb5c55df06f4c52c9de2bcbf695581718"
    "3995497cea956ae515d2261898fa051015728e5a8aacaa68ffffffffffffffff"
)


def _int_to_min_be(x: int) -> bytes:
    """num-bigint ``BigUint::to_bytes_be``: minimal length, zero -> [0]."""
    if x == 0:
        return b"\x00"
    return x.to_bytes((x.bit_length() + 7) // 8, "big")


def ext_gcd(a: int, b: int):
    """util.rs:18-25 (iterative form of the same recurrence)."""
    x0, x1, y0, y1 = 1, 0, 0, 1
    while b:
        # Python's // floors, Rust's BigInt `/` truncates; identical for the
        # non-negative operands the reference ever passes.
        qt = a // b
        a, b = b, a - qt * b
        x0, x1 = x1, x0 - qt * x1
        y0, y1 = y1, y0 - qt * y1
    return a, x0, y0


def mod_inverse(a: int, m: int):
    """util.rs:33-41 -- ``None`` unless gcd(a, m) == 1."""
    g, x, _ = ext_gcd(a % m if a >= 0 else a, m)
    if g != 1:
        return None
    return (x % m + m) % m


def lagrange_coefficient(i: int, values):
    """util.rs:47-64 -- (numerator, denominator) as signed integers."""
    if i not in values:
        return 0, 1
    num, den = 1, 1
    vs = set(values)
    for j in range(1, max(values) + 1):
        if j != i and j in vs:
            num *= j
            den *= j - i
    return num, den


class ModpGroup:
    """groups/modp.rs:42-197 -- Z_q^*, q the RFC 3526 2048-bit safe prime."""

    name = "modp"

    def __init__(self, q: int | None = None):
        self.q = int(RFC3526_2048_HEX, 16) if q is None else q  # modp.rs:47-58
        self.g = (self.q - 1) // 2          # modp.rs:59  subgroup order
        self.G = 2                          # modp.rs:64  main generator
        self.g_gen = pow(2, 2, self.q)      # modp.rs:65-66  subgroup generator
        self.q_minus_1 = self.q - 1         # modp.rs:67

    # -- trait Group (group.rs:24-124) --
    def order(self):                        # modp.rs:101-103
        return self.q_minus_1

    def subgroup_order(self):               # modp.rs:105-107
        return self.g

    def challenge_modulus(self):            # hash_to_scalar reduces mod g, modp.rs:145
        return self.g

    def generator(self):                    # modp.rs:109-111
        return self.G

    def subgroup_generator(self):           # modp.rs:113-116
        return self.g_gen

    def identity(self):                     # modp.rs:118-120
        return 1

    def exp(self, base, scalar):            # modp.rs:122-128
        return pow(base, scalar, self.q)

    def mul(self, a, b):                    # modp.rs:130-132
        return (a * b) % self.q

    def scalar_inverse(self, x):            # modp.rs:134-136
        return mod_inverse(x, self.q_minus_1)

    def element_inverse(self, x):           # modp.rs:138-140
        return mod_inverse(x, self.q)

    def hash_to_scalar(self, data: bytes):  # modp.rs:142-148
        return int.from_bytes(hashlib.sha256(data).digest(), "big") % self.g

    def element_to_bytes(self, e):          # modp.rs:150-152
        return _int_to_min_be(e)

    def bytes_to_element(self, b: bytes):   # modp.rs:154-156
        return int.from_bytes(b, "big")

    def scalar_to_bytes(self, s):           # modp.rs:158-160
        return _int_to_min_be(s)

    def generate_public_key(self, sk):      # modp.rs:176-178
        return self.exp(self.G, sk)

    def scalar_mul(self, a, b):             # modp.rs:180-182
        return (a * b) % self.q_minus_1

    def scalar_sub(self, a, b):             # modp.rs:184-192
        d = a - b
        if d < 0:
            return d + self.q_minus_1
        return d % self.q_minus_1

    def scalar_from_int(self, x: int):      # coefficients are used as-is
        return x

    def scalar_from_small(self, x: int):
        return x

    def mask_of(self, elem):
        """participant.rs:268-271 -- int(SHA-256(bytes(G^s))) mod q."""
        h = hashlib.sha256(self.element_to_bytes(elem)).digest()
        return int.from_bytes(h, "big") % self.q


# ---------------------------------------------------------------------------
# Secp256k1Group  (groups/secp256k1.rs; arithmetic = k256 0.13, SEC1/SEC2)
# ---------------------------------------------------------------------------

SECP_P = 2**256 - 2**32 - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141  # secp256k1.rs:47-52
SECP_GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
SECP_GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8


def _secp_add(P, Q):
    """Affine addition on y^2 = x^3 + 7; ``None`` is the identity."""
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    p = SECP_P
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return x3, (lam * (x1 - x3) - y1) % p


def _secp_jac_dbl(P):
    X, Y, Z = P
    p = SECP_P
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def _secp_jac_add_affine(P, Q):
    """Jacobian P + affine Q (Q not identity)."""
    X1, Y1, Z1 = P
    p = SECP_P
    if Z1 == 0:
        return (Q[0], Q[1], 1)
    x2, y2 = Q
    Z1Z1 = Z1 * Z1 % p
    U2 = x2 * Z1Z1 % p
    S2 = y2 * Z1 * Z1Z1 % p
    H = (U2 - X1) % p
    r = (S2 - Y1) % p
    if H == 0:
        if r == 0:
            return _secp_jac_dbl(P)
        return (1, 1, 0)
    HH = H * H % p
    HHH = H * HH % p
    V = X1 * HH % p
    X3 = (r * r - HHH - 2 * V) % p
    Y3 = (r * (V - X3) - Y1 * HHH) % p
    Z3 = Z1 * H % p
    return (X3, Y3, Z3)


def _secp_mul(k: int, P):
    """k*P by double-and-add in Jacobian coordinates; affine in/out."""
    k %= SECP_N
    if P is None or k == 0:
        return None
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _secp_jac_dbl(acc)
        if bit == "1":
            acc = _secp_jac_add_affine(acc, P)
    X, Y, Z = acc
    if Z == 0:
        return None
    zi = pow(Z, -1, SECP_P)
    zi2 = zi * zi % SECP_P
    return X * zi2 % SECP_P, Y * zi2 * zi % SECP_P


class Secp256k1Group:
    """groups/secp256k1.rs:33-188.  Elements: affine (x, y) or None (identity)."""

    name = "secp256k1"

    def __init__(self):
        self.n = SECP_N

    def order(self):                        # order_as_bigint, secp256k1.rs:186-188
        return self.n

    def subgroup_order(self):
        return self.n

    def challenge_modulus(self):
        return self.n

    def generator(self):                    # secp256k1.rs:78-80
        return (SECP_GX, SECP_GY)

    def subgroup_generator(self):           # secp256k1.rs:82-85
        return (SECP_GX, SECP_GY)

    def identity(self):                     # secp256k1.rs:87-89
        return None

    def exp(self, base, scalar):            # secp256k1.rs:91-100
        return _secp_mul(scalar, base)

    def mul(self, a, b):                    # secp256k1.rs:102-107
        return _secp_add(a, b)

    def scalar_inverse(self, x):            # secp256k1.rs:109-112
        x %= self.n
        return None if x == 0 else pow(x, -1, self.n)

    def element_inverse(self, e):           # secp256k1.rs:114-119
        return None if e is None else (e[0], (-e[1]) % SECP_P)

    def hash_to_scalar(self, data: bytes):  # secp256k1.rs:121-131
        return int.from_bytes(hashlib.sha256(data).digest(), "big") % self.n

    def element_to_bytes(self, e):          # secp256k1.rs:133-136 (SEC1 compressed)
        if e is None:
            # k256 encodes the identity as a 33-byte all-zero string (recalled from
            # upstream, not in tree); only reachable with degenerate inputs.
            return b"\x00" * 33
        return bytes([2 + (e[1] & 1)]) + e[0].to_bytes(32, "big")

    def bytes_to_element(self, b: bytes):   # secp256k1.rs:138-152
        if len(b) != 33 or b[0] not in (2, 3):
            return None
        x = int.from_bytes(b[1:], "big")
        if x >= SECP_P:
            return None
        y2 = (pow(x, 3, SECP_P) + 7) % SECP_P
        y = pow(y2, (SECP_P + 1) // 4, SECP_P)
        if y * y % SECP_P != y2:
            return None
        if (y & 1) != (b[0] & 1):
            y = SECP_P - y
        return (x, y)

    def scalar_to_bytes(self, s):           # secp256k1.rs:154-156
        return (s % self.n).to_bytes(32, "big")

    def generate_public_key(self, sk):      # secp256k1.rs:168-171
        return _secp_mul(sk, self.generator())

    def scalar_mul(self, a, b):             # secp256k1.rs:173-176
        return a * b % self.n

    def scalar_sub(self, a, b):             # secp256k1.rs:178-181
        return (a - b) % self.n

    def scalar_from_int(self, x: int):
        """participant.rs:1134-1143 -- right-aligned 32-byte BE -> from_repr().unwrap()."""
        if not 0 <= x < self.n:
            raise ValueError("Scalar::from_repr(...).unwrap() would panic")
        return x

    def scalar_from_small(self, x: int):    # Scalar::from(u64)
        return x % self.n

    def mask_of(self, elem):
        """participant.rs:1244-1259 -- SHA-256 -> from_repr().unwrap() -> mod n."""
        h = int.from_bytes(hashlib.sha256(self.element_to_bytes(elem)).digest(), "big")
        if h >= self.n:
            raise ValueError("Scalar::from_repr(...).unwrap() would panic")
        return h % self.n


# ---------------------------------------------------------------------------
# Ristretto255Group  (groups/ristretto255.rs; arithmetic = curve25519-dalek 4, RFC 9496)
# ---------------------------------------------------------------------------

ED_P = 2**255 - 19
ED_L = 2**252 + 27742317777372353535851937790883648493  # ristretto255.rs:55-60
ED_D = (-121665 * pow(121666, -1, ED_P)) % ED_P
ED_SQRT_M1 = pow(2, (ED_P - 1) // 4, ED_P)
ED_INVSQRT_A_MINUS_D = None  # filled below
ED_BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
ED_BY = 46316835694926478169428394003475163141307993866256225615783033603165251855960


def _is_neg(x: int) -> bool:
    return (x % ED_P) & 1 == 1


def _ct_abs(x: int) -> int:
    x %= ED_P
    return ED_P - x if x & 1 else x


def _sqrt_ratio_m1(u: int, v: int):
    """RFC 9496 §4.2 SQRT_RATIO_M1."""
    p = ED_P
    u %= p
    v %= p
    v3 = v * v % p * v % p
    v7 = v3 * v3 % p * v % p
    r = u * v3 % p * pow(u * v7 % p, (p - 5) // 8, p) % p
    check = v * r % p * r % p
    correct = check == u
    flipped = check == (-u) % p
    flipped_i = check == (-u * ED_SQRT_M1) % p
    if flipped or flipped_i:
        r = r * ED_SQRT_M1 % p
    r = _ct_abs(r)
    return (correct or flipped), r


ED_INVSQRT_A_MINUS_D = _sqrt_ratio_m1(1, (-1 - ED_D) % ED_P)[1]


def _ed_add(P, Q):
    """Extended twisted Edwards a=-1 unified addition (add-2008-hwcd-3)."""
    X1, Y1, Z1, T1 = P
    X2, Y2, Z2, T2 = Q
    p = ED_P
    A = (Y1 - X1) * (Y2 - X2) % p
    B = (Y1 + X1) * (Y2 + X2) % p
    C = T1 * 2 * ED_D % p * T2 % p
    D = Z1 * 2 * Z2 % p
    E, F, G, H = B - A, D - C, D + C, B + A
    return (E * F % p, G * H % p, F * G % p, E * H % p)


def _ed_mul(k: int, P):
    k %= ED_L
    acc = (0, 1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _ed_add(acc, acc)
        if bit == "1":
            acc = _ed_add(acc, P)
    return acc


def ristretto_encode(P) -> bytes:
    """RFC 9496 §4.3.2."""
    X0, Y0, Z0, T0 = P
    p = ED_P
    u1 = (Z0 + Y0) * (Z0 - Y0) % p
    u2 = X0 * Y0 % p
    _, invsqrt = _sqrt_ratio_m1(1, u1 * u2 % p * u2 % p)
    den1 = invsqrt * u1 % p
    den2 = invsqrt * u2 % p
    z_inv = den1 * den2 % p * T0 % p
    ix0 = X0 * ED_SQRT_M1 % p
    iy0 = Y0 * ED_SQRT_M1 % p
    enchanted = den1 * ED_INVSQRT_A_MINUS_D % p
    rotate = _is_neg(T0 * z_inv)
    if rotate:
        x, y, den_inv = iy0, ix0, enchanted
    else:
        x, y, den_inv = X0, Y0, den2
    if _is_neg(x * z_inv):
        y = (-y) % p
    s = _ct_abs(den_inv * (Z0 - y) % p)
    return s.to_bytes(32, "little")


def ristretto_decode(b: bytes):
    """RFC 9496 §4.3.1 -- ``None`` for non-canonical / invalid encodings."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= ED_P or _is_neg(s):
        return None
    p = ED_P
    ss = s * s % p
    u1 = (1 - ss) % p
    u2 = (1 + ss) % p
    u2_sqr = u2 * u2 % p
    v = (-(ED_D * u1 % p * u1) - u2_sqr) % p
    was_square, invsqrt = _sqrt_ratio_m1(1, v * u2_sqr % p)
    den_x = invsqrt * u2 % p
    den_y = invsqrt * den_x % p * v % p
    x = _ct_abs(2 * s * den_x % p)
    y = u1 * den_y % p
    t = x * y % p
    if (not was_square) or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def ristretto_eq(P, Q) -> bool:
    """RFC 9496 §4.3.3."""
    X1, Y1, _, _ = P
    X2, Y2, _, _ = Q
    return (X1 * Y2 - Y1 * X2) % ED_P == 0 or (Y1 * Y2 - X1 * X2) % ED_P == 0


class Ristretto255Group:
    """groups/ristretto255.rs:40-253.  Elements: extended (X, Y, Z, T) tuples."""

    name = "ristretto255"

    def __init__(self):
        self.l = ED_L

    def order(self):                        # order_as_bigint, ristretto255.rs:65-67
        return self.l

    def subgroup_order(self):
        return self.l

    def challenge_modulus(self):
        return self.l

    def generator(self):                    # ristretto255.rs:148-150
        return (ED_BX, ED_BY, 1, ED_BX * ED_BY % ED_P)

    def subgroup_generator(self):           # ristretto255.rs:152-155
        return self.generator()

    def identity(self):                     # ristretto255.rs:157-159
        return (0, 1, 1, 0)

    def exp(self, base, scalar):            # ristretto255.rs:161-170
        return _ed_mul(scalar, base)

    def mul(self, a, b):                    # ristretto255.rs:172-177
        return _ed_add(a, b)

    def scalar_inverse(self, x):            # ristretto255.rs:179-187
        x %= self.l
        return None if x == 0 else pow(x, -1, self.l)

    def element_inverse(self, e):           # ristretto255.rs:189-194
        X, Y, Z, T = e
        return ((-X) % ED_P, Y, Z, (-T) % ED_P)

    def hash_to_scalar(self, data: bytes):  # ristretto255.rs:196-205 (SHA-512, LE, wide)
        return int.from_bytes(hashlib.sha512(data).digest(), "little") % self.l

    def element_to_bytes(self, e):          # ristretto255.rs:207-210
        return ristretto_encode(e)

    def bytes_to_element(self, b: bytes):   # ristretto255.rs:212-220
        return ristretto_decode(b)

    def scalar_to_bytes(self, s):           # ristretto255.rs:222-225
        return (s % self.l).to_bytes(32, "little")

    def generate_public_key(self, sk):      # ristretto255.rs:238-242
        return _ed_mul(sk, self.generator())

    def scalar_mul(self, a, b):             # ristretto255.rs:244-247
        return a * b % self.l

    def scalar_sub(self, a, b):             # ristretto255.rs:249-252
        return (a - b) % self.l

    def scalar_from_int(self, x: int):
        """ristretto255.rs:78-105 -- BE bytes truncated to the first 32, reversed,
        ``from_bytes_mod_order``.  For x < 2^256 this is x mod l."""
        be = _int_to_min_be(x) if x else b""
        n = min(len(be), 32)
        le = bytes(be[n - 1 - i] for i in range(n)) + b"\x00" * (32 - n)
        return int.from_bytes(le, "little") % self.l

    def scalar_from_small(self, x: int):    # Scalar::from(u64)
        return x % self.l

    def mask_of(self, elem):
        """participant.rs:1696-1702 -- int_be(SHA-256(bytes)) mod l."""
        h = int.from_bytes(hashlib.sha256(self.element_to_bytes(elem)).digest(), "big")
        return h % self.l


def elements_equal(group, a, b) -> bool:
    """Element equality as the reference's ``Eq`` impls see it."""
    if group.name == "ristretto255":
        return ristretto_eq(a, b)
    return a == b


GROUPS = {"modp": ModpGroup, "secp256k1": Secp256k1Group, "ristretto255": Ristretto255Group}

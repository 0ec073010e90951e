"""CPU tests of the product's host-side helpers (bigint.h, sha2.h) against Python."""
import ctypes
import hashlib
import math
import random

import host_util


def test_bigint_divmod_modinv_mulmod():
    L = host_util.build()
    rng = random.Random(5)
    N = 520
    for it in range(1500):
        abits = rng.choice([1, 31, 32, 33, 64, 100, 2048, 4096, 4100])
        bbits = rng.choice([1, 5, 32, 33, 63, 64, 65, 2047, 2048])
        a = rng.getrandbits(abits)
        b = rng.getrandbits(bbits) | 1
        if it % 7 == 0:
            b = (1 << bbits) - 1
        if it % 11 == 0:
            a = (1 << abits) - 1
        q, r = (ctypes.c_uint8 * N)(), (ctypes.c_uint8 * N)()
        L.hc_divmod(host_util.le(a, N), N, host_util.le(b, N), N, q, r, N)
        assert int.from_bytes(bytes(q), "little") == a // b
        assert int.from_bytes(bytes(r), "little") == a % b
        m = rng.getrandbits(bbits) | (1 << (bbits - 1)) | 1
        if m == 1:
            continue
        out = (ctypes.c_uint8 * N)()
        rc = L.hc_modinv(host_util.le(a, N), N, host_util.le(m, N), N, out, N)
        if math.gcd(a, m) == 1:
            assert rc == 0 and int.from_bytes(bytes(out), "little") == pow(a, -1, m)
        else:
            assert rc == 1
        # binary extended Euclid for odd moduli (the root of the device-side batch inversion): same answers
        out2 = (ctypes.c_uint8 * N)()
        rc2 = L.hc_modinv_odd(host_util.le(a, N), N, host_util.le(m, N), N, out2, N)
        assert rc2 == rc and (rc or bytes(out2) == bytes(out))
        c = rng.getrandbits(2048)
        L.hc_mulmod(host_util.le(a, N), N, host_util.le(c, N), N, host_util.le(m, N), N, out, N)
        assert int.from_bytes(bytes(out), "little") == a * c % m
    # even modulus (q-1 of the MODP group is even): inverse exists iff coprime
    out = (ctypes.c_uint8 * N)()
    assert L.hc_modinv(host_util.le(7, N), N, host_util.le(30, N), N, out, N) == 0
    assert int.from_bytes(bytes(out), "little") == 13
    assert L.hc_modinv(host_util.le(6, N), N, host_util.le(30, N), N, out, N) == 1
    assert L.hc_modinv_odd(host_util.le(7, N), N, host_util.le(30, N), N, out, N) == 1      # even modulus: refused
    g = (int("ffffffffffffffffc90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74020bbea63b139b22514a08798e3404dd"
             "ef9519b3cd3a431b302b0a6df25f14374fe1356d6d51c245e485b576625e7ec6f44c42e9a637ed6b0bff5cb6f406b7ed"
             "ee386bfb5a899fa5ae9f24117c4b1fe649286651ece45b3dc2007cb8a163bf0598da48361c55d39a69163fa8fd24cf5f"
             "83655d23dca3ad961c62f356208552bb9ed529077096966d670c354e4abc9804f1746c08ca18217c32905e462e36ce3b"
             "e39e772c180e86039b2783a2ec07a28fb5c55df06f4c52c9de2bcbf6955817183995497cea956ae515d2261898fa0510"
             "15728e5a8aacaa68ffffffffffffffff", 16) - 1) // 2
    for a in (1, 2, g - 1, g - 2, rng.randrange(g), g + 5, 3 * g + 1):
        assert L.hc_modinv_odd(host_util.le(a, N), N, host_util.le(g, N), N, out, N) == 0
        assert int.from_bytes(bytes(out), "little") == pow(a, -1, g)
    assert L.hc_modinv_odd(host_util.le(2 * g, N), N, host_util.le(g, N), N, out, N) == 1   # a = 0 mod m


def test_sha2_matches_hashlib():
    L = host_util.build()
    rng = random.Random(6)
    for n in [0, 1, 55, 56, 63, 64, 65, 111, 112, 119, 120, 127, 128, 129, 1000, 4099]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        for chunk in (1, 7, 64, 5000):
            o = (ctypes.c_uint8 * 32)()
            L.hc_sha256(host_util.raw(d), n, chunk, o)
            assert bytes(o) == hashlib.sha256(d).digest()
            o = (ctypes.c_uint8 * 64)()
            L.hc_sha512(host_util.raw(d), n, chunk, o)
            assert bytes(o) == hashlib.sha512(d).digest()

"""Device-side per-share Fiat-Shamir transcripts (csrc/sha2_dev.cuh, ec::proof_body) executed on the CPU through
tests/emu, against hashlib and Python integers: framed rows -> SHA-256 -> hash_to_scalar -> (c, r) / verdict
(dleq.rs:58-61, 87-99, 42-50, 119-126; modp.rs:142-148; secp256k1.rs:121-131; ristretto255.rs:196-205)."""
import ctypes
import hashlib
import random

import numpy as np
import pytest

import emu_util as eu

U8P = ctypes.POINTER(ctypes.c_uint8)


def u8(a):
    return a.ctypes.data_as(U8P)


@pytest.fixture(scope="module")
def L():
    return eu.build_ec()


def framed_rows(rng, n, eb, minimal):
    """n rows of four frames len_u64_be || bytes in slots of 8 + eb bytes; `minimal` draws ModpGroup-like lengths
    (1..eb bytes, including the block-boundary cases), otherwise every frame is eb bytes long."""
    slot = 8 + eb
    rows = np.zeros((n, 4 * slot), dtype=np.uint8)
    msgs = []
    for i in range(n):
        msg = b""
        for e in range(4):
            ln = eb
            if minimal:
                ln = rng.choice([1, 2, eb - 1, eb, eb, eb, rng.randrange(1, eb + 1), 55 % eb + 1, 47, 48, 56, 64 % eb + 1])
                ln = min(max(ln, 1), eb)
            body = bytes(rng.getrandbits(8) for _ in range(ln))
            fr = ln.to_bytes(8, "big") + body
            rows[i, e * slot:e * slot + len(fr)] = np.frombuffer(fr, dtype=np.uint8)
            msg += fr
        msgs.append(msg)
    return rows, msgs


@pytest.mark.parametrize("eb,minimal,wide", [(256, True, 0), (33, False, 0), (32, False, 1), (7, True, 0), (120, True, 1)])
def test_row_hash_equals_hashlib(L, eb, minimal, wide):
    rng = random.Random(eb * 3 + wide)
    n = 40
    rows, msgs = framed_rows(rng, n, eb, minimal)
    stride = 64 if eb == 256 else (16 if wide else 8)
    out = np.full(n * stride, 0xDEADBEEF, dtype=np.uint32)
    dig = np.zeros(n * 32, dtype=np.uint8)
    L.emu_row_hash(u8(rows), rows.shape[1], 8 + eb, eu.P(out), stride, u8(dig), n, wide)
    for i, m in enumerate(msgs):
        d = hashlib.sha256(m).digest()
        assert bytes(dig[32 * i:32 * i + 32]) == d
        want = int.from_bytes(hashlib.sha512(d).digest(), "little") if wide else int.from_bytes(hashlib.sha256(d).digest(), "big")
        assert eu.from_limbs(out[stride * i:stride * (i + 1)]) == want      # including the zero fill above the hash


def test_every_message_length_pads_correctly(L):
    """lengths 0..200 cover every padding case (55 / 56 / 63 / 64 bytes modulo the block)."""
    for ln in range(0, 201):
        rows = np.zeros((1, 4 * 8 + 256), dtype=np.uint8)
        # one frame of ln - 32 + ... : build a row whose four frames add up to exactly `total` bytes
        lens = [max(0, min(ln, 60)), max(0, min(ln - 60, 60)), max(0, min(ln - 120, 60)), max(0, ln - 180)]
        slot = 8 + 64
        rows = np.zeros((1, 4 * slot), dtype=np.uint8)
        msg = b""
        for e, l in enumerate(lens):
            fr = l.to_bytes(8, "big") + bytes((7 * e + k) & 0xFF for k in range(l))
            rows[0, e * slot:e * slot + len(fr)] = np.frombuffer(fr, dtype=np.uint8)
            msg += fr
        out = np.zeros(8, dtype=np.uint32)
        dig = np.zeros(32, dtype=np.uint8)
        L.emu_row_hash(u8(rows), rows.shape[1], slot, eu.P(out), 8, u8(dig), 1, 0)
        assert bytes(dig) == hashlib.sha256(msg).digest(), ln


def test_box_hash_is_the_running_hash_over_all_rows(L):
    rng = random.Random(5)
    rows, msgs = framed_rows(rng, 37, 256, True)
    dig = np.zeros(32, dtype=np.uint8)
    L.emu_box_hash(u8(rows), rows.shape[1], 8 + 256, 37, u8(dig))
    assert bytes(dig) == hashlib.sha256(b"".join(msgs)).digest()


@pytest.mark.parametrize("m,wide,be", [(eu.SECP_N, 0, 1), (eu.ED_L, 1, 0)])
def test_proof_body_challenge_response_and_verdict(L, m, wide, be):
    rng = random.Random(wide + 11)
    n = 50
    M = eu.modulus_words(m)
    words = 16 if wide else 8
    H = [rng.getrandbits(32 * words) for _ in range(n)]
    H[0], H[1], H[2] = 0, (1 << (32 * words)) - 1, m            # extremes of the reduction
    if not wide:
        H[3], H[4] = m - 1, m + 1
    sk = [rng.randrange(m) for _ in range(n)]
    w = [rng.randrange(m) for _ in range(n)]
    sk[5], w[5] = m - 1, 0
    h = np.concatenate([eu.to_limbs(x, words) for x in H])
    skl = np.concatenate([eu.to_limbs(x, 8) for x in sk])
    wl = np.concatenate([eu.to_limbs(x, 8) for x in w])
    c_out, r_out = np.zeros(32 * n, dtype=np.uint8), np.zeros(32 * n, dtype=np.uint8)
    L.emu_ec_proof(eu.P(M), eu.P(h), eu.P(skl), eu.P(wl), None, u8(c_out), u8(r_out), None, n, wide, be)
    order = "big" if be else "little"
    for i in range(n):
        c = H[i] % m
        assert int.from_bytes(bytes(c_out[32 * i:32 * i + 32]), order) == c
        assert int.from_bytes(bytes(r_out[32 * i:32 * i + 32]), order) == (w[i] - sk[i] * c) % m
    # verdicts: the right challenge, a wrong one, and a non-canonical encoding of the right one (c + m) never match
    c_in = c_out.copy()
    c_in[32 * 7 + (31 if be else 0)] ^= 1
    if (H[8] % m) + m < (1 << 256):
        c_in[32 * 8:32 * 9] = np.frombuffer(((H[8] % m) + m).to_bytes(32, order), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint32)
    L.emu_ec_proof(eu.P(M), eu.P(h), None, None, u8(c_in), None, None, eu.P(ok), n, wide, be)
    for i in range(n):
        assert ok[i] == (0 if i == 7 or (i == 8 and (H[8] % m) + m < (1 << 256)) else 1), i

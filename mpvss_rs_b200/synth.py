"""Deterministic synthetic inputs for the PVSS hot path.

Everything is expanded from one 64-bit seed with SHA-256 in counter mode, so the
GPU path, the CPU oracle and the CPU baseline see identical private keys,
polynomial coefficients and DLEQ witnesses.  Scalars only: public keys
y_i = G^{sk_i} are produced by whichever implementation consumes them.

The reference draws these values from ``thread_rng`` (polynomial.rs:34-47,
participant.rs:223/1170/1626, modp.rs:162-174, secp256k1.rs:158-166,
ristretto255.rs:227-236); this module is the injection seam.
"""
from __future__ import annotations

import hashlib

DEFAULT_SEED = 0x6D70767373  # "mpvss"


def _stream(seed: int, label: str, index: int, nbytes: int) -> bytes:
    out = b""
    ctr = 0
    while len(out) < nbytes:
        out += hashlib.sha256(
            seed.to_bytes(8, "big") + label.encode() + index.to_bytes(8, "big") + ctr.to_bytes(4, "big")
        ).digest()
        ctr += 1
    return out[:nbytes]


def uniform_below(seed: int, label: str, index: int, bound: int) -> int:
    """Value in [0, bound): 64 extra bits then reduce (bias < 2^-64)."""
    nbytes = (bound.bit_length() + 7) // 8 + 8
    return int.from_bytes(_stream(seed, label, index, nbytes), "big") % bound


def coefficients(seed: int, t: int, order: int):
    """t polynomial coefficients, uniform below ``order`` (polynomial.rs:34-47)."""
    return [uniform_below(seed, "coeff", j, order) for j in range(t)]


def witnesses(seed: int, n: int, bound: int, label: str = "witness"):
    """n DLEQ witnesses uniform below ``bound`` (MODP: q, modp.rs:165-168; EC: order)."""
    return [uniform_below(seed, label, i, bound) for i in range(n)]


def private_keys(seed: int, n: int, group_name: str, order: int, modulus: int | None = None):
    """n distinct private keys.

    MODP (modp.rs:162-174): uniform below q with gcd(k, q-1) == 1; q-1 = 2g with g
    prime, so "odd and not g".  EC: uniform non-zero below the group order.
    """
    keys, seen = [], set()
    i = 0
    while len(keys) < n:
        if group_name == "modp":
            k = uniform_below(seed, "sk", i, modulus) | 1
            ok = k < modulus and k != (modulus - 1) // 2
        else:
            k = uniform_below(seed, "sk", i, order)
            ok = k != 0
        i += 1
        if ok and k not in seen:
            seen.add(k)
            keys.append(k)
    return keys

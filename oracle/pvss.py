"""CPU oracle: the five PVSS entry points of the reference, restated in Python.

TEST INFRASTRUCTURE ONLY (see oracle/groups.py header; parity status there).

Follows /root/reference/src/participant.rs (the three per-group specialisations
have the same shape; line triples are MODP / secp256k1 / ristretto255),
src/dleq.rs:37-126 and src/polynomial.rs:50-58.  Randomness the reference draws
internally from ``thread_rng`` (polynomial coefficients polynomial.rs:34-47,
DLEQ witnesses participant.rs:223/1170/1626) is an explicit *input* here.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

from .groups import lagrange_coefficient, mod_inverse, elements_equal
from math import gcd


# --- dleq.rs -----------------------------------------------------------------

def framed(b: bytes) -> bytes:
    """dleq.rs:58-61 -- 8-byte big-endian length, then the bytes."""
    return len(b).to_bytes(8, "big") + b


def append_transcript(group, h1, h2, a1, a2, hasher) -> None:
    """dleq.rs:87-99 -- order (h1, h2, a1, a2), each framed."""
    for e in (h1, h2, a1, a2):
        hasher.update(framed(group.element_to_bytes(e)))


def challenge_from(group, hasher):
    """participant.rs:251-252 / dleq.rs:119-126 -- the digest is hashed *again*
    inside hash_to_scalar."""
    return group.hash_to_scalar(hasher.digest())


def verifier_commitments(group, g1, h1, g2, h2, r, c):
    """dleq.rs:66-84 -- a1 = g1^r * h1^c, a2 = g2^r * h2^c."""
    a1 = group.mul(group.exp(g1, r), group.exp(h1, c))
    a2 = group.mul(group.exp(g2, r), group.exp(h2, c))
    return a1, a2


def prover_response(group, w, alpha, c):
    """dleq.rs:42-50 -- r = w - alpha*c."""
    return group.scalar_sub(w, group.scalar_mul(alpha, c))


# --- polynomial.rs -------------------------------------------------------------

def poly_get_value(coeffs, x: int) -> int:
    """polynomial.rs:50-58 -- sum a_i * x^i over the integers (unreduced)."""
    result = coeffs[0]
    for i in range(1, len(coeffs)):
        result += coeffs[i] * (x ** i)
    return result


def poly_eval_mod(coeffs, x: int, m: int) -> int:
    """Horner mod m; equals ``poly_get_value(coeffs, x) % m`` (asserted in tests)."""
    acc = 0
    for a in reversed(coeffs):
        acc = (acc * x + a) % m
    return acc


# --- sharebox.rs ---------------------------------------------------------------

@dataclass
class DistributionSharesBox:
    """sharebox.rs:75-86 -- maps keyed by serialised public key."""
    commitments: list = field(default_factory=list)
    positions: dict = field(default_factory=dict)
    shares: dict = field(default_factory=dict)
    publickeys: list = field(default_factory=list)
    challenge: int = 0
    responses: dict = field(default_factory=dict)
    U: int = 0
    # not part of the reference struct: per-participant values kept for parity tests
    trace: dict = field(default_factory=dict)


@dataclass
class ShareBox:
    """sharebox.rs:22-27."""
    publickey: object = None
    share: object = None
    challenge: int = 0
    response: int = 0
    trace: dict = field(default_factory=dict)


# --- X_i ---------------------------------------------------------------------

def x_reference_schedule(group, commitments, position: int):
    """participant.rs:207-215 / 1174-1184 / 1630-1642 (dealer) and
    :423-434 / 1411-1421 / 1854-1864 (verifier): X = prod_j C_j^(i^j mod ord)."""
    x_val = group.identity()
    exponent = group.scalar_from_small(1)
    pos = group.scalar_from_small(position)
    for c_j in commitments:
        x_val = group.mul(x_val, group.exp(c_j, exponent))
        exponent = group.scalar_mul(exponent, pos)
    return x_val


def x_horner_schedule(group, commitments, position: int):
    """Same group element as x_reference_schedule (asserted in tests), t-1 steps of
    "raise to the small integer i, multiply by C_j" -- the schedule the GPU runs."""
    acc = commitments[-1]
    for c_j in reversed(commitments[:-1]):
        acc = group.mul(group.exp(acc, position), c_j)
    return acc


# --- the five entry points -----------------------------------------------------

def distribute_secret(group, secret: int, publickeys, threshold: int, coeffs, witnesses,
                      x_schedule=x_reference_schedule) -> DistributionSharesBox:
    """participant.rs:160-286 / 1094-1274 / 1573-1717."""
    assert threshold <= len(publickeys)            # :166
    assert len(coeffs) == threshold and len(witnesses) == len(publickeys)
    sub_gen = group.subgroup_generator()
    main_gen = group.generator()
    order = group.order()
    box = DistributionSharesBox()
    hasher = hashlib.sha256()
    coeff_scalars = [group.scalar_from_int(a) for a in coeffs]
    box.commitments = [group.exp(sub_gen, a) for a in coeff_scalars]          # :189-193
    tr = {"X": [], "Y": [], "a1": [], "a2": [], "p": []}
    alphas = []
    for idx, pk in enumerate(publickeys):                                     # :196-248
        position = idx + 1
        kb = group.element_to_bytes(pk)
        box.positions[kb] = position
        p_i = group.scalar_from_int(poly_get_value(coeffs, position) % order)  # :202
        alphas.append(p_i)
        x_val = x_schedule(group, box.commitments, position)                  # :207-215
        y_val = group.exp(pk, p_i)                                            # :219
        box.shares[kb] = y_val
        w = witnesses[idx]
        a1 = group.exp(sub_gen, w)                                            # :236 -> dleq.rs:207
        a2 = group.exp(pk, w)                                                 # :237 -> dleq.rs:214
        append_transcript(group, x_val, y_val, a1, a2, hasher)                # :238-245
        for k, v in zip(("X", "Y", "a1", "a2", "p"), (x_val, y_val, a1, a2, p_i)):
            tr[k].append(v)
    box.challenge = challenge_from(group, hasher)                             # :251-252
    for idx, pk in enumerate(publickeys):                                     # :255-264
        kb = group.element_to_bytes(pk)
        alpha_c = group.scalar_mul(alphas[idx], box.challenge) % order
        box.responses[kb] = group.scalar_sub(witnesses[idx], alpha_c) % order
    s = group.scalar_from_int(poly_get_value(coeffs, 0) % order)              # :267
    g_s = group.exp(main_gen, s)                                              # :268
    box.U = secret ^ group.mask_of(g_s)                                       # :269-272
    box.publickeys = list(publickeys)
    tr["G_s"] = g_s
    box.trace = tr
    return box


def verify_distribution_shares(group, box: DistributionSharesBox,
                               x_schedule=x_reference_schedule, trace=None) -> bool:
    """participant.rs:399-455 / 1384-1442 / 1827-1885 (and mpvss.rs:90-144)."""
    sub_gen = group.subgroup_generator()
    hasher = hashlib.sha256()
    for pk in box.publickeys:
        kb = group.element_to_bytes(pk)
        position = box.positions.get(kb)
        response = box.responses.get(kb)
        y_val = box.shares.get(kb)
        if position is None or response is None or y_val is None:             # :415-420
            return False
        x_val = x_schedule(group, box.commitments, position)                  # :423-434
        a1, a2 = verifier_commitments(group, sub_gen, x_val, pk, y_val,
                                      response, box.challenge)                # :438-447
        append_transcript(group, x_val, y_val, a1, a2, hasher)
        if trace is not None:
            trace.setdefault("X", []).append(x_val)
            trace.setdefault("a1", []).append(a1)
            trace.setdefault("a2", []).append(a2)
    return challenge_from(group, hasher) == box.challenge                     # :451-454


def extract_secret_share(group, box: DistributionSharesBox, private_key, w):
    """participant.rs:294-353 / 1282-1338 / 1725-1781."""
    main_gen = group.generator()
    public_key = group.generate_public_key(private_key)                       # :306
    kb = group.element_to_bytes(public_key)
    y_val = box.shares.get(kb)                                                # :310
    if y_val is None:
        return None
    inv = group.scalar_inverse(private_key)                                   # :314
    if inv is None:
        return None
    share = group.exp(y_val, inv)                                             # :316
    hasher = hashlib.sha256()
    a1 = group.exp(main_gen, w)                                               # :331
    a2 = group.exp(share, w)                                                  # :332
    append_transcript(group, public_key, y_val, a1, a2, hasher)               # :333-340
    c = challenge_from(group, hasher)                                         # :342-343
    r = prover_response(group, w, private_key, c)                             # :347 -> dleq.rs:221-228
    return ShareBox(public_key, share, c, r, {"a1": a1, "a2": a2})


def verify_share(group, sharebox: ShareBox, box: DistributionSharesBox, publickey) -> bool:
    """participant.rs:361-386 / 1346-1371 / 1789-1814 -> dleq.rs:275-302."""
    kb = group.element_to_bytes(publickey)
    y_val = box.shares.get(kb)
    if y_val is None:
        return False
    hasher = hashlib.sha256()
    a1, a2 = verifier_commitments(group, group.generator(), publickey, sharebox.share, y_val,
                                  sharebox.response, sharebox.challenge)
    append_transcript(group, publickey, y_val, a1, a2, hasher)
    return challenge_from(group, hasher) == sharebox.challenge


def lagrange_factor(group, position: int, share, values):
    """participant.rs:526-561 (MODP: integers, gcd-reduce, inverse mod g, element
    inverse when negative) / 1518-1557 / 1955-2002 (EC: scalar field + negation)."""
    if group.name == "modp":
        num, den = lagrange_coefficient(position, values)
        negative = num * den < 0
        num, den = abs(num), abs(den)
        g = gcd(num, den)
        num //= g
        den //= g
        den_inv = mod_inverse(den, group.subgroup_order())
        if den_inv is None:
            return None
        factor = group.exp(share, (num * den_inv) % group.subgroup_order())
        if negative:
            factor = group.element_inverse(factor)
        return factor
    lam_num, lam_den, sign = 1, 1, 1
    n = group.order()
    for j in values:
        if j == position:
            continue
        lam_num = lam_num * j % n
        diff = j - position
        if diff < 0:
            sign = -sign
            diff = -diff
        lam_den = lam_den * diff % n
    inv = group.scalar_inverse(lam_den)
    lam = 0 if inv is None else lam_num * inv % n           # ristretto :1983-1989 falls back to 0
    factor = group.exp(share, lam)
    if sign < 0:
        factor = group.element_inverse(factor)
    return factor


def reconstruct(group, share_boxes, box: DistributionSharesBox, trace=None):
    """participant.rs:462-519 / 1452-1513 / 1895-1950."""
    if len(share_boxes) < len(box.commitments):                               # :469
        return None
    shares = {}
    for sb in share_boxes:                                                    # :476-482
        kb = group.element_to_bytes(sb.publickey)
        position = box.positions.get(kb)
        if position is None:
            return None
        shares[position] = sb.share
    values = sorted(shares)            # BTreeMap order (MODP); EC order is irrelevant (commutative)
    acc = group.identity()
    for position in values:                                                   # :490-509
        factor = lagrange_factor(group, position, shares[position], values)
        if factor is None:
            return None
        acc = group.mul(acc, factor)
    if trace is not None:
        trace["G_s"] = acc
    return group.mask_of(acc) ^ box.U                                         # :512-518


def string_to_secret(message: str) -> int:
    """lib.rs:49-52."""
    return int.from_bytes(message.encode(), "big")


def string_from_secret(secret: int) -> str:
    """lib.rs:54-57."""
    return secret.to_bytes((secret.bit_length() + 7) // 8, "big").decode()


__all__ = [n for n in dir() if not n.startswith("_")]
_ = elements_equal

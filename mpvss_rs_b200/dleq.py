"""Mirror of the reference's public `DLEQ<G>` (src/dleq.rs:153-335) and `PVSS<G>` (src/mpvss.rs:42-145)
wrappers on top of the batch C ABI: same method names and meaning, every group operation on the GPU.
The challenge/response arithmetic and the framed SHA-256 transcript are the host-side parts, restated
from dleq.rs:42-61, 87-126 and the groups' hash_to_scalar."""
from __future__ import annotations

import hashlib

from .participant import Group, Participant


def _framed(b: bytes) -> bytes:                 # dleq.rs:58-61
    return len(b).to_bytes(8, "big") + b


def hash_to_scalar(group: Group, data: bytes) -> int:
    """Group::hash_to_scalar (modp.rs:142-148, secp256k1.rs:121-131, ristretto255.rs:196-205)."""
    c = group.codec
    if c.name == "modp":
        return int.from_bytes(hashlib.sha256(data).digest(), "big") % ((c.key_bound - 1) // 2)
    if c.name == "secp256k1":
        return int.from_bytes(hashlib.sha256(data).digest(), "big") % c.order
    return int.from_bytes(hashlib.sha512(data).digest(), "little") % c.order


class DLEQ:
    """Chaum-Pedersen proof that log_g1(h1) = log_g2(h2) (dleq.rs:153-163)."""

    def __init__(self, group: Group):
        self.group = group
        self.g1 = self.h1 = self.g2 = self.h2 = None
        self.w = self.alpha = 0
        self.c = None
        self.r = None

    def init(self, g1, h1, g2, h2, alpha, w):   # dleq.rs:187-203
        self.g1, self.h1, self.g2, self.h2, self.alpha, self.w = g1, h1, g2, h2, alpha, w

    def get_a1(self):                           # dleq.rs:207-209
        return self.group.batch_exp(self.g1, [self.w])[0]

    def get_a2(self):                           # dleq.rs:214-216
        return self.group.batch_exp(self.g2, [self.w])[0]

    def get_r(self):                            # dleq.rs:221-228 -> Prover::response :42-50
        if self.c is None:
            return None
        order = self.group.codec.order
        return (self.w - (self.alpha * self.c) % order) % order

    @staticmethod
    def verifier_commitments(group, g1, h1, g2, h2, response, c):    # dleq.rs:232-244
        a1, a2 = group.dleq_verify_commit(g1, [h1], [g2], [h2], [response], c)
        return a1[0], a2[0]

    @staticmethod
    def append_transcript_hash(group, h1, h2, a1, a2, hasher):       # dleq.rs:247-256
        for e in (h1, h2, a1, a2):
            hasher.update(_framed(group.codec.key(e)))

    def update_hash(self, hasher):              # dleq.rs:306-321
        self.append_transcript_hash(self.group, self.h1, self.h2, self.get_a1(), self.get_a2(), hasher)

    def check(self, hasher) -> bool:            # dleq.rs:326-334 -> Verifier::check :119-126
        if self.c is None:
            return False
        return hash_to_scalar(self.group, hasher.copy().digest()) == self.c

    def verify(self) -> bool:                   # dleq.rs:275-302
        if self.c is None or self.r is None:
            return False
        hasher = hashlib.sha256()
        a1, a2 = self.verifier_commitments(self.group, self.g1, self.h1, self.g2, self.h2, self.r, self.c)
        self.append_transcript_hash(self.group, self.h1, self.h2, a1, a2, hasher)
        return self.check(hasher)


class PVSS:
    """src/mpvss.rs:42-61 holder with `verify_distribution_shares` (mpvss.rs:90-144)."""

    def __init__(self, group: Group):
        self.group = group

    def verify_distribution_shares(self, box) -> bool:
        return Participant(self.group).verify_distribution_shares(box)

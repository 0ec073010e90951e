"""A wire format for `DistributionSharesBox` / `ShareBox` in the structure-of-arrays layout the C ABI
consumes (the reference has none: sharebox.rs:21,74 derive only Debug/Clone).

A received box can be verified straight from its bytes: the arrays are sliced out of the blob and
handed to `mpvss_verify_distribution` without rebuilding the reference's HashMaps
(participant.rs:409-413) or converting a single element on the host.

Layout (all integers little-endian):
    magic "MPVB" | version u16 | group u16 | n u32 | t u32 | eb u16 | sb u16
    commitments  t  * eb      boundary encoding of include/mpvss_b200.h
    positions    n  * i64
    publickeys   n  * eb
    shares       n  * eb
    responses    n  * sb
    challenge    sb
    U            eb           big-endian, left-padded
ShareBox: magic "MPVS" | version u16 | group u16 | eb u16 | sb u16 | publickey | share | challenge | response
"""
from __future__ import annotations

import ctypes
import struct

from . import lib as _lib
from .participant import DistributionSharesBox, ShareBox

VERSION = 1
_BOX = struct.Struct("<4sHHIIHH")
_SB = struct.Struct("<4sHHHH")


def box_to_bytes(group, box: DistributionSharesBox) -> bytes:
    c = group.codec
    n, t = len(box.publickeys), len(box.commitments)
    keys = [c.key(pk) for pk in box.publickeys]
    out = [_BOX.pack(b"MPVB", VERSION, _lib.GROUP_IDS[c.name], n, t, c.eb, c.sb), c.enc_elems(box.commitments),
           struct.pack(f"<{n}q", *[box.positions[k] for k in keys]), c.enc_elems(box.publickeys),
           c.enc_elems([box.shares[k] for k in keys]), c.enc_scalars([box.responses[k] for k in keys]),
           c.enc_scalar(box.challenge), box.U.to_bytes(c.eb, "big")]
    return b"".join(out)


def _parse(group, blob: bytes):
    c = group.codec
    magic, ver, gid, n, t, eb, sb = _BOX.unpack_from(blob, 0)
    if magic != b"MPVB" or ver != VERSION:
        raise ValueError("not a DistributionSharesBox blob of this version")
    if gid != _lib.GROUP_IDS[c.name] or eb != c.eb or sb != c.sb:
        raise ValueError("blob belongs to another group")
    off = _BOX.size
    sizes = [t * eb, n * 8, n * eb, n * eb, n * sb, sb, eb]
    if len(blob) != off + sum(sizes):
        raise ValueError("truncated or oversized blob")
    parts = []
    for s in sizes:
        parts.append((off, s))
        off += s
    return n, t, parts


def box_from_bytes(group, blob: bytes) -> DistributionSharesBox:
    c = group.codec
    n, t, parts = _parse(group, blob)
    cut = lambda i: blob[parts[i][0]:parts[i][0] + parts[i][1]]
    box = DistributionSharesBox()
    box.commitments = c.dec_elems(cut(0), t)
    positions = struct.unpack(f"<{n}q", cut(1))
    box.publickeys = c.dec_elems(cut(2), n)
    ys, rs = c.dec_elems(cut(3), n), c.dec_scalars(cut(4), n)
    for pk, p, y, r in zip(box.publickeys, positions, ys, rs):
        k = c.key(pk)
        box.positions[k], box.shares[k], box.responses[k] = p, y, r
    box.challenge = c.dec_scalar(cut(5))
    box.U = int.from_bytes(cut(6), "big")
    return box


def verify_distribution_bytes(group, blob: bytes) -> bool:
    """Participant::verify_distribution_shares on a serialised box, zero host-side conversion."""
    n, t, parts = _parse(group, blob)
    base = (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)
    addr = ctypes.addressof(base)
    u8 = lambda i: ctypes.cast(addr + parts[i][0], ctypes.POINTER(ctypes.c_uint8))
    pos = ctypes.cast(addr + parts[1][0], ctypes.POINTER(ctypes.c_int64))
    if parts[1][0] % 8:                       # keep the i64 array aligned
        pos = (ctypes.c_int64 * n).from_buffer_copy(blob[parts[1][0]:parts[1][0] + 8 * n])
    ok = ctypes.c_int(0)
    ctx = group.ctx
    ctx.check(ctx.lib.mpvss_verify_distribution(ctx.h, n, t, u8(0), pos, u8(2), u8(3), u8(4), u8(5), ctypes.byref(ok),
                                                None, None, None, None))
    return bool(ok.value)


def sharebox_to_bytes(group, sb: ShareBox) -> bytes:
    c = group.codec
    return _SB.pack(b"MPVS", VERSION, _lib.GROUP_IDS[c.name], c.eb, c.sb) + c.enc_elem(sb.publickey) + \
        c.enc_elem(sb.share) + c.enc_scalar(sb.challenge) + c.enc_scalar(sb.response)


def sharebox_from_bytes(group, blob: bytes) -> ShareBox:
    c = group.codec
    magic, ver, gid, eb, sb = _SB.unpack_from(blob, 0)
    if magic != b"MPVS" or ver != VERSION or gid != _lib.GROUP_IDS[c.name] or eb != c.eb or sb != c.sb:
        raise ValueError("not a ShareBox blob for this group")
    if len(blob) != _SB.size + 2 * eb + 2 * sb:
        raise ValueError("truncated or oversized blob")
    o = _SB.size
    return ShareBox(c.dec_elem(blob[o:o + eb]), c.dec_elem(blob[o + eb:o + 2 * eb]),
                    c.dec_scalar(blob[o + 2 * eb:o + 2 * eb + sb]), c.dec_scalar(blob[o + 2 * eb + sb:]))

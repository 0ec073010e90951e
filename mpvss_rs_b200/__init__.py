"""mpvss_rs_b200 -- B200-native batched group-exponentiation hot path of mpvss-rs.

`csrc/` holds the CUDA kernels and the C ABI (libmpvss_b200.so, include/mpvss_b200.h);
`participant.py` mirrors the reference's Participant / DistributionSharesBox / ShareBox
API on top of it.  Nothing here computes on the CPU: without the built library and a
CUDA device every entry point raises.
"""
from .lib import Context, MpvssError, LIB_PATH, load  # noqa: F401
from .participant import (DistributionSharesBox, Group, Participant, ShareBox,  # noqa: F401
                          string_from_secret, string_to_secret)
from . import wire  # noqa: F401,E402
from .dleq import DLEQ, PVSS, hash_to_scalar  # noqa: F401,E402

"""ctypes access to the host-side helpers of the product (bigint.h, sha2.h) -- tests only."""
import ctypes, os, subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhost_check.so")
SRC = os.path.join(HERE, "host_check.cpp")
CSRC = os.path.join(HERE, "..", "..", "mpvss_rs_b200", "csrc")
_p, _s = ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t


def build():
    ni = os.path.join(CSRC, "sha256_ni.cpp")
    deps = [SRC, ni, os.path.join(CSRC, "bigint.h"), os.path.join(CSRC, "sha2.h"), os.path.join(CSRC, "modp_chain.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        obj = os.path.join(HERE, "sha256_ni.o")
        subprocess.check_call(["g++", "-O2", "-fPIC", "-msha", "-msse4.1", "-mssse3", "-c", ni, "-o", obj])
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", SO, SRC, obj])
    L = ctypes.CDLL(SO)
    L.hc_divmod.argtypes = [_p, _s, _p, _s, _p, _p, _s]
    L.hc_mulmod.argtypes = [_p, _s, _p, _s, _p, _s, _p, _s]
    L.hc_modinv.argtypes = [_p, _s, _p, _s, _p, _s]
    L.hc_modinv_odd.argtypes = [_p, _s, _p, _s, _p, _s]
    L.hc_sha256.argtypes = [_p, _s, _s, _p]
    L.hc_sha512.argtypes = [_p, _s, _s, _p]
    L.hc_chain_ops.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint16),
                               ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    return L


def le(x, n):
    return (ctypes.c_uint8 * n).from_buffer_copy(int(x).to_bytes(n, "little"))


def raw(d):
    return (ctypes.c_uint8 * max(len(d), 1)).from_buffer_copy(d or b"\0")

"""ctypes access to the lane-per-thread emulator build of the MODP kernel bodies."""
import ctypes, os, subprocess, numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libemu_modp.so")
SRC = os.path.join(HERE, "emu_modp.cpp")
CSRC = os.path.join(HERE, "..", "..", "mpvss_rs_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("simt.h", "modp_arith.cuh", "modp_kernels.cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-DMPVSS_SIMT_EMU", "-shared", "-fPIC", "-pthread",
                               "-o", SO, SRC])
    return ctypes.CDLL(SO)


def to_limbs(x, n=64):
    return np.frombuffer(int(x).to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def from_limbs(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def consts_block(q):
    R = 1 << 2048
    blk = np.zeros(260, dtype=np.uint32)
    blk[0:64] = to_limbs(q)
    blk[64:128] = to_limbs(R - q)
    blk[128:192] = to_limbs(R % q)
    blk[192:256] = to_limbs(R * R % q)
    blk[256] = (-pow(q, -1, 1 << 32)) % (1 << 32)
    return blk


def P(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))

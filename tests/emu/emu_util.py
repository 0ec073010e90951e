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
    blk = np.zeros(324, dtype=np.uint32)
    blk[0:64] = to_limbs(q)
    blk[64:128] = to_limbs(R - q)
    blk[128:192] = to_limbs(R % q)
    blk[192:256] = to_limbs(R * R % q)
    blk[256] = (-pow(q, -1, 1 << 32)) % (1 << 32)
    blk[260:324] = to_limbs((q + 1) // 2)
    return blk


def P(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


# ---- secp256k1 / 256-bit fields ---------------------------------------------------------
SECP_P = 2**256 - 2**32 - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
SECP_GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
SECP_GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
EC_SRC = os.path.join(HERE, "emu_ec.cpp")
EC_SO = os.path.join(HERE, "libemu_ec.so")
ED_P = 2**255 - 19
ED_L = 2**252 + 27742317777372353535851937790883648493


def build_ec(force=False):
    deps = [EC_SRC] + [os.path.join(CSRC, f) for f in ("simt.h", "fp256.cuh", "fpspecial.cuh", "secp.cuh", "rist.cuh", "ec_kernels.cuh", "sha2_dev.cuh")]
    if force or not os.path.exists(EC_SO) or any(os.path.getmtime(d) > os.path.getmtime(EC_SO) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-DMPVSS_SIMT_EMU", "-shared", "-fPIC", "-pthread",
                               "-o", EC_SO, EC_SRC])
    return ctypes.CDLL(EC_SO)


def modulus_words(m):
    """fp256::Modulus {m[8], r2[8], one[8], np}"""
    R = 1 << 256
    w = np.zeros(25, dtype=np.uint32)
    w[0:8] = to_limbs(m, 8)
    w[8:16] = to_limbs(R * R % m, 8)
    w[16:24] = to_limbs(R % m, 8)
    w[24] = (-pow(m, -1, 1 << 32)) % (1 << 32)
    return w


def secp_consts():
    R = 1 << 256
    p = SECP_P
    # base-field constants are in plain representation (fpspecial.cuh); scalar field stays Montgomery
    return np.concatenate([modulus_words(SECP_P), modulus_words(SECP_N), to_limbs(7, 8),
                           to_limbs(SECP_GX, 8), to_limbs(SECP_GY, 8), to_limbs((p + 1) // 4, 8)])


def rist_consts():
    from oracle import groups as og
    R, p = 1 << 256, ED_P
    mont = lambda v: to_limbs(v % p, 8)   # plain representation
    return np.concatenate([modulus_words(ED_P), modulus_words(ED_L), mont(og.ED_D), mont(2 * og.ED_D),
                           mont(og.ED_SQRT_M1), mont(og.ED_INVSQRT_A_MINUS_D), mont(og.ED_BX), mont(og.ED_BY),
                           to_limbs((p - 5) // 8, 8)])

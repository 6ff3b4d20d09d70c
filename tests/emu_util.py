"""Helpers shared by the emulator tests and the GPU parity tests: limb-entry packing and Montgomery constants
computed with Python ints (independent of the product's host bignum)."""
import ctypes
import os
import subprocess

import numpy as np

LW = 52
LMASK = (1 << LW) - 1
HERE = os.path.dirname(os.path.abspath(__file__))


def shape_id(L, TPI):
    return L * 10 + TPI


def lp(L):
    v = (L + 1) & ~1
    return v + 2 if v % 16 == 0 else v   # csrc/mont52.cuh: Pad<L> (no lane stride that is a multiple of 16 doubles)


def to_entry(val, L, TPI):
    """Python int -> padded limb entry [TPI][LP]: float64 holding the exact integer value of each 52-bit limb."""
    out = np.zeros((TPI, lp(L)), dtype=np.float64)
    for g in range(L * TPI):
        out[g // L, g % L] = float((val >> (LW * g)) & LMASK)
    assert val >> (LW * L * TPI) == 0
    return out.reshape(-1)


def from_entry(e, L, TPI):
    e = np.asarray(e, dtype=np.float64).reshape(TPI, lp(L))
    v = 0
    for g in range(L * TPI):
        v += int(e[g // L, g % L]) << (LW * g)
    return v


def mont_consts(N, L, TPI):
    R = 1 << (LW * L * TPI)
    assert N % 2 == 1 and N * (1 << 8) <= R
    return {
        "R": R,
        "n": to_entry(N, L, TPI),
        "n0inv": (-pow(N, -1, 1 << LW)) % (1 << LW),
        "r2": to_entry(R * R % N, L, TPI),
        "oneM": to_entry(R % N, L, TPI),
        "one": to_entry(1, L, TPI),
    }


def to_words(vals, nwords):
    buf = b"".join(int(v).to_bytes(4 * nwords, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), nwords).copy()


def from_words(arr):
    arr = np.ascontiguousarray(arr, dtype="<u4")
    raw = arr.tobytes()
    n = arr.shape[1] * 4
    return [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(arr.shape[0])]


_emu = None


def load_emu():
    global _emu
    if _emu is not None:
        return _emu
    src = os.path.join(HERE, "emu", "emu_driver.cpp")
    so = os.path.join(HERE, "emu", "libphe_emu.so")
    deps = [src] + [os.path.join(HERE, "..", "pailliercryptolib_python_b200", "csrc", f) for f in ("mont52.cuh", "paillier_items.cuh", "npair_items.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        # -frounding-math: the emulator runs under FE_TOWARDZERO (fma.rz.f64); no constant folding across that
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-frounding-math", "-ffp-contract=off", "-pthread", "-shared",
                               "-fPIC", "-o", so, src])
    _emu = ctypes.CDLL(so)
    return _emu


def P(a):
    """numpy uint32 array -> ctypes pointer (None passes NULL)."""
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def PD(a):
    """numpy float64 array (limb entries) -> ctypes pointer (None passes NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def P64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


NE_NAMES = ["N", "ONE", "D", "W00", "W01", "W10", "W11", "OM0", "OM1"]


def npair_consts(n, L, TPI, n_words):
    """Constant block of the n-adic pair engine (csrc/npair_items.cuh: NPairEntry) from Python ints."""
    K = L * TPI
    R = 1 << (LW * K)
    assert n % 2 == 1 and n << 8 <= R and 32 * n_words + 8 <= LW * K
    n2 = n * n
    D = -(-R // n) * n
    vals = {"N": n, "ONE": 1, "D": D % R}
    for c in range(2):
        w = (1 << (32 * n_words * c)) * R * R % n2
        vals["W%d0" % c], vals["W%d1" % c] = w % n, w // n
    om = R % n2
    vals["OM0"], vals["OM1"] = om % n, om // n
    block = np.concatenate([to_entry(vals[k], L, TPI) for k in NE_NAMES])
    return {"cst": block, "n0inv": (-pow(n, -1, 1 << LW)) % (1 << LW), "d_top": D >> (LW * K), "R": R}

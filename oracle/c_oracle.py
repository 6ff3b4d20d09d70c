"""ctypes loader for oracle/paillier_oracle.c (OpenSSL BIGNUM restatement of the reference's CPU path).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: imported only by tests/, __graft_entry__ and bench.py's
cpu_baseline / --impl reference legs.  Packed little-endian uint32 limb arrays in, packed arrays out.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "paillier_oracle.c")
LIB = os.path.join(HERE, "libpaillier_oracle.so")
_u32p = ctypes.POINTER(ctypes.c_uint32)
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-pthread", SRC, "-lcrypto", "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def _words(v, k):
    return np.frombuffer(int(v).to_bytes(4 * k, "little"), dtype="<u4").copy()


def encrypt(n, n_words, hs, m, r, threads=1):
    """m [N, n_words], r [N, r_words] or None, hs int or None (classic)."""
    m = np.ascontiguousarray(m, dtype=np.uint32)
    out = np.empty((m.shape[0], 2 * n_words), dtype=np.uint32)
    hs_w = _words(hs, 2 * n_words) if hs is not None else None
    rw = 0
    if r is not None:
        r = np.ascontiguousarray(r, dtype=np.uint32)
        rw = r.shape[1]
    rc = lib().oracle_encrypt(_p(_words(n, n_words)), n_words, _p(hs_w), _p(m), ctypes.c_size_t(m.shape[0]), _p(r), rw,
                              _p(out), threads)
    assert rc == 0
    return out


def decrypt(n, n_words, p, q, ct, threads=1):
    ct = np.ascontiguousarray(ct, dtype=np.uint32)
    out = np.empty((ct.shape[0], n_words), dtype=np.uint32)
    rc = lib().oracle_decrypt(_p(_words(n, n_words)), n_words, _p(_words(p, n_words // 2)), _p(_words(q, n_words // 2)),
                              _p(ct), ctypes.c_size_t(ct.shape[0]), _p(out), threads)
    assert rc == 0
    return out


def add(n, n_words, a, b, threads=1):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty_like(a)
    rc = lib().oracle_add(_p(_words(n, n_words)), n_words, _p(a), ctypes.c_size_t(a.shape[0]), _p(b),
                          ctypes.c_size_t(b.shape[0]), _p(out), threads)
    if rc == 2:
        raise ValueError("size mismatch")
    assert rc == 0
    return out


def mul(n, n_words, ct, e, threads=1):
    ct = np.ascontiguousarray(ct, dtype=np.uint32)
    e = np.ascontiguousarray(e, dtype=np.uint32)
    out = np.empty_like(ct)
    rc = lib().oracle_mul(_p(_words(n, n_words)), n_words, _p(ct), ctypes.c_size_t(ct.shape[0]), _p(e), e.shape[1],
                          ctypes.c_size_t(e.shape[0]), _p(out), threads)
    if rc == 2:
        raise ValueError("size mismatch")
    assert rc == 0
    return out


def modexp(base, exp, modulus, words, threads=1):
    base = np.ascontiguousarray(base, dtype=np.uint32)
    exp = np.ascontiguousarray(exp, dtype=np.uint32)
    out = np.empty_like(base)
    rc = lib().oracle_modexp(_p(base), _p(exp), _p(_words(modulus, words)), words, ctypes.c_size_t(base.shape[0]),
                             _p(out), threads)
    assert rc == 0
    return out

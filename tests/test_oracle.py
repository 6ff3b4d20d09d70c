"""CPU tests of the oracle (test infrastructure): the Python-int oracle and the C restatement (oracle/paillier_oracle.c)
against the committed golden fixtures (tests/golden, generated from the reference's own fixedpoint.py and from exact
integer arithmetic) and against two independent bignum libraries (libgmp, OpenSSL)."""
import ctypes
import ctypes.util
import json
import os
import random

import numpy as np
import pytest

import c_oracle as C
import paillier_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "paillier_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def fpgold():
    with open(os.path.join(GOLD, "fixedpoint.json")) as f:
        return json.load(f)


def _ints(xs):
    return [int(x, 16) for x in xs]


def test_fixedpoint_matches_reference_golden(fpgold):
    n, max_int = int(fpgold["n"], 16), int(fpgold["max_int"], 16)
    assert max_int == n // 3 - 1
    for c in fpgold["cases"]:
        if c["kind"] == "float":
            v = float.fromhex(c["value"])
        elif c["kind"] == "int":
            v = int(c["value"])
        else:
            v = getattr(np, c["kind"])(int(c["value"]))
        enc, expo = O.fp_encode(v, n, max_int)
        assert (enc, expo) == (int(c["encoding"], 16), c["exponent"]), c
        dec = O.fp_decode(enc, expo, n, max_int)
        if c["kind"] == "float" and c["decoded_is_float"]:
            assert float(dec).hex() == c["decoded"], c
        else:
            assert int(dec) == int(c["decoded"]), c
    for e in fpgold["errors"]:
        v = float.fromhex(e["value"]) if "0x" in e["value"] else int(e["value"])
        if e["raises"]:
            with pytest.raises(Exception) as ei:
                O.fp_encode(v, n, max_int)
            assert type(ei.value).__name__ == e["raises"]


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_python_oracle_reproduces_kat(kat, idx):
    k = kat["keys"][idx]
    n, p, q = int(k["n"], 16), int(k["p"], 16), int(k["q"], 16)
    pk = O.PubKey(n, k["bits"], k["djn"], int(k["hs"], 16), k["randbits"])
    sk = O.PrivKey(pk, p, q)
    assert O.encrypt_batch(pk, _ints(k["m"]), _ints(k["r"])) == _ints(k["ct"])
    assert O.encrypt_batch(pk, _ints(k["m"]), None) == _ints(k["ct_raw"])
    assert O.decrypt_batch(sk, _ints(k["dec_in"])) == _ints(k["dec_out"])
    assert O.decrypt_batch(sk, _ints(k["ct"])) == _ints(k["m"])
    assert O.add_batch(pk, _ints(k["ct"]), _ints(k["add_b"])) == _ints(k["add_out"])
    assert O.mul_batch(pk, _ints(k["ct"]), _ints(k["mul_e"])) == _ints(k["mul_out"])


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_c_oracle_reproduces_kat(kat, idx):
    k = kat["keys"][idx]
    n, p, q = int(k["n"], 16), int(k["p"], 16), int(k["q"], 16)
    nw = k["bits"] // 32
    hs = int(k["hs"], 16) if k["djn"] else None
    m = O.to_limbs(_ints(k["m"]), nw)
    rw = (k["randbits"] + 31) // 32 if k["djn"] else nw
    r = O.to_limbs(_ints(k["r"]), rw)
    ct = C.encrypt(n, nw, hs, m, r, threads=3)
    assert O.from_limbs(ct) == _ints(k["ct"])
    assert O.from_limbs(C.encrypt(n, nw, hs, m, None)) == _ints(k["ct_raw"])
    assert O.from_limbs(C.decrypt(n, nw, q, p, O.to_limbs(_ints(k["dec_in"]), 2 * nw), threads=2)) == _ints(k["dec_out"])
    b = O.to_limbs(_ints(k["add_b"]), 2 * nw)
    assert O.from_limbs(C.add(n, nw, ct, b)) == _ints(k["add_out"])
    assert O.from_limbs(C.add(n, nw, ct, b[:1])) == _ints(k["add_bcast_out"])
    e = O.to_limbs(_ints(k["mul_e"]), nw)
    assert O.from_limbs(C.mul(n, nw, ct, e, threads=4)) == _ints(k["mul_out"])
    assert O.from_limbs(C.mul(n, nw, ct, e[3:4])) == _ints(k["mul_bcast_out"])
    with pytest.raises(ValueError):
        C.add(n, nw, ct, b[:3])


def test_c_oracle_empty_and_single():
    pk, sk = O.seeded_keypair(1024, 77)
    assert C.encrypt(pk.n, 32, pk.hs, np.zeros((0, 32), np.uint32), np.zeros((0, 16), np.uint32)).shape == (0, 64)
    assert C.decrypt(pk.n, 32, sk.p, sk.q, np.zeros((0, 64), np.uint32)).shape == (0, 32)
    ct = C.encrypt(pk.n, 32, pk.hs, O.to_limbs([42], 32), O.to_limbs([99], 16), threads=8)
    assert O.from_limbs(C.decrypt(pk.n, 32, sk.p, sk.q, ct, threads=8)) == [42]


def _gmp():
    name = ctypes.util.find_library("gmp") or "libgmp.so.10"
    try:
        return ctypes.CDLL(name)
    except OSError:
        return None


def test_modexp_cross_check_gmp_and_openssl():
    """pow() vs OpenSSL (through the C oracle) vs libgmp mpz_powm on the hot-path shapes."""
    rng = random.Random(99)
    g = _gmp()

    class Mpz(ctypes.Structure):
        _fields_ = [("alloc", ctypes.c_int), ("size", ctypes.c_int), ("d", ctypes.c_void_p)]

    def gmp_powm(b, e, m):
        xs = [Mpz() for _ in range(4)]
        for x, v in zip(xs[1:], (b, e, m)):
            g.__gmpz_init_set_str(ctypes.byref(x), hex(v)[2:].encode(), 16)
        g.__gmpz_init(ctypes.byref(xs[0]))
        g.__gmpz_powm(*[ctypes.byref(x) for x in xs])
        buf = ctypes.create_string_buffer(4 * len(hex(m)))
        g.__gmpz_get_str(buf, 16, ctypes.byref(xs[0]))
        for x in xs:
            g.__gmpz_clear(ctypes.byref(x))
        return int(buf.value, 16)

    for mod_bits, exp_bits in ((4096, 1024), (2048, 1024), (4096, 53), (6144, 1536)):
        mod = rng.getrandbits(mod_bits) | 1 | (1 << (mod_bits - 1))
        words = mod_bits // 32
        bases = [rng.randrange(mod) for _ in range(3)]
        exps = [rng.getrandbits(exp_bits) for _ in range(3)]
        want = [pow(b, e, mod) for b, e in zip(bases, exps)]
        got = O.from_limbs(C.modexp(O.to_limbs(bases, words), O.to_limbs(exps, words), mod, words))
        assert got == want
        if g is not None:
            assert [gmp_powm(b, e, mod) for b, e in zip(bases, exps)] == want


def test_packing_rules():
    """BN2bytes / pyByte2BN / BNUtils.int2Bytes (ipcl_bindings.cpp:100-138, ipcl_python.py:936-964)."""
    assert O.bn_to_bytes(0) == b"\x00\x00\x00\x00"
    assert O.bn_to_bytes(1) == b"\x01\x00\x00\x00"
    assert O.bn_to_bytes(1 << 32) == b"\x00\x00\x00\x00\x01\x00\x00\x00"
    assert O.int_to_le_bytes(0x0102) == b"\x02\x01"
    for v in (0, 1, 255, 256, 2 ** 64 - 1, 2 ** 2048 - 1, 3 ** 500):
        assert O.bytes_to_int(O.bn_to_bytes(v)) == v
        assert O.bytes_to_int(O.int_to_le_bytes(v)) == v
        assert O.from_limbs(O.to_limbs([v], 70)) == [v]

"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/phe_b200.h declares, its
host-side helpers agree with Python integers, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from pailliercryptolib_python_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "phe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "libphe_b200.so does not export %s" % name
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in lib.phe_version()


def test_no_cpu_fallback():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    # key objects are host-side state (constructible, picklable); every compute entry point refuses to run
    n, p, q = capi.keygen(256)
    pk = capi.PubKey(n, 256, djn=True)
    sk = capi.PrivKey(pk, p, q)
    m = np.ones((2, 8), dtype=np.uint32)
    ct = np.ones((2, 16), dtype=np.uint32)
    for call in (lambda: pk.encrypt(m), lambda: pk.encrypt(m, None, make_secure=False), lambda: pk.add(ct, ct),
                 lambda: pk.mul(ct, m[:, :2]), lambda: pk.obfuscate(ct), lambda: sk.decrypt(ct),
                 lambda: capi.modexp([3], [5], 1000003, 1)):
        with pytest.raises(RuntimeError, match="no CUDA device"):
            call()


def test_host_modexp_and_mont_block():
    import random
    rng = random.Random(4)
    for bits in (1024, 2048, 4096):
        mod = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        b, e = rng.randrange(mod), rng.getrandbits(200)
        assert capi.host_modexp(b, e, mod, bits // 32) == pow(b, e, mod)


def test_keygen_host():
    import math
    n, p, q = capi.keygen(512)
    assert p * q == n and n.bit_length() == 512 and p % 4 == 3 and q % 4 == 3
    assert math.gcd(p - 1, q - 1) == 2
    with pytest.raises(RuntimeError):
        capi.keygen(100)


def test_packing_helpers_roundtrip():
    vals = [0, 1, 2 ** 2048 - 1, 12345678901234567890]
    arr = capi.ints_to_array(vals, 64)
    assert arr.dtype == np.uint32 and arr.shape == (4, 64)
    assert capi.array_to_ints(arr) == vals

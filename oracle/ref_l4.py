"""Harness that runs the UNMODIFIED reference Python layer (src/ipcl_python/ipcl_python.py + bindings/fixedpoint.py)
over a binding module of our choice -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two uses:
  * tests/golden/make_l4_flows.py imports the reference L4 (from /root/reference) over `MockBindings` below -- the
    pybind11 surface of ipcl_bindings restated on exact Python integers -- and records every ciphertext integer and
    exponent list the reference's own test flows produce -> tests/golden/l4_flows.json (reference-generated golden
    data: the L4 semantics -- alignment order, negative-plaintext rule, matmul index maps, padding of the add tree --
    are the reference's code, only the arithmetic underneath is the oracle's);
  * tests/test_l4_flows.py replays the same flows through this repo's ipcl_python + CUDA (bit for bit), and, when a copy
    of the reference L4 is available, through the reference L4 over the real pybind11 shim.

The mock restates what the reference's glue does with ipcl:: objects; each class cites the binding lines it follows
(/root/reference/src/ipcl_python/bindings/ipcl_bindings_classes.cpp unless noted).  Arithmetic = paillier_oracle.
"""
from __future__ import annotations

import importlib.util
import os
import random
import sys
import types

import paillier_oracle as O

REF_ROOT_CANDIDATES = (
    "/root/reference/src/ipcl_python",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "ipcl_python_ref"),
)


def reference_l4_dir():
    """Directory holding the reference's ipcl_python.py and bindings/fixedpoint.py, or None."""
    for d in REF_ROOT_CANDIDATES:
        if os.path.exists(os.path.join(d, "ipcl_python.py")) and os.path.exists(os.path.join(d, "bindings", "fixedpoint.py")):
            return d
    return None


# ------------------------------------------------------------------------------------------------ the mock bindings
class _BN:
    """ipclBigNumber (:380-491): an unsigned integer with little-endian byte / word views."""

    def __init__(self, v=0):
        if isinstance(v, _BN):
            self.v = v.v
        elif isinstance(v, (bytes, bytearray)):
            self.v = int.from_bytes(v, "little")            # pyByte2BN (ipcl_bindings.cpp:100-118)
        elif isinstance(v, int):
            self.v = v
        else:                                                 # list / array of little-endian u32 words (:389-400)
            self.v = sum(int(w) << (32 * i) for i, w in enumerate(v))
        if self.v < 0:
            raise ValueError("mock BigNumber is unsigned")

    def to_bytes(self):                                       # BN2bytes (ipcl_bindings.cpp:120-138): whole words
        words = max(1, (self.v.bit_length() + 31) // 32)
        return self.v.to_bytes(4 * words, "little")

    def __eq__(self, o): return isinstance(o, _BN) and self.v == o.v
    def __ne__(self, o): return not self == o
    def __lt__(self, o): return self.v < o.v
    def __le__(self, o): return self.v <= o.v
    def __gt__(self, o): return self.v > o.v
    def __ge__(self, o): return self.v >= o.v
    def __hash__(self): return hash(self.v)
    def __add__(self, o): return _BN(self.v + o.v)
    def __mul__(self, o): return _BN(self.v * (o.v if isinstance(o, _BN) else int(o)))
    def __str__(self): return str(self.v)
    def __repr__(self): return "<BigNumber val: %d>" % self.v
    def BitSize(self): return max(1, self.v.bit_length())
    def DwordSize(self): return max(1, (self.v.bit_length() + 31) // 32)


_BN.Zero, _BN.One, _BN.Two = _BN(0), _BN(1), _BN(2)


def _texts(data):
    if isinstance(data, _BN):
        return [data]
    if isinstance(data, int):
        return [_BN(data)]
    return [x if isinstance(x, _BN) else _BN(x) for x in data]


class _PlainText:
    """ipclPlainText (:165-268)."""

    def __init__(self, data):
        self.t = list(data.t) if isinstance(data, _PlainText) else _texts(data)

    def __len__(self): return len(self.t)
    def getSize(self): return len(self.t)
    def getTexts(self): return list(self.t)
    def __getitem__(self, k): return self.t[k]
    def rotate(self, shift): return _PlainText(_rotated(self.t, shift))


def _rotated(t, shift):
    """BaseText::rotate [UPSTREAM-RECALLED]: cyclic; only ever used on power-of-two lengths where element 0 of the
    rotate-and-add tree is the total whichever way it turns (ipcl_python.py:821-827)."""
    if not t:
        return []
    s = shift % len(t)
    return t[s:] + t[:s]


class _PublicKey:
    """ipclPublicKey (:14-91).  hs and the obfuscator exponents come from a seeded stream (mock only)."""

    def __init__(self, n, bits=1024, enable_DJN=False, _hs=None):
        self._n = n.v if isinstance(n, _BN) else int(n)
        self.length = bits
        rng = random.Random(self._n & 0xFFFFFFFF)
        if enable_DJN:
            if _hs is None:
                import math
                x = rng.getrandbits(bits + 128)
                while math.gcd(x, self._n) != 1:
                    x = rng.getrandbits(bits + 128)
                _hs = O.djn_hs(self._n, x)
            self.o = O.PubKey(self._n, bits, True, _hs, bits // 2)
        else:
            self.o = O.PubKey(self._n, bits, False)
        self._rng = rng

    n = property(lambda self: _BN(self._n))
    nsquare = property(lambda self: _BN(self._n * self._n))

    def __eq__(self, o): return isinstance(o, _PublicKey) and self._n == o._n      # :36-39
    def __ne__(self, o): return not self == o
    def __hash__(self): return hash(self._n)                                        # :40-46

    def _r(self):
        if self.o.djn:
            return self._rng.getrandbits(self.o.randbits)
        return self._rng.randrange(1, self._n)

    def encrypt(self, pt, make_secure):                                             # :53-60
        return _CipherText(self, [_BN(O.encrypt(self.o, m.v, self._r() if make_secure else None)) for m in pt.t])

    def encrypt_tolist(self, pt, make_secure):                                      # :61-70
        return self.encrypt(pt, make_secure).getTexts()

    def apply_obfuscator(self, x):                                                  # :71-83
        if isinstance(x, _BN):
            return _BN(x.v * O.obfuscator(self.o, self._r()) % (self._n * self._n))
        return [self.apply_obfuscator(c) for c in x.t]


class _PrivateKey:
    """ipclPrivateKey (:93-163)."""

    def __init__(self, pk, p, q):
        self.pk = pk
        self.o = O.PrivKey(pk.o, p.v, q.v)

    n = property(lambda self: _BN(self.pk._n))
    p = property(lambda self: _BN(self.o.p))
    q = property(lambda self: _BN(self.o.q))

    def __eq__(self, o): return self.o.q == o.o.q                                   # :110-113

    def decrypt(self, ct):                                                          # :127-133
        if ct.pk != self.pk:
            raise RuntimeError("decrypt: public key mismatch")
        return _PlainText([_BN(O.decrypt_crt(self.o, c.v)) for c in ct.t])

    def decrypt_tolist(self, ct):
        return self.decrypt(ct).getTexts()


class _CipherText:
    """ipclCipherText (:270-378): + is the product mod n^2 with size-1 broadcast (:318-323), * PlainText the
    element-wise power (:324-325)."""

    def __init__(self, pk, data):
        self.pk = pk
        self.t = list(data.t) if isinstance(data, _CipherText) else _texts(data)

    public_key = property(lambda self: self.pk)

    def __len__(self): return len(self.t)
    def getSize(self): return len(self.t)
    def getTexts(self): return list(self.t)
    def getCipherText(self, i): return _CipherText(self.pk, [self.t[i]])
    def rotate(self, shift): return _CipherText(self.pk, _rotated(self.t, shift))

    def __getitem__(self, k):
        if isinstance(k, slice):
            if k.step not in (None, 1):
                raise RuntimeError("Step size not supported")
            return self.t[k]                      # getChunk -> vector<BigNumber> [UPSTREAM-RECALLED]
        return self.t[k]

    def __add__(self, other):
        if isinstance(other, _PlainText):
            other = self.pk.encrypt(other, False)
        if other.pk != self.pk:
            raise RuntimeError("CipherText +: different public keys")
        if len(other) not in (len(self), 1):
            raise RuntimeError("CipherText +: size mismatch")
        b = other.t if len(other) == len(self) else other.t * len(self)
        return _CipherText(self.pk, [_BN(O.ct_add(self.pk.o, x.v, y.v)) for x, y in zip(self.t, b)])

    def __mul__(self, pt):
        if len(pt) not in (len(self), 1):
            raise RuntimeError("CipherText *: size mismatch")
        e = pt.t if len(pt) == len(self) else pt.t * len(self)
        return _CipherText(self.pk, [_BN(O.ct_mul(self.pk.o, x.v, y.v)) for x, y in zip(self.t, e)])


class _Keypair:
    """ipclKeypair.generate_keypair (ipcl_bindings.cpp:25-29): a seeded key here."""
    seed = 1

    @staticmethod
    def generate_keypair(n_length, enable_DJN):
        pk_o, sk_o = O.seeded_keypair(n_length, _Keypair.seed, enable_DJN)
        pk = _PublicKey(_BN(pk_o.n), n_length, enable_DJN, _hs=pk_o.hs if enable_DJN else None)
        return pk, _PrivateKey(pk, _BN(sk_o.p), _BN(sk_o.q))


def mock_bindings_module():
    m = types.ModuleType("ipcl_bindings")
    m.ipclKeypair, m.ipclPublicKey, m.ipclPrivateKey = _Keypair, _PublicKey, _PrivateKey
    m.ipclPlainText, m.ipclCipherText, m.ipclBigNumber = _PlainText, _CipherText, _BN
    return m


# ------------------------------------------------------------------------------------------------ loading the reference L4
def _gmpy2_stub():
    """The reference inverts ciphertexts with gmpy2.invert (ipcl_python.py:272-276); gmpy2 is not installed here."""
    m = types.ModuleType("gmpy2")

    def invert(x, mod):
        try:
            return pow(int(x), -1, int(mod))
        except ValueError as e:
            raise ZeroDivisionError("invert() no inverse exists") from e

    m.invert = invert
    return m


def load_reference_l4(bindings_module, name="_ref_ipcl_python", l4_dir=None):
    """Execute the reference's ipcl_python.py, unmodified, as package `name` with `bindings_module` standing in for
    the compiled ipcl_bindings.  Returns the ipcl_python module (PaillierKeypair, PaillierPublicKey, ...)."""
    l4_dir = l4_dir or reference_l4_dir()
    if l4_dir is None:
        raise FileNotFoundError("reference L4 sources not found (need /root/reference or baseline/_ref/ipcl_python_ref)")
    for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
        del sys.modules[k]
    if "gmpy2" not in sys.modules:
        try:
            import gmpy2  # noqa: F401
        except ImportError:
            sys.modules["gmpy2"] = _gmpy2_stub()
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    bpkg = types.ModuleType(name + ".bindings")
    bpkg.__path__ = []
    sys.modules[name], sys.modules[name + ".bindings"] = pkg, bpkg
    sys.modules[name + ".bindings.ipcl_bindings"] = bindings_module
    bpkg.ipcl_bindings = bindings_module

    def load(mod_name, path):
        spec = importlib.util.spec_from_file_location(mod_name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[mod_name] = mod
        spec.loader.exec_module(mod)
        return mod

    bpkg.fixedpoint = load(name + ".bindings.fixedpoint", os.path.join(l4_dir, "bindings", "fixedpoint.py"))
    pkg.ipcl_python = load(name + ".ipcl_python", os.path.join(l4_dir, "ipcl_python.py"))
    return pkg.ipcl_python

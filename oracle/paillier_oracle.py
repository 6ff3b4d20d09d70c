"""CPU oracle for the Paillier hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (pailliercryptolib_python_b200) never does.

PARITY STATUS: two halves.
  * The Python layer (ipcl_python.py: exponent alignment, negative-plaintext rule, matmul index maps, add-tree padding)
    IS pinned to the reference: oracle/ref_l4.py runs the reference's own ipcl_python.py + fixedpoint.py, unmodified, over
    a mock of its bindings built on the functions below, and tests/golden/l4_flows.json holds what it produced
    (tests/golden/make_l4_flows.py).
  * The arithmetic is *parity unpinned* at the ciphertext-bit level against IPCL itself.  The reference tree
    (/root/reference) contains no arithmetic: ipcl::PublicKey::encrypt, ipcl::PrivateKey::decrypt,
    ipcl::CipherText::operator+/*, ipcl::modExp and mbx_exp_mb8 live in the un-vendored dependencies
    intel/pailliercryptolib (branch `development`, unpinned; package version 2.0.0, /root/reference/lib/ipcl.cmake:6-7,
    CMakeLists.txt:6) and intel/ipp-crypto, which cannot be fetched or built offline, and the reference's own tests
    hold no golden vectors (/root/reference/tests/ipcl_python_test.py:21-66 are tolerance round trips with random keys).
What *is* pinned:
  * every function below has a unique canonical answer in [0, modulus) given its inputs,
    so exact Python-int arithmetic is bit-exact with any correct implementation;
  * the fixed-point codec and the byte packing are checked against the reference's own
    importable Python (bindings/fixedpoint.py) by tests/golden/make_golden.py;
  * the modexp core is cross-checked against two independent implementations
    (libgmp mpz_powm and OpenSSL BN_mod_exp) in tests/test_oracle.py.

Each function cites the reference call site it restates.
"""
from __future__ import annotations

import math
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------- key material

# The only fixed key material in the reference tree: bench/bench_ipcl_python.py:83-96
BENCH_P = int(
    "17907722236348068892950089903191692955407412936775759886364595"
    "52735277384518331167761570138552647970967958807251538217623805"
    "88199893129274771549316901998509025503556766712439571067562061"
    "82758501008605649830815202920954024506122402034968011655978902"
    "1149844414656481106116277049053335145991958168290159067444243"
)
BENCH_Q = int(
    "15364074494048192090239748141292366255531269713338718185264182"
    "86675686268115568620066283414819003320683895025898634379074026"
    "89773240679814850328978260611055592547225724264355875488478904"
    "93257704058129319548913255512313204302948601763310613641989076"
    "0822812194551465180127077927138009701322446602892596555566791"
)


@dataclass
class PubKey:
    """ipcl::PublicKey state as seen through ipcl_bindings.cpp:66-98 (pickle tuple)."""

    n: int
    bits: int
    djn: bool
    hs: int = 0
    randbits: int = 0

    @property
    def nsquare(self) -> int:
        return self.n * self.n

    @property
    def g(self) -> int:
        return self.n + 1


@dataclass
class PrivKey:
    """ipcl::PrivateKey(pk, p, q): stores p < q (swapped if needed), CRT constants.

    Restates upstream pri_key.cpp (UPSTREAM-RECALLED):
    hp = (L_p(g^(p-1) mod p^2))^-1 mod p, hq likewise, pinv = p^-1 mod q.
    Call site: ipcl_bindings_classes.cpp:96-101.
    """

    pk: PubKey
    p: int
    q: int

    def __post_init__(self):
        if self.p * self.q != self.pk.n:
            raise ValueError("PrivateKey: p * q != n")
        if self.p > self.q:
            self.p, self.q = self.q, self.p
        self.psq = self.p * self.p
        self.qsq = self.q * self.q
        self.hp = h_function(self.pk.g, self.p)
        self.hq = h_function(self.pk.g, self.q)
        self.pinv = pow(self.p, -1, self.q)


def l_function(u: int, x: int) -> int:
    """L_x(u) = (u - 1) / x (exact).  upstream computeLfun."""
    return (u - 1) // x


def h_function(g: int, x: int) -> int:
    """upstream computeHfun(x, x^2) = (L_x(g^(x-1) mod x^2))^-1 mod x."""
    return pow(l_function(pow(g % (x * x), x - 1, x * x), x), -1, x)


def djn_hs(n: int, x: int) -> int:
    """DJN generator: hs = (-x^2 mod n)^n mod n^2  (upstream PublicKey::enableDJN; x random, gcd(x,n)=1)."""
    if math.gcd(x, n) != 1:
        raise ValueError("x not coprime to n")
    h = (-(x % n) * (x % n)) % n
    return pow(h, n, n * n)


def make_pubkey(n: int, bits: int, djn: bool, x: Optional[int] = None, hs: Optional[int] = None) -> PubKey:
    """ipclPublicKey(n, bits, enable_DJN) (ipcl_bindings_classes.cpp:24-27); x or hs pins the DJN generator."""
    if not djn:
        return PubKey(n, bits, False)
    if hs is None:
        if x is None:
            raise ValueError("DJN key needs x or hs")
        hs = djn_hs(n, x)
    return PubKey(n, bits, True, hs, bits // 2)


# ----------------------------------------------------------------------------- the hot path


def raw_encrypt(pk: PubKey, m: int) -> int:
    """ct = (1 + m*n) mod n^2  -- g = n+1 so g^m needs no modexp.

    ipcl::PublicKey::encrypt(pt, make_secure=false); call site ipcl_bindings_classes.cpp:53-60,
    reached from ipcl_python.py:103-106,144.
    """
    return (1 + m * pk.n) % pk.nsquare


def obfuscator(pk: PubKey, r: int) -> int:
    """DJN: hs^r mod n^2 (r < 2^randbits); classic: r^n mod n^2 (r in [1, n-1]).

    ipcl::PublicKey::applyObfuscator -> getDJNObfuscator / getNormalObfuscator -> ippModExp;
    call site ipcl_bindings_classes.cpp:71-83.  r is drawn internally by the reference;
    here it is an input so results are deterministic.
    """
    nsq = pk.nsquare
    if pk.djn:
        return pow(pk.hs, r, nsq)
    return pow(r, pk.n, nsq)


def encrypt(pk: PubKey, m: int, r: Optional[int]) -> int:
    """ct = (1 + m n) * obf(r) mod n^2; r=None means make_secure=False."""
    ct = raw_encrypt(pk, m)
    if r is None:
        return ct
    return ct * obfuscator(pk, r) % pk.nsquare


def decrypt_crt(sk: PrivKey, c: int) -> int:
    """ipcl::PrivateKey::decrypt -> decryptCRT; call site ipcl_bindings_classes.cpp:127-133.

    mp = L_p(c^(p-1) mod p^2) * hp mod p ; mq likewise ; m = mp + ((mq - mp) * pinv mod q) * p.
    """
    p, q = sk.p, sk.q
    mp = l_function(pow(c % sk.psq, p - 1, sk.psq), p) * sk.hp % p
    mq = l_function(pow(c % sk.qsq, q - 1, sk.qsq), q) * sk.hq % q
    return mp + ((mq - mp) * sk.pinv % q) * p


def ct_add(pk: PubKey, a: int, b: int) -> int:
    """ipcl::CipherText::operator+(CipherText) -> raw_add: a*b mod n^2 (ipcl_bindings_classes.cpp:318-321)."""
    return a * b % pk.nsquare


def ct_mul(pk: PubKey, a: int, e: int) -> int:
    """ipcl::CipherText::operator*(PlainText) -> raw_mul -> modExp(a, e, n^2) (ipcl_bindings_classes.cpp:324-325)."""
    return pow(a, e, pk.nsquare)


def modexp(base: int, exp: int, mod: int) -> int:
    """ipcl::modExp / ippModExp element."""
    return pow(base, exp, mod)


# ----------------------------------------------------------------------------- batch helpers


def encrypt_batch(pk: PubKey, ms: Sequence[int], rs: Optional[Sequence[int]]) -> List[int]:
    if rs is None:
        return [raw_encrypt(pk, m) for m in ms]
    return [encrypt(pk, m, r) for m, r in zip(ms, rs)]


def decrypt_batch(sk: PrivKey, cs: Sequence[int]) -> List[int]:
    return [decrypt_crt(sk, c) for c in cs]


def add_batch(pk: PubKey, a: Sequence[int], b: Sequence[int]) -> List[int]:
    """b may have length 1 (broadcast), as in ipcl CipherText::operator+."""
    if len(b) == 1 and len(a) != 1:
        b = list(b) * len(a)
    if len(a) != len(b):
        raise ValueError("size mismatch")
    return [ct_add(pk, x, y) for x, y in zip(a, b)]


def mul_batch(pk: PubKey, a: Sequence[int], e: Sequence[int]) -> List[int]:
    if len(e) == 1 and len(a) != 1:
        e = list(e) * len(a)
    if len(a) != len(e):
        raise ValueError("size mismatch")
    return [ct_mul(pk, x, y) for x, y in zip(a, e)]


# ----------------------------------------------------------------------------- packing (BigNumber <-> bytes/limbs)


def int_to_le_bytes(val: int) -> bytes:
    """BNUtils.int2Bytes (ipcl_python.py:936-937): minimal-length little-endian bytes."""
    return val.to_bytes((val.bit_length() + 7) // 8, "little")


def bn_to_bytes(val: int) -> bytes:
    """BN2bytes (ipcl_bindings.cpp:121-129): LE bytes padded to BITSIZE_WORD(bitlen)*4; zero is one word."""
    words = max(1, (val.bit_length() + 31) // 32)
    return val.to_bytes(words * 4, "little")


def bytes_to_int(b: bytes) -> int:
    """pyByte2BN (ipcl_bindings.cpp:100-119) followed by BN2int: plain little-endian interpretation."""
    return int.from_bytes(b, "little")


def to_limbs(vals: Sequence[int], k: int):
    """Pack ints into a [N, k] uint32 little-endian limb array (the C-ABI layout)."""
    import numpy as np

    buf = b"".join(int(v).to_bytes(4 * k, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), k).copy()


def from_limbs(arr) -> List[int]:
    import numpy as np

    arr = np.ascontiguousarray(arr, dtype="<u4")
    k = arr.shape[1]
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * 4 * k : (i + 1) * 4 * k], "little") for i in range(arr.shape[0])]


# ----------------------------------------------------------------------------- fixed-point codec

FLOAT_MANTISSA_BITS = sys.float_info.mant_dig  # 53
BASE = 2


def fp_encode(scalar, n: int, max_int: int) -> Tuple[int, int]:
    """FixedPointNumber.encode (bindings/fixedpoint.py:55-96) with precision=None, max_exponent=None.

    Returns (encoding mod n, exponent).  ints (python / numpy integer) -> exponent 0; floats ->
    exponent = 53 - frexp(x)[1].
    """
    import numpy as np

    if abs(scalar) < 1e-200:
        scalar = 0
    if isinstance(scalar, (int, np.int16, np.int32, np.int64)) and not isinstance(scalar, bool):
        exponent = 0
    elif isinstance(scalar, (float, np.float16, np.float32, np.float64)):
        exponent = FLOAT_MANTISSA_BITS - math.frexp(scalar)[1]
    else:
        raise TypeError("Don't know the precision of type %s." % type(scalar))
    int_fixpoint = int(round(scalar * pow(BASE, exponent)))
    if abs(int_fixpoint) > max_int:
        raise ValueError("Integer needs to be within +/- %d" % max_int)
    return int_fixpoint % n, exponent


def fp_decode(encoding: int, exponent: int, n: int, max_int: int):
    """FixedPointNumber.decode (bindings/fixedpoint.py:98-115)."""
    if encoding >= n:
        raise ValueError("Attempted to decode corrupted number")
    if encoding <= max_int:
        mantissa = encoding
    elif encoding >= n - max_int:
        mantissa = encoding - n
    else:
        raise OverflowError("Overflow detected in decode number")
    return mantissa * pow(BASE, -exponent)


# ----------------------------------------------------------------------------- deterministic test keys


def _is_probable_prime(n: int, rng) -> bool:
    if n < 2:
        return False
    for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % sp == 0:
            return n == sp
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for _ in range(24):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def seeded_prime(bits: int, rng) -> int:
    """Random prime with the top two bits set and p = 3 (mod 4) (as upstream keygen requires)."""
    while True:
        c = rng.getrandbits(bits) | (3 << (bits - 2)) | 3
        if _is_probable_prime(c, rng):
            return c


def seeded_keypair(bits: int, seed: int, djn: bool = True) -> Tuple[PubKey, PrivKey]:
    """Deterministic key for tests / fixtures (any even bits, incl. 3072 which the reference rejects)."""
    import random

    rng = random.Random(seed)
    while True:
        p = seeded_prime(bits // 2, rng)
        q = seeded_prime(bits // 2, rng)
        if p != q and (p * q).bit_length() == bits and math.gcd(p - 1, q - 1) == 2:
            break
    n = p * q
    x = rng.getrandbits(bits + 128)
    while math.gcd(x, n) != 1:
        x = rng.getrandbits(bits + 128)
    pk = make_pubkey(n, bits, djn, x=x)
    return pk, PrivKey(pk, p, q)


def bench_keypair(djn: bool = True, x_seed: int = 20240611) -> Tuple[PubKey, PrivKey]:
    """The reference bench's fixed 2048-bit key (bench/bench_ipcl_python.py:83-101) with a seeded DJN x."""
    import random

    n = BENCH_P * BENCH_Q
    rng = random.Random(x_seed)
    x = rng.getrandbits(n.bit_length() + 128)
    while math.gcd(x, n) != 1:
        x = rng.getrandbits(n.bit_length() + 128)
    pk = make_pubkey(n, n.bit_length(), djn, x=x)
    return pk, PrivKey(pk, BENCH_P, BENCH_Q)

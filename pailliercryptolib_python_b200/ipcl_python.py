"""The ipcl_python API surface on top of the B200 engine: PaillierKeypair, PaillierPublicKey, PaillierPrivateKey,
PaillierEncryptedNumber, BNUtils -- same names, arguments, return types and error behaviour as the reference
(/root/reference/src/ipcl_python/ipcl_python.py), so existing callers switch by changing the import.

What is different underneath (B200-first, not a translation):
  * a PaillierEncryptedNumber keeps its ciphertexts as ONE packed [N, 2*n_words] uint32 matrix inside an ipclCipherText
    and its exponents as an int64 vector; every operator is a handful of whole-batch calls into libphe_b200.so
    (encrypt / decrypt / modmul / modexp kernels), never a Python loop over ipclBigNumber objects;
  * the fixed-point codec is vectorised (fixedpoint.encode_array / decode_array);
  * ciphertext batches stay in HBM from encrypt to decrypt: exponent alignment squares the rows that need scaling in
    place of a device copy (phe_scale_rows_dev; reference: ipcl_python.py:528-741, per-element lists), matmul gathers
    its operands by index on the device (phe_gather_rows_dev; :777-808), only exponents -- host metadata -- are looked at
    on the host;
  * negative plaintext factors still invert the ciphertext first (ipcl_python.py:272-276, 426-441, 470-479) so results
    are bit-identical to the reference's; the rows concerned are inverted on the device (phe_invert_rows_dev);
  * sum()/dot()/@ reduce with a log-depth tree of one-product HE-adds in one device call (phe_segsum_dev); sum() works
    for any length (the reference's passes a list where it needs a ciphertext container, ipcl_python.py:752-755).
"""
from typing import Optional, Tuple, Union

import numpy as np

from .bindings.ipcl_bindings import (
    ipclBigNumber,
    ipclCipherText,
    ipclKeypair,
    ipclPlainText,
    ipclPrivateKey,
    ipclPublicKey,
)
from .fixedpoint import FixedPointNumber, decode_mantissas, encode_array, encode_signed_array

_NUMBER = (int, float, np.integer, np.floating)


def _ints_to_limbs(vals, words):
    buf = b"".join(int(v).to_bytes(4 * words, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), words).astype(np.uint32)


def _limbs_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype="<u4")
    step = arr.shape[1] * 4
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * step:(i + 1) * step], "little") for i in range(arr.shape[0])]


class PaillierKeypair:
    @staticmethod
    def generate_keypair(n_length: int = 1024, enable_DJN: bool = True) -> Tuple["PaillierPublicKey", "PaillierPrivateKey"]:
        """Generate a key pair (ipcl_python.py:20-40).  Unlike the reference (<= 2048 bits) keys up to 3072 bits work."""
        pub, pri = ipclKeypair.generate_keypair(n_length, enable_DJN)
        return PaillierPublicKey(pub), PaillierPrivateKey(pri)


class PaillierPublicKey:
    def __init__(self, key: Union[ipclPublicKey, "PaillierPublicKey", int], n_length: Optional[int] = None,
                 enable_DJN: Optional[bool] = None):
        if isinstance(key, ipclPublicKey):
            self.pubkey = key
            self.n = BNUtils.BN2int(key.n)
        elif isinstance(key, PaillierPublicKey):
            # the reference's `self = key` leaves the object unset; share the handle instead
            self.pubkey, self.n = key.pubkey, key.n
        elif isinstance(key, int) and n_length is not None and enable_DJN is not None:
            self.n = key
            self.pubkey = ipclPublicKey(BNUtils.int2BN(key), n_length, enable_DJN)
        else:
            raise ValueError("PaillierPublicKey: PubKey should be either key value (n),"
                             "PaillierPublicKey or IPP-PaillierPublicKey object")
        self._derive()

    def _derive(self):
        self.max_int = self.n // 3 - 1
        self.nsquare = self.n * self.n
        self.n_words = (max(self.pubkey.length, self.n.bit_length()) + 31) // 32

    def __getstate__(self):
        return self.pubkey

    def __setstate__(self, state):
        self.pubkey = state
        self.n = BNUtils.BN2int(state.n)
        self._derive()

    def __repr__(self):
        return repr(self.pubkey)

    def __eq__(self, other):
        return self.n == other.n

    def __hash__(self):
        return hash(self.pubkey)

    def apply_obfuscator(self, x: Union[int, ipclBigNumber]):
        if isinstance(x, int):
            x = BNUtils.int2BN(x)
        return self.pubkey.apply_obfuscator(x)

    def raw_encrypt(self, plaintext) -> "PaillierEncryptedNumber":
        return self.encrypt(plaintext, apply_obfuscator=False)

    def encrypt(self, values: Union[np.ndarray, list, int, float], apply_obfuscator: bool = True) -> "PaillierEncryptedNumber":
        """Encrypt a scalar or a 1-D list/array of ints/floats into ONE PaillierEncryptedNumber holding the batch
        (ipcl_python.py:108-147)."""
        if np.isscalar(values):
            values = [values]
        arr = values if isinstance(values, np.ndarray) else None
        if arr is None or arr.dtype == object:
            if not all(isinstance(v, _NUMBER) for v in values):
                raise ValueError("PaillierPublicKey.encrypt: input value(s) should be integer or float")
        elif arr.ndim != 1 or arr.dtype.kind not in "iuf":
            raise ValueError("PaillierPublicKey.encrypt: input value(s) should be integer or float")
        limbs, expos = encode_array(values, self.n, self.max_int, self.n_words, compact=True)
        ct = self.pubkey.encrypt(ipclPlainText.from_packed(limbs), apply_obfuscator)
        return PaillierEncryptedNumber(self, ct, exponents=expos, length=len(expos))


class PaillierPrivateKey:
    def __init__(self, key: Union[ipclPrivateKey, ipclPublicKey, PaillierPublicKey], p: Optional[int] = None,
                 q: Optional[int] = None):
        if isinstance(key, ipclPrivateKey):
            self.prikey = key
            self.__n = BNUtils.BN2int(key.n)
        elif isinstance(key, ipclPublicKey) and p is not None and q is not None:
            self.prikey = ipclPrivateKey(key, BNUtils.int2BN(p), BNUtils.int2BN(q))
            self.__n = BNUtils.BN2int(key.n)
        elif isinstance(key, PaillierPublicKey) and p is not None and q is not None:
            self.prikey = ipclPrivateKey(key.pubkey, BNUtils.int2BN(p), BNUtils.int2BN(q))
            self.__n = key.n
        else:
            raise KeyError("PaillierPrivateKey: key should be either Private key or Public key (with p and q)")
        self.__max_int = self.__n // 3 - 1

    def __getstate__(self):
        return (self.prikey, self.__n, self.__max_int)

    def __setstate__(self, state):
        (self.prikey, self.__n, self.__max_int) = state

    def __eq__(self, other: "PaillierPrivateKey"):
        return (self.prikey.p == other.prikey.p) and (self.prikey.q == other.prikey.q)

    def __hash__(self):
        return hash(self.prikey)

    def __repr__(self):
        return repr(self.prikey)

    def _decrypt_packed(self, number: "PaillierEncryptedNumber", who: str) -> np.ndarray:
        if number.public_key.n != self.__n:
            raise ValueError("%s: Public key mismatch" % who)
        return self.prikey.decrypt(number.ciphertext()).to_packed()

    def raw_decrypt(self, ciphertext: "PaillierEncryptedNumber"):
        vals = _limbs_to_ints(self._decrypt_packed(ciphertext, "PaillierPrivateKey.raw_decrypt"))
        return vals if len(ciphertext) > 1 else vals[0]

    def decrypt(self, encrypted_number: "PaillierEncryptedNumber"):
        """Decrypt and decode: list of ints/floats, or a single value for a length-1 ciphertext (ipcl_python.py:219-245)."""
        if encrypted_number.public_key.n != self.__n:
            raise ValueError("PailierPrivateKey.decrypt: Public key mismatch")
        # the plaintexts are classified on the device (sign, 63-bit mantissa): 9 bytes per element come back instead of
        # 4 n_words, and only rows that are neither small positive nor small negative are decoded from their words
        mant, cls, rows = self.prikey.decrypt_mantissas(encrypted_number.ciphertext())
        vals = decode_mantissas(mant, cls, rows, encrypted_number._expo_array(), self.__n, self.__max_int)
        return vals if len(encrypted_number) > 1 else vals[0]


class PaillierEncryptedNumber:
    # numpy must not broadcast over this object: `ndarray + ct`, `ndarray * ct`, `ndarray @ ct` defer to the reflected
    # operators below (the reference only works with a *list* on the left, tests/ipcl_python_test.py:93-98)
    __array_ufunc__ = None

    def __init__(self, public_key: PaillierPublicKey, ciphertext: ipclCipherText, exponents, length: int):
        if ciphertext.public_key != public_key.pubkey:
            raise ValueError("PaillierEncryptedNumber: public key mismatch")
        self.public_key = public_key
        self.__ct = ciphertext
        self.__expo = np.asarray(exponents, dtype=np.int64).reshape(-1)
        self.__length = length

    # ---- plumbing ---------------------------------------------------------------------------------------------
    def _wrap(self, packed, exponents) -> "PaillierEncryptedNumber":
        ct = packed if isinstance(packed, ipclCipherText) else ipclCipherText.from_packed(self.public_key.pubkey, packed)
        return PaillierEncryptedNumber(self.public_key, ct, exponents, len(ct))

    def packed(self) -> np.ndarray:
        """[N, 2*n_words] uint32 little-endian limb matrix of the ciphertexts (a copy)."""
        return self.__ct.to_packed()

    def __repr__(self):
        return repr(self.__ct)

    def __getstate__(self) -> tuple:
        return (self.public_key, len(self), self.exponent(), _limbs_to_ints(self.packed()))

    def __setstate__(self, state: tuple):
        self.public_key, self.__length, expo, ints = state
        self.__expo = np.asarray(expo, dtype=np.int64).reshape(-1)
        self.__ct = ipclCipherText.from_packed(self.public_key.pubkey, _ints_to_limbs(ints, 2 * self.public_key.n_words))

    def __len__(self) -> int:
        return self.__length

    def length(self) -> int:
        return self.__length

    def ciphertext(self) -> ipclCipherText:
        return self.__ct

    def ciphertextBN(self, idx: Optional[int] = None):
        if idx is None:
            return self.__ct.getTexts()
        if not 0 <= idx < self.__length:
            raise IndexError("ciphertext: idx out of range")
        return self.__ct[idx]

    def _expo_array(self) -> np.ndarray:
        return self.__expo

    def exponent(self, idx: Optional[int] = None):
        if idx is None:
            return [int(e) for e in self.__expo]
        if not 0 <= idx < self.__length:
            raise IndexError("exponent: idx out of range")
        return int(self.__expo[idx])

    def apply_obfuscator(self):
        self.__ct = self.public_key.pubkey.apply_obfuscator_packed(self.__ct)

    def __getitem__(self, key: Union[int, slice]) -> "PaillierEncryptedNumber":
        if isinstance(key, (int, np.integer)):
            key = slice(int(key), int(key) + 1)
        start = 0 if key.start is None else key.start
        stop = len(self) if key.stop is None else key.stop
        if key.step not in (None, 1):
            raise RuntimeError("Step size not supported")
        if not 0 <= stop <= len(self) or not 0 <= start < len(self):
            raise IndexError("__getitem__: key out of range")
        return self._wrap(self.__ct[start:stop], self.__expo[start:stop])

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    # ---- homomorphic add ----------------------------------------------------------------------------------------
    def __add__(self, other):
        if self.__length == 1 and isinstance(other, PaillierEncryptedNumber) and len(other) > 1:
            return other.__raw_add(self)
        return self.__raw_add(other)

    def __radd__(self, other):
        return self + other

    def __sub__(self, other):
        if isinstance(other, list):
            other = np.array(other)
        return self.__raw_add(other * -1.0)

    def __rsub__(self, other):
        if isinstance(other, PaillierEncryptedNumber):
            return other - self
        return (self * (-1.0)).__raw_add(other)

    def __raw_add(self, other) -> "PaillierEncryptedNumber":
        if isinstance(other, (np.ndarray, list)):
            if self.__length != len(other):
                raise ValueError("PaillierEncryptedNumber.__raw_add: array(list) size mismatch with PaillierEncryptedNumber")
            other = self.public_key.encrypt(other, apply_obfuscator=False)
        elif np.isscalar(other) and isinstance(other, (int, float)):
            other = self.public_key.encrypt(other, apply_obfuscator=False)
        elif isinstance(other, PaillierEncryptedNumber):
            if self.public_key != other.public_key:
                raise ValueError("PaillierEncryptedNumber.__raw_add: PublicKey mismatch")
            if self.__length != len(other) and len(other) > 1:
                raise ValueError("PaillierEncryptedNumber.__raw_add: CipherText size mismatch with PaillierEncryptedNumber")
        x_ct, y_ct, expo = self.__align_exponent(self.__ct, self.__expo, other.ciphertext(), other.__expo)
        return self._wrap(x_ct + y_ct, expo)

    @staticmethod
    def _scale_rows(ct: ipclCipherText, rows: np.ndarray, deltas: np.ndarray) -> ipclCipherText:
        """ct with ct[rows] <- ct[rows] ^ (2^delta) mod n^2: delta squarings per row on the device, the batch never
        leaves HBM (the reference multiplies by the plaintext BASE^delta element by element: ipcl_python.py:551-560,
        602-606).  Any delta works (the reference's modExp has no exponent-size limit either)."""
        if rows.size == 0:
            return ct
        return ct.scale_rows(np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(deltas, dtype=np.int32))

    def increase_exponent_to(self, x_ct: ipclCipherText, x_expo, exponent: int) -> ipclCipherText:
        """Raise every element whose exponent is below `exponent` (ipcl_python.py:528-568)."""
        diff = int(exponent) - np.asarray(x_expo, dtype=np.int64)
        rows = np.nonzero(diff > 0)[0]
        return self._scale_rows(x_ct, rows, diff[rows])

    def __align_exponent(self, x_ct, x_expo, y_ct, y_expo):
        """Bring both operands to max(exponent) per element; y may be a single broadcast ciphertext
        (ipcl_python.py:570-741).  Row selection happens on the host (exponents are host metadata), the rows
        themselves are scaled in place of a device copy."""
        x_expo = np.asarray(x_expo, dtype=np.int64)
        y_expo = np.asarray(y_expo, dtype=np.int64)
        count = len(x_ct)
        y_bcast = np.broadcast_to(y_expo, (count,)) if len(y_ct) == 1 else y_expo
        out_expo = np.maximum(x_expo, y_bcast)
        x_rows = np.nonzero(x_expo < y_bcast)[0]
        y_rows = np.nonzero(y_bcast < x_expo)[0]
        x_ct = self._scale_rows(x_ct, x_rows, (y_bcast - x_expo)[x_rows])
        if y_rows.size:
            if len(y_ct) == 1 and count > 1:      # the broadcast operand needs different factors per row: expand it
                y_ct = y_ct.gather(np.zeros(count, dtype=np.int64))
            y_ct = self._scale_rows(y_ct, y_rows, (x_expo - y_bcast)[y_rows])
        return x_ct, y_ct, out_expo

    # ---- homomorphic multiply by plaintext ---------------------------------------------------------------------------
    def __invert_rows(self, ct: ipclCipherText, rows: np.ndarray) -> ipclCipherText:
        """ct with ct[rows] <- inverses modulo n^2, one batched device call (Montgomery's trick on the gathered rows);
        the reference calls gmpy2.invert per element (ipcl_python.py:272-276)."""
        if rows.size == 0:
            return ct
        try:
            if rows.size == len(ct):
                return ct.modinv()
            return ct.invert_rows(np.ascontiguousarray(rows, dtype=np.int64))
        except RuntimeError:
            # some element shares a factor with n: reproduce the element-wise error of the reference's gmpy2.invert
            nsq = self.public_key.nsquare
            packed = ct.to_packed()
            vals = _limbs_to_ints(packed[rows])
            try:
                packed[rows] = _ints_to_limbs([pow(v, -1, nsq) for v in vals], packed.shape[1])
            except ValueError as e:
                raise ZeroDivisionError("invert() no inverse exists") from e
            return ipclCipherText.from_packed(self.public_key.pubkey, packed)

    def _mul_encoded(self, ct: ipclCipherText, pt_limbs: np.ndarray, ct_expo, pt_expo):
        """ct[i] ^ pt[i] with the reference's negative-plaintext rule: if pt >= n - max_int use (ct^-1)^(n - pt) so the
        exponent stays short (ipcl_python.py:426-441, 470-479).  pt_limbs: [N or 1, n_words] (host: plaintexts start
        there).  The ciphertext batch stays on the device: the rows that meet a negative plaintext are inverted there."""
        pk = self.public_key
        n_l = _ints_to_limbs([pk.n], pk.n_words)[0]
        thr = _ints_to_limbs([pk.n - pk.max_int], pk.n_words)[0]
        # lexicographic compare pt >= thr from the top word down
        ge = np.ones(pt_limbs.shape[0], dtype=bool)
        decided = np.zeros(pt_limbs.shape[0], dtype=bool)
        for j in range(pk.n_words - 1, -1, -1):
            col = pt_limbs[:, j]
            gt, lt = (col > thr[j]) & ~decided, (col < thr[j]) & ~decided
            ge[lt] = False
            decided |= gt | lt
            if decided.all():
                break
        neg = np.nonzero(ge)[0]
        if neg.size:
            pt_limbs = pt_limbs.copy()
            borrow = np.zeros(neg.size, dtype=np.int64)
            for j in range(pk.n_words):
                v = np.int64(int(n_l[j])) - pt_limbs[neg, j].astype(np.int64) - borrow
                borrow = (v < 0).astype(np.int64)
                pt_limbs[neg, j] = (v & 0xFFFFFFFF).astype(np.uint32)
            if pt_limbs.shape[0] == 1 and len(ct) > 1:
                ct = self.__invert_rows(ct, np.arange(len(ct)))
            else:
                ct = self.__invert_rows(ct, neg)
        used = pk.n_words
        while used > 1 and not pt_limbs[:, used - 1].any():
            used -= 1
        res = ct * ipclPlainText.from_packed(np.ascontiguousarray(pt_limbs[:, :used]))
        return res, np.asarray(ct_expo, dtype=np.int64) + np.asarray(pt_expo, dtype=np.int64)

    def _mul_plain(self, ct: ipclCipherText, ct_expo, values):
        """ct * values (a 1-D sequence of len(ct) or 1 plaintext numbers): (result ciphertexts, result exponents)."""
        pk = self.public_key
        signed = encode_signed_array(values, pk.max_int)
        if signed is None:      # big Python ints, object arrays: the general path on n_words-wide encodings
            pt_limbs, pt_expo = encode_array(values, pk.n, pk.max_int, pk.n_words)
            return self._mul_encoded(ct, pt_limbs, ct_expo, pt_expo)
        # floats and 64-bit ints: magnitude and sign come straight from the codec -- the rows that meet a negative
        # plaintext are inverted on the device and raised to |mantissa| (ipcl_python.py:426-441, 470-479)
        mag, neg, pt_expo = signed
        rows = np.nonzero(neg)[0]
        if rows.size:
            ct = self.__invert_rows(ct, np.arange(len(ct)) if (mag.shape[0] == 1 and len(ct) > 1) else rows)
        used = 2 if mag[:, 1].any() else 1
        res = ct * ipclPlainText.from_packed(np.ascontiguousarray(mag[:, :used]))
        return res, np.asarray(ct_expo, dtype=np.int64) + pt_expo

    def __mul__(self, other) -> "PaillierEncryptedNumber":
        if np.isscalar(other):
            other = [other]
        elif len(other) != self.__length:
            raise ValueError("PaillierEncryptedNumber.__mul__: Multiply size mismatch")
        res, expo = self._mul_plain(self.__ct, self.__expo, other)
        return self._wrap(res, expo)

    def __rmul__(self, other):
        return self * other

    def __truediv__(self, other):
        if isinstance(other, list):
            other = np.array(other)
        return self * (1.0 / other)

    # ---- reductions ---------------------------------------------------------------------------------------------------
    def sum(self) -> "PaillierEncryptedNumber":
        """HE-sum of all elements: exponents aligned to the largest, then a log-depth tree of one-product additions on
        the device (phe_segsum_dev).  Works for any length; the reference's own sum() passes a list where it needs a
        ciphertext container (ipcl_python.py:752-755), its working form is the rotate-and-add tree of :810-827."""
        top = int(self.__expo.max())
        aligned = self.increase_exponent_to(self.__ct, self.__expo, top)
        return self._wrap(aligned.segsum(1, len(self)), [top])

    def mean(self) -> "PaillierEncryptedNumber":
        return self.sum() / len(self)

    def dot(self, other) -> "PaillierEncryptedNumber":
        if len(other) != len(self):
            raise ValueError("PaillierEncryptedNumber.dot: input size mismatch with ciphertext")
        return (self * other).sum()

    def __matmul(self, other: np.ndarray, m: int, n: int, k: int, rhs: bool) -> "PaillierEncryptedNumber":
        """(m x n) @ (n x k): one batched HE-mul over all m*k*n products, per-output exponent alignment, tree sum.
        Index maps as ipcl_python.py:777-808."""
        pk = self.public_key
        ii, jj, ll = np.meshgrid(np.arange(m), np.arange(k), np.arange(n), indexing="ij")
        if rhs:   # other (m x n) @ self (n x k)
            idx_self = (ll * k + jj).reshape(-1)
            pts = other[ii, ll].reshape(-1) if other.ndim == 2 else other[ll].reshape(-1)
        else:     # self (m x n) @ other (n x k)
            idx_self = (ii * n + ll).reshape(-1)
            pts = other[ll, jj].reshape(-1) if other.ndim == 2 else other[ll].reshape(-1)
        operands = self.__ct.gather(np.ascontiguousarray(idx_self, dtype=np.int64))       # device gather by the index map
        prod, expo = self._mul_plain(operands, self.__expo[idx_self], pts)
        expo = expo.reshape(m * k, n)
        top = expo.max(axis=1)
        delta = (top[:, None] - expo).reshape(-1)
        rows = np.nonzero(delta > 0)[0]
        prod = self._scale_rows(prod, rows, delta[rows])
        return self._wrap(prod.segsum(m * k, n), top)

    def __matmul__(self, other) -> "PaillierEncryptedNumber":
        if len(self) % len(other) != 0:
            raise ValueError("PaillierEncryptedNumber.__matmul__: matrix multiply size mismatch")
        other = np.array(other)
        if other.ndim not in (1, 2):
            raise NotImplementedError("PaillierEncryptedNumber.__matmul__: input ndim %dnot supported" % other.ndim)
        n = other.shape[0]
        k = other.shape[1] if other.ndim == 2 else 1
        return self.__matmul(other, len(self) // n, n, k, rhs=False)

    def __rmatmul__(self, other) -> "PaillierEncryptedNumber":
        other = np.array(other)
        if other.ndim not in (1, 2):
            raise NotImplementedError("PaillierEncryptedNumber.__rmatmul__: input ndim %d not supported" % other.ndim)
        m = other.shape[0] if other.ndim == 2 else 1
        n = other.shape[1] if other.ndim == 2 else other.shape[0]
        if len(self) % n != 0:
            raise ValueError("PaillierEncryptedNumber.__rmatmul__: matrix multiplysize mismatch")
        return self.__matmul(other, m, n, len(self) // n, rhs=True)

    def __imatmul__(self, other) -> "PaillierEncryptedNumber":
        return self @ other


class BNUtils:
    """Python int <-> ipclBigNumber through little-endian bytes (ipcl_python.py:933-977)."""

    @staticmethod
    def int2Bytes(val: int) -> bytes:
        return val.to_bytes((val.bit_length() + 7) // 8, byteorder="little")

    @staticmethod
    def bytes2Int(val: bytes) -> int:
        return int.from_bytes(val, "little")

    @staticmethod
    def int2BN(val: int) -> ipclBigNumber:
        if val in (0, 1, 2):
            return (ipclBigNumber.Zero, ipclBigNumber.One, ipclBigNumber.Two)[val]
        return ipclBigNumber(BNUtils.int2Bytes(val))

    @staticmethod
    def BN2int(val: ipclBigNumber) -> int:
        return BNUtils.bytes2Int(val.to_bytes())

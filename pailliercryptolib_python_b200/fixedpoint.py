"""Fixed-point codec of the ipcl_python API: float / int  <->  (mantissa * 2^exponent mod n, exponent).

Behavioural mirror of the reference's FATE-derived codec (/root/reference/src/ipcl_python/bindings/fixedpoint.py:26-187):
same class name, same encode/decode results (pinned by tests/golden/fixedpoint.json, which was generated with the
reference's own module), same error types.  New here: `encode_array` / `decode_array`, the vectorised numpy codec that
produces / consumes the packed little-endian uint32 limb matrices the CUDA path works on, so a batch of 100k values
never becomes 100k Python objects.
"""
import math
import sys

import numpy as np

_INT_TYPES = (int, np.int16, np.int32, np.int64)
_FLOAT_TYPES = (float, np.float16, np.float32, np.float64)


class FixedPointNumber(object):
    """A number encoded as `encoding = round(x * BASE^exponent) mod n`; negatives live in the top third of [0, n)."""

    BASE = 2
    LOG2_BASE = math.log(BASE, 2)
    FLOAT_MANTISSA_BITS = sys.float_info.mant_dig
    Q = 293973345475167247070445277780365744413 ** 2   # default modulus of the reference codec (fixedpoint.py:33)

    def __init__(self, encoding, exponent, n=None, max_int=None):
        self.n = FixedPointNumber.Q if n is None else n
        self.max_int = self.n // 2 if (n is None or max_int is None) else max_int
        self.encoding = encoding
        self.exponent = exponent

    # ------------------------------------------------------------------ scalar codec (fixedpoint.py:50-115)
    @classmethod
    def calculate_exponent_from_precision(cls, precision):
        return math.floor(math.log(precision, cls.BASE))

    @classmethod
    def encode(cls, scalar, n=None, max_int=None, precision=None, max_exponent=None):
        if np.abs(scalar) < 1e-200:      # flush tiny values (avoids int overflow in the scaling below)
            scalar = 0
        if n is None:
            n, max_int = cls.Q, cls.Q // 2
        if precision is not None:
            exponent = cls.calculate_exponent_from_precision(precision)
        elif isinstance(scalar, _INT_TYPES):
            exponent = 0
        elif isinstance(scalar, _FLOAT_TYPES):
            exponent = math.floor((cls.FLOAT_MANTISSA_BITS - math.frexp(scalar)[1]) / cls.LOG2_BASE)
        else:
            raise TypeError("Don't know the precision of type %s." % type(scalar))
        if max_exponent is not None:
            exponent = max(max_exponent, exponent)
        fixed = int(round(scalar * pow(cls.BASE, exponent)))
        if abs(fixed) > max_int:
            raise ValueError("Integer needs to be within +/- %d,but got %d,basic info, scalar=%s, base=%d, exponent=%d"
                             % (max_int, fixed, scalar, cls.BASE, exponent))
        return cls(fixed % n, exponent, n, max_int)

    def _mantissa(self):
        if self.encoding >= self.n:
            raise ValueError("Attempted to decode corrupted number")
        if self.encoding <= self.max_int:
            return self.encoding
        if self.encoding >= self.n - self.max_int:
            return self.encoding - self.n
        raise OverflowError("Overflow detected in decode number, encoding: %d, %d %d" % (self.encoding, self.exponent, self.n))

    def decode(self):
        return self._mantissa() * pow(self.BASE, -self.exponent)

    def increase_exponent_to(self, new_exponent):
        if new_exponent < self.exponent:
            raise ValueError("New exponent %i should be greater thanold exponent %i" % (new_exponent, self.exponent))
        scaled = self.encoding * pow(self.BASE, new_exponent - self.exponent) % self.n
        return FixedPointNumber(scaled, new_exponent, self.n, self.max_int)

    # ------------------------------------------------------------------ arithmetic on encodings (fixedpoint.py:129-187)
    def _same_modulus(self, other):
        if other.n != self.n:
            other = self.encode(other.decode(), n=self.n, max_int=self.max_int)
        return other

    def _aligned(self, other):
        x, y = self, other
        if x.exponent < y.exponent:
            x = x.increase_exponent_to(y.exponent)
        elif y.exponent < x.exponent:
            y = y.increase_exponent_to(x.exponent)
        return x, y

    def _combine(self, other, sign):
        if not isinstance(other, FixedPointNumber):
            other = self.encode(other, n=self.n, max_int=self.max_int)
        x, y = self._aligned(self._same_modulus(other))
        return FixedPointNumber((x.encoding + sign * y.encoding) % self.n, x.exponent, n=self.n, max_int=self.max_int)

    @staticmethod
    def _is_encrypted(obj):
        return type(obj).__name__ == "PaillierEncryptedNumber"

    def __add__(self, other):
        if self._is_encrypted(other):
            return other + self.decode()
        return self._combine(other, +1)

    __radd__ = __add__

    def __sub__(self, other):
        if self._is_encrypted(other):
            return (other - self.decode()) * -1
        if isinstance(other, FixedPointNumber):
            return self._combine(other, -1)
        return self._combine(-1 * other, +1)

    def __rsub__(self, other):
        if self._is_encrypted(other):
            return other - self.decode()
        return self.encode(-1 * self.__sub__(other).decode(), n=self.n, max_int=self.max_int)

    def __mul__(self, other):
        if self._is_encrypted(other):
            return other * self.decode()
        factor = other.decode() if isinstance(other, FixedPointNumber) else other
        return FixedPointNumber.encode(self.decode() * factor, n=self.n, max_int=self.max_int)

    __rmul__ = __mul__

    def __truediv__(self, other):
        divisor = other.decode() if isinstance(other, FixedPointNumber) else other
        return self.__mul__(1 / divisor)

    def __rtruediv__(self, other):
        return FixedPointNumber.encode(1.0 / self.__truediv__(other).decode(), n=self.n, max_int=self.max_int)

    def __abs__(self):
        if self.encoding <= self.max_int:
            return self
        if self.encoding >= self.n - self.max_int:
            return self * -1
        return None

    def __mod__(self, other):
        return FixedPointNumber(self.encoding % other, self.exponent, n=self.n, max_int=self.max_int)

    @staticmethod
    def _plain(v):
        return v.decode() if isinstance(v, FixedPointNumber) else v

    def __lt__(self, other): return bool(self.decode() < self._plain(other))
    def __gt__(self, other): return bool(self.decode() > self._plain(other))
    def __le__(self, other): return bool(self.decode() <= self._plain(other))
    def __ge__(self, other): return bool(self.decode() >= self._plain(other))
    def __eq__(self, other): return bool(self.decode() == self._plain(other))
    def __ne__(self, other): return bool(self.decode() != self._plain(other))
    __hash__ = None


# ---------------------------------------------------------------------------------------------- vectorised codec

def _n_limbs(n, words):
    return np.frombuffer(int(n).to_bytes(4 * words, "little"), dtype="<u4").astype(np.uint32)


def _sub_from_modulus(n_limbs, mag_lo, mag_hi, rows):
    """[len(rows), words] limbs of n - mag for 64-bit magnitudes (mag_lo, mag_hi uint32 arrays), mag > 0."""
    words = n_limbs.shape[0]
    out = np.empty((rows, words), dtype=np.uint32)
    borrow = np.zeros(rows, dtype=np.int64)
    for j in range(words):
        sub = mag_lo.astype(np.int64) if j == 0 else (mag_hi.astype(np.int64) if j == 1 else 0)
        v = np.int64(int(n_limbs[j])) - sub - borrow
        borrow = (v < 0).astype(np.int64)
        out[:, j] = (v & 0xFFFFFFFF).astype(np.uint32)
        if j >= 2 and not borrow.any():
            out[:, j + 1:] = n_limbs[j + 1:]
            break
    return out


def encode_signed_array(values, max_int):
    """The vector path of encode_array without the reduction modulo n: (|mantissa| as [N, 2] uint32 limbs, sign flags,
    exponents), or None when the input needs the scalar codec (object arrays, big Python ints, mixed lists).  The
    encoding is mantissa mod n, i.e. n - |mantissa| for the flagged rows -- HE mul wants exactly the magnitude and the
    sign (ipcl_python.py:426-441: a negative plaintext is n - pt with the ciphertext inverted), so building the n_words-wide
    n - |mantissa| on the host only to subtract it from n again is skipped."""
    if isinstance(values, np.ndarray):
        arr = values
    else:
        values = list(values)
        if values and all(type(v) is float or isinstance(v, np.floating) for v in values):
            arr = np.asarray(values, dtype=np.float64)
        elif values and all(type(v) is int and -(1 << 63) <= v < (1 << 63) for v in values):
            arr = np.asarray(values, dtype=np.int64)
        else:
            return None
    if arr.ndim != 1:
        raise ValueError("encode_array: need a 1-D sequence")
    if arr.dtype in (np.float64, np.float32, np.float16) and max_int >= (1 << 53):
        x = arr.astype(np.float64)
        if not np.isfinite(x).all():
            raise ValueError("encode_array: non-finite input")
        x = np.where(np.abs(x) < 1e-200, 0.0, x)
        mant, ex = np.frexp(x)
        expo = np.where(x == 0.0, 0, 53 - ex).astype(np.int64)
        mag = np.abs(np.round(np.ldexp(mant, 53))).astype(np.uint64)
        neg = (x < 0) & (mag > 0)
    elif arr.dtype in (np.int64, np.int32, np.int16) and max_int >= (1 << 63):
        x = arr.astype(np.int64)
        expo = np.zeros(arr.shape[0], dtype=np.int64)
        neg = x < 0
        mag = np.where(neg, -(x + 1), x).astype(np.uint64) + neg.astype(np.uint64)
    else:
        return None
    limbs = np.empty((arr.shape[0], 2), dtype=np.uint32)
    limbs[:, 0] = (mag & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    limbs[:, 1] = (mag >> np.uint64(32)).astype(np.uint32)
    return limbs, neg, expo


def encode_array(values, n, max_int, words, compact=False):
    """Vectorised FixedPointNumber.encode over a 1-D array: returns (limbs [N, words] uint32, exponents [N] int64).

    float arrays: exponent = 53 - frexp(x)[1], mantissa = round(x * 2^exponent) (|mantissa| < 2^53, exact);
    int16/32/64 arrays: exponent 0.  Anything else (object arrays, Python big ints, mixed lists) goes through the scalar
    codec element by element so the results are identical by construction.
    compact=True allows a [N, 2] result when every encoding is a non-negative 64-bit value (upper words all zero).
    """
    if isinstance(values, np.ndarray):
        arr = values
    else:
        # a Python sequence: the element TYPES decide the exponents (int -> 0, float -> 53 - frexp), so only a
        # homogeneous sequence may take a vector path; np.asarray would silently turn a mixed list into floats
        values = list(values)
        if values and all(type(v) is float or isinstance(v, np.floating) for v in values):
            arr = np.asarray(values, dtype=np.float64)
        elif values and all(type(v) is int and -(1 << 63) <= v < (1 << 63) for v in values):
            arr = np.asarray(values, dtype=np.int64)
        else:
            arr = np.empty(len(values), dtype=object)
            arr[:] = values
    if arr.ndim != 1:
        raise ValueError("encode_array: need a 1-D sequence")
    count = arr.shape[0]
    n_l = _n_limbs(n, words)
    if arr.dtype in (np.float64, np.float32, np.float16) and max_int >= (1 << 53):
        x = arr.astype(np.float64)
        if not np.isfinite(x).all():
            raise ValueError("encode_array: non-finite input")
        x = np.where(np.abs(x) < 1e-200, 0.0, x)
        mant, ex = np.frexp(x)
        expo = np.where(x == 0.0, 0, 53 - ex).astype(np.int64)   # a flushed / zero float encodes as the int 0
        mag = np.abs(np.round(np.ldexp(mant, 53))).astype(np.uint64)     # round(x * 2^expo), exact
        neg = (x < 0) & (mag > 0)
    elif arr.dtype in (np.int64, np.int32, np.int16) and max_int >= (1 << 63):
        x = arr.astype(np.int64)
        expo = np.zeros(count, dtype=np.int64)
        neg = x < 0
        mag = np.where(neg, -(x + 1), x).astype(np.uint64) + neg.astype(np.uint64)   # |x| without int64 overflow
    else:
        limbs = np.zeros((count, words), dtype=np.uint32)
        expo = np.zeros(count, dtype=np.int64)
        for i, v in enumerate(arr):
            f = FixedPointNumber.encode(v, n, max_int)
            limbs[i] = _n_limbs(f.encoding, words)
            expo[i] = f.exponent
        return limbs, expo
    lo = (mag & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (mag >> np.uint64(32)).astype(np.uint32)
    if compact and words > 2 and not neg.any():
        # all magnitudes fit two words: hand back [count, 2] (the C ABI takes short plaintext rows: 32x less to copy)
        limbs = np.empty((count, 2), dtype=np.uint32)
        limbs[:, 0] = lo
        limbs[:, 1] = hi
        return limbs, expo
    limbs = np.zeros((count, words), dtype=np.uint32)
    limbs[:, 0] = lo
    if words > 1:
        limbs[:, 1] = hi
    if neg.any():
        idx = np.nonzero(neg)[0]
        limbs[idx] = _sub_from_modulus(n_l, lo[idx], hi[idx], idx.shape[0])
    return limbs, expo


def decode_array(limbs, exponents, n, max_int):
    """Vectorised FixedPointNumber(...).decode() over packed plaintexts: returns a list of Python ints / floats with
    exactly the values and types the scalar codec gives (int when exponent == 0, else float)."""
    limbs = np.ascontiguousarray(limbs, dtype=np.uint32)
    count, words = limbs.shape
    expo = np.asarray(exponents, dtype=np.int64)
    n_l = _n_limbs(n, words)
    # small positive: everything above word 1 is zero and the value is < 2^63
    hi_zero = ~limbs[:, 2:].any(axis=1) if words > 2 else np.ones(count, dtype=bool)
    pos = hi_zero & (limbs[:, 1] < 0x80000000 if words > 1 else True)
    # small negative: n - value < 2^63.  d = n - value is only needed for the rows that are not small positives
    # (none at all for the usual all-positive batch: the 64-word borrow loop over every row was 60 % of decrypt's
    # host time at 100 000 elements)
    mant = np.zeros(count, dtype=np.int64)
    negs = np.zeros(count, dtype=bool)
    cand = np.nonzero(~pos)[0]
    if cand.size:
        sub = limbs[cand]
        d = np.empty_like(sub)
        borrow = np.zeros(cand.size, dtype=np.int64)
        for j in range(words):
            v = np.int64(int(n_l[j])) - sub[:, j].astype(np.int64) - borrow
            borrow = (v < 0).astype(np.int64)
            d[:, j] = (v & 0xFFFFFFFF).astype(np.uint32)
        d_hi_zero = ~d[:, 2:].any(axis=1) if words > 2 else np.ones(cand.size, dtype=bool)
        ok = (borrow == 0) & d_hi_zero & (d[:, 1] < 0x80000000 if words > 1 else True)
        negs[cand[ok]] = True
        if words > 1:
            mant[cand[ok]] = -(d[ok, 0].astype(np.int64) | (d[ok, 1].astype(np.int64) << 32))
        else:
            mant[cand[ok]] = -d[ok, 0].astype(np.int64)
    if words > 1:
        mant[pos] = (limbs[pos, 0].astype(np.int64) | (limbs[pos, 1].astype(np.int64) << 32))
    else:
        mant[pos] = limbs[pos, 0].astype(np.int64)
    fast = (pos | negs) & (np.abs(mant) <= min(max_int, (1 << 63) - 1))
    out = np.empty(count, dtype=object)      # object-array assignment stores Python ints / floats, not numpy scalars
    done = np.zeros(count, dtype=bool)
    as_int = fast & (expo == 0)
    if as_int.any():
        out[as_int] = mant[as_int]
        done |= as_int
    as_float = fast & (expo > 0) & (expo < 1100)
    if as_float.any():
        out[as_float] = mant[as_float].astype(np.float64) * np.ldexp(1.0, (-expo[as_float]).astype(np.int64))
        done |= as_float
    rest = np.nonzero(~done)[0]
    if rest.size:
        raw = limbs.tobytes()
        for i in rest:
            enc = int.from_bytes(raw[i * 4 * words:(i + 1) * 4 * words], "little")
            out[i] = FixedPointNumber(enc, int(expo[i]), n, max_int).decode()
    return out.tolist()


def classify_plain(limbs, n):
    """What phe_decrypt_mantissas computes on the device, in numpy (the CPU tests' stand-in for it): per row the signed
    mantissa when the plaintext m or n - m fits 63 bits, and the class 0 (positive) / 1 (negative) / 2 (neither)."""
    limbs = np.ascontiguousarray(limbs, dtype=np.uint32)
    count, words = limbs.shape
    n_l = _n_limbs(n, words)
    mant = np.zeros(count, dtype=np.int64)
    cls = np.full(count, 2, dtype=np.uint8)
    hi_zero = ~limbs[:, 2:].any(axis=1) if words > 2 else np.ones(count, dtype=bool)
    w1 = limbs[:, 1] if words > 1 else np.zeros(count, dtype=np.uint32)
    pos = hi_zero & (w1 < 0x80000000)
    mant[pos] = limbs[pos, 0].astype(np.int64) | (w1[pos].astype(np.int64) << 32)
    cls[pos] = 0
    d = np.empty_like(limbs)
    borrow = np.zeros(count, dtype=np.int64)
    for j in range(words):
        v = np.int64(int(n_l[j])) - limbs[:, j].astype(np.int64) - borrow
        borrow = (v < 0).astype(np.int64)
        d[:, j] = (v & 0xFFFFFFFF).astype(np.uint32)
    d1 = d[:, 1] if words > 1 else np.zeros(count, dtype=np.uint32)
    neg = ~pos & (borrow == 0) & (~d[:, 2:].any(axis=1) if words > 2 else True) & (d1 < 0x80000000)
    mant[neg] = -(d[neg, 0].astype(np.int64) | (d1[neg].astype(np.int64) << 32))
    cls[neg] = 1
    return mant, cls


def decode_mantissas(mant, cls, rows, exponents, n, max_int):
    """decode_array for plaintexts that were classified on the device (phe_decrypt_mantissas): mant / cls as above, rows
    the plaintext words (only read where cls == 2).  Same values and Python types as FixedPointNumber(...).decode()."""
    mant = np.asarray(mant, dtype=np.int64)
    cls = np.asarray(cls, dtype=np.uint8)
    expo = np.asarray(exponents, dtype=np.int64)
    count = mant.shape[0]
    fast = (cls < 2) & (np.abs(mant) <= min(max_int, (1 << 63) - 1))
    as_float = fast & (expo > 0) & (expo < 1100)
    if as_float.all():      # the usual batch: one multiplication, Python floats straight from tolist()
        return (mant.astype(np.float64) * np.ldexp(1.0, -expo)).tolist()
    as_int = fast & (expo == 0)
    if as_int.all():
        return mant.tolist()
    out = np.empty(count, dtype=object)      # object-array assignment stores Python ints / floats, not numpy scalars
    if as_int.any():
        out[as_int] = mant[as_int]
    if as_float.any():
        out[as_float] = mant[as_float].astype(np.float64) * np.ldexp(1.0, (-expo[as_float]).astype(np.int64))
    rest = np.nonzero(~(as_int | as_float))[0]
    if rest.size:
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        words = rows.shape[1]
        for i in rest:
            if cls[i] == 2:
                enc = int.from_bytes(rows[i].tobytes(), "little")
            else:
                enc = int(mant[i]) if cls[i] == 0 else n + int(mant[i])
            out[i] = FixedPointNumber(enc, int(expo[i]), n, max_int).decode()
    return out.tolist()


class FixedPointEndec(object):
    """Array/scalar encoder-decoder with a fixed precision (the reference's class of the same name,
    fixedpoint.py:304-367, minus its dependency on the FATE session tables)."""

    def __init__(self, n=None, max_int=None, precision=None, *args, **kwargs):
        if n is None:
            self.n, self.max_int = FixedPointNumber.Q, FixedPointNumber.Q // 2
        else:
            self.n = n
            self.max_int = n // 2 if max_int is None else max_int
        self.precision = precision

    def _encode(self, scalar):
        return FixedPointNumber.encode(scalar, n=self.n, max_int=self.max_int, precision=self.precision)

    @staticmethod
    def _decode(number):
        return number.decode()

    def _truncate(self, number):
        return FixedPointNumber.encode(number.decode(), n=self.n, max_int=self.max_int)

    @staticmethod
    def _map(tensor, op):
        if isinstance(tensor, np.ndarray):
            flat = [op(v) for v in tensor.flat]
            out = np.empty(len(flat), dtype=object)
            out[:] = flat
            return out.reshape(tensor.shape)
        if isinstance(tensor, (list, tuple)):
            return [op(v) for v in tensor]
        return op(tensor)

    def encode(self, float_tensor):
        return self._map(float_tensor, self._encode)

    def decode(self, integer_tensor):
        return self._map(integer_tensor, self._decode)

    def truncate(self, integer_tensor, *args, **kwargs):
        return self._map(integer_tensor, self._truncate)

"""ctypes view of the C ABI (include/phe_b200.h).  Used by the tests, bench.py and the numpy fast path.

The library has no CPU fallback: compute calls raise RuntimeError when no CUDA device is present.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHE_B200_LIB") or os.path.join(_HERE, "lib", "libphe_b200.so")   # override: A/B builds of the kernels

_u32p = ctypes.POINTER(ctypes.c_uint32)
_lib = None

# every symbol include/phe_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "phe_last_error", "phe_version", "phe_device_count", "phe_set_device", "phe_get_device", "phe_kernel_launches",
    "phe_pubkey_create", "phe_pubkey_destroy", "phe_pubkey_bits", "phe_pubkey_n_words", "phe_pubkey_is_djn",
    "phe_pubkey_randbits", "phe_pubkey_get_n", "phe_pubkey_get_nsquare", "phe_pubkey_get_hs",
    "phe_privkey_create", "phe_privkey_destroy", "phe_privkey_get_p", "phe_privkey_get_q", "phe_keygen",
    "phe_encrypt", "phe_obfuscate", "phe_decrypt", "phe_add", "phe_mul", "phe_modexp",
    "phe_encrypt_dev", "phe_decrypt_dev", "phe_add_dev", "phe_mul_dev",
    "phe_pubkey_set_comb_bits", "phe_pubkey_comb_bits", "phe_pubkey_comb_info", "phe_host_mont_block", "phe_host_modexp", "phe_host_shape_for_bits", "phe_host_powm_program", "phe_privkey_pair_block", "phe_privkey_pair_segments", "phe_pubkey_npair_block", "phe_dev_alloc", "phe_dev_free", "phe_copy", "phe_encrypt_dev_multi", "phe_enable_peer_access", "phe_ipc_export", "phe_ipc_open", "phe_ipc_close", "phe_chacha20_keystream", "phe_invert", "phe_encrypt_compact",
    "phe_decrypt_mantissas", "phe_gather_rows_dev", "phe_scatter_rows_dev", "phe_scale_rows_dev", "phe_invert_rows_dev", "phe_segsum_dev",
    "phe_timing_enable", "phe_timing_read", "phe_timing_kind_name", "phe_int_pipe_peak", "phe_fp64_pipe_peak",
    "phe_product_mix_peak",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libphe_b200.so is not built (run `python -m pailliercryptolib_python_b200.build`); "
                "there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.phe_last_error.restype = ctypes.c_char_p
        L.phe_version.restype = ctypes.c_char_p
        L.phe_kernel_launches.restype = ctypes.c_ulonglong
        for name in ("phe_encrypt", "phe_obfuscate", "phe_decrypt", "phe_add", "phe_mul", "phe_modexp",
                     "phe_encrypt_dev", "phe_decrypt_dev", "phe_add_dev", "phe_mul_dev", "phe_pubkey_create",
                     "phe_privkey_create", "phe_keygen"):
            getattr(L, name).restype = ctypes.c_int
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s: %s" % (what, lib().phe_last_error().decode()))


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):  # raw device pointer
        return ctypes.cast(ctypes.c_void_p(a), _u32p)
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"], "need a C-contiguous uint32 array"
    return a.ctypes.data_as(_u32p)


def int_to_words(v, nwords):
    return np.frombuffer(int(v).to_bytes(4 * nwords, "little"), dtype="<u4").copy()


def words_to_int(w):
    return int.from_bytes(np.ascontiguousarray(w, dtype="<u4").tobytes(), "little")


def ints_to_array(vals, nwords):
    buf = b"".join(int(v).to_bytes(4 * nwords, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), nwords).copy()


def array_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype="<u4")
    n = arr.shape[1] * 4
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(arr.shape[0])]


def device_count():
    return lib().phe_device_count()


def kernel_launches():
    return int(lib().phe_kernel_launches())


KERNEL_KINDS = ["k_modmul", "k_powm", "k_dec_prep", "k_dec_tail", "k_encrypt_comb", "k_encrypt_finish", "k_comb_build",
                "k_dec_pair", "k_dec_crt", "k_encrypt_npair", "k_mul_npair", "k_rows_move"]


def timing_enable(on=True):
    """Bracket every kernel launch with a cudaEvent pair on its stream (bench.py's roofline timing)."""
    _check(lib().phe_timing_enable(int(bool(on))), "phe_timing_enable")


def timing_read():
    """{kernel kind: (total device ms, launches)} since timing_enable()."""
    out = {}
    for k, name in enumerate(KERNEL_KINDS):
        ms = ctypes.c_double()
        n = ctypes.c_ulonglong()
        _check(lib().phe_timing_read(k, ctypes.byref(ms), ctypes.byref(n)), "phe_timing_read")
        out[name] = (ms.value, int(n.value))
    return out


def int_pipe_peak(reps=5):
    """Measured IMAD.WIDE.U32 issue rate of the current device in MAC/s (the roofline denominator)."""
    v = ctypes.c_double()
    _check(lib().phe_int_pipe_peak(int(reps), ctypes.byref(v)), "phe_int_pipe_peak")
    return v.value


def fp64_pipe_peak(reps=5):
    """Measured DFMA.RZ issue rate of the current device in lane operations/s."""
    v = ctypes.c_double()
    _check(lib().phe_fp64_pipe_peak(int(reps), ctypes.byref(v)), "phe_fp64_pipe_peak")
    return v.value


def product_mix_peak(reps=5):
    """Measured rate of the bare 52x52-bit limb-product instruction mix in products/s."""
    v = ctypes.c_double()
    _check(lib().phe_product_mix_peak(int(reps), ctypes.byref(v)), "phe_product_mix_peak")
    return v.value


class PubKey:
    """phe_pubkey handle.  n: Python int; hs: optional Python int (DJN generator)."""

    def __init__(self, n, bits, djn=True, hs=None, randbits=0):
        self.n_words = (bits + 31) // 32
        self.bits = bits
        self.n = int(n)
        h = ctypes.c_void_p()
        hs_arr = int_to_words(hs, 2 * self.n_words) if hs is not None else None
        _check(lib().phe_pubkey_create(_p(int_to_words(n, self.n_words)), self.n_words, bits, int(bool(djn)),
                                       _p(hs_arr), int(randbits), ctypes.byref(h)), "phe_pubkey_create")
        self.h = h
        self.djn = bool(djn)
        self.randbits = lib().phe_pubkey_randbits(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().phe_pubkey_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_comb_bits(self, bits):
        """Digit width of the DJN comb table (0 = automatic); see include/phe_b200.h."""
        _check(lib().phe_pubkey_set_comb_bits(self.h, int(bits)), "phe_pubkey_set_comb_bits")

    @property
    def comb_bits(self):
        return lib().phe_pubkey_comb_bits(self.h)

    @property
    def comb_info(self):
        """(bytes of device memory of the comb table in use, wall ms of its build); (0, 0.0) before one exists."""
        b, ms = ctypes.c_ulonglong(), ctypes.c_double()
        _check(lib().phe_pubkey_comb_info(self.h, ctypes.byref(b), ctypes.byref(ms)), "phe_pubkey_comb_info")
        return int(b.value), float(ms.value)

    @property
    def hs(self):
        out = np.zeros(2 * self.n_words, dtype=np.uint32)
        _check(lib().phe_pubkey_get_hs(self.h, _p(out)), "phe_pubkey_get_hs")
        return words_to_int(out)

    @property
    def nsquare(self):
        out = np.zeros(2 * self.n_words, dtype=np.uint32)
        _check(lib().phe_pubkey_get_nsquare(self.h, _p(out)), "phe_pubkey_get_nsquare")
        return words_to_int(out)

    # ---- host-buffer ops on packed arrays -------------------------------------------------------------
    def encrypt(self, m, r=None, make_secure=True, out=None):
        m = np.ascontiguousarray(m, dtype=np.uint32).reshape(-1, self.n_words)
        if out is None:
            out = np.empty((m.shape[0], 2 * self.n_words), dtype=np.uint32)
        assert out.shape == (m.shape[0], 2 * self.n_words)
        rw = 0
        if r is not None:
            r = np.ascontiguousarray(r, dtype=np.uint32)
            r = r.reshape(m.shape[0], -1)
            rw = r.shape[1]
        _check(lib().phe_encrypt(self.h, _p(m), ctypes.c_size_t(m.shape[0]), _p(r), rw, int(bool(make_secure)),
                                 _p(out)), "phe_encrypt")
        return out

    def obfuscate(self, ct, r=None):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, 2 * self.n_words).copy()
        rw = 0
        if r is not None:
            r = np.ascontiguousarray(r, dtype=np.uint32).reshape(ct.shape[0], -1)
            rw = r.shape[1]
        _check(lib().phe_obfuscate(self.h, _p(ct), ctypes.c_size_t(ct.shape[0]), _p(r), rw), "phe_obfuscate")
        return ct

    def add(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 2 * self.n_words)
        b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, 2 * self.n_words)
        out = np.empty_like(a)
        _check(lib().phe_add(self.h, _p(a), ctypes.c_size_t(a.shape[0]), _p(b), ctypes.c_size_t(b.shape[0]), _p(out)),
               "phe_add")
        return out

    def invert(self, ct):
        """Row-wise modular inverse modulo n^2 (batched on the device)."""
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, 2 * self.n_words)
        out = np.empty_like(ct)
        _check(lib().phe_invert(self.h, _p(ct), ctypes.c_size_t(ct.shape[0]), _p(out)), "phe_invert")
        return out

    def mul(self, ct, e):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, 2 * self.n_words)
        e = np.ascontiguousarray(e, dtype=np.uint32)
        if e.ndim == 1:
            e = e.reshape(1, -1)
        out = np.empty_like(ct)
        _check(lib().phe_mul(self.h, _p(ct), ctypes.c_size_t(ct.shape[0]), _p(e), e.shape[1],
                             ctypes.c_size_t(e.shape[0]), _p(out)), "phe_mul")
        return out

    # ---- device-pointer ops (ints are raw CUDA pointers, e.g. torch.Tensor.data_ptr()) ------------------
    def encrypt_dev(self, d_m, count, d_r, r_words, d_out, stream=0, make_secure=True):
        """make_secure with d_r None / 0: the library draws the obfuscator exponents itself (as the host call does)."""
        _check(lib().phe_encrypt_dev(self.h, _p(d_m), ctypes.c_size_t(count), _p(d_r) if d_r else None, r_words,
                                     int(bool(make_secure)), _p(d_out), ctypes.c_void_p(stream)), "phe_encrypt_dev")

    def encrypt_dev_multi(self, d_m, count, d_r, r_words, d_out, peer_outs, stream=0, make_secure=True):
        """DJN encrypt whose rows also go to the buffers in peer_outs (device pointers, e.g. peer-mapped gather buffers)."""
        arr = (ctypes.c_void_p * max(1, len(peer_outs)))(*[ctypes.c_void_p(int(q)) for q in peer_outs])
        _check(lib().phe_encrypt_dev_multi(self.h, _p(d_m), ctypes.c_size_t(count), _p(d_r) if d_r else None, r_words,
                                           int(bool(make_secure)), _p(d_out), ctypes.cast(arr, ctypes.c_void_p),
                                           len(peer_outs), ctypes.c_void_p(stream)),
               "phe_encrypt_dev_multi")

    # ---- row operations on device matrices (index / delta lists are host arrays) ----------------------------------
    @staticmethod
    def _i64(a):
        a = np.ascontiguousarray(a, dtype=np.int64)
        return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))

    def gather_rows_dev(self, d_src, src_rows, idx, d_dst, stream=0):
        idx, ip = self._i64(idx)
        _check(lib().phe_gather_rows_dev(self.h, _p(d_src), ctypes.c_size_t(src_rows), ip, ctypes.c_size_t(idx.size),
                                         _p(d_dst), ctypes.c_void_p(stream)), "phe_gather_rows_dev")

    def scatter_rows_dev(self, d_src, idx, d_dst, dst_rows, stream=0):
        idx, ip = self._i64(idx)
        _check(lib().phe_scatter_rows_dev(self.h, _p(d_src), ip, ctypes.c_size_t(idx.size), _p(d_dst),
                                          ctypes.c_size_t(dst_rows), ctypes.c_void_p(stream)), "phe_scatter_rows_dev")

    def scale_rows_dev(self, d_ct, rows, idx, delta, stream=0):
        idx, ip = self._i64(idx)
        delta = np.ascontiguousarray(delta, dtype=np.int32)
        _check(lib().phe_scale_rows_dev(self.h, _p(d_ct), ctypes.c_size_t(rows), ip,
                                        delta.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_size_t(idx.size),
                                        ctypes.c_void_p(stream)), "phe_scale_rows_dev")

    def invert_rows_dev(self, d_ct, rows, idx, stream=0):
        idx, ip = self._i64(idx)
        _check(lib().phe_invert_rows_dev(self.h, _p(d_ct), ctypes.c_size_t(rows), ip, ctypes.c_size_t(idx.size),
                                         ctypes.c_void_p(stream)), "phe_invert_rows_dev")

    def segsum_dev(self, d_ct, groups, width, d_out, stream=0):
        _check(lib().phe_segsum_dev(self.h, _p(d_ct), ctypes.c_size_t(groups), ctypes.c_size_t(width), _p(d_out),
                                    ctypes.c_void_p(stream)), "phe_segsum_dev")

    def add_dev(self, d_a, na, d_b, nb, d_out, stream=0):
        _check(lib().phe_add_dev(self.h, _p(d_a), ctypes.c_size_t(na), _p(d_b), ctypes.c_size_t(nb), _p(d_out),
                                 ctypes.c_void_p(stream)), "phe_add_dev")

    def mul_dev(self, d_ct, n, d_e, e_words, ne, exp_bits, d_out, stream=0):
        _check(lib().phe_mul_dev(self.h, _p(d_ct), ctypes.c_size_t(n), _p(d_e), e_words, ctypes.c_size_t(ne),
                                 exp_bits, _p(d_out), ctypes.c_void_p(stream)), "phe_mul_dev")


class PrivKey:
    def __init__(self, pk, p, q):
        self.pk = pk
        pw, qw = (max(1, (int(v).bit_length() + 31) // 32) for v in (p, q))   # unbalanced primes are allowed
        h = ctypes.c_void_p()
        _check(lib().phe_privkey_create(pk.h, _p(int_to_words(p, pw)), pw, _p(int_to_words(q, qw)), qw,
                                        ctypes.byref(h)), "phe_privkey_create")
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().phe_privkey_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def decrypt(self, ct, out=None):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, 2 * self.pk.n_words)
        if out is None:
            out = np.empty((ct.shape[0], self.pk.n_words), dtype=np.uint32)
        assert out.shape == (ct.shape[0], self.pk.n_words)
        _check(lib().phe_decrypt(self.h, _p(ct), ctypes.c_size_t(ct.shape[0]), _p(out)), "phe_decrypt")
        return out

    def decrypt_mantissas(self, ct):
        """(mant int64 [N], cls uint8 [N], rows uint32 [N, n_words]): see phe_decrypt_mantissas; rows is only filled where
        cls == 2."""
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, 2 * self.pk.n_words)
        mant = np.empty(ct.shape[0], dtype=np.int64)
        cls = np.empty(ct.shape[0], dtype=np.uint8)
        rows = np.zeros((ct.shape[0], self.pk.n_words), dtype=np.uint32)
        _check(lib().phe_decrypt_mantissas(self.h, _p(ct), ctypes.c_size_t(ct.shape[0]),
                                           mant.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                           cls.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), _p(rows)), "phe_decrypt_mantissas")
        return mant, cls, rows

    def decrypt_dev(self, d_ct, count, d_out, stream=0):
        _check(lib().phe_decrypt_dev(self.h, _p(d_ct), ctypes.c_size_t(count), _p(d_out), ctypes.c_void_p(stream)),
               "phe_decrypt_dev")


def modexp(base, exp, modulus, words):
    """Generic element-wise modexp (ipcl::modExp): lists of ints -> list of ints."""
    b = ints_to_array(base, words)
    e = ints_to_array(exp, words)
    out = np.empty_like(b)
    _check(lib().phe_modexp(_p(b), _p(e), _p(int_to_words(modulus, words)), words, ctypes.c_size_t(len(base)),
                            _p(out)), "phe_modexp")
    return array_to_ints(out)


def keygen(bits):
    n = np.zeros((bits + 31) // 32, dtype=np.uint32)
    p = np.zeros((bits // 2 + 31) // 32, dtype=np.uint32)
    q = np.zeros((bits // 2 + 31) // 32, dtype=np.uint32)
    _check(lib().phe_keygen(bits, _p(n), _p(p), _p(q)), "phe_keygen")
    return words_to_int(n), words_to_int(p), words_to_int(q)


def host_mont_block(modulus, mod_words, L, TPI):
    kp = lib().phe_host_mont_block(_p(int_to_words(modulus, mod_words)), mod_words, L, TPI, None, None)
    if kp <= 0:
        raise RuntimeError(lib().phe_last_error().decode())
    out = np.zeros(5 * kp, dtype=np.float64)
    n0 = ctypes.c_uint64()
    lib().phe_host_mont_block(_p(int_to_words(modulus, mod_words)), mod_words, L, TPI,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(n0))
    return out.reshape(5, kp), n0.value


class DeviceBuffer:
    """A cudaMalloc block on the key's device (phe_dev_alloc) that can be shared with the other ranks of the node through
    CUDA IPC; exposes __cuda_array_interface__ so that torch.as_tensor(buf) views it as a [rows, words] uint32->int32 matrix."""

    def __init__(self, pk, rows, words):
        self.rows, self.words = int(rows), int(words)
        p = _u32p()
        _check(lib().phe_dev_alloc(pk.h, ctypes.c_size_t(self.rows * self.words), ctypes.byref(p)), "phe_dev_alloc")
        self.ptr = ctypes.cast(p, ctypes.c_void_p).value
        self.__cuda_array_interface__ = {"shape": (self.rows, self.words), "typestr": "<i4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def ipc_handle(self):
        h = (ctypes.c_ubyte * 64)()
        _check(lib().phe_ipc_export(_p(self.ptr), h), "phe_ipc_export")
        return bytes(h)

    def free(self):
        if self.ptr:
            lib().phe_dev_free(_p(self.ptr))
            self.ptr = 0


def ipc_open(handle):
    """Map another process's DeviceBuffer (its ipc_handle()) on the current device; returns the device pointer."""
    h = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    p = _u32p()
    _check(lib().phe_ipc_open(h, ctypes.byref(p)), "phe_ipc_open")
    return ctypes.cast(p, ctypes.c_void_p).value


def npair_block(pk):
    """Constant block of the n-adic pair engine for this key: dict(L, TPI, cst [9][KP], n0inv, d_top), or None if the key
    does not use the engine (include/phe_b200.h: phe_pubkey_npair_block)."""
    L, TPI = ctypes.c_int(), ctypes.c_int()
    kp = lib().phe_pubkey_npair_block(pk.h, ctypes.byref(L), ctypes.byref(TPI), None, None, None)
    if kp < 0:
        raise RuntimeError(lib().phe_last_error().decode())
    if kp == 0:
        return None
    out = np.zeros(9 * kp, dtype=np.float64)
    n0, dt = ctypes.c_uint64(), ctypes.c_uint64()
    lib().phe_pubkey_npair_block(pk.h, None, None, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(n0),
                                 ctypes.byref(dt))
    return {"L": L.value, "TPI": TPI.value, "cst": out.reshape(9, kp), "n0inv": n0.value, "d_top": dt.value}


def pair_block(sk, y):
    """What the p-adic pair engine is given for x = p (y = 0) or q (y = 1): dict(L, n0inv, mod, cst, prog), or None if
    the key does not use the engine (include/phe_b200.h: phe_privkey_pair_block)."""
    n = lib().phe_privkey_pair_block(sk.h, int(y), None, None, None, None, None, 0)
    if n < 0:
        raise RuntimeError(lib().phe_last_error().decode())
    if n == 0:
        return None
    L, n0 = ctypes.c_int(), ctypes.c_uint64()
    lib().phe_privkey_pair_block(sk.h, int(y), ctypes.byref(L), ctypes.byref(n0), None, None, None, 0)
    mod = np.zeros(L.value, dtype=np.float64)
    cst = np.zeros(6 * 2 * L.value, dtype=np.float64)
    prog = np.zeros(n, dtype=np.uint32)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().phe_privkey_pair_block(sk.h, int(y), None, None, mod.ctypes.data_as(dp), cst.ctypes.data_as(dp), _p(prog), n)
    return {"L": L.value, "n0inv": n0.value, "mod": mod, "cst": cst, "prog": [int(v) for v in prog]}


def pair_segments(sk, y):
    """The pair-engine program of x = p / q cut into the time slices k_dec_pair runs: (prog, offsets), or None
    (include/phe_b200.h: phe_privkey_pair_segments)."""
    n = ctypes.c_int()
    nseg = lib().phe_privkey_pair_segments(sk.h, int(y), None, 0, None, 0, ctypes.byref(n))
    if nseg < 0:
        raise RuntimeError(lib().phe_last_error().decode())
    if nseg == 0:
        return None
    prog = np.zeros(n.value, dtype=np.uint32)
    off = np.zeros(nseg, dtype=np.int32)
    lib().phe_privkey_pair_segments(sk.h, int(y), _p(prog), n.value, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), nseg, None)
    return prog, [int(v) for v in off]


def host_powm_program(exponent, e_words):
    """Sliding-window program the library builds for a shared exponent (list of ints)."""
    e = int_to_words(exponent, e_words)
    n = lib().phe_host_powm_program(_p(e), e_words, None, 0)
    if n <= 0:
        raise RuntimeError(lib().phe_last_error().decode())
    out = np.zeros(n, dtype=np.uint32)
    lib().phe_host_powm_program(_p(e), e_words, _p(out), n)
    return [int(v) for v in out]


def host_modexp(base, exp, modulus, words):
    out = np.zeros(words, dtype=np.uint32)
    _check(lib().phe_host_modexp(_p(int_to_words(base, words)), _p(int_to_words(exp, words)),
                                 _p(int_to_words(modulus, words)), words, _p(out)), "phe_host_modexp")
    return words_to_int(out)

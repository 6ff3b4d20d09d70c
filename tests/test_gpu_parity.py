"""GPU parity tests: the CUDA path, called through the C ABI (libphe_b200.so), against the Python-int oracle
on the same seeded inputs.  Bit-exact (integer work).  Run with `pytest -m gpu` on a B200."""
import random

import numpy as np
import pytest

import paillier_oracle as O
from pailliercryptolib_python_b200 import capi

pytestmark = pytest.mark.gpu

SEED = 20240611


@pytest.fixture(scope="module")
def key2048():
    pk_o, sk_o = O.bench_keypair()
    pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
    return pk_o, sk_o, pk, sk


def _rand_cts(pk_o, rng, count):
    # valid ciphertexts without paying oracle modexps: (1 + m n) * s^n is overkill; any unit mod n^2 decrypts to something.
    return [rng.randrange(1, pk_o.nsquare) for _ in range(count)]


def test_requires_gpu():
    assert capi.device_count() >= 1


def test_encrypt_djn_pinned_r(key2048):
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED)
    ms = [0, 1, pk_o.n - 1, pk_o.n // 3 - 1] + [rng.randrange(pk_o.n) for _ in range(20)] + [rng.getrandbits(53) for _ in range(40)]
    rs = [0, 1, (1 << 1024) - 1] + [rng.getrandbits(1024) for _ in range(len(ms) - 3)]
    got = capi.array_to_ints(pk.encrypt(capi.ints_to_array(ms, 64), capi.ints_to_array(rs, 32)))
    assert got == O.encrypt_batch(pk_o, ms, rs)
    # make_secure = False
    got = capi.array_to_ints(pk.encrypt(capi.ints_to_array(ms, 64), None, make_secure=False))
    assert got == O.encrypt_batch(pk_o, ms, None)


@pytest.mark.parametrize("comb_bits", [1, 7, 8, 13, 16, 19])
def test_encrypt_comb_widths(comb_bits):
    """Every digit width of the fixed-base comb table gives the same ciphertexts (table rebuilt on the device)."""
    pk_o, sk_o = O.seeded_keypair(1024, 21)
    pk = capi.PubKey(pk_o.n, 1024, djn=True, hs=pk_o.hs)
    pk.set_comb_bits(comb_bits)
    rng = random.Random(SEED + comb_bits)
    ms = [0, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(30)]
    rs = [0, (1 << 512) - 1, 1, 1 << 511] + [rng.getrandbits(512) for _ in range(len(ms) - 4)]
    got = capi.array_to_ints(pk.encrypt(capi.ints_to_array(ms, 32), capi.ints_to_array(rs, 16)))
    assert pk.comb_bits == comb_bits
    assert got == O.encrypt_batch(pk_o, ms, rs)


def test_comb_table_promotion():
    """Automatic width: a key starts on the small comb table and moves to the wide one (as wide as a quarter of the
    free device memory allows, at most 20 bits) once it has encrypted 32768 elements; ciphertexts do not change."""
    pk_o, sk_o = O.bench_keypair()
    pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    rng = random.Random(SEED + 77)
    ms = [rng.getrandbits(53) for _ in range(40)]
    rs = [rng.getrandbits(1024) for _ in ms]
    m, r = capi.ints_to_array(ms, 64), capi.ints_to_array(rs, 32)
    want = O.encrypt_batch(pk_o, ms, rs)
    assert capi.array_to_ints(pk.encrypt(m, r)) == want
    assert pk.comb_bits == 12
    big = 33000
    bm = np.zeros((big, 64), dtype=np.uint32)
    bm[:, 0] = np.arange(big, dtype=np.uint32)
    br = np.random.default_rng(SEED).integers(0, 1 << 32, size=(big, 32), dtype=np.uint32)
    ct = pk.encrypt(bm, br)
    wide = pk.comb_bits
    assert 12 < wide <= 20
    idx = [0, 1, big // 2, big - 1]
    got = capi.array_to_ints(ct[idx])
    assert got == O.encrypt_batch(pk_o, [int(bm[i, 0]) for i in idx], capi.array_to_ints(br[idx]))
    assert capi.array_to_ints(pk.encrypt(m, r)) == want
    assert pk.comb_bits == wide


def test_decrypt_crt(key2048):
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + 1)
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(29)]
    cs = O.encrypt_batch(pk_o, ms, [rng.getrandbits(1024) for _ in ms])
    cs += _rand_cts(pk_o, rng, 32)
    got = capi.array_to_ints(sk.decrypt(capi.ints_to_array(cs, 128)))
    assert got[:len(ms)] == ms
    assert got == O.decrypt_batch(sk_o, cs)


@pytest.mark.parametrize("nseg", ["1", "8"])
def test_decrypt_time_sliced_units(key2048, nseg, monkeypatch):
    """k_dec_pair continues a unit segment by segment on whichever warp is free (phe_kernels.cuh); PHE_DEC_SEGMENTS pins
    the number of segments so that a small batch takes the sliced path too.  Ragged batch (last warp unit partly filled),
    units and non-units, both settings bit-equal to the oracle."""
    pk_o, sk_o, pk, sk = key2048
    monkeypatch.setenv("PHE_DEC_SEGMENTS", nseg)
    rng = random.Random(SEED + 77)
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(60)]
    cs = O.encrypt_batch(pk_o, ms, [rng.getrandbits(1024) for _ in ms])
    cs += _rand_cts(pk_o, rng, 70) + [0, sk_o.p, pk_o.n, pk_o.nsquare - 1]
    got = capi.array_to_ints(sk.decrypt(capi.ints_to_array(cs, 128)))
    assert got[:len(ms)] == ms
    assert got == O.decrypt_batch(sk_o, cs)


@pytest.mark.parametrize("bits", [512, 1024, 3072])
def test_decrypt_time_sliced_other_key_sizes(bits, monkeypatch):
    """Forced slicing on keys where the uncuttable preamble (conversion + window table) spans several segment targets,
    so the programs of p and q may cut into different numbers of segments (k_dec_pair takes a count per modulus)."""
    monkeypatch.setenv("PHE_DEC_SEGMENTS", "16")
    pk_o, sk_o = O.seeded_keypair(bits, 321)
    pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
    sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
    assert capi.pair_block(sk, 0) is not None
    segs = [len(capi.pair_segments(sk, y)[1]) for y in (0, 1)]
    assert all(2 <= v <= 16 for v in segs)
    rng = random.Random(SEED + bits + 5)
    nw = -(-bits // 32)
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(70)]
    cs = O.encrypt_batch(pk_o, ms, [rng.getrandbits(bits // 2) for _ in ms]) + [0, sk_o.p, pk_o.n]
    got = capi.array_to_ints(sk.decrypt(capi.ints_to_array(cs, 2 * nw)))
    assert got[:len(ms)] == ms
    assert got == O.decrypt_batch(sk_o, cs)


def test_decrypt_more_units_than_warps(key2048):
    """40 000 ciphertexts = 2500 warp units on 1184 resident warps: the automatic choice is the time-sliced launch, pieces
    of one unit run on different SMs.  Checked by round trip over the whole batch and against the oracle on a sample."""
    pk_o, sk_o, pk, sk = key2048
    N = 40000
    rng = np.random.default_rng(SEED + 78)
    m = np.zeros((N, 64), dtype=np.uint32)
    m[:, :2] = rng.integers(0, 1 << 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
    r = rng.integers(0, 1 << 32, size=(N, 32), dtype=np.uint64).astype(np.uint32)
    ct = pk.encrypt(m, r)
    back = sk.decrypt(ct)
    assert np.array_equal(back, m)
    idx = [0, 1, 31, 32, 33, N // 2, N - 33, N - 1]
    assert capi.array_to_ints(back[idx]) == O.decrypt_batch(sk_o, capi.array_to_ints(ct[idx]))


def test_add_and_broadcast(key2048):
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + 2)
    a = [0, 1, pk_o.nsquare - 1] + _rand_cts(pk_o, rng, 200)
    b = [5, 1, pk_o.nsquare - 1] + _rand_cts(pk_o, rng, 200)
    got = capi.array_to_ints(pk.add(capi.ints_to_array(a, 128), capi.ints_to_array(b, 128)))
    assert got == O.add_batch(pk_o, a, b)
    got = capi.array_to_ints(pk.add(capi.ints_to_array(a, 128), capi.ints_to_array(b[3:4], 128)))
    assert got == O.add_batch(pk_o, a, b[3:4])
    with pytest.raises(RuntimeError):
        pk.add(capi.ints_to_array(a, 128), capi.ints_to_array(b[:7], 128))


@pytest.mark.parametrize("ebits", [1, 7, 53, 160, 700, 2048, 2100])
def test_mul_variable_exponent(key2048, ebits):
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + ebits)
    n = 40 if ebits <= 160 else 12
    a = [0, 1] + _rand_cts(pk_o, rng, n)
    e = [3, 0] + [rng.getrandbits(ebits) for _ in range(n)]
    e[2] |= 1 << (ebits - 1)
    ew = (ebits + 31) // 32
    got = capi.array_to_ints(pk.mul(capi.ints_to_array(a, 128), capi.ints_to_array(e, ew)))
    assert got == O.mul_batch(pk_o, a, e)
    got = capi.array_to_ints(pk.mul(capi.ints_to_array(a, 128), capi.ints_to_array(e[2:3], ew)))
    assert got == O.mul_batch(pk_o, a, e[2:3])


def test_roundtrip_homomorphism_large(key2048):
    """Size-independent property at a batch far beyond what the oracle checks directly:
    D(E(a) * E(b) ) = a + b and D(E(a)^k) = k a (mod n), internal CSPRNG r."""
    pk_o, sk_o, pk, sk = key2048
    rng = np.random.Generator(np.random.PCG64(SEED))
    N = 20000
    a = rng.integers(0, 2**62, size=N, dtype=np.uint64)
    b = rng.integers(0, 2**62, size=N, dtype=np.uint64)
    k = rng.integers(1, 2**20, size=N, dtype=np.uint64)

    def pack(v):
        out = np.zeros((N, 64), dtype=np.uint32)
        out[:, 0] = (v & 0xFFFFFFFF).astype(np.uint32)
        out[:, 1] = (v >> 32).astype(np.uint32)
        return out

    ca, cb = pk.encrypt(pack(a)), pk.encrypt(pack(b))
    assert not np.array_equal(ca, pk.encrypt(pack(a)))  # randomised
    s = sk.decrypt(pk.add(ca, cb))
    want = a.astype(object) + b.astype(object)
    got = s[:, 0].astype(object) + (s[:, 1].astype(object) << 32) + (s[:, 2].astype(object) << 64)
    assert (s[:, 3:] == 0).all() and (got == want).all()
    ke = np.zeros((N, 1), dtype=np.uint32)
    ke[:, 0] = k.astype(np.uint32)
    p = sk.decrypt(pk.mul(ca, ke))
    want = a.astype(object) * k.astype(object)
    got = p[:, 0].astype(object) + (p[:, 1].astype(object) << 32) + (p[:, 2].astype(object) << 64)
    assert (p[:, 3:] == 0).all() and (got == want).all()


def test_obfuscate_matches_oracle(key2048):
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + 5)
    cs = _rand_cts(pk_o, rng, 9)
    rs = [rng.getrandbits(1024) for _ in cs]
    got = capi.array_to_ints(pk.obfuscate(capi.ints_to_array(cs, 128), capi.ints_to_array(rs, 32)))
    assert got == [c * O.obfuscator(pk_o, r) % pk_o.nsquare for c, r in zip(cs, rs)]


def test_classic_scheme_1024():
    pk_o, sk_o = O.seeded_keypair(1024, 5, djn=False)
    pk = capi.PubKey(pk_o.n, 1024, djn=False)
    sk = capi.PrivKey(pk, sk_o.q, sk_o.p)  # swapped on purpose: the key orders p < q itself
    rng = random.Random(SEED + 6)
    ms = [0, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(14)]
    rs = [1, pk_o.n - 1] + [rng.randrange(1, pk_o.n) for _ in range(14)]
    ct = pk.encrypt(capi.ints_to_array(ms, 32), capi.ints_to_array(rs, 32))
    assert capi.array_to_ints(ct) == O.encrypt_batch(pk_o, ms, rs)
    assert capi.array_to_ints(sk.decrypt(ct)) == ms
    # internal r
    assert capi.array_to_ints(sk.decrypt(pk.encrypt(capi.ints_to_array(ms, 32)))) == ms


@pytest.mark.parametrize("bits", [1024, 3072])
def test_djn_other_key_sizes(bits):
    pk_o, sk_o = O.seeded_keypair(bits, 77)
    pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
    sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
    rng = random.Random(SEED + bits)
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(13)]
    rs = [rng.getrandbits(bits // 2) for _ in ms]
    ct = pk.encrypt(capi.ints_to_array(ms, bits // 32), capi.ints_to_array(rs, bits // 64))
    assert capi.array_to_ints(ct) == O.encrypt_batch(pk_o, ms, rs)
    assert capi.array_to_ints(sk.decrypt(ct)) == ms == O.decrypt_batch(sk_o, capi.array_to_ints(ct))
    a, b = capi.array_to_ints(ct[:8]), capi.array_to_ints(ct[8:])
    assert capi.array_to_ints(pk.add(ct[:8], ct[8:])) == O.add_batch(pk_o, a, b)
    e = [rng.getrandbits(53) for _ in range(8)]
    assert capi.array_to_ints(pk.mul(ct[:8], capi.ints_to_array(e, 2))) == O.mul_batch(pk_o, a, e)


def test_generated_djn_key_roundtrip():
    """ipclPublicKey(n, bits, True) without hs: the library draws x and builds hs on the device."""
    n, p, q = capi.keygen(1024)
    pk = capi.PubKey(n, 1024, djn=True)
    hs = pk.hs
    assert 1 < hs < n * n and pow(hs, 1, n * n) == hs
    sk = capi.PrivKey(pk, p, q)
    pk_o = O.PubKey(n, 1024, True, hs, 512)
    rng = random.Random(SEED + 9)
    ms = [rng.randrange(n) for _ in range(8)]
    rs = [rng.getrandbits(512) for _ in ms]
    ct = pk.encrypt(capi.ints_to_array(ms, 32), capi.ints_to_array(rs, 16))
    assert capi.array_to_ints(ct) == O.encrypt_batch(pk_o, ms, rs)
    assert capi.array_to_ints(sk.decrypt(ct)) == ms


def test_generic_modexp():
    rng = random.Random(SEED + 10)
    for bits in (512, 1024, 2048, 4096):
        mod = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        base = [0, 1, mod - 1, mod + 5] + [rng.randrange(mod) for _ in range(6)]
        exp = [5, 0, 2, 3] + [rng.getrandbits(bits) for _ in range(6)]
        assert capi.modexp(base, exp, mod, bits // 32) == [pow(b, e, mod) for b, e in zip(base, exp)]


def test_empty_batches(key2048):
    pk_o, sk_o, pk, sk = key2048
    assert pk.encrypt(np.zeros((0, 64), dtype=np.uint32)).shape == (0, 128)
    assert sk.decrypt(np.zeros((0, 128), dtype=np.uint32)).shape == (0, 64)


# ---- committed golden vectors (tests/golden/paillier_kat.json) through the C ABI ---------------------------------
def _kat():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "paillier_kat.json")) as f:
        return json.load(f)["keys"]


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_golden_vectors(idx):
    k = _kat()[idx]
    ints = lambda xs: [int(x, 16) for x in xs]  # noqa: E731
    n, p, q, bits = int(k["n"], 16), int(k["p"], 16), int(k["q"], 16), k["bits"]
    nw = bits // 32
    pk = capi.PubKey(n, bits, djn=k["djn"], hs=int(k["hs"], 16) if k["djn"] else None)
    sk = capi.PrivKey(pk, p, q)
    rw = (k["randbits"] + 31) // 32 if k["djn"] else nw
    m = capi.ints_to_array(ints(k["m"]), nw)
    ct = pk.encrypt(m, capi.ints_to_array(ints(k["r"]), rw))
    assert capi.array_to_ints(ct) == ints(k["ct"])
    assert capi.array_to_ints(pk.encrypt(m, None, make_secure=False)) == ints(k["ct_raw"])
    assert capi.array_to_ints(sk.decrypt(capi.ints_to_array(ints(k["dec_in"]), 2 * nw))) == ints(k["dec_out"])
    b = capi.ints_to_array(ints(k["add_b"]), 2 * nw)
    assert capi.array_to_ints(pk.add(ct, b)) == ints(k["add_out"])
    assert capi.array_to_ints(pk.add(ct, b[:1])) == ints(k["add_bcast_out"])
    e = capi.ints_to_array(ints(k["mul_e"]), nw)
    assert capi.array_to_ints(pk.mul(ct, e)) == ints(k["mul_out"])
    assert capi.array_to_ints(pk.mul(ct, e[3:4])) == ints(k["mul_bcast_out"])


def test_batch_vs_c_oracle_ragged_sizes(key2048):
    """Ragged batch sizes around the CTA / wave boundaries, every element checked against the C oracle."""
    import os

    import c_oracle
    pk_o, sk_o, pk, sk = key2048
    rng = np.random.Generator(np.random.PCG64(SEED + 11))
    th = os.cpu_count() or 1
    for count in (1, 2, 31, 33, 63, 65, 127, 129, 1000):
        m = np.zeros((count, 64), dtype=np.uint32)
        m[:, :3] = rng.integers(0, 2**32, size=(count, 3), dtype=np.uint64).astype(np.uint32)
        r = rng.integers(0, 2**32, size=(count, 32), dtype=np.uint64).astype(np.uint32)
        ct = pk.encrypt(m, r)
        assert np.array_equal(ct, c_oracle.encrypt(pk_o.n, 64, pk_o.hs, m, r, threads=th))
        assert np.array_equal(sk.decrypt(ct), m)
        e = rng.integers(0, 2**32, size=(count, 2), dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(pk.mul(ct, e), c_oracle.mul(pk_o.n, 64, ct, e, threads=th))
        assert np.array_equal(pk.add(ct, ct[::-1].copy()), c_oracle.add(pk_o.n, 64, ct, ct[::-1].copy(), threads=th))


def test_timing_hooks_and_pipe_peak(key2048):
    pk_o, sk_o, pk, sk = key2048
    capi.timing_enable(True)
    ct = pk.encrypt(np.zeros((64, 64), dtype=np.uint32), np.ones((64, 32), dtype=np.uint32))
    sk.decrypt(ct)
    t = capi.timing_read()
    capi.timing_enable(False)
    assert t["k_encrypt_npair"][1] == 1 and t["k_encrypt_comb"][1] == 0 and t["k_dec_pair"][1] == 1 and t["k_dec_pair"][0] > 0 and t["k_dec_crt"][1] == 1
    peak = capi.int_pipe_peak(2)
    assert 4e12 < peak < 12e12   # IMAD.WIDE.U32: one warp instruction per 4 cycles per SM sub-partition
    assert 10e12 < capi.fp64_pipe_peak(2) < 20e12      # DFMA: one warp instruction per 2 cycles per sub-partition
    assert 2e12 < capi.product_mix_peak(2) < 7e12      # 2 DFMA + DADD + IADD3 + IADD3.X per limb product


def test_decrypt_generic_path_and_unbalanced_key(monkeypatch):
    """Keys whose primes are not both exactly bits/2 long cannot use the p-adic pair engine and take the generic
    k_dec_prep / k_powm_prog / k_dec_tail path; PHE_NO_PAIR_ENGINE=1 forces that path for any key.  Same results."""
    rng = random.Random(SEED + 99)
    pk_o, sk_o = O.bench_keypair()
    monkeypatch.setenv("PHE_NO_PAIR_ENGINE", "1")
    pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
    assert capi.pair_block(sk, 0) is None
    monkeypatch.delenv("PHE_NO_PAIR_ENGINE")
    sk2 = capi.PrivKey(pk, sk_o.p, sk_o.q)
    assert capi.pair_block(sk2, 0)["L"] == 20
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(29)]
    cs = O.encrypt_batch(pk_o, ms, [rng.getrandbits(1024) for _ in ms]) + _rand_cts(pk_o, rng, 32) + [0, sk_o.p, pk_o.n]
    want = O.decrypt_batch(sk_o, cs)
    capi.timing_enable(True)
    got = capi.array_to_ints(sk.decrypt(capi.ints_to_array(cs, 128)))
    t = capi.timing_read()
    capi.timing_enable(False)
    assert got == want and t["k_powm"][1] == 1 and t["k_dec_pair"][1] == 0
    assert capi.array_to_ints(sk2.decrypt(capi.ints_to_array(cs, 128))) == want
    # unbalanced primes: 1020-bit p, 1024-bit q (grossly unbalanced ones, q^2 > 2^2048, are rejected)
    import sympy
    with pytest.raises(RuntimeError):
        capi.PrivKey(capi.PubKey((2 ** 1000 - 1) * (2 ** 1048 - 1), 2048, djn=False), 2 ** 1000 - 1, 2 ** 1048 - 1)
    p = sympy.nextprime(rng.getrandbits(1020) | (1 << 1019))
    q = sympy.nextprime(rng.getrandbits(1024) | (1 << 1023))
    n = p * q
    x = rng.getrandbits(2200) % n
    hs = pow((-x * x) % n, n, n * n)
    pk_u = O.PubKey(n, 2048, True, hs, 1024)
    sk_u = O.PrivKey(pk_u, p, q)
    cpk = capi.PubKey(n, 2048, djn=True, hs=hs)
    csk = capi.PrivKey(cpk, p, q)
    assert capi.pair_block(csk, 0) is None
    ms = [0, n - 1] + [rng.randrange(n) for _ in range(14)]
    rs = [rng.getrandbits(1024) for _ in ms]
    ct = cpk.encrypt(capi.ints_to_array(ms, 64), capi.ints_to_array(rs, 32))
    assert capi.array_to_ints(ct) == O.encrypt_batch(pk_u, ms, rs)
    assert capi.array_to_ints(csk.decrypt(ct)) == ms == O.decrypt_batch(sk_u, capi.array_to_ints(ct))


def _chacha20_block(key, counter, nonce):
    """RFC 8439 section 2.3 block function (reference implementation for the known-answer test)."""
    def rotl(v, c): return ((v << c) | (v >> (32 - c))) & 0xFFFFFFFF
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [counter] + list(nonce)
    x = list(s)
    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)
    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, s)]


def test_device_csprng_chacha20_kat_and_internal_r(key2048):
    """The device keystream is ChaCha20: RFC 8439 2.3.2 test vector + a longer run against the reference above; an
    encrypt with library-drawn r decrypts correctly, never repeats, and r stays below 2^randbits."""
    import ctypes
    key = np.frombuffer(bytes(range(32)), dtype="<u4").copy()
    nonce = np.frombuffer(bytes.fromhex("000000090000004a00000000"), dtype="<u4").copy()
    out = np.zeros(16 * 37 + 5, dtype=np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    rc = capi.lib().phe_chacha20_keystream(key.ctypes.data_as(u32p), nonce.ctypes.data_as(u32p), 1,
                                           out.ctypes.data_as(u32p), ctypes.c_size_t(out.size))
    assert rc == 0
    assert out[:16].tobytes().hex() == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                                        "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    want = sum((_chacha20_block(key.tolist(), 1 + b, nonce.tolist()) for b in range(38)), [])[:out.size]
    assert out.tolist() == want
    pk_o, sk_o, pk, sk = key2048
    m = capi.ints_to_array([5, 6, 7, 5, 5, 5, 5, 5], 64)
    c1, c2 = pk.encrypt(m), pk.encrypt(m)
    assert np.array_equal(sk.decrypt(c1), m) and np.array_equal(sk.decrypt(c2), m)
    rows = {c.tobytes() for c in np.concatenate([c1, c2])}
    assert len(rows) == 16                      # fresh r for every element of every call


@pytest.mark.parametrize("count", [1, 5, 16, 17, 1000, 20011])
def test_batched_modular_inverse(key2048, count):
    """phe_invert (Montgomery's trick on the device, recursive over block totals) == pow(c, -1, n^2) row by row."""
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + count)
    cs = _rand_cts(pk_o, rng, min(count, 40))
    cs = (cs * (count // len(cs) + 1))[:count]          # repeated rows are fine: every row is inverted on its own
    cs[0] = 1
    cs[-1] = pk_o.nsquare - 1
    got = pk.invert(capi.ints_to_array(cs, 128))
    inv = {c: pow(c, -1, pk_o.nsquare) for c in set(cs)}
    assert capi.array_to_ints(got) == [inv[c] for c in cs]
    if count >= 17:
        bad = list(cs)
        bad[count // 2] = sk_o.p * 12345                # shares the factor p with n
        with pytest.raises(RuntimeError):
            pk.invert(capi.ints_to_array(bad, 128))


@pytest.mark.parametrize("bits", [1024, 2048, 3072])
def test_npair_engine_and_n2_engine_agree(monkeypatch, bits):
    """HE mul and DJN encrypt run on the n-adic pair engine by default; PHE_NO_NPAIR_ENGINE=1 keeps a key on the
    Montgomery engine mod n^2.  Both must give the oracle's bits."""
    pk_o, sk_o = O.seeded_keypair(bits, 9)
    nw = bits // 32
    rng = random.Random(SEED + bits)
    ms = [0, 1, pk_o.n - 1] + [rng.randrange(pk_o.n) for _ in range(40)] + [rng.getrandbits(53) for _ in range(60)]
    rs = [0, 1, (1 << (bits // 2)) - 1] + [rng.getrandbits(bits // 2) for _ in range(len(ms) - 3)]
    cts = [0, 1, pk_o.nsquare - 1, pk_o.n] + [rng.randrange(pk_o.nsquare) for _ in range(60)]
    want_ct = O.encrypt_batch(pk_o, ms, rs)
    for ebits in (1, 53, 200, bits):
        es = [rng.getrandbits(ebits) for _ in cts]
        es[0] |= 1 << (ebits - 1)
        es[1] = 0
        want_mul = [pow(c, e, pk_o.nsquare) for c, e in zip(cts, es)]
        for off in ("0", "1"):
            monkeypatch.setenv("PHE_NO_NPAIR_ENGINE", off)
            pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
            assert (capi.npair_block(pk) is None) == (off == "1")
            pk.set_comb_bits(6)
            got = capi.array_to_ints(pk.mul(capi.ints_to_array(cts, 2 * nw), capi.ints_to_array(es, (ebits + 31) // 32)))
            assert got == want_mul
            if ebits == 53:
                got = capi.array_to_ints(pk.encrypt(capi.ints_to_array(ms, nw), capi.ints_to_array(rs, (bits // 2 + 31) // 32)))
                assert got == want_ct
                got = capi.array_to_ints(pk.encrypt(capi.ints_to_array(ms, nw), None, make_secure=False))
                assert got == O.encrypt_batch(pk_o, ms, None)
                # broadcast exponent
                got = capi.array_to_ints(pk.mul(capi.ints_to_array(cts, 2 * nw), capi.ints_to_array(es[:1], 2)))
                assert got == [pow(c, es[0], pk_o.nsquare) for c in cts]


def test_encrypt_dev_multi_stores_every_row_to_all_buffers(key2048):
    """phe_encrypt_dev_multi (the fused gather of BASELINE config 4): the encrypt kernel writes every row to the main
    output and to each extra buffer -- here three buffers on the same GPU stand in for the peer-mapped ones."""
    import torch
    pk_o, sk_o, pk, sk = key2048
    rng = np.random.default_rng(SEED + 5)
    count = 1000
    m_np = np.zeros((count, 64), dtype=np.uint32)
    m_np[:, :2] = rng.integers(0, 1 << 32, size=(count, 2), dtype=np.uint64).astype(np.uint32)
    r_np = rng.integers(0, 1 << 32, size=(count, 32), dtype=np.uint64).astype(np.uint32)
    dev = torch.device("cuda", 0)
    m = torch.from_numpy(m_np.view(np.int32)).to(dev)
    r = torch.from_numpy(r_np.view(np.int32)).to(dev)
    outs = [torch.zeros((count + 7, 128), dtype=torch.int32, device=dev) for _ in range(4)]
    row0 = 7   # the caller points every buffer at its first row
    pk.encrypt_dev_multi(m.data_ptr(), count, r.data_ptr(), 32, outs[0].data_ptr() + row0 * 512,
                         [o.data_ptr() + row0 * 512 for o in outs[1:]], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = pk.encrypt(m_np, r_np)
    for o in outs:
        got = o.cpu().numpy().view(np.uint32)
        assert not got[:row0].any()
        assert np.array_equal(got[row0:], want)
    idx = [0, 1, count - 1]
    assert capi.array_to_ints(want[idx]) == O.encrypt_batch(pk_o, capi.array_to_ints(m_np[idx]), capi.array_to_ints(r_np[idx]))


def test_two_streams_share_one_key(key2048):
    """Stream rule of include/phe_b200.h: calls on one key may come from different (non-blocking) streams; the library
    chains them with events so that they never overlap on the key's scratch.  Two different batches are decrypted and
    multiplied concurrently from two streams; both must match the oracle."""
    import torch
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + 30)
    dev = torch.device("cuda", 0)
    n_rows = 3000
    batches = []
    for b in range(2):
        ms = [rng.getrandbits(53) + b for _ in range(n_rows)]
        ct_np = np.zeros((n_rows, 128), dtype=np.uint32)
        ct_np[:, :66] = capi.ints_to_array([1 + m * pk_o.n for m in ms], 66)     # raw ciphertexts 1 + m n
        batches.append((ms, torch.from_numpy(ct_np.view(np.int32)).to(dev)))
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    outs = [torch.empty((n_rows, 64), dtype=torch.int32, device=dev) for _ in range(2)]
    prods = [torch.empty((n_rows, 128), dtype=torch.int32, device=dev) for _ in range(2)]
    e = torch.from_numpy(capi.ints_to_array([3, 0], 2)[:1].view(np.int32).copy()).to(dev)
    torch.cuda.synchronize()
    for rep in range(3):
        for b in (0, 1):
            s = streams[b].cuda_stream
            sk.decrypt_dev(batches[b][1].data_ptr(), n_rows, outs[b].data_ptr(), s)
            pk.mul_dev(batches[b][1].data_ptr(), n_rows, e.data_ptr(), 2, 1, 2, prods[b].data_ptr(), s)
    torch.cuda.synchronize()
    for b in (0, 1):
        ms = batches[b][0]
        assert capi.array_to_ints(outs[b].cpu().numpy().view(np.uint32)[:, :2]) == ms
        got = capi.array_to_ints(prods[b].cpu().numpy().view(np.uint32)[:16])
        assert got == [pow(1 + m * pk_o.n, 3, pk_o.nsquare) for m in ms[:16]]


def test_device_encrypt_draws_r_itself(key2048):
    """phe_encrypt_dev(make_secure=1, d_r=NULL) draws the obfuscator exponents on the device, exactly as phe_encrypt does:
    ciphertexts differ from the raw ones and from each other, and decrypt to the plaintexts; make_secure=0 is 1 + m n."""
    import torch
    pk_o, sk_o, pk, sk = key2048
    dev = torch.device("cuda", 0)
    ms = [5, 5, 7, 123456789]
    m = torch.from_numpy(capi.ints_to_array(ms, 64).view(np.int32)).to(dev)
    ct = torch.empty((4, 128), dtype=torch.int32, device=dev)
    pk.encrypt_dev(m.data_ptr(), 4, None, 0, ct.data_ptr(), 0, make_secure=True)
    torch.cuda.synchronize()
    sec = capi.array_to_ints(ct.cpu().numpy().view(np.uint32))
    assert sec[0] != sec[1] and all(c != 1 + v * pk_o.n for c, v in zip(sec, ms))
    assert O.decrypt_batch(sk_o, sec) == ms
    pk.encrypt_dev(m.data_ptr(), 4, None, 0, ct.data_ptr(), 0, make_secure=False)
    torch.cuda.synchronize()
    assert capi.array_to_ints(ct.cpu().numpy().view(np.uint32)) == [1 + v * pk_o.n for v in ms]


@pytest.mark.parametrize("bits", [200, 1000, 1028, 2044])
def test_reference_key_lengths(bits):
    """The reference accepts any n_length % 4 == 0 in [200, 2048]: generate, encrypt with pinned r,
    add, multiply, decrypt -- all against the oracle.  (Keys whose primes do not fill their words decrypt on the generic
    Montgomery path instead of the p-adic pair engine.)"""
    n, p, q = capi.keygen(bits)
    assert n.bit_length() == bits and p * q == n
    nw = (bits + 31) // 32
    pk = capi.PubKey(n, bits, djn=True)
    sk = capi.PrivKey(pk, p, q)
    pk_o = O.PubKey(n, bits, True, pk.hs, bits // 2)
    sk_o = O.PrivKey(pk_o, p, q)
    rng = random.Random(SEED + bits)
    ms = [0, 1, n - 1] + [rng.randrange(n) for _ in range(13)]
    rs = [rng.getrandbits(bits // 2) for _ in ms]
    ct = pk.encrypt(capi.ints_to_array(ms, nw), capi.ints_to_array(rs, (bits // 2 + 31) // 32))
    assert capi.array_to_ints(ct) == O.encrypt_batch(pk_o, ms, rs)
    assert capi.array_to_ints(sk.decrypt(ct)) == ms
    a, b = capi.array_to_ints(ct[:8]), capi.array_to_ints(ct[8:])
    assert capi.array_to_ints(pk.add(ct[:8], ct[8:])) == O.add_batch(pk_o, a, b)
    e = [rng.getrandbits(53) for _ in range(8)]
    assert capi.array_to_ints(pk.mul(ct[:8], capi.ints_to_array(e, 2))) == O.mul_batch(pk_o, a, e)


def _dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int32)).to("cuda:0")


def _host(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("engine", ["npair", "n2"])
def test_row_operations_on_device(key2048, engine, monkeypatch):
    """phe_gather_rows_dev / phe_scatter_rows_dev / phe_scale_rows_dev / phe_invert_rows_dev / phe_segsum_dev and the
    one-product broadcast add, against Python ints (both engines for the scaling)."""
    import torch
    pk_o, sk_o, pk, sk = key2048
    if engine == "n2":
        monkeypatch.setenv("PHE_NO_NPAIR_ENGINE", "1")
        pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    n2 = pk_o.nsquare
    rng = random.Random(SEED + 40)
    rows = 37
    vals = _rand_cts(pk_o, rng, rows)
    d = _dev(capi.ints_to_array(vals, 128))
    # gather (with repeats: a broadcast) and scatter
    idx = [5, 5, 0, 36, 7, 5]
    g = torch.zeros((len(idx), 128), dtype=torch.int32, device="cuda:0")
    pk.gather_rows_dev(d.data_ptr(), rows, idx, g.data_ptr())
    assert capi.array_to_ints(_host(g)) == [vals[i] for i in idx]
    tgt = d.clone()
    sidx = [3, 30, 1]
    pk.scatter_rows_dev(g.data_ptr(), sidx, tgt.data_ptr(), rows)   # first three rows of g
    want = list(vals)
    for k, i in enumerate(sidx):
        want[i] = vals[idx[k]]
    assert capi.array_to_ints(_host(tgt)) == want
    with pytest.raises(RuntimeError, match="out of range"):
        pk.gather_rows_dev(d.data_ptr(), rows, [0, rows], g.data_ptr())
    # exponent alignment in place: unsorted deltas, zero, and one beyond 32 * n_words bits
    work = d.clone()
    sc_idx = [2, 11, 36, 0, 17, 9]
    deltas = [3, 0, 2100 if engine == "npair" else 70, 53, 1, 53]
    pk.scale_rows_dev(work.data_ptr(), rows, sc_idx, deltas)
    want = list(vals)
    for i, dl in zip(sc_idx, deltas):
        want[i] = pow(vals[i], 1 << dl, n2)
    assert capi.array_to_ints(_host(work)) == want
    with pytest.raises(RuntimeError, match="duplicate"):
        pk.scale_rows_dev(work.data_ptr(), rows, [1, 1], [1, 2])
    # inverse of a subset of rows in place
    work = d.clone()
    inv_idx = [0, 4, 8, 36, 20]
    pk.invert_rows_dev(work.data_ptr(), rows, inv_idx)
    want = list(vals)
    for i in inv_idx:
        want[i] = pow(vals[i], -1, n2)
    assert capi.array_to_ints(_host(work)) == want
    # add trees: every width up to 9, then a long one; groups > 1
    for groups, width in [(1, 1), (3, 1), (1, 2), (4, 2), (2, 3), (5, 5), (1, 7), (3, 8), (4, 9), (1, 37)]:
        use = vals[: groups * width] if groups * width <= rows else [vals[i % rows] for i in range(groups * width)]
        src = _dev(capi.ints_to_array(use, 128))
        out = torch.empty((groups, 128), dtype=torch.int32, device="cuda:0")
        pk.segsum_dev(src.data_ptr(), groups, width, out.data_ptr())
        want = []
        for gi in range(groups):
            acc = 1
            for v in use[gi * width:(gi + 1) * width]:
                acc = acc * v % n2
            want.append(acc)
        assert capi.array_to_ints(_host(out)) == want, (groups, width)
    # broadcast add: one product per element
    out = torch.empty_like(d)
    pk.add_dev(d.data_ptr(), rows, d[7:8].contiguous().data_ptr(), 1, out.data_ptr())
    assert capi.array_to_ints(_host(out)) == [v * vals[7] % n2 for v in vals]


def test_segsum_large(key2048):
    """100 000 rows into one sum, and a 64 x 64 block of 64-wide sums, against products of Python ints."""
    import torch
    pk_o, sk_o, pk, sk = key2048
    n2 = pk_o.nsquare
    rng = random.Random(SEED + 41)
    base = _rand_cts(pk_o, rng, 64)
    rows = 100000
    src = _dev(capi.ints_to_array(base, 128))[torch.arange(rows, device="cuda:0") % 64].contiguous()
    out = torch.empty((1, 128), dtype=torch.int32, device="cuda:0")
    pk.segsum_dev(src.data_ptr(), 1, rows, out.data_ptr())
    want = 1
    for j, v in enumerate(base):
        want = want * pow(v, rows // 64 + (1 if j < rows % 64 else 0), n2) % n2
    assert capi.array_to_ints(_host(out)) == [want]
    out = torch.empty((64 * 24, 128), dtype=torch.int32, device="cuda:0")
    pk.segsum_dev(src.data_ptr(), 64 * 24, 64, out.data_ptr())        # every group of 64 = the same 64 values
    tot = 1
    for v in base:
        tot = tot * v % n2
    got = capi.array_to_ints(_host(out[:3])) + capi.array_to_ints(_host(out[-1:]))
    assert got == [tot] * 4


def test_comb_table_is_shared_between_objects_of_one_key(key2048):
    """Two phe_pubkey objects of the same key use one fixed-base table: the second neither builds nor holds a copy."""
    pk_o, sk_o, pk, sk = key2048
    rng = random.Random(SEED + 50)
    ms = [rng.getrandbits(53) for _ in range(64)]
    rs = [rng.getrandbits(1024) for _ in ms]
    m, r = capi.ints_to_array(ms, 64), capi.ints_to_array(rs, 32)
    a = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    a.set_comb_bits(14)
    want = a.encrypt(m, r)                                  # builds the 14-bit table (1 + 13 build launches + encrypt)
    assert a.comb_bits == 14 and a.comb_info[0] > 0
    b = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    b.set_comb_bits(14)
    before = capi.kernel_launches()
    got = b.encrypt(m, r)
    assert capi.kernel_launches() - before == 1             # the encrypt kernel only: no table build
    assert np.array_equal(got, want) and b.comb_info == a.comb_info
    del a                                                   # the table stays alive through b
    assert np.array_equal(b.encrypt(m, r), want)
    c = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs + 0)  # a different width is a different table
    c.set_comb_bits(9)
    assert np.array_equal(c.encrypt(m, r), want) and c.comb_info[0] != b.comb_info[0]


def test_decrypt_mantissas_classification(key2048):
    """phe_decrypt_mantissas: decrypt + per-row sign / 63-bit mantissa on the device, against the numpy restatement
    (fixedpoint.classify_plain) and the plaintexts themselves."""
    from pailliercryptolib_python_b200.fixedpoint import classify_plain
    pk_o, sk_o, pk, sk = key2048
    n = pk_o.n
    rng = random.Random(SEED + 60)
    ms = [0, 1, (1 << 63) - 1, 1 << 63, n - 1, n - (1 << 63) + 1, n - (1 << 63), n - (1 << 63) - 1, n // 2, n // 3]
    ms += [rng.getrandbits(53) for _ in range(40)] + [n - rng.getrandbits(60) - 1 for _ in range(40)] + [rng.randrange(n) for _ in range(10)]
    ct = capi.ints_to_array([1 + m * n for m in ms], 128)        # raw ciphertexts: decrypt gives ms back
    mant, cls, rows = sk.decrypt_mantissas(ct)
    want_mant, want_cls = classify_plain(capi.ints_to_array(ms, 64), n)
    assert np.array_equal(cls, want_cls) and np.array_equal(mant, want_mant)
    for i, m in enumerate(ms):
        if cls[i] == 0:
            assert int(mant[i]) == m
        elif cls[i] == 1:
            assert n + int(mant[i]) == m
        else:
            assert capi.words_to_int(rows[i]) == m
    assert sorted(set(cls.tolist())) == [0, 1, 2]

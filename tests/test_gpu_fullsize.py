"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run 10^5 - 10^6 modexps
in test time): round trips, homomorphic linearity checked after decryption, a checksum of checksums across the two
engines, and oracle comparison of seeded samples.  All through the C ABI on device-resident batches."""
import random

import numpy as np
import pytest

import c_oracle
import paillier_oracle as O
from pailliercryptolib_python_b200 import capi

pytestmark = pytest.mark.gpu
SEED = 20240611


def _dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int32)).to("cuda:0")


def _host(t):
    return np.ascontiguousarray(t.cpu().numpy().view(np.uint32))


@pytest.fixture(scope="module")
def key():
    pk_o, sk_o = O.bench_keypair()
    pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
    return pk_o, sk_o, pk, capi.PrivKey(pk, sk_o.p, sk_o.q)


def test_config2_100k_encrypt_decrypt_and_linearity(key):
    """configs[1]: 100 000 obfuscated encrypts and decrypts.  D(E(a)) = a; D(E(a) (+) E(b)) = a + b mod n;
    D(E(a)^k) = k a mod n for 53-bit k -- all 100 000 rows; 2 000 seeded ciphertext rows against the CPU oracle."""
    import torch
    pk_o, sk_o, pk, sk = key
    N, n = 100000, pk_o.n
    rng = np.random.Generator(np.random.PCG64(SEED))
    a = np.zeros((N, 64), dtype=np.uint32)
    b = np.zeros((N, 64), dtype=np.uint32)
    a[:, :2] = rng.integers(0, 1 << 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
    b[:, :3] = rng.integers(0, 1 << 32, size=(N, 3), dtype=np.uint64).astype(np.uint32)     # 96-bit plaintexts
    r = rng.integers(0, 1 << 32, size=(2, N, 32), dtype=np.uint64).astype(np.uint32)
    k = np.zeros((N, 2), dtype=np.uint32)
    k[:, 0] = rng.integers(0, 1 << 32, size=N, dtype=np.uint64).astype(np.uint32)
    k[:, 1] = rng.integers(0, 1 << 21, size=N, dtype=np.uint64).astype(np.uint32)
    da, db, dk = _dev(a), _dev(b), _dev(k)
    ca = torch.empty((N, 128), dtype=torch.int32, device="cuda:0")
    cb, cs, cp = torch.empty_like(ca), torch.empty_like(ca), torch.empty_like(ca)
    out = torch.empty((N, 64), dtype=torch.int32, device="cuda:0")
    pk.encrypt_dev(da.data_ptr(), N, _dev(r[0]).data_ptr(), 32, ca.data_ptr())
    pk.encrypt_dev(db.data_ptr(), N, _dev(r[1]).data_ptr(), 32, cb.data_ptr())
    sk.decrypt_dev(ca.data_ptr(), N, out.data_ptr())
    assert torch.equal(out, da)
    sel = np.random.Generator(np.random.PCG64(1)).choice(N, size=2000, replace=False)
    assert np.array_equal(_host(ca)[sel], c_oracle.encrypt(n, 64, pk_o.hs, a[sel], r[0][sel], threads=8))
    # linearity: a + b < n (96-bit + 64-bit values), k a < n (53 + 64 bits)
    pk.add_dev(ca.data_ptr(), N, cb.data_ptr(), N, cs.data_ptr())
    sk.decrypt_dev(cs.data_ptr(), N, out.data_ptr())
    got = _host(out)
    av = a[:, 0].astype(object) + (a[:, 1].astype(object) << 32)
    bv = b[:, 0].astype(object) + (b[:, 1].astype(object) << 32) + (b[:, 2].astype(object) << 64)
    want = capi.ints_to_array(list(av + bv), 64)
    assert np.array_equal(got, want)
    pk.mul_dev(ca.data_ptr(), N, dk.data_ptr(), 2, N, 53, cp.data_ptr())
    sk.decrypt_dev(cp.data_ptr(), N, out.data_ptr())
    kv = k[:, 0].astype(object) + (k[:, 1].astype(object) << 32)
    assert np.array_equal(_host(out), capi.ints_to_array(list(av * kv), 64))


def test_config3_1m_add_mul_both_engines_agree(key, monkeypatch):
    """configs[2]: 1 M HE adds and HE muls.  The n-adic pair engine and the Montgomery engine mod n^2 are two independent
    implementations of the same map: their 1 M outputs must be identical (compared on the device), and 1 000 seeded rows
    equal the CPU oracle."""
    import torch
    pk_o, sk_o, pk, sk = key
    M, n = 1 << 20, pk_o.n
    g = torch.Generator(device="cuda:0")
    g.manual_seed(SEED)
    a = torch.randint(0, 2**31 - 1, (M, 128), device="cuda:0", dtype=torch.int32, generator=g)
    b = torch.randint(0, 2**31 - 1, (M, 128), device="cuda:0", dtype=torch.int32, generator=g)
    a[:, 127] &= 0x0FFFFFFF
    b[:, 127] &= 0x0FFFFFFF
    e = torch.randint(0, 2**31 - 1, (M, 2), device="cuda:0", dtype=torch.int32, generator=g)
    e[:, 1] &= (1 << 21) - 1
    monkeypatch.setenv("PHE_NO_NPAIR_ENGINE", "1")
    pk2 = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)       # Montgomery engine mod n^2
    o1, o2 = torch.empty_like(a), torch.empty_like(a)
    sel = torch.from_numpy(np.random.Generator(np.random.PCG64(2)).choice(M, size=1000, replace=False)).to("cuda:0")
    pk.mul_dev(a.data_ptr(), M, e.data_ptr(), 2, M, 53, o1.data_ptr())
    pk2.mul_dev(a.data_ptr(), M, e.data_ptr(), 2, M, 53, o2.data_ptr())
    assert torch.equal(o1, o2)
    assert np.array_equal(_host(o1[sel]), c_oracle.mul(n, 64, _host(a[sel]), _host(e[sel]), threads=8))
    pk.add_dev(a.data_ptr(), M, b.data_ptr(), M, o1.data_ptr())
    assert np.array_equal(_host(o1[sel]), c_oracle.add(n, 64, _host(a[sel]), _host(b[sel]), threads=8))
    # the add tree against the chained adds: product of 16 rows per group, 65 536 groups
    pk.segsum_dev(a.data_ptr(), M // 16, 16, o2.data_ptr())
    acc = a.view(M // 16, 16, 128)[:, 0].contiguous()
    tmp = torch.empty_like(acc)
    for j in range(1, 16):
        pk.add_dev(acc.data_ptr(), M // 16, a.view(M // 16, 16, 128)[:, j].contiguous().data_ptr(), M // 16, tmp.data_ptr())
        acc, tmp = tmp, acc
    assert torch.equal(o2[: M // 16], acc)


def test_config5_3072_bit_100k_round_trip():
    """configs[4]: 3072-bit key, 100 000 encrypt + decrypt: full round trip and 64 rows against the Python-int oracle
    (the reference itself stops at 2048 bits, ipcl_python.py:29-30)."""
    import torch
    bits, N = 3072, 100000
    nw = bits // 32
    pk_o, sk_o = O.seeded_keypair(bits, 77)
    pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
    sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
    rng = np.random.Generator(np.random.PCG64(SEED))
    m = np.zeros((N, nw), dtype=np.uint32)
    m[:, :2] = rng.integers(0, 1 << 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
    m[:8] = capi.ints_to_array([0, 1, pk_o.n - 1, pk_o.n // 3, pk_o.n - 2, 2, 3, 4], nw)
    r = rng.integers(0, 1 << 32, size=(N, nw // 2), dtype=np.uint64).astype(np.uint32)
    dm, dr = _dev(m), _dev(r)
    ct = torch.empty((N, 2 * nw), dtype=torch.int32, device="cuda:0")
    out = torch.empty((N, nw), dtype=torch.int32, device="cuda:0")
    pk.encrypt_dev(dm.data_ptr(), N, dr.data_ptr(), nw // 2, ct.data_ptr())
    sk.decrypt_dev(ct.data_ptr(), N, out.data_ptr())
    assert torch.equal(out, dm)
    idx = list(range(8)) + random.Random(5).sample(range(N), 56)
    want = O.encrypt_batch(pk_o, capi.array_to_ints(m[idx]), capi.array_to_ints(r[idx]))
    assert capi.array_to_ints(_host(ct)[idx]) == want


def test_decrypt_across_launch_boundary(key):
    """A decrypt of more than 2^17 ciphertexts runs as several launches of k_dec_pair (the per-unit window tables bound a
    launch): 140 000 = one time-sliced launch of 131 072 + one of 8 928 with whole units.  Round trip over every row, the
    rows either side of the boundary and the ragged end against the oracle."""
    import torch
    pk_o, sk_o, pk, sk = key
    N = 140000
    rng = np.random.Generator(np.random.PCG64(SEED + 9))
    m = np.zeros((N, 64), dtype=np.uint32)
    m[:, :2] = rng.integers(0, 1 << 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
    r = rng.integers(0, 1 << 32, size=(N, 32), dtype=np.uint64).astype(np.uint32)
    dm = _dev(m)
    ct = torch.empty((N, 128), dtype=torch.int32, device="cuda:0")
    out = torch.empty((N, 64), dtype=torch.int32, device="cuda:0")
    pk.encrypt_dev(dm.data_ptr(), N, _dev(r).data_ptr(), 32, ct.data_ptr())
    sk.decrypt_dev(ct.data_ptr(), N, out.data_ptr())
    assert torch.equal(out, dm)
    idx = [0, 131071, 131072, 131073, N - 33, N - 1]
    cts = capi.array_to_ints(_host(ct[idx]))
    assert capi.array_to_ints(_host(out[idx])) == O.decrypt_batch(sk_o, cts)

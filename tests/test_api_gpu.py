"""GPU tests of the ipcl_python-compatible API: the reference's own test flows (tests/ipcl_python_test.py:21-119:
test_add, test_mul, test_matmul, test_rmatmul, test_imatmul -- here with the assertions the reference forgot) plus
bit-level parity of every operator against the Python-int oracle on pinned keys."""
import pickle
import random

import numpy as np
import pytest

import paillier_oracle as O
from pailliercryptolib_python_b200 import (BNUtils, PaillierEncryptedNumber, PaillierKeypair, PaillierPrivateKey,
                                           PaillierPublicKey)
from pailliercryptolib_python_b200.fixedpoint import FixedPointNumber

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def keys():
    return PaillierKeypair.generate_keypair(2048, True)      # as the reference's setUp (ipcl_python_test.py:11-15)


@pytest.fixture(scope="module")
def bench():
    pk_o, sk_o = O.bench_keypair()
    from pailliercryptolib_python_b200.bindings.ipcl_bindings import ipclPublicKey
    pub = PaillierPublicKey(ipclPublicKey.create(BNUtils.int2BN(pk_o.n), 2048, BNUtils.int2BN(pk_o.hs), 1024))
    pri = PaillierPrivateKey(pub, sk_o.p, sk_o.q)
    return pk_o, sk_o, pub, pri


def _cts(x):
    return [BNUtils.BN2int(b) for b in x.ciphertextBN()]


def test_add_reference_flow(keys):
    pub, pri = keys
    rng = np.random.RandomState(1)
    a = np.arange(100) * 1.5
    b = rng.rand(100) * 1000 - 500
    c = list(range(100))
    d = rng.rand(100)
    ct = pub.encrypt(a) + pub.encrypt(b) + pub.encrypt(c) + pub.encrypt(d)
    got = pri.decrypt(ct)
    for g, w in zip(got, a + b + np.array(c) + d):
        assert g == pytest.approx(w, abs=1e-7)
    # plaintext operands: list, array, scalar; radd; sub; rsub
    e = pub.encrypt(a)
    assert pri.decrypt(e + list(b)) == pytest.approx(list(a + b), abs=1e-7)
    assert pri.decrypt(e + 5) == pytest.approx(list(a + 5), abs=1e-9)
    assert pri.decrypt(2.5 + e) == pytest.approx(list(a + 2.5), abs=1e-9)
    assert pri.decrypt(e - b) == pytest.approx(list(a - b), abs=1e-7)
    assert pri.decrypt(e - pub.encrypt(b)) == pytest.approx(list(a - b), abs=1e-7)
    assert pri.decrypt(10 - e) == pytest.approx(list(10 - a), abs=1e-9)
    # broadcast of a length-1 ciphertext, in both orders
    one = pub.encrypt(0.25)
    assert pri.decrypt(e + one) == pytest.approx(list(a + 0.25), abs=1e-9)
    assert pri.decrypt(one + e) == pytest.approx(list(a + 0.25), abs=1e-9)
    with pytest.raises(ValueError):
        e + [1.0, 2.0]
    with pytest.raises(ValueError):
        e + pub.encrypt([1.0, 2.0])


def test_mul_reference_flow(keys):
    pub, pri = keys
    x = np.arange(1, 51) * 0.75
    y = -(np.arange(1, 51) * 1.25)           # negative plaintexts: the ciphertext-inversion branch
    z = np.arange(50) * 3.0
    t = 0.5
    res = (pub.encrypt(x) * y + z) * t
    assert pri.decrypt(res) == pytest.approx(list((x * y + z) * t), rel=1e-9)
    s = pub.encrypt(1234.5)
    want = 1234.5
    for _ in range(10):
        s = s + 5000
        s = s - 0.2
        want = want + 5000 - 0.2
        assert pri.decrypt(s) == pytest.approx(want, abs=1e-6)
    e = pub.encrypt(x)
    assert pri.decrypt(e * 3) == pytest.approx(list(x * 3))
    assert pri.decrypt(-2 * e) == pytest.approx(list(x * -2))
    assert pri.decrypt(e / 4) == pytest.approx(list(x / 4))
    assert pri.decrypt(e * list(y)) == pytest.approx(list(x * y))
    with pytest.raises(ValueError):
        e * [1.0, 2.0]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matmul_reference_flows(keys, seed):
    pub, pri = keys
    rng = np.random.RandomState(seed)
    m, n, k = rng.randint(1, 7, size=3)
    a = rng.rand(m, n) * 20 - 10
    b = rng.rand(n, k) * 20 - 10
    ct_a = pub.encrypt(a.flatten())
    got = np.array(pri.decrypt(ct_a @ b)).reshape(m, k)
    assert np.allclose(got, a @ b)
    ct_b = pub.encrypt(b.flatten())
    got = np.array(pri.decrypt(a.tolist() @ ct_b)).reshape(m, k)      # list on the left, as the reference test
    assert np.allclose(got, a @ b)
    got = np.array(pri.decrypt(a @ ct_b)).reshape(m, k)               # ndarray on the left defers to __rmatmul__
    assert np.allclose(got, a @ b)
    ct_a @= b
    assert np.allclose(np.array(pri.decrypt(ct_a)).reshape(m, k), a @ b)
    v = rng.rand(n)
    assert np.allclose(np.array(pri.decrypt(pub.encrypt(a.flatten()) @ v)).reshape(m), a @ v)
    with pytest.raises(ValueError):
        pub.encrypt(np.arange(7.0)) @ np.ones((3, 2))


def test_sum_mean_dot_slicing_pickle(keys):
    pub, pri = keys
    x = np.array([1.5, -2.25, 1000.125, 3e-3, 7.0])
    e = pub.encrypt(x)
    assert pri.decrypt(e.sum()) == pytest.approx(x.sum())
    assert pri.decrypt(e.mean()) == pytest.approx(x.mean())
    w = np.array([2.0, -1.0, 0.5, 10.0, -3.0])
    assert pri.decrypt(e.dot(w)) == pytest.approx(float(x @ w))
    assert pri.decrypt(pub.encrypt([4]).sum()) == 4
    assert len(e) == e.length() == 5 and len(e[1:3]) == 2
    assert pri.decrypt(e[1:3]) == pytest.approx([-2.25, 1000.125])
    assert pri.decrypt(e[4]) == 7.0
    assert [pri.decrypt(c) for c in e] == pytest.approx(list(x))
    with pytest.raises(IndexError):
        e[5]
    with pytest.raises(IndexError):
        e.exponent(9)
    e2 = pickle.loads(pickle.dumps(e))
    assert _cts(e2) == _cts(e) and e2.exponent() == e.exponent() and pri.decrypt(e2) == pytest.approx(list(x))
    pri2 = pickle.loads(pickle.dumps(pri))
    assert pri2.decrypt(e) == pytest.approx(list(x))
    ints = pub.encrypt([1, -7, 2 ** 60])
    assert pri.decrypt(ints) == [1, -7, 2 ** 60] and all(isinstance(v, int) for v in pri.decrypt(ints))
    assert pri.raw_decrypt(pub.encrypt(5)) == 5
    other_pub, other_pri = PaillierKeypair.generate_keypair(1024, True)
    with pytest.raises(ValueError):
        other_pri.decrypt(e)
    with pytest.raises(ValueError):
        e + other_pub.encrypt(list(x))


def test_bit_parity_with_oracle(bench):
    """Every operator's ciphertext bits against exact integer arithmetic on the reference bench key."""
    pk_o, sk_o, pub, pri = bench
    assert BNUtils.BN2int(pub.pubkey.hs) == pk_o.hs
    n, nsq, max_int = pk_o.n, pk_o.nsquare, pk_o.n // 3 - 1
    xs = [1234.5678 * (i + 11) for i in range(12)] + [-3.5, 0.0, 7, -9]
    enc = [FixedPointNumber.encode(v, n, max_int) for v in xs]
    raw = pub.raw_encrypt(xs)
    assert _cts(raw) == [O.raw_encrypt(pk_o, f.encoding) for f in enc]
    assert raw.exponent() == [f.exponent for f in enc]
    assert pri.decrypt(raw) == pytest.approx(xs)
    assert pri.raw_decrypt(raw) == [f.encoding for f in enc]
    sec = pub.encrypt(xs)
    assert _cts(sec) != _cts(raw) and O.decrypt_batch(sk_o, _cts(sec)) == [f.encoding for f in enc]
    # ct * negative / positive plaintext list: (ct^-1)^(n - pt) resp. ct^pt, exponents add
    ys = [(-1) ** i * 1.3872 * (32768 - i) for i in range(len(xs))]
    yenc = [FixedPointNumber.encode(v, n, max_int) for v in ys]
    prod = raw * ys
    want = []
    for c, f in zip(_cts(raw), yenc):
        want.append(pow(pow(c, -1, nsq), n - f.encoding, nsq) if f.encoding >= n - max_int else pow(c, f.encoding, nsq))
    assert _cts(prod) == want
    assert prod.exponent() == [a.exponent + b.exponent for a, b in zip(enc, yenc)]
    # scalar broadcast, negative scalar
    f = FixedPointNumber.encode(-2.5, n, max_int)
    assert _cts(raw * -2.5) == [pow(pow(c, -1, nsq), n - f.encoding, nsq) for c in _cts(raw)]
    # ct + ct with exponent alignment: the lower-exponent side is raised to 2^delta first
    a, b = raw[0:8], pub.raw_encrypt([0.001 * (i + 1) for i in range(8)])
    s = a + b
    want, wexp = [], []
    for ca, ea, cb, eb in zip(_cts(a), a.exponent(), _cts(b), b.exponent()):
        if ea < eb:
            ca = pow(ca, 1 << (eb - ea), nsq)
        elif eb < ea:
            cb = pow(cb, 1 << (ea - eb), nsq)
        want.append(ca * cb % nsq)
        wexp.append(max(ea, eb))
    assert _cts(s) == want and s.exponent() == wexp
    # sum: product of the aligned ciphertexts
    tot = raw.sum()
    top = max(raw.exponent())
    acc = 1
    for c, e in zip(_cts(raw), raw.exponent()):
        acc = acc * pow(c, 1 << (top - e), nsq) % nsq
    assert _cts(tot) == [acc] and tot.exponent() == [top]
    # apply_obfuscator keeps the plaintext, changes the bits
    before = _cts(raw)
    raw.apply_obfuscator()
    assert _cts(raw) != before and pri.decrypt(raw) == pytest.approx(xs)
    ob = pub.apply_obfuscator(before[0])
    assert O.decrypt_crt(sk_o, BNUtils.BN2int(ob)) == enc[0].encoding


def test_large_batch_vectorised_path(bench):
    pk_o, sk_o, pub, pri = bench
    x = (np.arange(20000) + 11) * 1234.5678      # the reference bench's generator (bench_ipcl_python.py:27)
    y = (32768 - np.arange(20000)) * 1.3872
    e = pub.encrypt(x)
    assert np.allclose(pri.decrypt(e), x, rtol=0, atol=1e-6)
    assert np.allclose(pri.decrypt(e + pub.encrypt(y)), x + y)
    assert np.allclose(pri.decrypt(e * y), x * y)


def test_device_resident_ciphertexts(bench):
    """encrypt / + / * leave their results in HBM; every way of looking at the words (packed, indexing, slices,
    getTexts, rotate, pickle) brings the same bits to the host, and host-built and device-resident operands mix."""
    from pailliercryptolib_python_b200.bindings.ipcl_bindings import ipclCipherText
    pk_o, sk_o, pub, pri = bench
    x = np.arange(1, 41, dtype=np.int64)
    ct = pub.encrypt(x, apply_obfuscator=False)              # deterministic: 1 + m n
    want = [(1 + int(v) * pk_o.n) % pk_o.nsquare for v in x]
    assert _cts(ct) == want                                   # getTexts on a device-resident batch
    s = ct + ct                                               # device + device
    assert _cts(s) == [w * w % pk_o.nsquare for w in want]
    p3 = s * 3                                                # device * plaintext
    want3 = [pow(w * w % pk_o.nsquare, 3, pk_o.nsquare) for w in want]
    packed = p3.packed()
    assert np.array_equal(packed, p3.packed())                # second look: same host copy
    assert [int.from_bytes(r.tobytes(), "little") for r in packed] == want3
    assert BNUtils.BN2int(p3.ciphertext()[7]) == want3[7]
    assert _cts(p3[5:9]) == want3[5:9]
    rot = p3.ciphertext().rotate(3)
    assert [BNUtils.BN2int(b) for b in rot.getTexts()] == want3[3:] + want3[:3]
    back = pickle.loads(pickle.dumps(p3))
    assert _cts(back) == want3
    host_ct = ipclCipherText.from_packed(pub.pubkey, packed)  # host-built operand + device-resident operand
    mixed = host_ct + s.ciphertext()
    assert [BNUtils.BN2int(b) for b in mixed.getTexts()] == [a * b % pk_o.nsquare for a, b in zip(want3, [w * w % pk_o.nsquare for w in want])]
    assert pri.decrypt(p3) == [6 * int(v) for v in x]
    assert pri.decrypt(back) == [6 * int(v) for v in x]


def test_alignment_and_reductions_stay_on_the_device(bench):
    """mixed-exponent add, multiplication by signed plaintexts, sum, dot and @ never bring the
    ciphertext batch to the host -- the results are device resident with no host copy until somebody looks -- and
    they are bit-identical to the Python-int formulas."""
    pk_o, sk_o, pub, pri = bench
    n2 = pk_o.nsquare
    rs = np.random.RandomState(5)
    a = (rs.rand(48) - 0.5) * 10.0 ** rs.randint(-5, 6, size=48)     # exponents differ row by row
    b = rs.randint(-50, 50, size=48)
    ea = pub.encrypt(a, apply_obfuscator=False)
    eb = pub.encrypt(b, apply_obfuscator=False)
    res = {
        "add": ea + eb,
        "add_broadcast": ea + pub.encrypt(0.5, apply_obfuscator=False),
        "mul_signed": ea * (b * 1.5),
        "sum": ea.sum(),
        "dot": ea.dot(b * 0.25),
        "matmul": ea @ rs.rand(8, 3),
        "rmatmul": rs.rand(2, 6).tolist() @ ea,
        "slice": ea[3:20],
    }
    for name, r in res.items():
        c = r.ciphertext()
        assert c.on_device and not c.host_valid, "%s came back through the host" % name
    # values: decrypt and compare with numpy; bits: re-derive two of them with Python ints
    assert pri.decrypt(res["add"]) == pytest.approx(list(a + b), rel=1e-12, abs=1e-12)
    assert pri.decrypt(res["sum"]) == pytest.approx(float(a.sum()), rel=1e-9)
    assert pri.decrypt(res["dot"]) == pytest.approx(float(a @ (b * 0.25)), rel=1e-9)
    cts, ex = _cts(ea), ea.exponent()
    top = max(ex)
    want = 1
    for c, e in zip(cts, ex):
        want = want * pow(c, 1 << (top - e), n2) % n2
    assert _cts(res["sum"]) == [want] and res["sum"].exponent() == [top]
    assert _cts(res["slice"]) == cts[3:20]
    # a small key and exponents further apart than its 32 * n_words bits: the alignment has no exponent-size limit
    # (the reference's modExp has none either; an HE-mul by the plaintext 2^delta would not fit n here)
    small_pub, small_pri = PaillierKeypair.generate_keypair(256, True)
    s = small_pub.encrypt(0.0) + small_pub.encrypt(1e-100)        # exponents 0 and 385: delta = 385 > 256
    assert s.exponent() == [385]
    assert small_pri.decrypt(s) == 1e-100

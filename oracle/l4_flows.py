"""The reference's own test flows (/root/reference/tests/ipcl_python_test.py:21-119: test_add, test_mul, test_matmul,
test_rmatmul, test_imatmul) plus the operator variants of ipcl_python.py:365-410, with seeded inputs and the obfuscator
off, written against the PUBLIC API only -- TEST INFRASTRUCTURE.

`run_flows(api, keyspec)` works on any module that exposes the reference's classes (PaillierPublicKey,
PaillierPrivateKey, PaillierEncryptedNumber, BNUtils): the unmodified reference L4 over mock or real bindings
(oracle/ref_l4.py) and this repo's ipcl_python.  It returns {step name: {count, ct_sha256, ct_first, ct_last, expo, dec}}: a
digest of every ciphertext integer of every intermediate result, its exponents and the decrypted values.
"""
from __future__ import annotations

import hashlib
import pickle

import numpy as np

import paillier_oracle as O


def keyspecs():
    """Named deterministic keys: (n, p, q, bits)."""
    pk1, sk1 = O.seeded_keypair(1024, 7)
    pk2, sk2 = O.bench_keypair()
    return {"seeded1024": (pk1.n, sk1.p, sk1.q, 1024), "bench2048": (pk2.n, sk2.p, sk2.q, 2048)}


def digest(cts):
    """A batch of ciphertext integers as it is kept in the fixture: count, SHA-256 over all of them (bit-exactness of
    every element), and the first and last in full (a readable anchor when something differs)."""
    hexes = ["%x" % c for c in cts]
    return {"count": len(cts), "ct_sha256": hashlib.sha256("\n".join(hexes).encode()).hexdigest(),
            "ct_first": hexes[0] if hexes else "", "ct_last": hexes[-1] if hexes else ""}


class _Rec:
    def __init__(self, api, pri):
        self.api, self.pri, self.out = api, pri, {}

    def __call__(self, name, en):
        cts = [self.api.BNUtils.BN2int(b) for b in en.ciphertextBN()]
        dec = self.pri.decrypt(en)
        if not isinstance(dec, list):
            dec = [dec]
        self.out[name] = {
            **digest(cts),
            "expo": [int(e) for e in en.exponent()],
            "dec": [v if isinstance(v, int) else float(v) for v in dec],
        }
        return en


def flow_add(api, pub, pri, count):
    """tests/ipcl_python_test.py:21-38 (seeded)."""
    rec = _Rec(api, pri)
    rs = np.random.RandomState(11)
    x = np.ones(count) * rs.randint(100)
    y = np.ones(count) * rs.randint(1000)
    z = np.ones(count) * rs.rand()
    t = list(range(count))
    ex, ey = rec("en_x", pub.encrypt(x, apply_obfuscator=False)), rec("en_y", pub.encrypt(y, apply_obfuscator=False))
    ez, et = rec("en_z", pub.encrypt(z, apply_obfuscator=False)), rec("en_t", pub.encrypt(t, apply_obfuscator=False))
    s1 = rec("x+y", ex + ey)
    s2 = rec("x+y+z", s1 + ez)
    rec("x+y+z+t", s2 + et)
    return rec.out


def flow_mul(api, pub, pri, count, loops):
    """tests/ipcl_python_test.py:40-66 (seeded): negative plaintext factors, array and list operands, the scalar
    add / sub loop."""
    rec = _Rec(api, pri)
    rs = np.random.RandomState(12)
    x = np.ones(count) * rs.randint(1, 100)
    y = np.ones(count) * rs.randint(1, 1000) * -1
    z = np.ones(count) * rs.rand()
    t = list(range(count))
    ex = rec("en_x", pub.encrypt(x, apply_obfuscator=False))
    a = rec("x*y", ex * y)
    b = rec("x*y+z", a + z)
    rec("(x*y+z)*t", b * t)
    en = rec("en_9", pub.encrypt(9, apply_obfuscator=False))
    for i in range(loops):
        en = en + 5000
        en = en - 0.2
        rec("loop%d" % i, en)
    return rec.out


def flow_matmul(api, pub, pri, shapes):
    """tests/ipcl_python_test.py:68-119 (seeded shapes and values): ct @ pt, pt @ ct (list on the left, as the reference
    test), ct @= pt; signed entries so that the inversion branch of __matmul runs (ipcl_python.py:851-857)."""
    rec = _Rec(api, pri)
    rs = np.random.RandomState(13)
    for (m, n, k) in shapes:
        x = rs.rand(m, n) - 0.3
        y = rs.rand(n, k) - 0.3
        tag = "%dx%dx%d" % (m, n, k)
        ex = pub.encrypt(x.flatten(), apply_obfuscator=False)
        rec("matmul " + tag, ex @ y)
        rec("matmul1d " + tag, ex @ y[:, 0])
        ey = pub.encrypt(y.flatten(), apply_obfuscator=False)
        rec("rmatmul " + tag, x.tolist() @ ey)
        ex @= y
        rec("imatmul " + tag, ex)
    return rec.out


def flow_ops(api, pub, pri, count):
    """Operator variants: ipcl_python.py:365-410 (__add__/__radd__/__sub__/__rsub__/__rmul__/__truediv__), the
    length-1 broadcast in both orders (:369-375, :599-667), mixed exponents in both directions (:672-741),
    __getitem__ / __iter__ (:348-363), pickling (:281-298), increase_exponent_to (:528-568)."""
    rec = _Rec(api, pri)
    rs = np.random.RandomState(14)
    a = (rs.rand(count) - 0.5) * 10.0 ** rs.randint(-6, 7, size=count)      # exponents all over the place
    b = rs.randint(-1000, 1000, size=count)                                   # ints: exponent 0
    c = (rs.rand(count) - 0.5) * 1e-120                                       # far-away exponents (delta ~ 400)
    ea = rec("en_a", pub.encrypt(a, apply_obfuscator=False))
    eb = rec("en_b", pub.encrypt(b, apply_obfuscator=False))
    ec = rec("en_c", pub.encrypt(c, apply_obfuscator=False))
    rec("a+b", ea + eb)
    rec("b+a", eb + ea)
    rec("a+c", ea + ec)
    rec("a+list", ea + [float(v) for v in b])
    rec("a+array", ea + c)
    rec("a+int", ea + 7)
    rec("float+a", 2.5 + ea)
    rec("a-b", ea - eb)
    rec("a-array", ea - a[::-1].copy())
    rec("a-list", ea - [float(v) for v in a])
    rec("int-a", 10 - ea)
    rec("a*scalar", ea * 3.25)
    rec("a*negscalar", ea * -3)
    rec("scalar*a", -0.125 * ea)
    rec("a*array", ea * c)
    rec("a*intlist", ea * [int(v) for v in b])
    rec("a/scalar", ea / 8.0)
    rec("a/array", ea / (np.abs(a) + 1.0))
    one = rec("en_one", pub.encrypt(0.25, apply_obfuscator=False))
    rec("a+one", ea + one)
    rec("one+a", one + ea)
    big = rec("en_big", pub.encrypt(2.0 ** 70, apply_obfuscator=False))       # smaller exponent than every a[i]
    rec("a+big", ea + big)
    rec("big+a", big + ea)
    rec("a[3]", ea[3])
    rec("a[2:7]", ea[2:7])
    rec("iter", sum((e for e in ea[0:4]), pub.encrypt(0, apply_obfuscator=False)))
    rec("pickle", pickle.loads(pickle.dumps(ea)))
    top = max(ea.exponent()) + 3
    raised = ea.increase_exponent_to(ea.ciphertext(), ea.exponent(), top)
    rec("raised", api.PaillierEncryptedNumber(pub, raised, [top] * len(ea), len(ea)))
    # obfuscated encrypt: r is drawn inside, only the round trip is comparable
    eo = pub.encrypt(a)
    dec = pri.decrypt(eo)
    rec.out["obfuscated roundtrip"] = {**digest([]), "expo": [int(e) for e in eo.exponent()], "dec": [float(v) for v in dec]}
    return rec.out


def run_flows(api, keyname):
    n, p, q, bits = keyspecs()[keyname]
    pub = api.PaillierPublicKey(n, bits, True)
    pri = api.PaillierPrivateKey(pub, p, q)
    small = keyname != "seeded1024"
    out = {}
    out["add"] = flow_add(api, pub, pri, 24 if small else 100)
    out["mul"] = flow_mul(api, pub, pri, 16 if small else 100, 3 if small else 10)
    out["matmul"] = flow_matmul(api, pub, pri, [(2, 3, 2)] if small else [(1, 1, 1), (3, 5, 2), (4, 8, 3), (2, 7, 1), (5, 3, 4)])
    out["ops"] = flow_ops(api, pub, pri, 8 if small else 20)
    return out

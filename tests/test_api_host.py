"""CPU tests of the reference-facing Python surface that need no GPU: the pybind11 module's BigNumber / container /
key objects (construction, accessors, pickling wire formats, error types) and the drop-in `ipcl_python` alias."""
import pickle

import numpy as np
import pytest

import paillier_oracle as O
from pailliercryptolib_python_b200 import BNUtils, PaillierKeypair, PaillierPrivateKey, PaillierPublicKey
from pailliercryptolib_python_b200.bindings.ipcl_bindings import (context, hybridControl, hybridMode, ipclBigNumber,
                                                                  ipclCipherText, ipclPlainText, ipclPrivateKey,
                                                                  ipclPublicKey)


def test_alias_package_exports_reference_names():
    import ipcl_python
    from ipcl_python.bindings.fixedpoint import FixedPointNumber  # noqa: F401
    from ipcl_python.bindings.ipcl_bindings import ipclKeypair  # noqa: F401
    for name in ("PaillierKeypair", "PaillierPublicKey", "PaillierPrivateKey", "PaillierEncryptedNumber", "context",
                 "hybridControl", "hybridMode"):
        assert hasattr(ipcl_python, name)


def test_bignumber_matches_python_ints():
    vals = [0, 1, 2, 255, 2 ** 32 - 1, 2 ** 32, 3 ** 200, 2 ** 2048 - 1]
    for v in vals:
        bn = BNUtils.int2BN(v)
        assert BNUtils.BN2int(bn) == v and str(bn) == str(v)
        assert bn.to_bytes() == O.bn_to_bytes(v)                    # BN2bytes layout (ipcl_bindings.cpp:121-129)
        assert bn.DwordSize() == max(1, (v.bit_length() + 31) // 32)
        assert bn.BitSize() == max(1, v.bit_length())
        n, words = bn.data()
        assert n == len(words) and sum(w << (32 * i) for i, w in enumerate(words)) == v
        assert pickle.loads(pickle.dumps(bn)) == bn
    a, b = BNUtils.int2BN(3 ** 200), BNUtils.int2BN(7 ** 90)
    assert BNUtils.BN2int(a + b) == 3 ** 200 + 7 ** 90 and BNUtils.BN2int(a * b) == 3 ** 200 * 7 ** 90
    assert BNUtils.BN2int(a - b) == 3 ** 200 - 7 ** 90 and str(b - a) == str(7 ** 90 - 3 ** 200)
    assert BNUtils.BN2int(a * 12345) == 3 ** 200 * 12345
    assert (a > b) and (b < a) and (a >= a) and (a <= a) and (a != b) and not (a == b)
    assert a[0] == (3 ** 200) & 0xFFFFFFFF
    with pytest.raises(IndexError):
        a[1000]
    # bytes whose length is not a multiple of 4, list and numpy constructors
    assert BNUtils.BN2int(ipclBigNumber(b"\x01\x02\x03\x04\x05")) == 0x0504030201
    assert BNUtils.BN2int(ipclBigNumber([1, 2])) == 1 + (2 << 32)
    assert BNUtils.BN2int(ipclBigNumber(np.array([5, 0, 7], dtype=np.uint32))) == 5 + (7 << 64)
    assert BNUtils.BN2int(ipclBigNumber.Zero) == 0 and BNUtils.BN2int(ipclBigNumber.Two) == 2
    assert BNUtils.int2Bytes(0x0102) == b"\x02\x01" and BNUtils.bytes2Int(b"\x02\x01") == 0x0102


def test_plaintext_container():
    vals = [5, 2 ** 70 + 3, 0, 99]
    pt = ipclPlainText([BNUtils.int2BN(v) for v in vals])
    assert len(pt) == pt.getSize() == 4
    assert [BNUtils.BN2int(x) for x in pt.getTexts()] == vals
    assert BNUtils.BN2int(pt[1]) == vals[1] and [BNUtils.BN2int(x) for x in pt[1:3].getTexts()] == vals[1:3]
    assert [BNUtils.BN2int(x) for x in pt.rotate(1).getTexts()] == vals[1:] + vals[:1]
    assert pt.getElementVec(1) == [3, 0, 64] and pt.getElementHex(0) == "0x00000005"
    assert pt == ipclPlainText(pt)
    with pytest.raises(RuntimeError):
        pt == ipclPlainText(BNUtils.int2BN(5))
    with pytest.raises(RuntimeError):
        pt[0:4:2]
    with pytest.raises(IndexError):
        pt[7]
    assert [BNUtils.BN2int(x) for x in pickle.loads(pickle.dumps(pt)).getTexts()] == vals
    n, items = pt.__getstate__()                      # wire format (ipcl_bindings_classes.cpp:248-265)
    assert n == 4 and items[1] == O.bn_to_bytes(vals[1])
    packed = ipclPlainText.from_packed(np.arange(6, dtype=np.uint32).reshape(2, 3))
    assert BNUtils.BN2int(packed[1]) == 3 + (4 << 32) + (5 << 64)
    assert packed.to_packed(5).shape == (2, 5)
    assert [BNUtils.BN2int(x) for x in ipclPlainText(np.array([7, 8], dtype=np.uint32)).getTexts()] == [7, 8]
    assert BNUtils.BN2int(ipclPlainText(9)[0]) == 9


def test_keys_construct_and_pickle_without_gpu():
    pk_o, sk_o = O.seeded_keypair(1024, 77)
    pub = PaillierPublicKey(pk_o.n, 1024, True)
    assert pub.n == pk_o.n and pub.max_int == pk_o.n // 3 - 1 and pub.nsquare == pk_o.nsquare
    assert pub.pubkey.length == 1024 and pub.pubkey.isDJN and pub.pubkey.randbits == 512
    hs = BNUtils.BN2int(pub.pubkey.hs)
    assert 1 < hs < pk_o.nsquare
    # hs = (-x^2)^n mod n^2 is an n-th power residue: hs^lambda = 1 (mod n^2), lambda = lcm(p-1, q-1)
    import math
    lam = (sk_o.p - 1) * (sk_o.q - 1) // math.gcd(sk_o.p - 1, sk_o.q - 1)
    assert pow(hs, lam, pk_o.nsquare) == 1
    state = pub.pubkey.__getstate__()                 # (scheme, n, bits, hs, randbits) (ipcl_bindings.cpp:66-81)
    assert state[0] == 1 and state[1] == O.bn_to_bytes(pk_o.n) and state[2] == 1024 and state[4] == 512
    pub2 = pickle.loads(pickle.dumps(pub))
    assert pub2 == pub and BNUtils.BN2int(pub2.pubkey.hs) == hs and hash(pub2) == hash(pub)
    classic = PaillierPublicKey(pk_o.n, 1024, False)
    assert classic.pubkey.__getstate__()[0] == 0 and not classic.pubkey.isDJN
    pri = PaillierPrivateKey(pub, sk_o.q, sk_o.p)     # the key orders p < q itself
    assert BNUtils.BN2int(pri.prikey.p) == sk_o.p and BNUtils.BN2int(pri.prikey.q) == sk_o.q
    pri2 = pickle.loads(pickle.dumps(pri))
    assert pri2 == pri and BNUtils.BN2int(pri2.prikey.n) == pk_o.n
    with pytest.raises(RuntimeError):
        ipclPrivateKey(pub.pubkey, BNUtils.int2BN(sk_o.p), BNUtils.int2BN(sk_o.q + 2))
    with pytest.raises(ValueError):
        PaillierPublicKey("nope")
    with pytest.raises(KeyError):
        PaillierPrivateKey(pub)
    assert ipclPublicKey(BNUtils.int2BN(pk_o.n), 1024) == pub.pubkey


def test_ciphertext_container_and_generated_keys():
    pub, pri = PaillierKeypair.generate_keypair(512, True)
    assert pub.n.bit_length() == 512 and BNUtils.BN2int(pri.prikey.p) * BNUtils.BN2int(pri.prikey.q) == pub.n
    vals = [3, pub.nsquare - 1, 12345]
    ct = ipclCipherText(pub.pubkey, [BNUtils.int2BN(v) for v in vals])
    assert len(ct) == 3 and ct.public_key == pub.pubkey
    assert ct.to_packed().shape == (3, 32)
    assert [BNUtils.BN2int(x) for x in ct.rotate(2).getTexts()] == vals[2:] + vals[:2]
    assert BNUtils.BN2int(ct.getCipherText(1)[0]) == vals[1]
    ct2 = pickle.loads(pickle.dumps(ct))
    assert [BNUtils.BN2int(x) for x in ct2.getTexts()] == vals and ct2.public_key == pub.pubkey
    with pytest.raises(RuntimeError):
        ipclCipherText(pub.pubkey, [BNUtils.int2BN(1 << 1100)])   # does not fit n^2
    with pytest.raises(RuntimeError):
        PaillierKeypair.generate_keypair(100)


def test_context_and_hybrid_names():
    assert context.initializeContext("QAT") in (True, False)
    assert context.isQATRunning() is False and context.isQATActive() is False
    hybridControl.setHybridMode(hybridMode.OPTIMAL)
    assert hybridControl.getHybridMode() == hybridMode.OPTIMAL
    hybridControl.setHybridOff()
    assert hybridControl.getHybridMode() == hybridMode.UNDEFINED
    assert len(hybridMode.__members__) == 13
    context.terminateContext()

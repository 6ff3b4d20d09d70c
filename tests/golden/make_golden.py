"""Generates the committed golden fixtures under tests/golden/.  Run in the authoring container only
(it imports the reference's own Python from /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

fixedpoint.json  -- outputs of the REFERENCE's FixedPointNumber.encode/.decode
                    (/root/reference/src/ipcl_python/bindings/fixedpoint.py:55-115, imported here unmodified)
                    on the reference bench/test input generators (bench/bench_ipcl_python.py:27,37,48-49;
                    tests/ipcl_python_test.py:22-31,41-46) plus edge cases.  Pins oracle.fp_encode/fp_decode and the
                    product's vectorised codec.
paillier_kat.json -- known-answer vectors for encrypt (pinned r) / decrypt / add / mul at 1024, 2048 (the reference
                    bench key, bench/bench_ipcl_python.py:83-96) and 3072 bits, computed with exact Python integers
                    (oracle/paillier_oracle.py).  The reference's arithmetic backend cannot be built offline, so these
                    pin the C oracle and the CUDA path against drift, not against IPCL itself ("parity unpinned").
"""
import importlib.util
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import paillier_oracle as O  # noqa: E402

REF_FP = "/root/reference/src/ipcl_python/bindings/fixedpoint.py"


def load_reference_fixedpoint():
    spec = importlib.util.spec_from_file_location("ref_fixedpoint", REF_FP)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FixedPointNumber


def fixedpoint_fixture():
    FPN = load_reference_fixedpoint()
    pk, _ = O.bench_keypair()
    n, max_int = pk.n, pk.n // 3 - 1
    k = np.arange(24)
    floats = []
    floats += list((k + 11) * 1234.5678)
    floats += list((k + 1) * 1234.5678)
    floats += list((k + 11) * 5111.2834)
    floats += list((32768 - k) * 1.3872)
    floats += [0.0, -0.0, 1.0, -1.0, 0.5, -0.2, 5000.0, 1e-201, -1e-201, 1e-200, 3.141592653589793, -2.718281828459045,
               1e15, -1e15, 1e300, 1.7976931348623157e308, 5e-324, 2.0 ** -1074, 123456789.125, -987654321.0625]
    rng = np.random.RandomState(20240611)
    floats += list(rng.rand(16) * 100 - 50)
    ints = [0, 1, -1, 5000, -5000, 2 ** 31 - 1, -(2 ** 31), 2 ** 53, -(2 ** 53) + 1, 2 ** 62, 10 ** 30, -(10 ** 30), max_int, -max_int]
    np_ints = [np.int32(7), np.int32(-7), np.int64(2 ** 40), np.int64(-(2 ** 40)), np.int16(-300)]
    cases = []
    for v in floats:
        v = float(v)
        f = FPN.encode(v, n, max_int)
        dec = f.decode()
        cases.append({"kind": "float", "value": v.hex(), "encoding": hex(f.encoding), "exponent": f.exponent,
                      "decoded": float(dec).hex() if isinstance(dec, float) else int(dec), "decoded_is_float": isinstance(dec, float)})
    for v in ints:
        f = FPN.encode(v, n, max_int)
        cases.append({"kind": "int", "value": str(v), "encoding": hex(f.encoding), "exponent": f.exponent, "decoded": str(f.decode())})
    for v in np_ints:
        f = FPN.encode(v, n, max_int)
        cases.append({"kind": type(v).__name__, "value": str(int(v)), "encoding": hex(f.encoding), "exponent": f.exponent,
                      "decoded": str(f.decode())})
    errors = []
    for v in (max_int + 1, -(max_int + 1), float(2 ** 1000) * 2.0 ** 23):
        try:
            FPN.encode(v, n, max_int)
            errors.append({"value": str(v) if isinstance(v, int) else float(v).hex(), "raises": None})
        except Exception as e:  # noqa: BLE001
            errors.append({"value": str(v) if isinstance(v, int) else float(v).hex(), "raises": type(e).__name__})
    return {"source": "reference FixedPointNumber (fixedpoint.py:55-115), n = bench key, max_int = n//3 - 1 (ipcl_python.py:76)",
            "n": hex(n), "max_int": hex(max_int), "cases": cases, "errors": errors}


def kat_for(pk, sk, seed, count, r_bits):
    rng = random.Random(seed)
    n, nsq = pk.n, pk.nsquare
    ms = [0, 1, n - 1, n // 3 - 1] + [rng.randrange(n) for _ in range(count - 6)] + [rng.getrandbits(53), rng.getrandbits(20)]
    if pk.djn:
        rs = [0, 1, (1 << r_bits) - 1] + [rng.getrandbits(r_bits) for _ in range(count - 3)]
    else:
        rs = [1, n - 1] + [rng.randrange(1, n) for _ in range(count - 2)]
    cts = O.encrypt_batch(pk, ms, rs)
    raw = O.encrypt_batch(pk, ms, None)
    others = [rng.randrange(1, nsq) for _ in range(4)]
    dec_in = cts + others
    add_b = [cts[(i * 7 + 3) % count] for i in range(count)]
    exps = [0, 1, 2, (1 << 53) - 1, n - 1] + [rng.getrandbits(53) for _ in range(count - 7)] + [rng.getrandbits(pk.bits) % n, 1 << 40]
    h = lambda xs: [hex(x) for x in xs]  # noqa: E731
    return {
        "bits": pk.bits, "djn": pk.djn, "n": hex(n), "p": hex(sk.p), "q": hex(sk.q), "hs": hex(pk.hs), "randbits": pk.randbits,
        "m": h(ms), "r": h(rs), "ct": h(cts), "ct_raw": h(raw),
        "dec_in": h(dec_in), "dec_out": h(O.decrypt_batch(sk, dec_in)),
        "add_b": h(add_b), "add_out": h(O.add_batch(pk, cts, add_b)), "add_bcast_out": h(O.add_batch(pk, cts, add_b[:1])),
        "mul_e": h(exps), "mul_out": h(O.mul_batch(pk, cts, exps)), "mul_bcast_out": h(O.mul_batch(pk, cts, exps[3:4])),
    }


def paillier_fixture():
    out = {"source": "oracle/paillier_oracle.py (exact Python integers); keys: bench key (bench_ipcl_python.py:83-96) and seeded keys",
           "keys": []}
    pk, sk = O.bench_keypair()
    out["keys"].append(kat_for(pk, sk, 1, 12, 1024))
    pk, sk = O.seeded_keypair(1024, 77)
    out["keys"].append(kat_for(pk, sk, 2, 10, 512))
    pk, sk = O.seeded_keypair(3072, 77)
    out["keys"].append(kat_for(pk, sk, 3, 8, 1536))
    pk, sk = O.seeded_keypair(1024, 5, djn=False)
    out["keys"].append(kat_for(pk, sk, 4, 8, 0))
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "fixedpoint.json"), "w") as f:
        json.dump(fixedpoint_fixture(), f, indent=0)
    with open(os.path.join(HERE, "paillier_kat.json"), "w") as f:
        json.dump(paillier_fixture(), f, indent=0)
    print("wrote", os.listdir(HERE))

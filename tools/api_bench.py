#!/usr/bin/env python
"""Per-call wall times of the Python API (the calls a user of the reference makes) on one GPU: encrypt, decrypt, add with
aligned and with mixed exponents, multiplication by positive and mixed-sign plaintexts, sum, dot, matmul.
    python tools/api_bench.py [--count 100000] [--reps 6]
Prints one JSON object; times are host wall clock around each call (the call a user waits for), ms."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipcl_python import PaillierKeypair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--count", type=int, default=100000)
ap.add_argument("--reps", type=int, default=6)
args = ap.parse_args()
N = args.count
pub, pri = PaillierKeypair.generate_keypair(2048, True)
x = (np.arange(N) + 11) * 1234.5678
rs = np.random.RandomState(3)
mixed = (rs.rand(N) - 0.5) * 10.0 ** rs.randint(-4, 5, size=N)      # exponents differ row by row
signed = np.where(np.arange(N) % 2 == 0, -2.5, 3.0)
pub.encrypt(x[:16]); pub.encrypt(x)         # warm: promotes the key to the wide comb table (once)
out = {"count": N, "reps": args.reps}


def timed(name, fn):
    ts = []
    res = None
    for _ in range(args.reps):
        t0 = time.perf_counter()
        res = fn()
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
    out[name] = {"ms_each": ts, "ms_median": float(np.median(ts))}
    return res


ct = timed("encrypt", lambda: pub.encrypt(x))
y = timed("decrypt", lambda: pri.decrypt(ct))
assert np.array_equal(np.asarray(y), x)
ct_m = pub.encrypt(mixed)
timed("add_aligned", lambda: ct + ct)
s = timed("add_mixed_exponents", lambda: ct + ct_m)
timed("mul_scalar", lambda: ct * 2.5)
p = timed("mul_mixed_sign", lambda: ct * signed)
assert np.allclose(np.asarray(pri.decrypt(p), dtype=float), x * signed)
assert np.allclose(np.asarray(pri.decrypt(s), dtype=float), x + mixed, rtol=1e-12)
tot = timed("sum", lambda: ct.sum())
assert abs(pri.decrypt(tot) - x.sum()) <= 1e-6 * abs(x.sum())
timed("dot", lambda: ct.dot(signed))
a64 = pub.encrypt(rs.rand(64 * 64))
b64 = rs.rand(64, 64) - 0.5
mm = timed("matmul_64x64_by_64x64", lambda: a64 @ b64)
out["matmul_on_device"] = bool(mm.ciphertext().on_device and not mm.ciphertext().host_valid)
print(json.dumps(out))

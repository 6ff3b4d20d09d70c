#!/usr/bin/env python
"""Small driver for ncu captures of the n-adic pair engine kernels: one DJN encrypt, one HE mul (53-bit exponents) and one
decrypt of `count` elements under a `bits`-bit key, comb width fixed so that the table build stays short.
  ncu --set full -k regex:k_encrypt_npair ... python tools/npair_profile.py --bits 2048 --count 37888"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import paillier_oracle as O  # noqa: E402  (keys only)
from pailliercryptolib_python_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bits", type=int, default=2048)
ap.add_argument("--count", type=int, default=37888)
ap.add_argument("--comb-bits", type=int, default=12)
args = ap.parse_args()
bits, N = args.bits, args.count
nw = bits // 32
pk_o, sk_o = O.bench_keypair() if bits == 2048 else O.seeded_keypair(bits, 77)
pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
pk.set_comb_bits(args.comb_bits)
sk = capi.PrivKey(pk, sk_o.p, sk_o.q)
rng = np.random.Generator(np.random.PCG64(1))
m = np.zeros((N, nw), dtype=np.uint32)
m[:, :2] = rng.integers(0, 1 << 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
m[:, 1] &= (1 << 21) - 1
r = rng.integers(0, 1 << 32, size=(N, nw // 2), dtype=np.uint64).astype(np.uint32)
ct = pk.encrypt(m, r)
e = m[:, :2].copy()
prod = pk.mul(ct, e)
back = sk.decrypt(ct)
assert np.array_equal(back, m)
print("ok", bits, N, pk.comb_bits)

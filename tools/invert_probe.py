#!/usr/bin/env python
"""Where the time of the batched ciphertext inverse goes: wall time of phe_invert_rows_dev on device-resident rows against
the summed kernel time (phe_timing), for a few batch sizes."""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import paillier_oracle as O  # noqa: E402  (key only)
from pailliercryptolib_python_b200 import capi  # noqa: E402

pk_o, sk_o = O.bench_keypair()
pk = capi.PubKey(pk_o.n, 2048, djn=True, hs=pk_o.hs)
rng = random.Random(1)
base = capi.ints_to_array([rng.randrange(1, pk_o.nsquare) for _ in range(64)], 128)
out = {}
for n in (1000, 10000, 50000, 100000):
    d = torch.from_numpy(base.view(np.int32)).to("cuda:0")[torch.arange(n, device="cuda:0") % 64].contiguous()
    idx = np.arange(n, dtype=np.int64)
    pk.invert_rows_dev(d.data_ptr(), n, idx)
    torch.cuda.synchronize()
    capi.timing_enable(True)
    t0 = time.perf_counter()
    pk.invert_rows_dev(d.data_ptr(), n, idx)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    kt = capi.timing_read()
    capi.timing_enable(False)
    out[n] = {"wall_ms": round(wall, 2), "kernels_ms": {k: (round(v[0], 3), v[1]) for k, v in kt.items() if v[1]}}
print(json.dumps(out))

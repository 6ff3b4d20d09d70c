#!/usr/bin/env python
"""How much of k_dec_pair's time at N = 100 000 is the partial last wave?  One wave is 296 CTAs x 128 lanes = 37 888
ciphertexts per modulus: 100 000 is 2.64 waves, 113 664 exactly 3.  Prints one JSON line.
    python tools/tail_probe.py [N,N,...]        PHE_DEC_SEGMENTS=1: whole work units instead of time-sliced ones"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bench_key, make_workload  # noqa: E402
from pailliercryptolib_python_b200 import capi  # noqa: E402

n, p, q, hs = bench_key()
pk = capi.PubKey(n, 2048, djn=True, hs=hs)
sk = capi.PrivKey(pk, p, q)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
out = {}
sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [37888, 75776, 100000, 113664]
for i, N in enumerate(sizes):
    m_np, r_np = make_workload(N, 5)
    m = torch.from_numpy(m_np.view(np.int32)).to(dev)
    r = torch.from_numpy(r_np.view(np.int32)).to(dev)
    ct = torch.empty((N, 128), dtype=torch.int32, device=dev)
    res = torch.empty((N, 64), dtype=torch.int32, device=dev)
    pk.encrypt_dev(m.data_ptr(), N, r.data_ptr(), 32, ct.data_ptr(), stream)
    sk.decrypt_dev(ct.data_ptr(), N, res.data_ptr(), stream)
    torch.cuda.synchronize()
    capi.timing_enable(True)
    for _ in range(2):
        sk.decrypt_dev(ct.data_ptr(), N, res.data_ptr(), stream)
    torch.cuda.synchronize()
    t = capi.timing_read()
    capi.timing_enable(False)
    assert torch.equal(res, m)
    ms = t["k_dec_pair"][0] / 2
    out["%d:%d" % (i, N)] = {"waves_per_launch": N / 37888.0, "k_dec_pair_ms": ms, "decrypt_per_s": N / (ms * 1e-3)}
print(json.dumps(out))

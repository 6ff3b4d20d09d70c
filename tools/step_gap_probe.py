#!/usr/bin/env python
"""Where does the ~1 ms per step between the kernel times and bench.py's ms_per_step go?  Times 5 device-resident steps
(encrypt_dev + decrypt_dev, 100 000 x 2048 bits) with the per-kernel timing hooks on and off, with and without the L2
flush, and each call alone.  One JSON line."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bench_key, make_workload  # noqa: E402
from pailliercryptolib_python_b200 import capi  # noqa: E402

N = 100000
n, p, q, hs = bench_key()
pk = capi.PubKey(n, 2048, djn=True, hs=hs)
sk = capi.PrivKey(pk, p, q)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
m_np, r_np = make_workload(N, 5)
m = torch.from_numpy(m_np.view(np.int32)).to(dev)
r = torch.from_numpy(r_np.view(np.int32)).to(dev)
ct = torch.empty((N, 128), dtype=torch.int32, device=dev)
res = torch.empty((N, 64), dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def enc():
    pk.encrypt_dev(m.data_ptr(), N, r.data_ptr(), 32, ct.data_ptr(), stream)


def dec():
    sk.decrypt_dev(ct.data_ptr(), N, res.data_ptr(), stream)


def timed(fn, steps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for _ in range(3):
    enc(); dec()
torch.cuda.synchronize()
out = {}
for hooks in (False, True):
    capi.timing_enable(hooks)
    tag = "hooks_on" if hooks else "hooks_off"
    out[tag] = {"step_with_flush": timed(lambda: (flush.zero_(), enc(), dec())), "step_no_flush": timed(lambda: (enc(), dec())),
                "encrypt_only": timed(enc), "decrypt_only": timed(dec), "flush_only": timed(lambda: flush.zero_())}
    if hooks:
        t = capi.timing_read()
        out["kernel_ms_per_launch"] = {k: v[0] / v[1] for k, v in t.items() if v[1]}
    capi.timing_enable(False)
assert torch.equal(res, m)
print(json.dumps(out))

import sys, time, random
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import paillier_oracle as O
from pailliercryptolib_python_b200 import capi
pk_o, sk_o = O.bench_keypair()
t = time.time(); pk = capi.PubKey(pk_o.n, 2048, True, hs=pk_o.hs); sk = capi.PrivKey(pk, sk_o.p, sk_o.q); print('key setup %.3fs' % (time.time() - t))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = torch.Generator(device='cuda'); g.manual_seed(1)
def rnd(shape): return torch.randint(0, 2**31 - 1, shape, device='cuda', dtype=torch.int32, generator=g)
m = torch.zeros((N, 64), dtype=torch.int32, device='cuda'); m[:, :2] = rnd((N, 2))
r = rnd((N, 32))
ct = torch.empty((N, 128), dtype=torch.int32, device='cuda'); out = torch.empty((N, 64), dtype=torch.int32, device='cuda')
ct2 = torch.empty_like(ct)
e53 = torch.zeros((N, 2), dtype=torch.int32, device='cuda'); e53[:, 0] = rnd((N,)); e53[:, 1] = rnd((N,)) & 0x1fffff
def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts); print('%-14s %9.2f ms  %12.0f ops/s' % (name, ms, N / ms * 1e3)); return ms
timeit('encrypt_djn', lambda: pk.encrypt_dev(m.data_ptr(), N, r.data_ptr(), 32, ct.data_ptr()))
timeit('decrypt_crt', lambda: sk.decrypt_dev(ct.data_ptr(), N, out.data_ptr()))
print('roundtrip ok:', bool((out == m).all()))
timeit('add', lambda: pk.add_dev(ct.data_ptr(), N, ct.data_ptr(), N, ct2.data_ptr()))
timeit('mul53', lambda: pk.mul_dev(ct.data_ptr(), N, e53.data_ptr(), 2, N, 53, ct2.data_ptr()))
print('launches', capi.kernel_launches())

import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
from ipcl_python import PaillierKeypair
import bench
pub, pri = PaillierKeypair.generate_keypair(2048, True)
for N in (1000, 100000):
    x = (np.arange(N) + 11) * 1234.5678
    pub.encrypt(x[:16]); pub.encrypt(x)   # warm: the second call promotes the key to the wide comb table (0.2 s, once)
    t0 = time.perf_counter(); ct = pub.encrypt(x); t1 = time.perf_counter(); y = pri.decrypt(ct); t2 = time.perf_counter()
    ct2 = ct + ct; t3 = time.perf_counter(); ct3 = ct * 2.5; t4 = time.perf_counter()
    ok = np.allclose(np.asarray(y, dtype=float), x)
    print(N, 'encrypt %.1f ms  decrypt %.1f ms  add %.1f ms  mul %.1f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), ok)
# negative multipliers: the reference rule inverts the ciphertext first (one batched phe_invert call here)
N = 100000
x = (np.arange(N) + 11) * 1234.5678
ct = pub.encrypt(x)
w = np.where(np.arange(N) % 2 == 0, -2.5, 3.0)
t0 = time.perf_counter(); ct4 = ct * w; t1 = time.perf_counter()
y = pri.decrypt(ct4)
print(N, 'mul by mixed-sign plaintexts %.1f ms' % ((t1 - t0) * 1e3), np.allclose(np.asarray(y, dtype=float), x * w))

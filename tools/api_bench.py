import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
from ipcl_python import PaillierKeypair
import bench
pub, pri = PaillierKeypair.generate_keypair(2048, True)
for N in (1000, 100000):
    x = (np.arange(N) + 11) * 1234.5678
    pub.encrypt(x[:16]); 
    t0 = time.perf_counter(); ct = pub.encrypt(x); t1 = time.perf_counter(); y = pri.decrypt(ct); t2 = time.perf_counter()
    ct2 = ct + ct; t3 = time.perf_counter(); ct3 = ct * 2.5; t4 = time.perf_counter()
    ok = np.allclose(np.asarray(y, dtype=float), x)
    print(N, 'encrypt %.1f ms  decrypt %.1f ms  add %.1f ms  mul %.1f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), ok)

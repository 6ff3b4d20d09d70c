import cProfile, pstats, io, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from ipcl_python import PaillierKeypair
pub, pri = PaillierKeypair.generate_keypair(2048, True)
N = 100000
x = (np.arange(N) + 11) * 1234.5678
ct = pub.encrypt(x[:64]); pri.decrypt(ct)
ct = pub.encrypt(x)
w = np.where(np.arange(N) % 2 == 0, -2.5, 3.0)
ct * w
for name, fn in (("mul mixed-sign", lambda: ct * w), ("mul positive", lambda: ct * 2.5), ("add", lambda: ct + ct)):
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable(); fn(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(16)
    print("=====", name, "%.1f ms" % ((t1 - t0) * 1e3)); print("\n".join(s.getvalue().split("\n")[4:28]))

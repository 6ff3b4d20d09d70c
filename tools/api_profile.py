import cProfile, pstats, io, sys, numpy as np
sys.path.insert(0, '/root/repo')
from ipcl_python import PaillierKeypair
pub, pri = PaillierKeypair.generate_keypair(2048, True)
N = 100000
x = (np.arange(N) + 11) * 1234.5678
ct = pub.encrypt(x[:64]); pri.decrypt(ct)
ct = pub.encrypt(x)
for name, fn in (("encrypt", lambda: pub.encrypt(x)), ("decrypt", lambda: pri.decrypt(ct))):
    pr = cProfile.Profile(); pr.enable(); fn(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
    print("=====", name); print("\n".join(s.getvalue().split("\n")[4:26]))

// p-adic pair engine instantiations: L = limbs of p, q (10: 1024-bit keys, 20: 2048-bit, 30: 3072-bit).
#include "phe_launch.cuh"
namespace phe {
extern const PairOps g_pair_10 = PairLaunch<10>::ops();
extern const PairOps g_pair_20 = PairLaunch<20>::ops();
extern const PairOps g_pair_30 = PairLaunch<30>::ops();
}

// Lane-group shape L=15 limbs/lane, TPI=8 lanes/bignum (capacity 6240 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_15_8 = Launch<15, 8>::ops(); }

// Lane-group shape L=20 limbs/lane, TPI=1 lanes/bignum (capacity 1040 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_20_1 = Launch<20, 1>::ops(); }

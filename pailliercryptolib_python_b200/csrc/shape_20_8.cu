// Lane-group shape L=20 limbs/lane, TPI=8 lanes/bignum (capacity 8320 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_20_8 = Launch<20, 8>::ops(); }

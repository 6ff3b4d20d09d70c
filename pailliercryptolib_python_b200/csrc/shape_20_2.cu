// Lane-group shape L=20 limbs/lane, TPI=2 lanes/bignum (capacity 2080 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_20_2 = Launch<20, 2>::ops(); }

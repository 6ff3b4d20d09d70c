// Lane-group shape L=28 limbs/lane, TPI=8 lanes/bignum (capacity 6272 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_28_8 = Launch<28, 8>::ops(); }

// Lane-group shape L=28 limbs/lane, TPI=2 lanes/bignum (capacity 1568 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_28_2 = Launch<28, 2>::ops(); }

// Lane-group shape L=28 limbs/lane, TPI=4 lanes/bignum (capacity 3136 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_28_4 = Launch<28, 4>::ops(); }

// Lane-group shape L=15 limbs/lane, TPI=4 lanes/bignum (capacity 3120 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_15_4 = Launch<15, 4>::ops(); }

// Lane-group shape L=20 limbs/lane, TPI=4 lanes/bignum (capacity 4160 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_20_4 = Launch<20, 4>::ops(); }

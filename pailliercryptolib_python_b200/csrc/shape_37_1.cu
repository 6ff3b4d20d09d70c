// Lane-group shape L=37 limbs/lane, TPI=1 lanes/bignum (capacity 1036 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_37_1 = Launch<37, 1>::ops(); }

// Lane-group shape L=37 limbs/lane, TPI=4 lanes/bignum (capacity 4144 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_37_4 = Launch<37, 4>::ops(); }

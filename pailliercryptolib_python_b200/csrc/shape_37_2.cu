// Lane-group shape L=37 limbs/lane, TPI=2 lanes/bignum (capacity 2072 bits).
#include "phe_launch.cuh"
namespace phe { extern const ShapeOps g_ops_37_2 = Launch<37, 2>::ops(); }

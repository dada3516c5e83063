// Kernels and launch sequences of controller shape Shape<0, 3, 2, 2> (plant, n_y, n_u, n_controllers).
#include "shape_ops.cuh"
CMPC_DEFINE_SHAPE_OPS(kOps_coop_par, 0, 3, 2, 2)

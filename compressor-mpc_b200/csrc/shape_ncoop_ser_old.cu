// Kernels and launch sequences of controller shape Shape<1, 3, 2, 2> (plant, n_y, n_u, n_controllers).
#include "shape_ops.cuh"
CMPC_DEFINE_SHAPE_OPS(kOps_ncoop_ser_old, 1, 3, 2, 2)

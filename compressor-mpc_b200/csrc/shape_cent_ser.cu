// Kernels and launch sequences of controller shape Shape<1, 4, 4, 1> (plant, n_y, n_u, n_controllers).
#include "shape_ops.cuh"
CMPC_DEFINE_SHAPE_OPS(kOps_cent_ser, 1, 4, 4, 1)

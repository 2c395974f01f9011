// Register-network select kernels, padded sizes 128 (see coord_select_impl.cuh).
#include "coord_select_impl.cuh"
BL_SELECT_LAUNCHER(16)

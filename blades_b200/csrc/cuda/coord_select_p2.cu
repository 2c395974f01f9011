// Register-network select kernels, padded sizes 72, 80 (see coord_select_impl.cuh).
#include "coord_select_impl.cuh"
BL_SELECT_LAUNCHER(9) BL_SELECT_LAUNCHER(10)

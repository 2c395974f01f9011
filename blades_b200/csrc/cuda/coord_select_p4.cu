// Register-network select kernels, padded sizes 104, 112 (see coord_select_impl.cuh).
#include "coord_select_impl.cuh"
BL_SELECT_LAUNCHER(13) BL_SELECT_LAUNCHER(14)

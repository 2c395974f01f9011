// Register-network select kernels, padded sizes 88, 96 (see coord_select_impl.cuh).
#include "coord_select_impl.cuh"
BL_SELECT_LAUNCHER(11) BL_SELECT_LAUNCHER(12)

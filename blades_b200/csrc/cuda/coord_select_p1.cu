// Register-network select kernels, padded sizes 8, 16, 24, 32, 40, 48, 56, 64 (see coord_select_impl.cuh).
#include "coord_select_impl.cuh"
BL_SELECT_LAUNCHER(1) BL_SELECT_LAUNCHER(2) BL_SELECT_LAUNCHER(3) BL_SELECT_LAUNCHER(4) BL_SELECT_LAUNCHER(5) BL_SELECT_LAUNCHER(6) BL_SELECT_LAUNCHER(7) BL_SELECT_LAUNCHER(8)

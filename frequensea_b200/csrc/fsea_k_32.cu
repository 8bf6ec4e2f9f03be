// fsea_k_32.cu -- 32-point instantiations (N = 32 x 1 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(32, 32, 1, 1)

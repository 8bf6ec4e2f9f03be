// fsea_k_64.cu -- 64-point instantiations (N = 8 x 8 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(64, 8, 8, 1)

// fsea_k_2048.cu -- 2048-point instantiations (N = 32 x 8 x 8)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(2048, 32, 8, 8)

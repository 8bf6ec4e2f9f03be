// fsea_k_128.cu -- 128-point instantiations (N = 16 x 8 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(128, 16, 8, 1)

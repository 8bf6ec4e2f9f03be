// fsea_k_1024.cu -- 1024-point instantiations (N = 32 x 32 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(1024, 32, 32, 1)

// fsea_k_512.cu -- 512-point instantiations (N = 32 x 16 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(512, 32, 16, 1)

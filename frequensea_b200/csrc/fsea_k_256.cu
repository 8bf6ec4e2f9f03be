// fsea_k_256.cu -- 256-point instantiations (N = 16 x 16 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(256, 16, 16, 1)

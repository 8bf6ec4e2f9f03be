// fsea_k_16.cu -- 16-point instantiations (N = 16 x 1 x 1)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(16, 16, 1, 1)

// fsea_k_4096.cu -- 4096-point instantiations (N = 32 x 32 x 4)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(4096, 32, 32, 4)

// fsea_k_4096.cu -- 4096-point instantiations (N = 16 x 16 x 16)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(4096, 16, 16, 16)

// fsea_k_8192.cu -- 8192-point instantiations (N = 32 x 16 x 16)
#include "fsea_instantiate.cuh"
FSEA_DEFINE_SIZE(8192, 32, 16, 16)

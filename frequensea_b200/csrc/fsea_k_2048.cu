// fsea_k_2048.cu -- 2048-point instantiations: packed pairs N = 32 x 8 x 8, scalar N = 64 x 32,
// frame per unit 2 x (32 x 32)   (-DFSEA_S2048_R0=.. _R1 _R2 for A/B builds of the packed shape)
#include "fsea_instantiate.cuh"
#ifndef FSEA_S2048_R0
#define FSEA_S2048_R0 32
#define FSEA_S2048_R1 8
#define FSEA_S2048_R2 8
#endif
#define FSEA_DEF(...) FSEA_DEFINE_SIZE3(__VA_ARGS__)
FSEA_DEF(2048, FSEA_S2048_R0, FSEA_S2048_R1, FSEA_S2048_R2, 64, 32, 1, 32, 32, 1)

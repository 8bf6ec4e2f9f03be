// fsea_k_4096.cu -- 4096-point instantiations: packed pairs N = 32 x 32 x 4, scalar N = 64 x 64,
// frame per unit 2 x (16 x 16 x 8)   (-DFSEA_S4096_R0=.. _R1 _R2 for A/B builds of the packed shape, -DFSEA_H4096_R0.. of the frame-per-unit shape)
#include "fsea_instantiate.cuh"
#ifndef FSEA_S4096_R0
#define FSEA_S4096_R0 32
#define FSEA_S4096_R1 32
#define FSEA_S4096_R2 4
#endif
#ifndef FSEA_H4096_R0
#define FSEA_H4096_R0 16
#define FSEA_H4096_R1 16
#define FSEA_H4096_R2 8
#endif
#define FSEA_DEF(...) FSEA_DEFINE_SIZE3(__VA_ARGS__)
FSEA_DEF(4096, FSEA_S4096_R0, FSEA_S4096_R1, FSEA_S4096_R2, 64, 64, 1, FSEA_H4096_R0, FSEA_H4096_R1, FSEA_H4096_R2)

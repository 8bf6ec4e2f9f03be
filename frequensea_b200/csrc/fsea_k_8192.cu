// fsea_k_8192.cu -- 8192-point instantiations: packed pairs N = 32 x 16 x 16, frame per unit
// 2 x (32 x 8 x 16)   (-DFSEA_S8192_R0=.. _R1 _R2 for A/B builds of the packed shape, -DFSEA_H8192_R0.. of the frame-per-unit shape)
#include "fsea_instantiate.cuh"
#ifndef FSEA_S8192_R0
#define FSEA_S8192_R0 32
#define FSEA_S8192_R1 16
#define FSEA_S8192_R2 16
#endif
#ifndef FSEA_H8192_R0
#define FSEA_H8192_R0 32
#define FSEA_H8192_R1 8
#define FSEA_H8192_R2 16
#endif
#define FSEA_DEF(...) FSEA_DEFINE_SIZE3(__VA_ARGS__)
FSEA_DEF(8192, FSEA_S8192_R0, FSEA_S8192_R1, FSEA_S8192_R2, FSEA_S8192_R0, FSEA_S8192_R1, FSEA_S8192_R2, FSEA_H8192_R0, FSEA_H8192_R1, FSEA_H8192_R2)

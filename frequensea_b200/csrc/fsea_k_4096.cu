// fsea_k_4096.cu -- 4096-point instantiations: packed pairs N = 32 x 32 x 4, scalar N = 64 x 64,
// frame per unit 2 x (16 x 16 x 8)   (-DFSEA_S4096_R0=.. _R1 _R2 for A/B builds of the packed shape, -DFSEA_H4096_R0.. of the frame-per-unit shape)
#include "fsea_instantiate.cuh"
#include "fsea_k64.cuh"
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
#define fsea_get_kernel_4096 fsea_get_kernel_4096_shapes
FSEA_DEF(4096, FSEA_S4096_R0, FSEA_S4096_R1, FSEA_S4096_R2, 64, 64, 1, FSEA_H4096_R0, FSEA_H4096_R1, FSEA_H4096_R2)
#undef fsea_get_kernel_4096

bool fsea_get_kernel_4096(int variant, fsea_kernel_info* out) {
    if (variant != FSEA_VAR_F32X2X) return fsea_get_kernel_4096_shapes(variant, out);
    out->fn = (const void*)&fsea_fft4096x_kernel;
    out->threads = FSEA_K64_T * FSEA_K64_G;
    out->groups = FSEA_K64_G;
    out->lanes = 1;
    out->half = 0;
    out->x64 = 1;
    out->smem = FSEA_K64_SMEM;
    out->n = 4096; out->r0 = 64; out->r1 = 64; out->r2 = 1;
    out->tw1 = 0; out->tw2 = 0;
    out->scalar_bytes = 4;
    out->name = "fft<4096=64x64,f32x2 in-thread even/odd, one exchange>";
    return true;
}

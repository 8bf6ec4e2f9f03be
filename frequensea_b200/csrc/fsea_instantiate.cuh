// fsea_instantiate.cuh -- per-length instantiation helper (included by fsea_k_<N>.cu).
#pragma once
#include "fsea_kernels.cuh"
#include "fsea_dispatch.h"

template <class SH, typename V, int FMT, int GOV = 0, int TWSOV = -1>
struct fsea_cfg {
    static const int UNIT_BYTES = 2 * (int)sizeof(V);
    static const int SB = (int)sizeof(typename fsea_vt<V>::scalar);
    static const bool COLS = fsea_cols_geom<FMT>::COLS;
    static const bool STAGED = SH::NPASS > 1 && SB == 4 && (FMT == FSEA_IN_U8 || FMT == FSEA_IN_U8H || COLS);
    static const int RINGB = COLS ? 4 * fsea_cols_geom<FMT>::RND * SH::N : fsea_vt<V>::lanes * 2 * SH::N;     // (COLS: per group, of the block's tile ring)
    static const int XBYTES = SH::NPASS > 1 ? SH::UNITS * (int)sizeof(V) + (STAGED ? RINGB : 0) : 0;   // exchange plane + raw ring per group
    static const int TWBYTES = (SH::TW1 + SH::TW2) * 2 * SB;
    static const bool TWS = TWSOV >= 0 ? (TWSOV != 0) : (SH::NPASS > 1 && TWBYTES <= 16384);
    static const bool TILED = COLS;
    static const int TPB = (TILED && SH::P <= 16) ? 256 : 128;                    // target threads per block
    static const int G_T = (TPB / SH::T) > 1 ? (TPB / SH::T) : 1;
    static const int G_S = XBYTES ? ((96 * 1024) / XBYTES > 1 ? (96 * 1024) / XBYTES : 1) : 1024;
    static const int G_0 = G_T < G_S ? G_T : G_S;
    static const int G_1 = (SH::T > 32 && G_0 > 15) ? 15 : G_0;                 // named barriers 1..15
    static const int G_2 = TILED ? (G_1 >= 16 ? 16 : (G_1 >= 8 ? 8 : (G_1 >= 4 ? 4 : (G_1 >= 2 ? 2 : 1)))) : G_1;   // 2G rows divide 32
    static const int G = GOV > 0 ? GOV : G_2;
    static const size_t SMEM = (size_t)G * XBYTES + (COLS ? 128 : 0) + (TWS ? TWBYTES : 0);
    // resident blocks per SM the register allocator must allow: what shared memory
    // permits, capped so that a thread keeps >= REGCAP registers
    static const int REGCAP_0 = (sizeof(V) >= 8 ? (SH::P >= 32 ? 168 : (SH::P >= 16 ? 112 : 80)) : (SH::P >= 64 ? 168 : (SH::P >= 32 ? 96 : 64)));
    static const int REGCAP = REGCAP_0;
    static const int THREADS = SH::T * G;
    static const int B_SMEM = SMEM ? (int)((227 * 1024) / (SMEM + 1024)) : 16;
    static const int B_REG = 65536 / (THREADS * REGCAP);
    static const int B_THR = 2048 / THREADS;
    static const int MINB_0 = B_SMEM < B_REG ? B_SMEM : B_REG;
    static const int MINB_1 = MINB_0 < B_THR ? MINB_0 : B_THR;
    static const int MINB = MINB_1 < 1 ? 1 : (MINB_1 > 16 ? 16 : MINB_1);
};

template <class SH, typename V, int FMT, int OK = FSEA_OK_SPECTRUM, int NTOT = SH::N, bool WIN = false, int GOV = 0, int TWSOV = -1>
static bool fsea_fill(fsea_kernel_info* o, const char* name) {
    typedef fsea_cfg<SH, V, FMT, GOV, TWSOV> C;
    o->fn = (const void*)&fsea_fft_kernel<SH, V, FMT, OK, NTOT, C::G, C::TWS, C::MINB, WIN>;
    o->threads = SH::T * C::G;
    o->groups = C::G;
    o->lanes = (FMT == FSEA_IN_U8H) ? 1 : fsea_vt<V>::lanes;     // frames per unit
    o->half = (FMT == FSEA_IN_U8H) ? 1 : 0;                       // SH is the shape of the half-length transform
    o->smem = C::SMEM;
    o->n = SH::N; o->r0 = SH::R0; o->r1 = SH::R1; o->r2 = SH::R2;
    o->tw1 = SH::TW1; o->tw2 = SH::TW2;
    o->scalar_bytes = C::SB;
    o->name = name;
    return true;
}

// frame-per-unit variants: groups per block / twiddle tables in shared memory (0 / -1: derived, see fsea_cfg)
#ifndef FSEA_HGOV
#define FSEA_HGOV 0
#endif
#ifndef FSEA_HTWS
#define FSEA_HTWS -1
#endif
#define FSEA_STR2(x) #x
#define FSEA_STR(x) FSEA_STR2(x)
// SH: shape of the packed (f32x2), float2-sample and fp64 variants; SHS: shape of the scalar fp32
// variants (a thread may hold 64 points there: 2048 = 64 x 32 and 4096 = 64 x 64 need ONE exchange)
#define FSEA_SHAPE_NAME(N, R0, R1, R2) FSEA_STR(N) ":" FSEA_STR(R0) "x" FSEA_STR(R1) "x" FSEA_STR(R2)
// SHH: shape of the (N/2)-point transform of the "half" variants (one frame per unit, H0 = 0: none)
#define FSEA_DEFINE_SIZE3(N, R0, R1, R2, S0, S1, S2, H0, H1, H2)                                         \
    bool fsea_get_kernel_##N(int variant, fsea_kernel_info* out) {                                       \
        typedef fsea_shape<N, R0, R1, R2> SH;                                                            \
        typedef fsea_shape<N, S0, S1, S2> SHS;                                                           \
        typedef fsea_shape<(H0 ? N / 2 : 64), (H0 ? H0 : 8), (H0 ? H1 : 8), (H0 ? H2 : 1)> SHH;          \
        switch (variant) {                                                                               \
            case FSEA_VAR_F32X2H:                                                                        \
                if constexpr (H0 == 0) return false; else                                               \
                return fsea_fill<SHH, f2, FSEA_IN_U8H, FSEA_OK_HALF, (H0 ? N / 2 : 64), false, FSEA_HGOV, FSEA_HTWS>(out, "fft<" FSEA_STR(N) "=2x(" FSEA_STR(H0) "x" FSEA_STR(H1) "x" FSEA_STR(H2) "),f32x2 frame/unit>");      \
            case FSEA_VAR_F32X2H_WIN:                                                                    \
                if constexpr (H0 == 0) return false; else                                               \
                return fsea_fill<SHH, f2, FSEA_IN_U8H, FSEA_OK_HALF, N / 2, true, FSEA_HGOV, FSEA_HTWS>(out, "fft<" FSEA_STR(N) "=2x(" FSEA_STR(H0) "x" FSEA_STR(H1) "x" FSEA_STR(H2) "),f32x2 frame/unit,win>");      \
            case FSEA_VAR_F32:                                                                           \
                return fsea_fill<SHS, float, FSEA_IN_U8>(out, "fft<" FSEA_SHAPE_NAME(N, S0, S1, S2) ",f32>");      \
            case FSEA_VAR_F32X2:                                                                         \
                return fsea_fill<SH, f2, FSEA_IN_U8>(out, "fft<" FSEA_SHAPE_NAME(N, R0, R1, R2) ",f32x2>");       \
            case FSEA_VAR_F32_WIN:                                                                       \
                return fsea_fill<SHS, float, FSEA_IN_U8, FSEA_OK_SPECTRUM, N, true>(out, "fft<" FSEA_SHAPE_NAME(N, S0, S1, S2) ",f32,win>");      \
            case FSEA_VAR_F32X2_WIN:                                                                     \
                return fsea_fill<SH, f2, FSEA_IN_U8, FSEA_OK_SPECTRUM, N, true>(out, "fft<" FSEA_SHAPE_NAME(N, R0, R1, R2) ",f32x2,win>");       \
            case FSEA_VAR_F32C:                                                                          \
                return fsea_fill<SH, float, FSEA_IN_F32C>(out, "fft<" FSEA_SHAPE_NAME(N, R0, R1, R2) ",f32c>");   \
            case FSEA_VAR_F64:                                                                           \
                return fsea_fill<SH, double, FSEA_IN_F64C>(out, "fft<" FSEA_SHAPE_NAME(N, R0, R1, R2) ",f64>");   \
        }                                                                                                \
        return false;                                                                                    \
    }
#define FSEA_DEFINE_SIZE2(N, R0, R1, R2, S0, S1, S2) FSEA_DEFINE_SIZE3(N, R0, R1, R2, S0, S1, S2, 0, 0, 0)
#define FSEA_DEFINE_SIZE(N, R0, R1, R2) FSEA_DEFINE_SIZE2(N, R0, R1, R2, R0, R1, R2)

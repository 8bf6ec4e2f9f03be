// fsea_vec.cuh -- the arithmetic vocabulary of the FFT kernels.
//
// V is the element type a kernel computes in:
//   float   one sequence per thread-group, fp32
//   double  one sequence per thread-group, fp64 (NUT_BUFFER_F64 inputs, src/nrf.c:607-609)
//   f2      TWO independent sequences ("a pair": two frames, or two columns of
//           a four-step transform) packed in one 64-bit register pair and
//           driven by Blackwell's packed fp32 instructions FADD2 / FMUL2 /
//           FFMA2 (PTX add/mul/fma.rn.f32x2, sm_100+): one issue slot does
//           the same fp32 operation for both sequences.
// Constants (c) are compile-time scalars and become instruction immediates;
// "s" variants take a run-time scalar (a twiddle) that FFMA2 broadcasts to
// both halves (SASS operand form  R.F32).
#pragma once

#if defined(__CUDACC__) || defined(FSEA_EMU)      // (FSEA_EMU: tests/emu/, the kernel source on a CPU fiber machine)
#include <cuda_runtime.h>
#define FSEA_HD __host__ __device__ __forceinline__
#define FSEA_D __device__ __forceinline__
#else
#include <cmath>
#define FSEA_HD inline
#define FSEA_D inline
#endif

template <typename V> struct fsea_vt;
template <> struct fsea_vt<float>  { typedef float scalar;  static const int lanes = 1; };
template <> struct fsea_vt<double> { typedef double scalar; static const int lanes = 1; };

// ---- scalar fp32 / fp64 ----------------------------------------------------
#define FSEA_SCALAR_OPS(T)                                                        \
    FSEA_HD T vadd(T a, T b) { return a + b; }                                    \
    FSEA_HD T vsub(T a, T b) { return a - b; }                                    \
    FSEA_HD T vmul(T a, T b) { return a * b; }                                    \
    FSEA_HD T vfma(T a, T c, T b) { return a * c + b; }   /* a*c + b */           \
    FSEA_HD T vfnma(T a, T c, T b) { return b - a * c; }  /* b - a*c */           \
    FSEA_HD T vfms(T a, T c, T b) { return a * c - b; }   /* a*c - b */           \
    FSEA_HD T vmuls(T a, T s) { return a * s; }                                   \
    FSEA_HD T vfmas(T a, T s, T b) { return a * s + b; }                          \
    FSEA_HD T vfnmas(T a, T s, T b) { return b - a * s; }                         \
    FSEA_HD T vaddc(T a, T c) { return a + c; }                                   \
    FSEA_HD T vsplat(T s, T) { return s; }
FSEA_SCALAR_OPS(float)
FSEA_SCALAR_OPS(double)
#undef FSEA_SCALAR_OPS

// ---- packed pair of fp32 ---------------------------------------------------
#if defined(__CUDACC__) || defined(FSEA_EMU)
struct f2 {
    float2 v;
};
template <> struct fsea_vt<f2> { typedef float scalar; static const int lanes = 2; };

FSEA_D f2 mk2(float a, float b) { f2 r; r.v = make_float2(a, b); return r; }
FSEA_D float2 fsea_neg2(float2 a) { return make_float2(-a.x, -a.y); }
FSEA_D float2 fsea_bc2(float s) { return make_float2(s, s); }
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
FSEA_D f2 vadd(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
FSEA_D f2 vsub(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, fsea_neg2(b.v)); return r; }
FSEA_D f2 vmul(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
FSEA_D f2 vfma(f2 a, float c, f2 b) { f2 r; r.v = __ffma2_rn(a.v, fsea_bc2(c), b.v); return r; }
FSEA_D f2 vfnma(f2 a, float c, f2 b) { f2 r; r.v = __ffma2_rn(a.v, fsea_bc2(-c), b.v); return r; }
FSEA_D f2 vfms(f2 a, float c, f2 b) { f2 r; r.v = __ffma2_rn(a.v, fsea_bc2(c), fsea_neg2(b.v)); return r; }
FSEA_D f2 vfma2(f2 a, f2 c, f2 b) { f2 r; r.v = __ffma2_rn(a.v, c.v, b.v); return r; }
FSEA_D f2 vmuls(f2 a, float s) { f2 r; r.v = __fmul2_rn(a.v, fsea_bc2(s)); return r; }
#else   // host pass of nvcc (never executed) / pre-sm_100 parse
FSEA_D f2 vadd(f2 a, f2 b) { return mk2(a.v.x + b.v.x, a.v.y + b.v.y); }
FSEA_D f2 vsub(f2 a, f2 b) { return mk2(a.v.x - b.v.x, a.v.y - b.v.y); }
FSEA_D f2 vmul(f2 a, f2 b) { return mk2(a.v.x * b.v.x, a.v.y * b.v.y); }
FSEA_D f2 vfma(f2 a, float c, f2 b) { return mk2(a.v.x * c + b.v.x, a.v.y * c + b.v.y); }
FSEA_D f2 vfnma(f2 a, float c, f2 b) { return mk2(b.v.x - a.v.x * c, b.v.y - a.v.y * c); }
FSEA_D f2 vfms(f2 a, float c, f2 b) { return mk2(a.v.x * c - b.v.x, a.v.y * c - b.v.y); }
FSEA_D f2 vfma2(f2 a, f2 c, f2 b) { return mk2(a.v.x * c.v.x + b.v.x, a.v.y * c.v.y + b.v.y); }
FSEA_D f2 vmuls(f2 a, float s) { return mk2(a.v.x * s, a.v.y * s); }
#endif
FSEA_D f2 vaddc(f2 a, float c) { return vadd(a, mk2(c, c)); }
FSEA_D f2 vfmas(f2 a, float s, f2 b) { return vfma(a, s, b); }
FSEA_D f2 vfnmas(f2 a, float s, f2 b) { return vfnma(a, s, b); }
FSEA_D f2 vsplat(float s, f2) { return mk2(s, s); }
FSEA_D float vfma2(float a, float c, float b) { return a * c + b; }
FSEA_D double vfma2(double a, double c, double b) { return a * c + b; }
#endif  // __CUDACC__ || FSEA_EMU

// complex multiply of the value (re, im) by the run-time scalar twiddle (wr + i wi)
template <typename V, typename S>
FSEA_HD void vcmul(V& re, V& im, S wr, S wi) {
    const V r = vfnmas(im, wi, vmuls(re, wr));   // re*wr - im*wi
    const V i = vfmas(re, wi, vmuls(im, wr));    // re*wi + im*wr
    re = r;
    im = i;
}

// fsea_dispatch.h -- how the engine finds the kernel instantiation for a plan.
// Each fsea_k_<N>.cu translation unit instantiates fsea_fft_kernel for one
// transform length and registers its variants through fsea_get_kernel_<N>().
#pragma once
#include <stddef.h>

enum {
    FSEA_VAR_F32 = 0,    // 8-bit IQ, one frame per group, scalar fp32
    FSEA_VAR_F32X2 = 1,  // 8-bit IQ, frame PAIR per group, packed FADD2/FMUL2/FFMA2
    FSEA_VAR_F32C = 2,   // float2 samples, scalar fp32
    FSEA_VAR_F64 = 3,    // double2 samples, fp64
    FSEA_VAR_F32_WIN = 4,    // as F32 / F32X2 with the Hann multiply compiled in
    FSEA_VAR_F32X2_WIN = 5,
    FSEA_VAR_F32X2H = 6,     // 8-bit IQ, ONE frame per group: lanes = even / odd samples, radix-2 across the lanes at the end
    FSEA_VAR_F32X2H_WIN = 7,
    FSEA_VAR_F32X2X = 8,     // 8-bit IQ, N = 4096 = 64 x 64: ONE frame per group of 64 threads, 64 points per thread (fsea_k64.cuh)
    FSEA_VAR_COUNT = 9
};

struct fsea_kernel_info {
    const void* fn;      // __global__ entry (cudaLaunchKernel)
    int threads;         // block size = T * G
    int groups;          // G
    int lanes;           // sequences per unit (1 or 2)
    size_t smem;         // dynamic shared memory per block
    int n, r0, r1, r2;   // shape
    int tw1, tw2;        // twiddle table entries
    int scalar_bytes;    // 4 or 8
    int half;            // 1: n, r0.. describe the (N/2)-point transform of a "half" kernel
    int x64;             // 1: fsea_fft4096x_kernel (its twiddle table goes where the half kernels keep theirs)
    const char* name;
};

typedef bool (*fsea_get_kernel_fn)(int variant, fsea_kernel_info* out);

#define FSEA_DECL_SIZE(N) bool fsea_get_kernel_##N(int variant, fsea_kernel_info* out);
FSEA_DECL_SIZE(16)
FSEA_DECL_SIZE(32)
FSEA_DECL_SIZE(64)
FSEA_DECL_SIZE(128)
FSEA_DECL_SIZE(256)
FSEA_DECL_SIZE(512)
FSEA_DECL_SIZE(1024)
FSEA_DECL_SIZE(2048)
FSEA_DECL_SIZE(4096)
FSEA_DECL_SIZE(8192)
#undef FSEA_DECL_SIZE

// Four-step plans (N = N1 x N2, N >= 8192, 8-bit input): one cluster kernel per frame batch
// (fsea_fused4_kernel) -- step A transforms the N2 interleaved sub-sequences of a frame, step B
// combs the sub-spectra; the intermediate stays in L2.
struct fsea_kernel_info4 {
    const void* fn;          // __global__ entry
    const void* fn_sums;     // the same kernel leaving per-frame sums of |X| instead of spectra (fsea_kparams::frame_sums), or null
    int cs;                  // blocks per cluster (1: plain launch, one block per frame)
    int ws;                  // > 0: warp-specialised kernel (no L2 intermediate, one block per SM), frames per block iteration
    int tma;                 // 1: the kernel reads a tensor map of the input (its second argument); 2: takes the argument, ignores it
    int threads;             // block size
    size_t smem;             // dynamic shared memory per block
    int n1, n2;              // sub-spectrum length, sub-sequences
    int r0, r1, r2;          // shape of the N1-point transform (twiddle tables)
    int tw1, tw2;
    const char* name;
};
// alt: 0 = the shipped configuration of this length, 1.. = alternatives kept for A/B runs (FSEA_FOUR_ALT)
bool fsea_get_kernel4(int n, bool window, int out_mode, int alt, fsea_kernel_info4* out);
// warp-specialised block-per-frame kernel with the intermediate in shared memory (fsea_ws4.cuh)
bool fsea_get_kernel_ws4(int n, bool window, int out_mode, fsea_kernel_info4* out);
// ... spread over a cluster, intermediate in distributed shared memory, raw tiles by TMA (fsea_wsc.cuh): N = 32768, 65536
bool fsea_get_kernel_wsc(int n, bool window, int out_mode, fsea_kernel_info4* out);

// fsea_kparams.h -- the kernel argument block, shared by the launch code
// (fsea_engine.cu) and the kernels (fsea_kernels.cuh).
#pragma once
#include <stdint.h>

// FSEA_OUT_LOGMAG / PIX: out = mul * log2(|X|^2 + eps) + add.  Default 10 log10(|X|^2 + 1e-20)
// (c/fft-batch.c:88-89); log_mult / power_gain of fsea_config move the three (c/fft.c:191-196).
struct fsea_logp { float mul, eps, add; double dmul, deps, dadd; };   // (d*: the fp64 kernels)

struct fsea_kparams;
// byte offset of output row `f` (frame index) from kparams.out
#if defined(__CUDACC__) || defined(FSEA_EMU)
#define FSEA_KP_HD __host__ __device__ inline
#else
#define FSEA_KP_HD inline
#endif
struct fsea_kparams {
    const void* in;        // raw IQ (8-bit pairs), float2 or double2 samples
    void* out;             // float / uint8 / double rows of N
    long n_seq;            // sequences (frames) in the batch
    long seq_stride;       // complex samples between sequence starts (hop)
    const void* tw1;       // pass-1 twiddles [j-1][k], k < R0            (S2 = float2/double2)
    const void* tw2;       // pass-2 twiddles [j-1][k], k < R0*R1
    const float* window;   // N floats or nullptr
    float pix_scale;       // S of FSEA_OUT_PIX (c/fft-batch.c: 10, c/fft-batch-broad.c: 5)
    fsea_logp lg;          // log-magnitude constants
    int flags;             // FSEA_FLAG_*
    int dc_term;           // 1: add N/2 (1+i) to bin k=0 (uncentred 8-bit input, no patch)
    int out_mode;          // FSEA_OUT_*
    unsigned* ticket;      // 2 counters: work tickets, finished blocks (self re-arming)
    const void* tw4;       // four-step twiddles W_N^{n2 k1}, float2 [N1][N2]
    int n1, n2;            // four-step factors (0 otherwise)
    void* dump;            // N x 8 bytes: where sequences beyond the batch are written
    uint32_t xor_mask;     // 0x8080 for two's-complement input (FSEA_IN_S8), else 0
    const void* twh;       // "half" kernels: W_2N^i, i < M (butterflies of the last pass), float2
    void* scratch;         // fused four-step: resident clusters x N float2 (the L2-resident intermediate)
    int tmap_arg;          // 1: the kernel takes a second argument, the input's tensor map (fsea_ws4 / fsea_wsc kernels)
    long out_group;        // > 0: output rows come in groups of this many frames (a sweep frequency's image), group g starting
    long out_group_stride; //      g * out_group_stride bytes into `out` (images side by side in a stitched canvas); 0: one run of rows
    long out_stride;       // bytes between the starts of consecutive output rows (N x element size when rows are packed;
                           // larger when the rows land inside a wider picture, e.g. a sweep's stitched canvas)
    double* frame_sums;    // kernels built with SUMS (fsea_ws4.cuh): no spectrum is stored; comb warp w of frame f leaves the sum of
    long sums_stride;      //   the |X| it computed at frame_sums[f * sums_stride + sums_offset + w] (the sweep gate,
    long sums_offset;      //   c/fft-batch-broad.c:80-98, without a magnitude row ever reaching memory)
};

FSEA_KP_HD size_t fsea_out_row(const fsea_kparams& kp, long f) {
    if (kp.out_group > 0) {
        const long g = f / kp.out_group;
        return (size_t)g * (size_t)kp.out_group_stride + (size_t)(f - g * kp.out_group) * (size_t)kp.out_stride;
    }
    return (size_t)f * (size_t)kp.out_stride;
}

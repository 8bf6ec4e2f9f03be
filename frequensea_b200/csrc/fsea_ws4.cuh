// fsea_ws4.cuh -- N = 1024 x N2 (N2 = 4, 8, 16: N = 4096, 8192, 16384) with BOTH four-step phases
// inside one block and the intermediate in SHARED memory: a warp-specialised kernel.
//
// Decimation in time as in fsea_fused4_kernel (n = N2 n1 + n2): step A = 1024-point transforms
// of the N2 interleaved sub-sequences of a frame, step B = twiddle + radix-N2 across the
// sub-spectra.  There the intermediate Y[n2][k] made a round trip through L2 (8 + 8 bytes per
// point through the LSU, a release/acquire cluster barrier, 4-byte global stores: ncu showed the
// kernel waiting on exactly those -- STG 10 %, ERRBAR 8 %, UCGABAR 5 %, LDG 6 % of all stall
// samples, nothing saturated).  Here a block of 12 warps owns FPB = 16 / N2 frames at a time:
//
//   warps 0..7  (A)  one unit each = the PAIR of sub-sequences (2u, 2u+1) of one frame, the two
//                    packed lanes: raw words -> radix-32 -> a warp-private exchange plane (no
//                    block barrier) -> twiddle -> radix-32, i.e. the 32 x 32 warp-synchronous
//                    transform of the 1024-point kernel; the finished sub-spectra go to
//                    Y[frame][n2][k] in shared memory.  The unit's raw words (4 bytes per n1,
//                    2 N2 bytes apart in the frame) are loaded by the warp itself straight into
//                    registers (32 LDG.32 gathers; the comb warps have pulled the frame into L2
//                    an iteration earlier, the 8 units of a frame share its lines in L1): no
//                    staging in shared memory -- a cp.async gather costs one shared-memory
//                    wavefront per 128-byte line it touches, 8 x the data (ncu: 29 % of all
//                    wavefronts of the kernel);
//   warps 8..11 (B)  comb: per bin pair (k, k+1) 2 N2 conflict-free LDS.64 of Y, the four-step
//                    twiddles, one packed radix-N2 codelet, magnitude / log / pixel epilogue and
//                    coalesced stores of the spectrum.
//
// A computes frames f+1 while B combs frames f: all 12 warps run at once, as in the 1024-point
// kernel, and the only waits are mbarriers in shared memory.  The comb reads Y in PHASES (BIT
// passes of its 128 threads over the block's bins: the lower and the upper half of the bins for
// N2 = 16, one frame per phase for N2 <= 8), and each phase has its own pair
//   full[p]  (the transform warps whose values the phase reads -> B)
//   free[p]  (B -> those warps; a warp holds its finished unit in registers until the comb has
//             read the previous frames' values of that phase out of Y)
// so a transform warp stores the half the comb reads first while the comb is still busy with the
// other half of the previous frame, and -- waiting on an mbarrier, not a bar.sync -- never waits
// for its SIBLINGS: the transform warps drift apart into different phases (load / arithmetic /
// exchange) the way the independent warps of the 1024-point kernel do.
// Shared memory: 8 exchange planes (66 KB) + Y (128 KB) + the pass-1 twiddle table (8 KB): one
// block per SM, 384 threads at 168 registers.
//
// What replaces what: as fsea_kernels.cuh (src/nrf.c:598-631, c/fft-batch-broad.c:63-118).
#pragma once
#include "fsea_kernels.cuh"

#ifndef FSEA_WS_EARLY
#define FSEA_WS_EARLY 0   /* measured: 20.8 vs 24.0 M frames/s at 16384 -- the 32 extra live registers spill */
#endif
#define FSEA_WS_NA 8
#define FSEA_WS_NB 4
#define FSEA_WS_THREADS ((FSEA_WS_NA + FSEA_WS_NB) * 32)
#define FSEA_WS_PLANE (1056 * 8)
#define FSEA_WS_Y (2 * 16 * 1024 * 4)
#define FSEA_WS_TW (31 * 32 * 8)
#define FSEA_WS_STAGE (2 * 1024 * 16)
#define FSEA_WS_SMEM (FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y + FSEA_WS_TW + 128)       /* + mbarriers */
#define FSEA_WS_SMEM_TMA (FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y + FSEA_WS_STAGE + 128)



// step B loads: one bin pair of all N2 sub-spectra from Y in shared memory (rows of 1024 floats)
template <int N2>
__device__ __forceinline__ void fsea_ws_comb_load(const float* y, int k, f2* re, f2* im) {
#pragma unroll
    for (int j = 0; j < N2; j++) {
        const float2 a = *(const float2*)(y + j * 1024 + k);
        const float2 b = *(const float2*)(y + (N2 + j) * 1024 + k);
        re[j] = mk2(a.x, a.y); im[j] = mk2(b.x, b.y);
    }
}

// ... and of FOUR neighbouring bins k .. k+3 (k a multiple of 4): one LDS.128 per sub-spectrum and plane; the two bin
// pairs land in two packed register sets
template <int N2>
__device__ __forceinline__ void fsea_ws_comb_load4(const float* y, int k, f2* re0, f2* im0, f2* re1, f2* im1) {
#pragma unroll
    for (int j = 0; j < N2; j++) {
        const float4 a = *(const float4*)(y + j * 1024 + k);
        const float4 b = *(const float4*)(y + (N2 + j) * 1024 + k);
        re0[j] = mk2(a.x, a.y); re1[j] = mk2(a.z, a.w);
        im0[j] = mk2(b.x, b.y); im1[j] = mk2(b.z, b.w);
    }
}

// radix-N2 + epilogue of four neighbouring bins: spectrum positions k .. k+3 of every row m leave as ONE 16-byte store
// (float32) or ONE 4-byte store (u8 pixels: c/fft-batch.c:85-92 rows, four bins per store instead of one)
// SUMS (float32 magnitudes only): nothing is stored, the 4 N2 magnitudes are added into `acc` instead
template <int OUT, int N1, int N2, bool SUMS = false>
__device__ __forceinline__ void fsea_ws_comb_finish4(const fsea_kparams& kp, int k, uint8_t* row, bool dc_patch,
                                                     f2* re0, f2* im0, f2* re1, f2* im1, float* acc = nullptr) {
    constexpr int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;
    fsea_dft<N2, f2>::run(re0, im0);
    fsea_dft<N2, f2>::run(re1, im1);
    if (kp.dc_term && k == 0) { re0[0].v.x += (float)(N2 / 2) * (float)N1; im0[0].v.x += (float)(N2 / 2) * (float)N1; }   // N/2 (1+i)
#pragma unroll
    for (int m = 0; m < N2; m++) {
        const f2 p0 = vfma2(re0[m], re0[m], vmul(im0[m], im0[m]));
        const f2 p1 = vfma2(re1[m], re1[m], vmul(im1[m], im1[m]));
        uint8_t* dst = row + ((size_t)k + (size_t)N1 * (m ^ (N2 / 2))) * ESZ;
        __align__(16) unsigned char tmp[16];
        fsea_tile_put<OUT>(tmp, p0, kp.pix_scale, kp.lg);
        fsea_tile_put<OUT>(tmp + 2 * ESZ, p1, kp.pix_scale, kp.lg);
        if constexpr (SUMS) {
            static_assert(!SUMS || OUT == FSEA_OUT_MAG, "sums are sums of magnitudes");
            *acc += (((const float*)tmp)[0] + ((const float*)tmp)[1]) + (((const float*)tmp)[2] + ((const float*)tmp)[3]);
            continue;
        }
        if (m == 0 && dc_patch && k == 0) {                                  // bin 0 is not stored (src/nrf.c:621-622)
            if (ESZ == 4) { ((float*)dst)[1] = ((const float*)tmp)[1]; ((float*)dst)[2] = ((const float*)tmp)[2]; ((float*)dst)[3] = ((const float*)tmp)[3]; }
            else { dst[1] = tmp[1]; dst[2] = tmp[2]; dst[3] = tmp[3]; }
        } else if (ESZ == 4) {
            fsea_st_stream((float4*)dst, *(const float4*)tmp);
        } else {
            *(uchar4*)dst = *(const uchar4*)tmp;
        }
        if (m == N2 - 1 && dc_patch && k == N1 - 4) {                        // bin N-1 also fills the DC slot N/2
            uint8_t* dc = row + (size_t)(N2 / 2) * N1 * ESZ;
            if (ESZ == 4) ((float*)dc)[0] = ((const float*)tmp)[3]; else dc[0] = tmp[3];
        }
    }
}

// cos / sin(2 pi j / 32), j a compile-time constant after unrolling (folded into FFMA immediates)
__device__ __forceinline__ float fsea_cos32(int j) {
    switch (j & 31) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654757f;
        case 5: return 0.55557023301960229f;
        case 6: return 0.38268343236508984f;
        case 7: return 0.19509032201612833f;
        case 8: return 6.123233995736766e-17f;
        case 9: return -0.19509032201612819f;
        case 10: return -0.38268343236508973f;
        case 11: return -0.55557023301960196f;
        case 12: return -0.70710678118654746f;
        case 13: return -0.83146961230254535f;
        case 14: return -0.92387953251128674f;
        case 15: return -0.98078528040323043f;
        case 16: return -1.0f;
        case 17: return -0.98078528040323043f;
        case 18: return -0.92387953251128685f;
        case 19: return -0.83146961230254546f;
        case 20: return -0.70710678118654768f;
        case 21: return -0.55557023301960218f;
        case 22: return -0.38268343236509034f;
        case 23: return -0.19509032201612866f;
        case 24: return -1.8369701987210297e-16f;
        case 25: return 0.1950903220161283f;
        case 26: return 0.38268343236509f;
        case 27: return 0.55557023301960184f;
        case 28: return 0.70710678118654735f;
        case 29: return 0.83146961230254524f;
        case 30: return 0.92387953251128652f;
        case 31: return 0.98078528040323032f;
    }
    return 1.0f;
}
__device__ __forceinline__ float fsea_sin32(int j) { return fsea_cos32(j + 24); }      // sin x = cos(x - pi/2) = cos(x + 3 pi/2)

// TMA (N2 = 16): the frame's raw IQ -- 1024 rows of 32 bytes -- is fetched by cp.async.bulk.tensor into a staging
// buffer [16-byte chunk][n1] a whole iteration ahead (one elected thread, 8 boxes, bytes counted in on an mbarrier)
// instead of being gathered by the transform warps: a 4-byte gather pays one L1 wavefront per 128-byte line it touches
// (2048 per frame, 22 % of this kernel's L1 traffic; staged: 256 in + 1024 out with 4-way conflicts) and 32 loads +
// their address arithmetic per thread.  The 32 KB come out of the pass-1 twiddle table's place (read through L1 instead).
template <int N2, int OUT, bool WIN, bool TMA, bool SUMS = false>
__global__ void __launch_bounds__(FSEA_WS_THREADS, 1)
fsea_ws4_kernel(const fsea_kparams kp, const __grid_constant__ fsea_tmap tm) {
    static_assert(!TMA || N2 >= 8, "a TMA box is at least 16 bytes wide: rows of 2 N2 bytes");
    static_assert(!SUMS || (N2 == 16 && OUT == FSEA_OUT_MAG), "per-frame magnitude sums: one frame per block iteration");
    constexpr int CH = (2 * N2) / 16 > 0 ? (2 * N2) / 16 : 1;     // (TMA) 16-byte chunks per row of the frame
    constexpr int N1 = 1024;
    constexpr int FPB = 16 / N2;                            // frames a block works on at a time
    constexpr int UPF = N2 / 2;                             // units (= transform warps) per frame
    typedef fsea_comb<OUT, N1, N2> COMB;
    constexpr int U = COMB::U > 4 ? 4 : COMB::U;            // bin pairs per comb call (64 points; 32 for N2 = 4: registers)
    constexpr int TB = FSEA_WS_NB * 32;
    constexpr int CPF = (N1 / 2) / U;                       // comb calls per frame
    constexpr int BIT = FPB * CPF / TB;
    static_assert(FPB * UPF == FSEA_WS_NA, "one unit per transform warp");
    static_assert(BIT * TB == FPB * CPF && BIT >= 1, "the comb must divide over its warps");

    FSEA_DYN_SMEM(fsea_smem);
    const int tid = threadIdx.x;
    const int w = tid >> 5, t = tid & 31;
    float* Y = reinterpret_cast<float*>(fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE);       // [frame][re | im][n2][1024]
    float2* stw = reinterpret_cast<float2*>(fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y);      // (!TMA)
    unsigned char* stage = fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y;                        // (TMA) [frame of the group][chunk][n1][16 bytes] = 32 KB
    const long frames = kp.n_seq;
    const long groups = (frames + FPB - 1) / FPB;           // block iterations in the batch
    const long step = gridDim.x;
    const size_t frame_bytes = (size_t)kp.seq_stride * 2;
    const uint32_t x4 = kp.xor_mask | (kp.xor_mask << 16);

    // phases of the comb: BIT passes; phase p reads the values of APP transform warps
    constexpr int NPH = BIT;
    constexpr int APP = (N2 == 16) ? FSEA_WS_NA : FSEA_WS_NA / NPH;          // transform warps whose values a phase reads
    static_assert(N2 == 16 ? NPH == 2 : NPH == FPB, "one phase per frame of the group (N2 <= 8) or per half of the bins (N2 = 16)");
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y + (TMA ? FSEA_WS_STAGE : FSEA_WS_TW));
    unsigned long long* full = bars;                        // [NPH]
    unsigned long long* fre = bars + 4;                     // [NPH]
    unsigned long long *raw_full = bars + 8, *raw_empty = bars + 9;     // (TMA)
    if (!TMA) for (int q = tid; q < 31 * 32; q += FSEA_WS_THREADS) stw[q] = ((const float2*)kp.tw1)[q];
    if (tid == 0) {
#pragma unroll
        for (int p = 0; p < NPH; p++) { fsea_mbar_init(full + p, APP * 32); fsea_mbar_init(fre + p, TB); }
        if (TMA) { fsea_mbar_init(raw_full, 1); fsea_mbar_init(raw_empty, FSEA_WS_NA); }
    }
    __syncthreads();
    auto fetch = [&](long g) {                              // one thread: the raw bytes of group g (frames past the end of the batch read as zeros)
        fsea_mbar_expect_tx(raw_full, FSEA_WS_STAGE);
#pragma unroll
        for (int ff = 0; ff < FPB; ff++)
#pragma unroll
            for (int h = 0; h < CH; h++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    fsea_tma_load_3d(stage + ((ff * CH + h) * 4 + q) * 4096, &tm, 16 * h, 256 * q, (int)(g * FPB + ff), raw_full);
    };

    if (w < FSEA_WS_NA) {
        // ------------------------------------------------ A: transform warps
        const int fi = w / UPF, u = w - fi * UPF;           // frame of the group, unit of the frame
        f2* xp = reinterpret_cast<f2*>(fsea_smem + (size_t)w * FSEA_WS_PLANE);
        float* yre = Y + (size_t)fi * (2 * N2 * 1024) + (2 * u) * 1024 + t;
        float* yim = yre + N2 * 1024;
        auto unit_src = [&](long g) {                       // frames past the end of the batch: transform the last one again (never stored)
            long f = g * FPB + fi;
            if (f >= frames) f = frames - 1;
            return (const uint8_t*)kp.in + (size_t)f * frame_bytes + 4 * u;
        };
        unsigned it = 0;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(stage + (fi * CH + (u >> 2)) * 16384) + (u & 3);   // (TMA) word of row n1 at [4 n1]
        const float2* tw1g = (const float2*)kp.tw1;
        f2 hc = mk2(1.0f, 1.0f), hs = mk2(0.0f, 0.0f);     // (WIN) cos / sin of 2 pi n0 / N for the thread's first sample pair n0 = N2 t + 2u (+1)
        if (WIN) {
            float c0, s0, c1, s1;
            sincospif(2.0f * (float)(N2 * t + 2 * u) / (float)(N1 * N2), &s0, &c0);
            sincospif(2.0f * (float)(N2 * t + 2 * u + 1) / (float)(N1 * N2), &s1, &c1);
            hc = mk2(c0, c1); hs = mk2(s0, s1);
        }
        if (TMA && w == 0 && t == 0 && (long)blockIdx.x < groups) fetch(blockIdx.x);
        uint32_t raw[32];
        auto gather = [&](long g) {
            const uint8_t* src = unit_src(g) + (size_t)t * (2 * N2);
#pragma unroll
            for (int j = 0; j < 32; j++) raw[j] = __ldg((const uint32_t*)(src + (size_t)j * (32 * 2 * N2)));
        };
        constexpr bool EARLY = (N2 == 16) && FSEA_WS_EARLY && !TMA;  // the next frame's gathers are issued between the two halves of the Y store
        if (EARLY && (long)blockIdx.x < groups) gather(blockIdx.x);
        (void)words; (void)tw1g;
        for (long g = blockIdx.x; g < groups; g += step, it++) {
            f2 re[32], im[32];
            if constexpr (TMA) {
                fsea_mbar_wait(raw_full, it & 1);
#pragma unroll
                for (int j = 0; j < 32; j++) raw[j] = words[4 * (t + 32 * j)];
                __syncwarp();
                if (t == 0) fsea_mbar_arrive(raw_empty);               // this warp has taken its words
                if (w == 0 && t == 0 && g + step < groups) {           // the next frame's tile, a whole iteration ahead
                    fsea_mbar_wait(raw_empty, it & 1);
                    fetch(g + step);
                }
                __syncwarp();
            } else if (!EARLY) gather(g);
            if (WIN) {      // the 32 window values of a thread are the same for every frame: recompute them here rather than keep 64 registers
                fsea_opaque(hc.v.x); fsea_opaque(hc.v.y); fsea_opaque(hs.v.x); fsea_opaque(hs.v.y);
            }
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t v = raw[j] ^ x4;
                re[j] = mk2(__uint_as_float(__byte_perm(v, 0x47000000u, 0x7504)), __uint_as_float(__byte_perm(v, 0x47000000u, 0x7524)));
                im[j] = mk2(__uint_as_float(__byte_perm(v, 0x47000000u, 0x7514)), __uint_as_float(__byte_perm(v, 0x47000000u, 0x7534)));
                if (WIN) {
                    // Hann without a table: sample n = N2 (t + 32 j) + 2u (+1), so 2 pi n / N = a + 2 pi j / 32 with a of this thread's
                    // first sample pair: w = 1/2 - 1/2 (cos a cos(2 pi j/32) - sin a sin(2 pi j/32)), the j-part as immediates
                    const f2 cw = vfnma(hs, fsea_sin32(j), vmuls(hc, fsea_cos32(j)));      // cos(a + 2 pi j / 32)
                    const f2 w2 = vfnma(cw, 0.5f, mk2(0.5f, 0.5f));
                    const f2 off = mk2(32896.0f, 32896.0f);
                    re[j] = vmul(vsub(re[j], off), w2);
                    im[j] = vmul(vsub(im[j], off), w2);
                }
            }
            if (WIN) fsea_dft32<f2>(re, im);
            else fsea_dft32_raw<f2>(re, im);
            // exchange (fsea_kernels.cuh): write a' = t 33 + m, read a' = t + 33 j; re plane, then im plane
            {
                f2* wp = xp + t * 33;
#pragma unroll
                for (int m = 0; m < 32; m++) wp[m] = re[m];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j++) re[j] = xp[t + 33 * j];
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 32; m++) wp[m] = im[m];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j++) im[j] = xp[t + 33 * j];
                __syncwarp();
            }
            re[0] = vmuls(re[0], 1.0f / 256.0f);
            im[0] = vmuls(im[0], 1.0f / 256.0f);
#pragma unroll
            for (int j = 1; j < 32; j++) {
                const float2 tw = TMA ? __ldg(tw1g + (j - 1) * 32 + t) : stw[(j - 1) * 32 + t];
                vcmul(re[j], im[j], tw.x, tw.y);
            }
            fsea_dft32<f2>(re, im);
            if constexpr (N2 == 16) {
                // bins t + 32 m: m < 16 is the half the comb reads first
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    fsea_mbar_wait(fre + h, (it & 1) ^ 1);             // the comb has read this half of the previous frame
#pragma unroll
                    for (int m = 16 * h; m < 16 * h + 16; m++) {
                        yre[32 * m] = re[m].v.x;
                        yre[1024 + 32 * m] = re[m].v.y;
                        yim[32 * m] = im[m].v.x;
                        yim[1024 + 32 * m] = im[m].v.y;
                    }
                    fsea_mbar_arrive(full + h);
                    // half of the result registers are free: the next frame's raw words start their trip now, behind the
                    // wait for the comb and the second half of the stores, instead of at the top of the loop
                    if (!TMA && h == 0 && g + step < groups) gather(g + step);
                }
            } else {
                fsea_mbar_wait(fre + fi, (it & 1) ^ 1);                // the comb has read this frame slot of the previous group
#pragma unroll
                for (int m = 0; m < 32; m++) {
                    yre[32 * m] = re[m].v.x;
                    yre[1024 + 32 * m] = re[m].v.y;
                    yim[32 * m] = im[m].v.x;
                    yim[1024 + 32 * m] = im[m].v.y;
                }
                fsea_mbar_arrive(full + fi);
            }
        }
    } else {
        // ------------------------------------------------ B: comb warps
        const int bt = tid - FSEA_WS_NA * 32;
        const bool dc_patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
        long g = blockIdx.x;
        // the raw bytes of a group (FPB frames of 2 N bytes = 256 lines), pulled into L2 two iterations ahead (L1 is the
        // 28 KB left beside 228 KB of shared memory: a frame does not fit)
        auto prefetch_l2 = [&](long gg) {
            if (gg >= groups) return;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int line = bt + TB * q;                              // 256 lines: frame line / (256 / FPB)
                const int pf = line / (256 / FPB), pl = line - pf * (256 / FPB);
                long f = gg * FPB + pf;
                if (f >= frames) f = frames - 1;
                fsea_prefetch_l2((const uint8_t*)kp.in + (size_t)f * frame_bytes + (size_t)pl * 128);
            }
        };
        unsigned gi = 0;
        for (; g < groups; g += step, gi++) {
            if (!TMA) prefetch_l2(g + 2 * step);
            double fsum = 0.0;                                         // (SUMS) this thread's share of sum |X| of frame g
#pragma unroll 1
            for (int it = 0; it < BIT; it++) {
                const int c = it * TB + bt;                            // comb call: frame c / CPF, pairs (c % CPF) + uu CPF
                const int fi = c / CPF, p0 = c - fi * CPF;
                const long f = g * FPB + fi;
                const float* yf = Y + (size_t)fi * (2 * N2 * 1024);
                uint8_t* row = f < frames ? (uint8_t*)kp.out + fsea_out_row(kp, f) : (uint8_t*)kp.dump;
                // a call works on U / 2 groups of FOUR neighbouring bins: 4 (p0 + v CPF) .. + 3
                static_assert(U % 2 == 0, "bin pairs are combed two at a time");
                fsea_mbar_wait(full + it, gi & 1);                     // the values this phase reads are in Y
                f2 re[U][N2], im[U][N2];
#pragma unroll
                for (int v = 0; v < U / 2; v++) fsea_ws_comb_load4<N2>(yf, 4 * (p0 + v * CPF), re[2 * v], im[2 * v], re[2 * v + 1], im[2 * v + 1]);
                // (issuing the twiddle rows' loads BEFORE the wait -- they do not depend on Y -- was measured: 24.4 vs 25.7 M
                // frames/s; the 32 extra live registers cost more than the L2 latency they hide)
#pragma unroll
                for (int uu = 0; uu < U; uu++) COMB::twiddle(kp, 4 * (p0 + (uu / 2) * CPF) + 2 * (uu & 1), re[uu], im[uu]);
                fsea_mbar_arrive(fre + it);                            // this phase's part of Y has been read
                float acc = 0.0f;
#pragma unroll
                for (int v = 0; v < U / 2; v++)
                    fsea_ws_comb_finish4<OUT, N1, N2, SUMS>(kp, 4 * (p0 + v * CPF), row, dc_patch, re[2 * v], im[2 * v], re[2 * v + 1], im[2 * v + 1], &acc);
                if (SUMS) fsum += (double)acc;                         // (float over one call's 4 N2 U/2 magnitudes, double beyond)
            }
            if (SUMS) {                                                // lanes by shuffle in a fixed order, one value per comb warp
                for (int o = 16; o > 0; o >>= 1) fsum += __shfl_down_sync(0xffffffffu, fsum, o);
                if ((bt & 31) == 0 && g < frames) kp.frame_sums[(size_t)g * kp.sums_stride + kp.sums_offset + (bt >> 5)] = fsum;
            }
        }
    }
}

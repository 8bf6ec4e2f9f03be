// fsea_wsc.cuh -- N = 1024 x N2 for N2 = 32, 64 (N = 32768, 65536): the warp-specialised kernel of
// fsea_ws4.cuh spread over a thread-block CLUSTER, the intermediate in DISTRIBUTED shared memory.
//
// A frame no longer fits one SM: its intermediate Y[n2][k] is 8 N bytes = 256 / 512 KB.  A cluster of
// CS = N2 / 16 blocks owns a frame; block c
//   * transforms the 8 unit pairs (sub-sequences 16 c .. 16 c + 15) of the frame with its 8 transform
//     warps, exactly as fsea_ws4_kernel does (32 x 32 warp-synchronous, private exchange plane);
//   * keeps the slice Y[all n2][k in c KL .. (c+1) KL), KL = 1024 / CS, in ITS shared memory (128 KB),
//     so every transform warp of the cluster stores 1 / CS of its bins locally and the rest straight into
//     the other blocks' shared memory (st.shared::cluster, 128-byte rows per warp);
//   * combs its KL bins with its 4 comb warps and stores their N2 KL spectrum points.
// The round trip of the intermediate through L2 of fsea_fused4_kernel (8 + 8 bytes per point through the
// LSU, a release / acquire cluster barrier per frame, 4-byte global stores) is gone; what crosses the SM
// boundary is (CS - 1) / CS of 8 bytes per point over the cluster's DSMEM network.
//
// Raw IQ: a block's tile is a 32-byte column slice of the frame's 1024 rows of 2 N2 bytes.  It is fetched by
// TMA (cp.async.bulk.tensor, one elected thread, 8 boxes of 16 bytes x 256 rows per frame) into a staging
// buffer while the previous frame is being transformed; an mbarrier counts the bytes in.
//
// Synchronisation is all mbarriers (no block or cluster barrier inside the loop, and no fence):
//   RAW_FULL   TMA -> transform warps (bytes counted in)      RAW_EMPTY  transform warps -> the thread that issues the TMA
//   Y_FULL     the sub-spectra are written with st.async, which counts every byte in on the DESTINATION block's barrier:
//              its comb warps wait for the 128 KB of a frame, wherever in the cluster they come from
//   Y_FREE     every comb warp of the cluster -> this block's transform warps (relaxed remote arrives)
#pragma once
#include "fsea_ws4.cuh"

// Y stores: 0 = plain st.shared::cluster (immediate offsets) + one cluster-scope release fence and relaxed arrives per
// thread and frame; 1 = st.async, every byte counted in on the destination's barrier (no fence, but STAS takes its two
// addresses from a register pair: an address computation per store)
#ifndef FSEA_WSC_STAS
#define FSEA_WSC_STAS 0
#endif
#define FSEA_WSC_STAGE (2 * 1024 * 16)                       // [chunk][n1][16 bytes]
#define FSEA_WSC_SMEM (FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y + FSEA_WSC_STAGE + 64)

// a transform thread's 32 bins of both sub-spectra -> Y of the blocks that comb them.  Bins t + 32 m, m in [C MPB, (C+1) MPB),
// belong to block C: ordinary shared-memory stores where that is this block, stores into the peer's shared memory
// otherwise (compile-time recursion over m: every byte offset is an immediate).
template <int N2, int CS, int KL, bool PAIRED, bool LOCAL, int C, int MM>
__device__ __forceinline__ void fsea_ws_store_y_block(fsea_cptr ybase, float* ylocal, fsea_cptr bar, const f2* re, const f2* im) {
    constexpr int MPB = 32 / CS;                                       // values of m per destination block
    if constexpr (MM < MPB) {
        constexpr int M = C * MPB + MM, KOFF = 32 * MM;                // local bin offset from t
        if constexpr (LOCAL) {
            if constexpr (PAIRED) {
                ((float2*)ylocal)[KOFF] = make_float2(re[M].v.x, re[M].v.y);
                ((float2*)ylocal)[(N2 / 2) * KL + KOFF] = make_float2(im[M].v.x, im[M].v.y);
            } else {
                ylocal[KOFF] = re[M].v.x;
                ylocal[KL + KOFF] = re[M].v.y;
                ylocal[N2 * KL + KOFF] = im[M].v.x;
                ylocal[(N2 + 1) * KL + KOFF] = im[M].v.y;
            }
        } else {
#if FSEA_WSC_STAS
            if constexpr (PAIRED) {
                fsea_st_async_at<KOFF * 8>(ybase, make_float2(re[M].v.x, re[M].v.y), bar);
                fsea_st_async_at<((N2 / 2) * KL + KOFF) * 8>(ybase, make_float2(im[M].v.x, im[M].v.y), bar);
            } else {
                fsea_st_async_at<KOFF * 4>(ybase, re[M].v.x, bar);
                fsea_st_async_at<(KL + KOFF) * 4>(ybase, re[M].v.y, bar);
                fsea_st_async_at<(N2 * KL + KOFF) * 4>(ybase, im[M].v.x, bar);
                fsea_st_async_at<((N2 + 1) * KL + KOFF) * 4>(ybase, im[M].v.y, bar);
            }
#else
            if constexpr (PAIRED) {
                fsea_st_cluster_at<KOFF * 8>(ybase, make_float2(re[M].v.x, re[M].v.y));
                fsea_st_cluster_at<((N2 / 2) * KL + KOFF) * 8>(ybase, make_float2(im[M].v.x, im[M].v.y));
            } else {
                fsea_st_cluster_at<KOFF * 4>(ybase, re[M].v.x);
                fsea_st_cluster_at<(KL + KOFF) * 4>(ybase, re[M].v.y);
                fsea_st_cluster_at<(N2 * KL + KOFF) * 4>(ybase, im[M].v.x);
                fsea_st_cluster_at<((N2 + 1) * KL + KOFF) * 4>(ybase, im[M].v.y);
            }
#endif
        }
        fsea_ws_store_y_block<N2, CS, KL, PAIRED, LOCAL, C, MM + 1>(ybase, ylocal, bar, re, im);
    }
}
template <int N2, int CS, int KL, bool PAIRED, int C>
__device__ __forceinline__ void fsea_ws_store_y(const fsea_cptr* ybase, float* ylocal, const fsea_cptr* bar, unsigned crank, const f2* re, const f2* im) {
    if constexpr (C < CS) {
#if FSEA_WSC_STAS
        fsea_ws_store_y_block<N2, CS, KL, PAIRED, false, C, 0>(ybase[C], ylocal, bar[C], re, im);      // (every byte is counted in: local ones too)
#else
        if (crank == (unsigned)C) fsea_ws_store_y_block<N2, CS, KL, PAIRED, true, C, 0>(ybase[C], ylocal, bar[C], re, im);
        else fsea_ws_store_y_block<N2, CS, KL, PAIRED, false, C, 0>(ybase[C], ylocal, bar[C], re, im);
#endif
        fsea_ws_store_y<N2, CS, KL, PAIRED, C + 1>(ybase, ylocal, bar, crank, re, im);
    }
}

template <int N2, int OUT, bool WIN>
__global__ void __launch_bounds__(FSEA_WS_THREADS, 1)
fsea_wsc_kernel(const fsea_kparams kp, const __grid_constant__ fsea_tmap tm) {
    constexpr int N1 = 1024;
    constexpr int CS = N2 / 16;                             // blocks per cluster = per frame
    constexpr int KL = N1 / CS;                             // bins of every sub-spectrum this block combs
    constexpr bool PAIRED = (N2 == 64);                     // Y rows of a unit interleaved (fsea_out_cplx): the comb's lanes are (2j, 2j+1)
    typedef fsea_comb<OUT, N1, N2> COMB;
    constexpr int TB = FSEA_WS_NB * 32;
    constexpr int CALLS = PAIRED ? KL : (KL / 2) / COMB::U;  // comb calls per block and frame (64 points each)
    constexpr int BIT = CALLS / TB;
    static_assert(BIT * TB == CALLS && BIT >= 1, "the comb must divide over its warps");
    constexpr int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;

    FSEA_DYN_SMEM(fsea_smem);
    const int tid = threadIdx.x;
    const int w = tid >> 5, t = tid & 31;
    float* Y = reinterpret_cast<float*>(fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE);       // [re | im][n2][KL]  (PAIRED: [re | im][n2 / 2][KL][2])
    unsigned char* stage = fsea_smem + FSEA_WS_NA * FSEA_WS_PLANE + FSEA_WS_Y;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage + FSEA_WSC_STAGE);
    unsigned long long *raw_full = bars, *raw_empty = bars + 1, *y_full = bars + 2, *y_free = bars + 3;
    const unsigned crank = fsea_cluster_ctarank();
    const long cid = (long)fsea_cluster_id(), step = (long)fsea_cluster_count();
    const long frames = kp.n_seq;
    const uint32_t x4 = kp.xor_mask | (kp.xor_mask << 16);

    if (tid == 0) {
        fsea_mbar_init(raw_full, 1);                                   // the issuing thread's expect_tx arrive
        fsea_mbar_init(raw_empty, FSEA_WS_NA);                         // one lane per transform warp
#if FSEA_WSC_STAS
        fsea_mbar_init(y_full, 1);                                     // one arming arrive per frame + the frame's 128 KB (st.async)
#else
        fsea_mbar_init(y_full, CS * FSEA_WS_NA * 32);                  // every transform thread of the cluster
#endif
        fsea_mbar_init(y_free, CS * FSEA_WS_NB * 32);                  // every comb thread of the cluster
        fsea_mbar_init_fence();
    }
    fsea_cluster_sync_all();                                           // barriers exist before anyone arrives on them

    auto fetch = [&](long f) {                                         // one thread: the block's tile of frame f
        fsea_mbar_expect_tx(raw_full, FSEA_WSC_STAGE);
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int q = 0; q < 4; q++)
                fsea_tma_load_3d(stage + h * 16384 + q * 4096, &tm, 32 * (int)crank + 16 * h, 256 * q, (int)f, raw_full);
    };

    if (w < FSEA_WS_NA) {
        // ------------------------------------------------ A: transform warps
        const int u = w;                                               // unit of the block: sub-sequences 16 crank + 2u, + 1
        f2* xp = reinterpret_cast<f2*>(fsea_smem + (size_t)w * FSEA_WS_PLANE);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(stage + (u >> 2) * 16384) + (u & 3);    // word of row n1 at [4 n1]
        const float2* tw1 = (const float2*)kp.tw1;
        const int gu = (int)crank * FSEA_WS_NA + u;                    // unit of the frame
        // where this thread's bins go: bin k = t + 32 m belongs to block k / KL, local bin k % KL.  One base address per
        // destination block (this thread's first element there); everything else is an immediate of the store
        fsea_cptr ybase[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) ybase[c] = fsea_mapa(Y, (unsigned)c) + (PAIRED ? ((gu * KL + t) * 2) * 4 : ((2 * gu) * KL + t) * 4);
        float* ylocal = Y + (PAIRED ? (gu * KL + t) * 2 : (2 * gu) * KL + t);     // the same element in this block's own Y
        fsea_cptr full_dst[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) full_dst[c] = fsea_mapa(y_full, (unsigned)c);
        if (w == 0 && t == 0 && cid < frames) fetch(cid);
        unsigned it = 0;
        for (long f = cid; f < frames; f += step, it++) {
            f2 re[32], im[32];
            fsea_mbar_wait(raw_full, it & 1);
            const float* wsrc = WIN ? kp.window + (size_t)N2 * t + 2 * gu : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t v = words[4 * (t + 32 * j)] ^ x4;
                re[j] = mk2(__uint_as_float(__byte_perm(v, 0x47000000u, 0x7504)), __uint_as_float(__byte_perm(v, 0x47000000u, 0x7524)));
                im[j] = mk2(__uint_as_float(__byte_perm(v, 0x47000000u, 0x7514)), __uint_as_float(__byte_perm(v, 0x47000000u, 0x7534)));
                if (WIN) {
                    const float2 wv = __ldg((const float2*)(wsrc + (size_t)N2 * 32 * j));
                    const f2 w2 = mk2(wv.x, wv.y), off = mk2(32896.0f, 32896.0f);
                    re[j] = vmul(vsub(re[j], off), w2);
                    im[j] = vmul(vsub(im[j], off), w2);
                }
            }
            __syncwarp();
            if (t == 0) fsea_mbar_arrive(raw_empty);                   // this warp has taken its words
            if (w == 0 && t == 0 && f + step < frames) {               // the tile of the next frame, a whole iteration ahead
                fsea_mbar_wait(raw_empty, it & 1);
                fetch(f + step);
            }
            __syncwarp();
            if (WIN) fsea_dft32<f2>(re, im);
            else fsea_dft32_raw<f2>(re, im);
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
                const float2 tw = __ldg(tw1 + (j - 1) * 32 + t);
                vcmul(re[j], im[j], tw.x, tw.y);
            }
            fsea_dft32<f2>(re, im);
            fsea_mbar_wait(y_free, (it & 1) ^ 1);                      // every comb warp of the cluster has read the previous frame
            fsea_ws_store_y<N2, CS, KL, PAIRED, 0>(ybase, ylocal, full_dst, crank, re, im);
#if !FSEA_WSC_STAS
            fsea_fence_cluster();                                      // ONE release for all of this thread's stores ...
#pragma unroll
            for (int c = 0; c < CS; c++) fsea_mbar_arrive_cluster_relaxed(full_dst[c]);   // ... then a plain signal to every consumer
#endif
        }
    } else {
        // ------------------------------------------------ B: comb warps
        const int bt = tid - FSEA_WS_NA * 32;
        const bool dc_patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
        fsea_cptr free_dst[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) free_dst[c] = fsea_mapa(y_free, (unsigned)c);
        unsigned it = 0;
#if FSEA_WSC_STAS
        if (bt == 0 && cid < frames) fsea_mbar_expect_tx(y_full, FSEA_WS_Y);      // arm: the first frame's 128 KB
#endif
        for (long f = cid; f < frames; f += step, it++) {
            fsea_mbar_wait_cluster(y_full, it & 1);                    // this frame's slice of Y is complete, wherever in the cluster it came from
#if FSEA_WSC_STAS
            if (bt == 0 && f + step < frames) fsea_mbar_expect_tx(y_full, FSEA_WS_Y);   // arm the next frame (its bytes cannot come before Y_FREE)
#endif
            uint8_t* row = (uint8_t*)kp.out + fsea_out_row(kp, f);
#pragma unroll 1
            for (int bi = 0; bi < BIT; bi++) {
                const int c = bi * TB + bt;                            // comb call of this block
                if constexpr (PAIRED) {
                    // one bin k of all 64 sub-spectra; lanes = sub-spectra (2j, 2j+1)  (fsea_comb<.., 64>)
                    const int kl = c, k = (int)crank * KL + kl;
                    f2 re[32], im[32];
                    const float2* y = (const float2*)Y + kl;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float2 a = y[j * KL], b = y[(32 + j) * KL];
                        re[j] = mk2(a.x, a.y); im[j] = mk2(b.x, b.y);
                    }
                    fsea_comb64_twiddle(kp, N1, k, re, im);
                    if (bi == BIT - 1) {
#pragma unroll
                        for (int cc = 0; cc < CS; cc++) fsea_mbar_arrive_cluster_relaxed(free_dst[cc]);
                    }
                    fsea_comb64_finish<OUT, N1>(kp, k, row, dc_patch, re, im);
                } else {
                    constexpr int U = COMB::U;
                    const int p0 = c;                                  // local pairs p0 + uu CALLS
                    f2 re[U][N2], im[U][N2];
#pragma unroll
                    for (int uu = 0; uu < U; uu++) {
                        const int kl = 2 * (p0 + uu * CALLS);
#pragma unroll
                        for (int j = 0; j < N2; j++) {
                            const float2 a = *(const float2*)(Y + j * KL + kl);
                            const float2 b = *(const float2*)(Y + (N2 + j) * KL + kl);
                            re[uu][j] = mk2(a.x, a.y); im[uu][j] = mk2(b.x, b.y);
                        }
                    }
#pragma unroll
                    for (int uu = 0; uu < U; uu++) COMB::twiddle(kp, (int)crank * KL + 2 * (p0 + uu * CALLS), re[uu], im[uu]);
                    if (bi == BIT - 1) {
#pragma unroll
                        for (int cc = 0; cc < CS; cc++) fsea_mbar_arrive_cluster_relaxed(free_dst[cc]);
                    }
#pragma unroll
                    for (int uu = 0; uu < U; uu++) COMB::finish(kp, (int)crank * KL + 2 * (p0 + uu * CALLS), row, dc_patch, re[uu], im[uu]);
                }
            }
        }
    }
    fsea_cluster_sync_all();                                           // nobody leaves while a peer may still store into its memory
    (void)ESZ;
}

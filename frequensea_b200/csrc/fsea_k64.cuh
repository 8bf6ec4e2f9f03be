// fsea_k64.cuh -- N = 4096 = 64 x 64 with ONE shared-memory exchange (8-bit input, no window).
//
// A thread holds 64 points of ONE frame in 32 packed registers: lane x = even, lane y = odd index
// of its radix-64 butterfly.  A 64-point DFT is then the packed 32-point codelet (both lanes, same
// constants) followed by the radix-2 across the lanes (fsea_cross<32>: X[m] = A_m + W64^m B_m in
// lane x, X[m + 32] = A_m - W64^m B_m in lane y) -- the arrangement fsea_comb<.., 64> uses for the
// four-step comb.  Two such passes with one exchange between them replace the three packed passes
// and two exchanges of the frame-per-unit shape 2 x (16 x 16 x 8), whose shared-memory pipe is the
// binding resource.
//
//   pass 0  thread t < 64: samples n = t + 64 j, j < 64 (lanes: j = 2e, 2e + 1 -- two 16-bit words
//           of the raw ring), raw radix-32 codelet + cross  ->  Y_t[k0], k0 = m (x), m + 32 (y)
//   exchange  plane[k0][t] (row stride 66 floats): 4-byte stores (lanes of a register go to rows m and
//           m + 32), 8-byte loads of (t = 2e, 2e + 1); re plane, then im plane
//   pass 1  thread k0 < 64: twiddle W_4096^{k0 t} / 256 (packed table [e][k0], through L1), radix-32
//           codelet + cross  ->  X[k0 + 64 k1], k1 = m (x), m + 32 (y); |X|^2, stored at
//           k0 + 64 (k1 ^ 32)  (the reference's (-1)^n, src/nrf.c:611-612)
// Block = G groups of 64 threads, one frame per group at a time, the next frame's bytes arriving by
// cp.async while this one is transformed; frames are handed out by the plan's ticket counter.
#pragma once
#include "fsea_kernels.cuh"

#define FSEA_K64_N 4096
#define FSEA_K64_T 64
#define FSEA_K64_G 2
#define FSEA_K64_ROW 66                                   // floats per exchange row (64 + 2: 8-byte loads conflict-free)
#define FSEA_K64_XB (64 * FSEA_K64_ROW * 4)               // exchange plane, bytes
#define FSEA_K64_RING (2 * FSEA_K64_N)                    // raw frame, bytes
#define FSEA_K64_SMEM (FSEA_K64_G * (FSEA_K64_XB + FSEA_K64_RING))

template <int OUT>
__device__ __forceinline__ void fsea_k64_put(uint8_t* row, int pos, float pwr, float pix_scale, const fsea_logp& lg) {
    if (OUT == FSEA_OUT_MAG) ((float*)row)[pos] = fsea_sqrt_approx(pwr);
    else if (OUT == FSEA_OUT_LOGMAG) ((float*)row)[pos] = fsea_db(pwr, lg);
    else row[pos] = fsea_pix(fsea_db(pwr, lg), pix_scale);
}

template <int OUT>
__device__ __forceinline__ void fsea_k64_store(const fsea_kparams& kp, uint8_t* row, int k0, const f2* re, const f2* im, bool dc_patch) {
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const f2 pw = vfma2(re[m], re[m], vmul(im[m], im[m]));
        // lane x: bin k0 + 64 m -> position k0 + 64 (m + 32); lane y: bin k0 + 64 (m + 32) -> position k0 + 64 m
        if (!(m == 0 && dc_patch && k0 == 0)) fsea_k64_put<OUT>(row, k0 + 64 * (m + 32), pw.v.x, kp.pix_scale, kp.lg);   // bin 0 is not stored (src/nrf.c:621-622)
        fsea_k64_put<OUT>(row, k0 + 64 * m, pw.v.y, kp.pix_scale, kp.lg);
        if (m == 31 && dc_patch && k0 == 63) fsea_k64_put<OUT>(row, FSEA_K64_N / 2, pw.v.y, kp.pix_scale, kp.lg);      // bin N-1 also fills the DC slot
    }
}

__global__ void __launch_bounds__(FSEA_K64_T* FSEA_K64_G, 3)
fsea_fft4096x_kernel(const fsea_kparams kp) {
    constexpr int N = FSEA_K64_N, T = FSEA_K64_T, G = FSEA_K64_G, ROW = FSEA_K64_ROW;
    FSEA_DYN_SMEM(fsea_smem);
    const int tid = threadIdx.x;
    const int g = tid / T;
    const int t = tid - g * T;
    float* plane = reinterpret_cast<float*>(fsea_smem + (size_t)g * FSEA_K64_XB);
    unsigned char* ring = fsea_smem + (size_t)G * FSEA_K64_XB + (size_t)g * FSEA_K64_RING;
    const float2* twc = (const float2*)kp.twh;            // [e][k0]: cos of W^{k0 (2e)}, W^{k0 (2e+1)} (x 1/256)
    const float2* tws = twc + 32 * 64;                    //          sin
    const bool dc_patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
    const uint32_t x4 = kp.xor_mask | (kp.xor_mask << 16);
    const long frames = kp.n_seq;
    const size_t in_stride = (size_t)kp.seq_stride * 2;
    const int esz = (kp.out_mode == FSEA_OUT_PIX) ? 1 : 4;

    // Frames are handed out by the plan's ticket counter (fsea_kernels.cuh: the SM's warp arbiter is
    // not fair; a static split left 9.8 of 12 warps resident): the first one per group is fixed, every
    // further one is first + tickets taken.  The leader takes a ticket at the top of an iteration and
    // publishes it after the iteration's second barrier; the group reads it behind the next top barrier.
    FSEA_STATIC_SMEM(unsigned, slot, FSEA_K64_G);
    unsigned* ticket = kp.ticket;
    const long base = (long)gridDim.x * G;
    long f = (long)blockIdx.x * G + g;
    if (t == 0) slot[g] = fsea_take_ticket(ticket);
    if (f < frames) fsea_async_frame<N, T>((const uint8_t*)kp.in + (size_t)f * in_stride, ring, t);
    fsea_cp_async_commit();
    fsea_group_sync<T>(g);
    long fnext = base + (long)slot[g];
    while (f < frames) {
        f2 re[32], im[32];
        unsigned tk = 0;
        fsea_cp_async_wait_all();
        fsea_group_sync<T>(g);                                         // (also: everyone has read slot[g])
        if (t == 0) tk = fsea_take_ticket(ticket);
        // ---- pass 0: samples t + 128 e (lane x) and t + 128 e + 64 (lane y)
        {
            const uint16_t* rw = reinterpret_cast<const uint16_t*>(ring);
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const uint32_t a = rw[t + 128 * e], b = rw[t + 128 * e + 64];
                fsea_in<FSEA_IN_U8H, f2>::unpack((a | (b << 16)) ^ x4, re[e], im[e]);
            }
        }
        fsea_group_sync<T>(g);                                         // ring free: fetch the next frame
        if (fnext < frames) fsea_async_frame<N, T>((const uint8_t*)kp.in + (size_t)fnext * in_stride, ring, t);
        fsea_cp_async_commit();
        fsea_dft32_raw<f2>(re, im);
        fsea_cross<32>(re, im);                                        // Y_t[m] (x), Y_t[m + 32] (y)
        // ---- exchange (re plane, then im plane) and the twiddles of pass 1
#pragma unroll
        for (int m = 0; m < 32; m++) { plane[m * ROW + t] = re[m].v.x; plane[(m + 32) * ROW + t] = re[m].v.y; }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int e = 0; e < 32; e++) { const float2 v = *reinterpret_cast<const float2*>(plane + t * ROW + 2 * e); re[e] = mk2(v.x, v.y); }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int m = 0; m < 32; m++) { plane[m * ROW + t] = im[m].v.x; plane[(m + 32) * ROW + t] = im[m].v.y; }
        if (t == 0) slot[g] = tk;                                      // the ticket has had two barriers and a codelet to arrive
        fsea_group_sync<T>(g);
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const float2 v = *reinterpret_cast<const float2*>(plane + t * ROW + 2 * e);
            const float2 c = __ldg(twc + e * 64 + t), s = __ldg(tws + e * 64 + t);
            const f2 xi = mk2(v.x, v.y), wc = mk2(c.x, c.y), ws = mk2(s.x, s.y);
            const f2 xr = re[e];
            re[e] = vsub(vmul(xr, wc), vmul(xi, ws));                  // (xr + i xi)(wc + i ws)
            im[e] = vfma2(xr, ws, vmul(xi, wc));
        }
        const long fafter = base + (long)slot[g];
        // ---- pass 1 (this thread is k0 = t)
        fsea_dft32<f2>(re, im);
        fsea_cross<32>(re, im);                                        // X[t + 64 m] (x), X[t + 64 (m + 32)] (y)
        if (kp.dc_term && t == 0) { re[0].v.x += (float)(N / 2); im[0].v.x += (float)(N / 2); }   // N/2 (1+i) in FFT bin 0
        uint8_t* row = (uint8_t*)kp.out + fsea_out_row(kp, f);
        if (kp.out_mode == FSEA_OUT_MAG) fsea_k64_store<FSEA_OUT_MAG>(kp, row, t, re, im, dc_patch);
        else if (kp.out_mode == FSEA_OUT_LOGMAG) fsea_k64_store<FSEA_OUT_LOGMAG>(kp, row, t, re, im, dc_patch);
        else fsea_k64_store<FSEA_OUT_PIX>(kp, row, t, re, im, dc_patch);
        // (the next iteration's first barrier also orders this frame's im-plane reads before the next writes)
        f = fnext;
        fnext = fafter;
    }
    fsea_cp_async_wait_all();
    // last block out re-arms the counters for the next launch that uses this slot
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ticket + 1, 1u) == gridDim.x - 1) {
            ticket[0] = 0u;
            ticket[1] = 0u;
            __threadfence();
        }
    }
}

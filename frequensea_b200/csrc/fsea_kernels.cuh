// fsea_kernels.cuh -- the sm_100a IQ -> spectrum kernels.
//
// One kernel template serves every transform length: a mixed-radix Stockham
// FFT whose butterflies are the straight-line register codelets of
// fsea_codelets.cuh (radix 8/16/32, 6-op "tangent" butterflies) and whose
// only inter-thread traffic is one shared-memory exchange per pass boundary.
//
//   N = R0 * R1 * R2   (R2 = 1: two passes, R1 = R2 = 1: one pass)
//   P = max radix      points a thread holds in registers
//   T = N / P          threads that cooperate on one sequence ("a group")
//
// A "unit" of work is one sequence (V = float / double) or a PAIR of
// independent sequences (V = f2: two frames, or two neighbouring columns /
// rows of a four-step transform) whose fp32 arithmetic is issued as packed
// FADD2 / FMUL2 / FFMA2.  Pass p (radix R, p_prev = product of earlier
// radices, M = N / R butterflies, butterfly i, k = i mod p_prev):
//     u_j  = x[i + j M] * W_{p_prev R}^{j k}        j = 0..R-1
//     v    = DFT_R(u)
//     y[(i - k) R + k + m p_prev] = v_m              m = 0..R-1
// (autosort: the last pass leaves natural order).  The exchange buffer is
// addressed through a' = a + (a >> log2 R0), which makes every STS/LDS of
// every pass bank-conflict free while keeping all per-element offsets
// compile-time immediates.
//
// What replaces what: the unpack + (-1)^n + fftw_execute + magnitude + DC
// patch of src/nrf.c:598-631 and the per-row work of c/fft-batch.c:61-92 /
// c/fft-batch-broad.c:63-118.  The (-1)^n pre-multiply of the reference is
// realised as the equivalent output rotation by N/2 (store index), and its
// uncentred u8/256 scaling as centred samples plus the analytic DC term
// N/2 (1+i) in bin N/2 (exact by linearity, keeps fp32 error at 1e-7).
#pragma once
#include <stdint.h>
#include "fsea_ptx.cuh"
#include "fsea_codelets.cuh"
#include "../../include/fsea.h"
#include "fsea_kparams.h"

// internal input / output kinds of the four-step (N = N1 x N2) kernels.  The loader format
// carries its geometry: N2 interleaved sub-sequences, tiles transformed in RND rounds.
#define FSEA_IN_COLS_FMT(N2, RND) (100 + (N2) + 1000 * (RND))
template <int FMT> struct fsea_cols_geom {
    static const bool COLS = FMT >= 100;
    static const int N2 = COLS ? (FMT % 1000) - 100 : 32;
    static const int RND = COLS ? FMT / 1000 : 1;
};
enum { FSEA_OK_SPECTRUM = 0, FSEA_OK_CPLX = 3, FSEA_OK_HALF = 4 };
// "half" kernels: ONE frame of 2 N samples per unit, the packed lanes being its even and odd
// samples -- the N-point transform runs on both at once and the output stage finishes the
// radix-2 decimation in time across the lanes (fsea_out_half).  Halves the threads and the
// shared memory a frame needs: 2048 = 2 x 1024 stays inside one warp.
enum { FSEA_IN_U8H = 20 };


template <int X> struct fsea_log2 { static const int v = 1 + fsea_log2<X / 2>::v; };
template <> struct fsea_log2<1> { static const int v = 0; };

template <int N_, int R0_, int R1_, int R2_>
struct fsea_shape {
    static const int N = N_, R0 = R0_, R1 = R1_, R2 = R2_;
    static const int P = (R0 > R1 ? (R0 > R2 ? R0 : R2) : (R1 > R2 ? R1 : R2));
    static const int T = N / P;
    static const int NPASS = (R1 == 1 ? 1 : (R2 == 1 ? 2 : 3));
    static const int L0 = fsea_log2<R0>::v;
    static const int UNITS = N + (N >> L0);          // padded exchange length
    static const int TW1 = (R1 - 1) * R0;            // pass-1 table entries
    static const int TW2 = (R2 - 1) * R0 * R1;       // pass-2 table entries
    static_assert(R0 * R1 * R2 == N, "radices must multiply to N");
    static_assert(R0 >= 8 || N < 8, "first radix >= 8 keeps the exchange conflict-free");
};

// ---- scalar helpers ---------------------------------------------------------
__device__ __forceinline__ float fsea_db(float pwr, const fsea_logp& lg) {   // mult log10(gain pwr + 1e-20); default 10 log10(pwr + 1e-20)
    return fmaf(lg.mul, __log2f(pwr + lg.eps), lg.add);
}
__device__ __forceinline__ double fsea_db(double pwr, const fsea_logp& lg) {
    return lg.dmul * log2(pwr + lg.deps) + lg.dadd;
}
__device__ __forceinline__ uint8_t fsea_pix(float db, float scale) {   // c/fft-batch.c:33-35,90-91
    int v = __float2int_rz(db * scale);
    v = max(0, min(255, v));
    return (uint8_t)v;
}

template <typename V> struct fsea_cplx;
template <> struct fsea_cplx<float>  { typedef float2 t; };
template <> struct fsea_cplx<f2>     { typedef float2 t; };
template <> struct fsea_cplx<double> { typedef double2 t; };

// lane accessors: sequence `l` of the unit
__device__ __forceinline__ float fsea_lane(float v, int) { return v; }
__device__ __forceinline__ double fsea_lane(double v, int) { return v; }
__device__ __forceinline__ float fsea_lane(f2 v, int l) { return l ? v.v.y : v.v.x; }

// ---- input side: element e of the unit's sequence(s) -> (re, im) ------------
// 8-bit IQ -> fp32 without cvt: PRMT drops the byte into the mantissa of
// 2^15 (0x47000000 | b = 32768 + b/256), one FADD removes 32768.5:
// (b - 128)/256 exactly.  Two's-complement input is flipped to offset binary
// first (b ^ 0x80 == (b + 128) mod 256, src/nrf.c:102-105).
// FMT here is FSEA_IN_U8 for both 8-bit conventions (the XOR mask is a kernel
// argument) or FSEA_IN_F32C.
template <int FMT, typename V> struct fsea_in;

// every thread of the group touches its share of the sequence's 128-byte lines
template <int BYTES, int T>
__device__ __forceinline__ void fsea_prefetch_seq(const uint8_t* p, int t) {
    constexpr int LINES = (BYTES + 127) / 128;
#pragma unroll
    for (int l = 0; l < (LINES + T - 1) / T; l++) {
        const int line = t + l * T;
        if (line < LINES) fsea_prefetch_l2(p + (size_t)line * 128);
    }
}

template <int FMT> struct fsea_in<FMT, float> {
    const uint8_t* p;
    uint32_t xm;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool& live) {
        long s = unit < kp.n_seq ? unit : kp.n_seq - 1;
        live = unit < kp.n_seq;
        xm = kp.xor_mask;
        win = kp.window;
        p = (const uint8_t*)kp.in + s * kp.seq_stride * (FMT == FSEA_IN_F32C ? 8 : 2);
    }
    __device__ __forceinline__ void load(int e, float& re, float& im) const {
        if (FMT == FSEA_IN_F32C) {
            float2 v = __ldg((const float2*)p + e);
            re = v.x; im = v.y;
        } else {
            const uint32_t w = (uint32_t)__ldg((const unsigned short*)p + e) ^ xm;
            re = __uint_as_float(__byte_perm(w, 0x47000000u, 0x7650)) - 32768.5f;
            im = __uint_as_float(__byte_perm(w, 0x47000000u, 0x7651)) - 32768.5f;
        }
    }
    const float* win;
    __device__ __forceinline__ float window(int e) const { return __ldg(win + e); }
    template <int N, int T> __device__ __forceinline__ void prefetch(int t) const {
        fsea_prefetch_seq<N * (FMT == FSEA_IN_F32C ? 8 : 2), T>(p, t);
    }
    // 8-bit only: coalesced 16-byte loads of the whole frame into shared memory
    template <int N, int T> __device__ __forceinline__ void async_stage(int t, unsigned char* sm) const;
    // Staged kernels work on INTEGER-scaled samples (byte in mantissa bits 8..15:
    // 0x47000000 | b << 8 = 32768 + b); the 1/256 of src/nrf.c:605-606 is folded into the
    // pass-1 twiddle table.  load_staged: centred value b - 128; load_staged_raw: the magic
    // float itself (the first butterfly stage of fsea_dft*_raw removes the bias).
    template <int N> __device__ __forceinline__ void load_staged(const unsigned char* sm, int e, float& re, float& im) const {
        load_staged_raw<N>(sm, e, re, im);
        re -= 32896.0f; im -= 32896.0f;
    }
    template <int N> __device__ __forceinline__ void load_staged_raw(const unsigned char* sm, int e, float& re, float& im) const {
        const uint32_t w = (uint32_t)((const unsigned short*)sm)[e] ^ xm;
        re = __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504));
        im = __uint_as_float(__byte_perm(w, 0x47000000u, 0x7514));
    }
};

template <int FMT> struct fsea_in<FMT, f2> {
    const uint8_t *pa, *pb;
    uint32_t xm;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool& live) {
        long s0 = 2 * unit, s1 = 2 * unit + 1;
        live = s0 < kp.n_seq;
        if (s0 >= kp.n_seq) s0 = kp.n_seq - 1;
        if (s1 >= kp.n_seq) s1 = kp.n_seq - 1;
        xm = kp.xor_mask;
        win = kp.window;
        const long b = (FMT == FSEA_IN_F32C ? 8 : 2);
        pa = (const uint8_t*)kp.in + s0 * kp.seq_stride * b;
        pb = (const uint8_t*)kp.in + s1 * kp.seq_stride * b;
    }
    __device__ __forceinline__ void load(int e, f2& re, f2& im) const {
        if (FMT == FSEA_IN_F32C) {
            float2 a = __ldg((const float2*)pa + e), b = __ldg((const float2*)pb + e);
            re = mk2(a.x, b.x); im = mk2(a.y, b.y);
        } else {
            const uint32_t wa = (uint32_t)__ldg((const unsigned short*)pa + e) ^ xm;
            const uint32_t wb = (uint32_t)__ldg((const unsigned short*)pb + e) ^ xm;
            f2 r = mk2(__uint_as_float(__byte_perm(wa, 0x47000000u, 0x7650)),
                       __uint_as_float(__byte_perm(wb, 0x47000000u, 0x7650)));
            f2 i = mk2(__uint_as_float(__byte_perm(wa, 0x47000000u, 0x7651)),
                       __uint_as_float(__byte_perm(wb, 0x47000000u, 0x7651)));
            const f2 off = mk2(32768.5f, 32768.5f);
            re = vsub(r, off); im = vsub(i, off);
        }
    }
    const float* win;
    __device__ __forceinline__ f2 window(int e) const { const float w = __ldg(win + e); return mk2(w, w); }
    template <int N, int T> __device__ __forceinline__ void prefetch(int t) const {
        fsea_prefetch_seq<N * (FMT == FSEA_IN_F32C ? 8 : 2), T>(pa, t);
        fsea_prefetch_seq<N * (FMT == FSEA_IN_F32C ? 8 : 2), T>(pb, t);
    }
    template <int N, int T> __device__ __forceinline__ void async_stage(int t, unsigned char* sm) const;
    template <int N> __device__ __forceinline__ void load_staged(const unsigned char* sm, int e, f2& re, f2& im) const {
        load_staged_raw<N>(sm, e, re, im);
        const f2 off = mk2(32896.0f, 32896.0f);
        re = vsub(re, off); im = vsub(im, off);
    }
    template <int N> __device__ __forceinline__ void load_staged_raw(const unsigned char* sm, int e, f2& re, f2& im) const {
        const uint32_t wa = (uint32_t)((const unsigned short*)sm)[e] ^ xm;
        const uint32_t wb = (uint32_t)((const unsigned short*)(sm + 2 * N))[e] ^ xm;
        re = mk2(__uint_as_float(__byte_perm(wa, 0x47000000u, 0x7504)),
                 __uint_as_float(__byte_perm(wb, 0x47000000u, 0x7504)));
        im = mk2(__uint_as_float(__byte_perm(wa, 0x47000000u, 0x7514)),
                 __uint_as_float(__byte_perm(wb, 0x47000000u, 0x7514)));
    }
};

template <> struct fsea_in<FSEA_IN_F64C, double> {
    const double2* p;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool& live) {
        long s = unit < kp.n_seq ? unit : kp.n_seq - 1;
        live = unit < kp.n_seq;
        win = kp.window;
        p = (const double2*)kp.in + s * kp.seq_stride;
    }
    __device__ __forceinline__ void load(int e, double& re, double& im) const {
        double2 v = __ldg(p + e);
        re = v.x; im = v.y;
    }
    const float* win;
    __device__ __forceinline__ double window(int e) const { return (double)__ldg(win + e); }
    template <int N, int T> __device__ __forceinline__ void prefetch(int t) const {
        fsea_prefetch_seq<N * 16, T>((const uint8_t*)p, t);
    }
    template <int N, int T> __device__ __forceinline__ void async_stage(int, unsigned char*) const {}
    template <int N> __device__ __forceinline__ void load_staged(const unsigned char*, int, double&, double&) const {}
    template <int N> __device__ __forceinline__ void load_staged_raw(const unsigned char*, int, double&, double&) const {}
};

// ---- one frame per unit, lanes = (even, odd) samples: element e of the packed sequence is the
// four adjacent bytes I,Q,I,Q of samples 2e and 2e+1 ------------------------------------------
template <> struct fsea_in<FSEA_IN_U8H, f2> {
    const uint8_t* p;
    const float* win;
    uint32_t x4;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool& live) {
        live = unit < kp.n_seq;
        const long s = live ? unit : kp.n_seq - 1;
        p = (const uint8_t*)kp.in + s * kp.seq_stride * 2;
        win = kp.window;
        x4 = kp.xor_mask | (kp.xor_mask << 16);
    }
    static __device__ __forceinline__ void unpack(uint32_t w, f2& re, f2& im) {
        re = mk2(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7504)), __uint_as_float(__byte_perm(w, 0x47000000u, 0x7524)));
        im = mk2(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7514)), __uint_as_float(__byte_perm(w, 0x47000000u, 0x7534)));
    }
    __device__ __forceinline__ void load(int e, f2& re, f2& im) const {              // (unstaged: single-pass shapes)
        unpack(__ldg((const uint32_t*)p + e) ^ x4, re, im);
        const f2 off = mk2(32896.0f, 32896.0f), sc = mk2(1.0f / 256.0f, 1.0f / 256.0f);
        re = vmul(vsub(re, off), sc); im = vmul(vsub(im, off), sc);
    }
    __device__ __forceinline__ f2 window(int e) const {
        const float2 w = __ldg((const float2*)win + e);
        return mk2(w.x, w.y);
    }
    template <int N, int T> __device__ __forceinline__ void prefetch(int) const {}
    template <int N, int T> __device__ __forceinline__ void async_stage(int t, unsigned char* sm) const;   // 4 N bytes
    template <int N> __device__ __forceinline__ void load_staged_raw(const unsigned char* sm, int e, f2& re, f2& im) const {
        unpack(((const uint32_t*)sm)[e] ^ x4, re, im);
    }
    template <int N> __device__ __forceinline__ void load_staged(const unsigned char* sm, int e, f2& re, f2& im) const {
        load_staged_raw<N>(sm, e, re, im);
        const f2 off = mk2(32896.0f, 32896.0f);
        re = vsub(re, off); im = vsub(im, off);
    }
};

// ---- four-step, decimation in time (N = N1 x N2, n = N2 n1 + n2) ------------------
// Step A is this file's transform on the N2 interleaved sub-sequences x[N2 n1 + n2] of a frame:
// a unit is the PAIR of neighbouring sub-sequences (n2, n2+1) -- four adjacent bytes per n1,
// the two packed lanes.  A block takes a TILE of G x RND units of one frame and transforms
// them in RND rounds of G; the tile's bytes (8 G RND / 2 per n1) reach shared memory as
// 8-byte pieces laid out ring[h][n1] (h = unit pair), so a group reads its samples with one
// LDS.32 per point.  (Narrower tiles cost LSU time out of proportion: an LDGSTS is replayed
// once per 128-byte line it gathers from.)  The sub-spectra go to the intermediate as planes
// Yre[n2][k1], Yim[n2][k1]; step B (fsea_comb*, below) twiddles them and does the radix-N2
// across n2 in registers.
template <int N2_, int RND_> struct fsea_in_cols {
    static const int N2 = N2_, RND = RND_;
    const uint8_t* p;      // the unit's first byte in its frame
    const float* win;
    uint32_t x4;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool& live) {
        long s0 = 2 * unit;                                    // sub-sequence index in the batch
        live = s0 < kp.n_seq;
        if (!live) s0 = kp.n_seq - 2;
        const long f = s0 / N2;
        const int c = (int)(s0 - f * N2);
        p = (const uint8_t*)kp.in + 2 * (f * kp.seq_stride + c);
        win = kp.window + c;
        x4 = kp.xor_mask | (kp.xor_mask << 16);
    }
    __device__ __forceinline__ void load(int, f2&, f2&) const {}
    __device__ __forceinline__ f2 window(int e) const {
        const float2 w = __ldg((const float2*)(win + N2 * e));
        return mk2(w.x, w.y);
    }
    template <int N, int T> __device__ __forceinline__ void prefetch(int) const {}
    template <int N, int T> __device__ __forceinline__ void async_stage(int, unsigned char*) const {}
    // ring geometry: H pieces of 8 bytes per n1; row h of the ring is N + 16/H pieces long
    // (the pad keeps the LDGSTS writes of a half-warp on distinct banks)
    template <int N, int G> struct ring {
        static const int H = (G * RND / 2) > 0 ? (G * RND / 2) : 1;
        static const int ROW = N + (16 / H > 0 ? 16 / H : 1);  // pieces per ring row
        static const int BYTES = H * ROW * 8;
    };
    // all TPB threads of the block fetch the tile whose first unit is this (round-0) unit of group g
    template <int N, int TPB, int G>
    __device__ __forceinline__ void async_tile(int tid, int g, unsigned char* rg) const {
        typedef ring<N, G> R;
        static_assert((N * R::H) % TPB == 0, "tile pieces must divide over the block");
        const uint8_t* pt = p - 4 * g;
#pragma unroll
        for (int c = 0; c < (N * R::H) / TPB; c++) {
            const int q = tid + c * TPB;
            const int n1 = q / R::H, h = q - n1 * R::H;
            fsea_cp_async8(rg + ((size_t)h * R::ROW + n1) * 8, pt + (2 * N2) * (size_t)n1 + 8 * h);
        }
    }
    // where unit `slot` (0 .. G RND - 1) of the tile finds its words in the ring
    template <int N, int G>
    static __device__ __forceinline__ const unsigned char* ring_of(const unsigned char* rg, int slot) {
        return rg + (size_t)(slot >> 1) * ring<N, G>::ROW * 8 + (slot & 1) * 4;
    }
    template <int N> __device__ __forceinline__ void load_staged_raw(const unsigned char* sm, int e, f2& re, f2& im) const {
        const uint32_t w = ((const uint32_t*)sm)[2 * e] ^ x4;
        re = mk2(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7504)), __uint_as_float(__byte_perm(w, 0x47000000u, 0x7524)));
        im = mk2(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7514)), __uint_as_float(__byte_perm(w, 0x47000000u, 0x7534)));
    }
    template <int N> __device__ __forceinline__ void load_staged(const unsigned char* sm, int e, f2& re, f2& im) const {
        load_staged_raw<N>(sm, e, re, im);
        const f2 off = mk2(32896.0f, 32896.0f);
        re = vsub(re, off); im = vsub(im, off);
    }
};
#define FSEA_COLS_SPEC(N2, RND) template <> struct fsea_in<FSEA_IN_COLS_FMT(N2, RND), f2> : fsea_in_cols<N2, RND> {};
FSEA_COLS_SPEC(8, 1)
FSEA_COLS_SPEC(16, 1)
FSEA_COLS_SPEC(32, 1)
FSEA_COLS_SPEC(32, 2)
FSEA_COLS_SPEC(64, 1)
FSEA_COLS_SPEC(64, 2)
#undef FSEA_COLS_SPEC

// ---- output side: bin k of the unit's sequence(s) ---------------------------
// The reference's (-1)^n input modulation (src/nrf.c:611-612) is the output
// rotation by N/2: FFT bin k is stored at position (k + N/2) mod N.  In the
// last pass a thread owns bins i + m*M (i < M), so the position is
// i + ((m*M + N/2) mod N): a per-thread base pointer plus an immediate.
template <int N, typename V> struct fsea_out {
    typedef typename fsea_vt<V>::scalar S;
    static const int LANES = fsea_vt<V>::lanes;
    uint8_t* row[LANES];
    bool live[LANES];
    float pix_scale;
    fsea_logp lg;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool unit_live, int esz) {
        pix_scale = kp.pix_scale;
        lg = kp.lg;
#pragma unroll
        for (int l = 0; l < LANES; l++) {
            long s = unit * LANES + l;
            live[l] = unit_live && s < kp.n_seq;
            // sequences past the end of the batch (odd tail, idle group) write to a dump row:
            // the stores themselves stay unconditional
            row[l] = live[l] ? (uint8_t*)kp.out + fsea_out_row(kp, s) : (uint8_t*)kp.dump;
        }
    }
    template <int OUT> __device__ __forceinline__ void put(int l, int pos, S pwr) const {
        if (OUT == FSEA_OUT_MAG) {
            if (sizeof(S) == 8) ((double*)row[l])[pos] = sqrt((double)pwr);
            else ((float*)row[l])[pos] = fsea_sqrt_approx((float)pwr);
        } else if (OUT == FSEA_OUT_LOGMAG) {
            if (sizeof(S) == 8) ((double*)row[l])[pos] = fsea_db((double)pwr, lg);
            else ((float*)row[l])[pos] = fsea_db((float)pwr, lg);
        } else {
            if (sizeof(S) == 8) {
                int v = (int)(fsea_db((double)pwr, lg) * (double)pix_scale);
                row[l][pos] = (uint8_t)max(0, min(255, v));
            } else {
                row[l][pos] = fsea_pix(fsea_db((float)pwr, lg), pix_scale);
            }
        }
    }
    // bins i + m*M, m = 0..R-1, of one butterfly; dc_patch: bin N-1 also fills
    // the DC slot N/2 and bin 0 is not stored (src/nrf.c:621-622).
    template <int OUT, int R, int M>
    __device__ __forceinline__ void store_bfly(int i, const V* re, const V* im, bool dc_patch) const {
#pragma unroll
        for (int m = 0; m < R; m++) {
            constexpr int NN = N;
            const int off = (m * M + NN / 2) & (NN - 1);
            if (m == 0 && dc_patch && i == 0) continue;
            const V pw = vfma2(re[m], re[m], vmul(im[m], im[m]));
#pragma unroll
            for (int l = 0; l < LANES; l++) {
                put<OUT>(l, i + off, fsea_lane(pw, l));
                if (m == R - 1 && dc_patch && i == M - 1) put<OUT>(l, NN / 2, fsea_lane(pw, l));
            }
        }
    }
    template <int R, int M>
    __device__ __forceinline__ void store_sel(int mode, int i, const V* re, const V* im, bool dc_patch) const {
        if (mode == FSEA_OUT_MAG) store_bfly<FSEA_OUT_MAG, R, M>(i, re, im, dc_patch);
        else if (mode == FSEA_OUT_LOGMAG) store_bfly<FSEA_OUT_LOGMAG, R, M>(i, re, im, dc_patch);
        else store_bfly<FSEA_OUT_PIX, R, M>(i, re, im, dc_patch);
    }
    // uncentred 8-bit input: the analytic DC term N/2 (1+i) belongs to bin 0 (thread 0, m = 0)
    __device__ __forceinline__ void dc_adjust(const fsea_kparams& kp, int t, V& re0, V& im0) const {
        if (kp.dc_term && t == 0) { re0 = vadd(re0, vsplat((S)(N / 2), re0)); im0 = vadd(im0, vsplat((S)(N / 2), im0)); }
    }
};

// Output side of the "half" kernels.  After the last pass a thread holds, per butterfly i, the
// bins k = i + m M (m < R) of the even-sample transform E (lane x) and of the odd-sample
// transform O (lane y) of ONE frame of 2 N samples.  X[k] = E[k] + W_2N^k O[k],
// X[k + N] = E[k] - W_2N^k O[k] with W_2N^k = W_2N^i W_2R^m: the thread's own factor W_2N^i
// comes from a table (kp.twh, M entries), the W_2R^m are the constants of fsea_cross<R>.
// Shifted positions (src/nrf.c:611-612): bin k -> k + N, bin k + N -> k.
template <int N> struct fsea_out_half {
    uint8_t* row;
    const float2* twh;
    float pix_scale, dc;
    fsea_logp lg;
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool unit_live, int esz) {
        pix_scale = kp.pix_scale;
        lg = kp.lg;
        twh = (const float2*)kp.twh;
        dc = kp.dc_term ? (float)N : 0.0f;                                 // N/2 (1+i) of the 2N-point frame
        row = unit_live ? (uint8_t*)kp.out + fsea_out_row(kp, unit) : (uint8_t*)kp.dump;
    }
    template <int OUT> __device__ __forceinline__ void put(int pos, float pwr) const {
        if (OUT == FSEA_OUT_MAG) ((float*)row)[pos] = fsea_sqrt_approx(pwr);
        else if (OUT == FSEA_OUT_LOGMAG) ((float*)row)[pos] = fsea_db(pwr, lg);
        else row[pos] = fsea_pix(fsea_db(pwr, lg), pix_scale);
    }
    template <int OUT, int R, int M>
    __device__ __forceinline__ void store_bfly(int i, f2* re, f2* im, bool dc_patch) const {
        const float2 w = __ldg(twh + i);                                   // W_2N^i
#pragma unroll
        for (int m = 0; m < R; m++) {                                      // O <- W_2N^i O
            const float br = re[m].v.y, bi = im[m].v.y;
            re[m].v.y = fmaf(br, w.x, -bi * w.y);
            im[m].v.y = fmaf(br, w.y, bi * w.x);
        }
        fsea_cross<R>(re, im);                                             // lanes: (X[k], X[k + N])
        if (i == 0) { re[0].v.x += dc; im[0].v.x += dc; }
#pragma unroll
        for (int m = 0; m < R; m++) {
            const int k = i + m * M;
            const f2 pw = vfma2(re[m], re[m], vmul(im[m], im[m]));
            if (!(m == 0 && dc_patch && i == 0)) put<OUT>(k + N, pw.v.x);      // bin 0 is not stored (src/nrf.c:621-622)
            put<OUT>(k, pw.v.y);
            if (m == R - 1 && dc_patch && i == M - 1) put<OUT>(N, pw.v.y);     // bin 2N-1 also fills the DC slot
        }
    }
    template <int R, int M>
    __device__ __forceinline__ void store_sel(int mode, int i, f2* re, f2* im, bool dc_patch) const {
        if (mode == FSEA_OUT_MAG) store_bfly<FSEA_OUT_MAG, R, M>(i, re, im, dc_patch);
        else if (mode == FSEA_OUT_LOGMAG) store_bfly<FSEA_OUT_LOGMAG, R, M>(i, re, im, dc_patch);
        else store_bfly<FSEA_OUT_PIX, R, M>(i, re, im, dc_patch);
    }
    __device__ __forceinline__ void dc_adjust(const fsea_kparams&, int, f2&, f2&) const {}
};

// values of one bin pair in the output element type
template <int OUT> __device__ __forceinline__ void fsea_tile_put(unsigned char* p, f2 pw, float pix_scale, const fsea_logp& lg) {
    if (OUT == FSEA_OUT_MAG) *(float2*)p = make_float2(fsea_sqrt_approx(pw.v.x), fsea_sqrt_approx(pw.v.y));
    else if (OUT == FSEA_OUT_LOGMAG) *(float2*)p = make_float2(fsea_db(pw.v.x, lg), fsea_db(pw.v.y, lg));
    else *(uchar2*)p = make_uchar2(fsea_pix(fsea_db(pw.v.x, lg), pix_scale), fsea_pix(fsea_db(pw.v.y, lg), pix_scale));
}

// Four-step step A: the two sub-spectra of a unit as rows of the intermediate, which is kept as
// two float planes per frame, Yre[N2][N1] then Yim[N2][N1]: step A stores lanes of its packed
// registers as they are and step B loads them straight into packed registers (an interleaved
// re/im layout cost ~450 MOVs per thread of repacking in step B).
// N2 = 64 (PAIRED): step B's lanes are the two sub-spectra (2u, 2u+1) step A produced together,
// so the rows of a unit are stored interleaved, Yre[N2/2][N1][2]: one 8-byte store per packed
// register here and one 8-byte load per packed register there (half the LSU instructions).
template <int N, int N2> struct fsea_out_cplx {
    static const bool PAIRED = (N2 == 64);
    float *ra, *rb;                                       // re rows of the unit's two sub-sequences; im at + N2 N
    __device__ __forceinline__ void init(const fsea_kparams& kp, long unit, bool unit_live, int) {
        const long f = unit / (N2 / 2);
        const int u = (int)(unit - f * (N2 / 2));
        ra = unit_live ? (float*)kp.out + (size_t)f * 2 * N2 * N + (size_t)(2 * u) * N : (float*)kp.dump;
        rb = ra + N;
    }
    __device__ __forceinline__ void set(float* frame_base, int u) {
        ra = frame_base + (size_t)(2 * u) * N;
        rb = ra + N;
    }
    template <int R, int M>
    __device__ __forceinline__ void store_sel(int, int i, const f2* re, const f2* im, bool) const {
#pragma unroll
        for (int m = 0; m < R; m++) {
            if (PAIRED) {
                ((float2*)ra)[i + m * M] = make_float2(re[m].v.x, re[m].v.y);
                ((float2*)(ra + N2 * N))[i + m * M] = make_float2(im[m].v.x, im[m].v.y);
            } else {
                ra[i + m * M] = re[m].v.x;
                rb[i + m * M] = re[m].v.y;
                ra[N2 * N + i + m * M] = im[m].v.x;
                rb[N2 * N + i + m * M] = im[m].v.y;
            }
        }
    }
    __device__ __forceinline__ void dc_adjust(const fsea_kparams&, int, f2&, f2&) const {}
};

// (NTOT: the four-step factor N2 for complex rows, unused for spectra)
template <int OK, int N, int NTOT, typename V> struct fsea_out_sel { typedef fsea_out<N, V> type; };
template <int N, int NTOT> struct fsea_out_sel<FSEA_OK_CPLX, N, NTOT, f2> { typedef fsea_out_cplx<N, NTOT> type; };
template <int N, int NTOT> struct fsea_out_sel<FSEA_OK_HALF, N, NTOT, f2> { typedef fsea_out_half<N> type; };

// ---- group barrier ----------------------------------------------------------
template <int T> __device__ __forceinline__ void fsea_group_sync(int group) {
    if (T <= 32) {
        __syncwarp();
    } else {
        fsea_bar_sync<T>(group + 1);
    }
}

// one frame of 8-bit IQ (2N bytes) -> shared memory, 16-byte pieces spread over the group
template <int N, int T>
__device__ __forceinline__ void fsea_async_frame(const uint8_t* p, unsigned char* sm, int t) {
    constexpr int CHUNKS = (2 * N) / 16;
#pragma unroll
    for (int c = 0; c < (CHUNKS + T - 1) / T; c++) {
        const int q = t + c * T;
        if (q < CHUNKS) fsea_cp_async16(sm + 16 * q, p + 16 * q);
    }
}

template <int FMT> template <int N, int T>
__device__ __forceinline__ void fsea_in<FMT, float>::async_stage(int t, unsigned char* sm) const {
    fsea_async_frame<N, T>(p, sm, t);
}
template <int FMT> template <int N, int T>
__device__ __forceinline__ void fsea_in<FMT, f2>::async_stage(int t, unsigned char* sm) const {
    fsea_async_frame<N, T>(pa, sm, t);
    fsea_async_frame<N, T>(pb, sm + 2 * N, t);
}
template <int N, int T>
__device__ __forceinline__ void fsea_in<FSEA_IN_U8H, f2>::async_stage(int t, unsigned char* sm) const {
    fsea_async_frame<2 * N, T>(p, sm, t);
}

// ---- dynamic work distribution ------------------------------------------------
__device__ __forceinline__ unsigned fsea_take_ticket(unsigned* counter) { return fsea_atom_inc(counter); }

// Units are handed out by a global ticket counter instead of a fixed stride:
// the SM's warp arbiter is not fair, and with a static split the favoured
// blocks finish early and leave their SM under-occupied for the rest of the
// launch (ncu: 9.4 of 12 resident warps on average).  The group leader takes
// the ticket at the top of an iteration and the group reads it at the bottom,
// so the atomic's round trip to L2 is hidden behind a whole transform.
template <int T>
__device__ __forceinline__ unsigned fsea_ticket_share(unsigned v, int t, int g, volatile unsigned* slot, int late = 0) {
    if (T <= 32) {
        // The broadcast must stay at the bottom of the iteration: hoisted next to the atomic
        // (which is what the compiler does with it, volatile or not) the warp sits out the
        // atomic's whole round trip (ncu: 6.7 % of all stall samples on this one SHFL).
        // `late` is a value computed from the last stored bin that is always 0 but not
        // provably so; adding it to the source lane pins the shuffle behind the epilogue.
        return fsea_shfl_idx(v, (int)(threadIdx.x & 31) - t + late);
    } else {
        if (t == 0) slot[g] = v;
        fsea_group_sync<T>(g);
        v = slot[g];
        fsea_group_sync<T>(g);
        return v;
    }
}

// ---- one unit through all passes ---------------------------------------------------
// Shared by the batched kernel (fsea_fft_kernel) and the fused four-step kernel.  Returns
// a float of the last butterfly register (the ticket broadcast is pinned behind it).
template <class SH, typename V, int FMT, int OK, int NTOT, int G, bool WIN, class OUTP>
__device__ __forceinline__ float fsea_transform(const fsea_kparams& kp, const fsea_in<FMT, V>& in,
                                                const fsea_in<FMT, V>& in_next, bool more, const OUTP& out,
                                                V* xp, unsigned char* raw,
                                                const typename fsea_cplx<V>::t* tw1, const typename fsea_cplx<V>::t* tw2,
                                                int t, int g, int tid, bool dc_patch, int out_mode,
                                                unsigned char* smem_base,
                                                const typename fsea_vt<V>::scalar* wreg = nullptr,
                                                int slot = 0, bool ring_first = true, bool ring_last = true) {
    typedef typename fsea_vt<V>::scalar S;
    typedef typename fsea_cplx<V>::t S2;
    constexpr int N = SH::N, R0 = SH::R0, R1 = SH::R1, R2 = SH::R2, P = SH::P, T = SH::T;
    constexpr int L0 = SH::L0, NPASS = SH::NPASS;
    constexpr bool COLS = fsea_cols_geom<FMT>::COLS;
    constexpr bool STAGED = (FMT == FSEA_IN_U8 || FMT == FSEA_IN_U8H || COLS) && NPASS > 1 && sizeof(S) == 4;
    constexpr int XB = NPASS > 1 ? SH::UNITS * (int)sizeof(V) : 0;
    (void)tw1; (void)tw2; (void)xp; (void)raw; (void)smem_base; (void)tid; (void)more; (void)slot; (void)ring_first; (void)ring_last;
    V re[P], im[P];

    // ---------------- pass 0: radix R0, p_prev = 1 ----------------
    {
        constexpr int M = N / R0, B = P / R0;
        bool raw_stage = false;                      // pass 0 ran on biased magic floats
        if (STAGED) {
            // (COLS: the ring holds a tile of the BLOCK, filled once for both rounds)
            if constexpr (COLS) { if (ring_first) { fsea_cp_async_wait_all(); __syncthreads(); } }
            else { fsea_cp_async_wait_all(); fsea_group_sync<T>(g); }
            const unsigned char* rg = raw;
            if constexpr (COLS) rg = fsea_in<FMT, V>::template ring_of<N, G>(raw, slot);
            raw_stage = !WIN;
            if (raw_stage) {
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const int i = t + b * T;
#pragma unroll
                    for (int j = 0; j < R0; j++) in.template load_staged_raw<N>(rg, i + j * M, re[b * R0 + j], im[b * R0 + j]);
                }
            } else {
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const int i = t + b * T;
#pragma unroll
                    for (int j = 0; j < R0; j++) in.template load_staged<N>(rg, i + j * M, re[b * R0 + j], im[b * R0 + j]);
                }
            }
            // ring is free again: fetch the next unit / tile
            if constexpr (COLS) {
                if (ring_last) {
                    __syncthreads();
                    if (more) in_next.template async_tile<N, T * G, G>(tid, g, raw);
                    fsea_cp_async_commit();
                }
            } else {
                fsea_group_sync<T>(g);
                if (more) {
                    in_next.template async_stage<N, T>(t, raw);
                    fsea_cp_async_commit();
                }
            }
        } else {
            if (more) in_next.template prefetch<N, T>(t);
#pragma unroll
            for (int b = 0; b < B; b++) {
                const int i = t + b * T;
#pragma unroll
                for (int j = 0; j < R0; j++) in.load(i + j * M, re[b * R0 + j], im[b * R0 + j]);
            }
        }
        if (WIN) {
#pragma unroll
            for (int b = 0; b < B; b++) {
                const int i = t + b * T;
#pragma unroll
                for (int j = 0; j < R0; j++) {
                    const V w = in.window(i + j * M);
                    re[b * R0 + j] = vmul(re[b * R0 + j], w);
                    im[b * R0 + j] = vmul(im[b * R0 + j], w);
                }
            }
        }
        if (STAGED && R0 >= 8 && raw_stage) {
#pragma unroll
            for (int b = 0; b < B; b++) fsea_dft_raw<(R0 >= 8 ? R0 : 8), V>::run(re + b * R0, im + b * R0);
        } else {
#pragma unroll
            for (int b = 0; b < B; b++) fsea_dft<R0, V>::run(re + b * R0, im + b * R0);
        }
        if (NPASS == 1) {
            // N = R0, T = 1: bins m
            out.dc_adjust(kp, 0, re[0], im[0]);
            out.template store_sel<R0, 1>(out_mode, 0, re, im, dc_patch);
        }
    }
    if constexpr (NPASS > 1) {
        // ---------------- exchange 0 -> 1 and pass 1: radix R1, p_prev = R0 ----------------
        constexpr int B0 = P / R0;
        constexpr int M = N / R1, B = P / R1;
        constexpr int JSTR = M + (M >> L0);                    // padded stride of j
        // write a' = i (R0 + 1) + m ; read a' = i + (i >> L0) + j JSTR  (one plane at a time)
#pragma unroll
        for (int b = 0; b < B0; b++) {
            V* w = xp + (t + b * T) * (R0 + 1);
#pragma unroll
            for (int m = 0; m < R0; m++) w[m] = re[b * R0 + m];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int i = t + b * T;
            const V* r = xp + i + (i >> L0);
#pragma unroll
            for (int j = 0; j < R1; j++) re[b * R1 + j] = r[j * JSTR];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B0; b++) {
            V* w = xp + (t + b * T) * (R0 + 1);
#pragma unroll
            for (int m = 0; m < R0; m++) w[m] = im[b * R0 + m];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int i = t + b * T;
            const int k = i & (R0 - 1);
            const V* r = xp + i + (i >> L0);
#pragma unroll
            for (int j = 0; j < R1; j++) im[b * R1 + j] = r[j * JSTR];
            if (STAGED) {                                      // integer-scaled samples: tw1 carries 1/256,
                re[b * R1] = vmuls(re[b * R1], (S)(1.0 / 256.0));   // the untwiddled j = 0 input takes it here
                im[b * R1] = vmuls(im[b * R1], (S)(1.0 / 256.0));
            }
#pragma unroll
            for (int j = 1; j < R1; j++) {
                const S2 w = tw1[(j - 1) * R0 + k];
                vcmul(re[b * R1 + j], im[b * R1 + j], w.x, w.y);
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) fsea_dft<R1, V>::run(re + b * R1, im + b * R1);
        if constexpr (NPASS == 2) {
            // N = R0 R1: i < R0 = M, k = i, bins i + m R0
            out.dc_adjust(kp, t, re[0], im[0]);
#pragma unroll
            for (int b = 0; b < B; b++)
                out.template store_sel<R1, M>(out_mode, t + b * T, re + b * R1, im + b * R1, dc_patch);
        }
    }
    if constexpr (NPASS > 2) {
        // ---------------- exchange 1 -> 2 and pass 2: radix R2, p_prev = R0 R1 ----------------
        constexpr int B1 = P / R1;
        constexpr int M = N / R2, B = P / R2;
        constexpr int JSTR = M + (M >> L0);
        // write a' = hi (R0 R1 + R1) + k + m (R0 + 1), i = hi R0 + k ; read a' = i + (i >> L0) + j JSTR
        fsea_group_sync<T>(g);                                 // all pass-1 reads done (WAR)
#pragma unroll
        for (int b = 0; b < B1; b++) {
            const int i = t + b * T;
            V* w = xp + (i >> L0) * (R0 * R1 + R1) + (i & (R0 - 1));
#pragma unroll
            for (int m = 0; m < R1; m++) w[m * (R0 + 1)] = re[b * R1 + m];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int i = t + b * T;
            const V* r = xp + i + (i >> L0);
#pragma unroll
            for (int j = 0; j < R2; j++) re[b * R2 + j] = r[j * JSTR];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B1; b++) {
            const int i = t + b * T;
            V* w = xp + (i >> L0) * (R0 * R1 + R1) + (i & (R0 - 1));
#pragma unroll
            for (int m = 0; m < R1; m++) w[m * (R0 + 1)] = im[b * R1 + m];
        }
        fsea_group_sync<T>(g);
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int i = t + b * T;                           // k = i
            const V* r = xp + i + (i >> L0);
#pragma unroll
            for (int j = 0; j < R2; j++) im[b * R2 + j] = r[j * JSTR];
#pragma unroll
            for (int j = 1; j < R2; j++) {
                const S2 w = tw2[(j - 1) * M + i];
                vcmul(re[b * R2 + j], im[b * R2 + j], w.x, w.y);
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) fsea_dft<R2, V>::run(re + b * R2, im + b * R2);
        out.dc_adjust(kp, t, re[0], im[0]);
#pragma unroll
        for (int b = 0; b < B; b++)
            out.template store_sel<R2, M>(out_mode, t + b * T, re + b * R2, im + b * R2, dc_patch);
    }
    return (float)fsea_lane(re[P - 1], 0);
}

// ---- the kernel -------------------------------------------------------------
// block = G groups of T threads, each group transforms one unit at a time,
// grid-stride over units.  Dynamic shared memory per block:
//   G x exchange plane   UNITS x sizeof(V)   (re and im cross in two half-exchanges,
//                                             so STS/LDS move whole V registers: no repacking)
//   G x raw ring         lanes x 2N bytes    (8-bit formats: the NEXT unit's IQ bytes land
//                                             here by cp.async while this one is transformed)
//   twiddle tables       when TWS
template <class SH, typename V, int FMT, int OK, int NTOT, int G, bool TWS, int MINB, bool WIN>
__global__ void __launch_bounds__(SH::T* G, MINB)
fsea_fft_kernel(const fsea_kparams kp) {
    typedef typename fsea_vt<V>::scalar S;
    typedef typename fsea_cplx<V>::t S2;
    constexpr int N = SH::N, R0 = SH::R0, R1 = SH::R1, R2 = SH::R2, P = SH::P, T = SH::T;
    constexpr int L0 = SH::L0, NPASS = SH::NPASS;
    constexpr int LANES = fsea_vt<V>::lanes;
    constexpr bool COLS = fsea_cols_geom<FMT>::COLS;
    constexpr bool STAGED = (FMT == FSEA_IN_U8 || FMT == FSEA_IN_U8H || COLS) && NPASS > 1 && sizeof(S) == 4;
    // four-step kernels hand out TILES of G neighbouring units per block (adjacent columns /
    // rows -> contiguous bytes across the block); the others hand out single units per group
    constexpr bool TILED = COLS;
    constexpr int XB = NPASS > 1 ? SH::UNITS * (int)sizeof(V) : 0;     // exchange plane bytes per group
    constexpr int RB = STAGED ? (COLS ? 0 : LANES * 2 * N) : 0;        // raw ring bytes per group (COLS: one per block)

    FSEA_DYN_SMEM(fsea_smem);
    const int tid = threadIdx.x;
    const int g = tid / T;
    const int t = tid - g * T;
    V* xp = reinterpret_cast<V*>(fsea_smem + (size_t)g * XB);
    unsigned char* raw = fsea_smem + (size_t)G * XB + (COLS ? 0 : (size_t)g * RB);   // COLS: one ring per block

    const S2* tw1 = (const S2*)kp.tw1;
    const S2* tw2 = (const S2*)kp.tw2;
    if (TWS && NPASS > 1) {
        S2* stw = reinterpret_cast<S2*>(fsea_smem + (size_t)G * (XB + RB) + (COLS ? fsea_in<COLS ? FMT : FSEA_IN_COLS_FMT(32, 1), f2>::template ring<N, G>::BYTES : 0));
        for (int q = tid; q < SH::TW1 + SH::TW2; q += T * G)
            stw[q] = (q < SH::TW1) ? tw1[q] : tw2[q - SH::TW1];
        __syncthreads();
        tw1 = stw;
        tw2 = stw + SH::TW1;
    }

    const bool dc_patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
    const int out_mode = kp.out_mode;
    const int out_esz = (out_mode == FSEA_OUT_PIX) ? 1 : (int)sizeof(S);
    const long n_units = (FMT == FSEA_IN_U8H) ? kp.n_seq : (kp.n_seq + LANES - 1) / LANES;   // (half: a frame per unit)
    unsigned* ticket = kp.ticket;                                 // [0] next unit, [1] blocks finished
    FSEA_STATIC_SMEM(unsigned, fsea_slot, G > 0 ? G : 1);
    FSEA_STATIC_SMEM(unsigned, fsea_tile_slot_, 1);
    unsigned& fsea_tile_slot = fsea_tile_slot_[0];
    // Scheduling entity and ticket size: a block (TILED: one tile of G units), a group
    // (T >= 32: one unit), or -- for small transforms, where one atomic per unit would make
    // the ticket counter itself the bottleneck (N = 128: 0.9 G units/s) -- a WARP, which takes
    // KB units for each of its GW = 32/T groups per ticket.
    constexpr int GW = (!TILED && T < 32) ? 32 / T : 1;
    constexpr int KB = COLS ? fsea_cols_geom<FMT>::RND : ((!TILED && T < 32) ? 8 : 1);   // units per group and ticket
    constexpr int STEP = TILED ? G : GW;                                        // ... and how far apart they are
    const int gw = (GW > 1) ? ((tid & 31) / T) : 0;               // group within its warp
    const bool leader = TILED ? (tid == 0) : (GW > 1 ? ((tid & 31) == 0) : (t == 0));

    auto share = [&](unsigned v, int late = 0) -> long {           // -> base unit of the ticket
        if (TILED) {                                              // one ticket per block = G units
            if (tid == 0) fsea_tile_slot = v;
            __syncthreads();
            const unsigned tile = fsea_tile_slot;
            __syncthreads();
            return (long)tile * (G * KB) + g;
        }
        if (GW > 1) {
            const unsigned r = fsea_shfl_idx(v, late);
            return (long)r * (GW * KB) + gw;
        }
        return (long)fsea_ticket_share<T>(v, t, g, fsea_slot, late);
    };
    const long n_claims = TILED ? (n_units + G * KB - 1) / (G * KB) * (G * KB) : n_units;   // tiles run to completion

    unsigned pend = 0;
    if (leader) pend = fsea_take_ticket(ticket);
    long cur = share(pend);
    if (leader) pend = fsea_take_ticket(ticket);
    long nxt_base = share(pend);
    // two tickets are always in flight: the one read at the bottom of an iteration was
    // requested a whole iteration earlier, wherever the compiler schedules the broadcast
    unsigned pend_prev = 0;
    if (leader) pend_prev = fsea_take_ticket(ticket);
    int kb = 0;                                                   // position inside the current ticket

    // A thread multiplies the same P window values into every unit; where one block per SM is
    // all that fits anyway (MINB == 1) they live in registers instead of being re-read
    // (ncu, N = 8192 Hann: the window loads were the top long-scoreboard stall).
    constexpr bool WREG = WIN && MINB == 1 && FMT == FSEA_IN_U8;
    S wreg[WREG ? P : 1];
    if (WREG) {
        constexpr int M0 = N / R0, B0 = P / R0;
#pragma unroll
        for (int b = 0; b < B0; b++)
#pragma unroll
            for (int j = 0; j < R0; j++) wreg[b * R0 + j] = (S)__ldg(kp.window + (t + b * T) + j * M0);
    }

    fsea_in<FMT, V> in;
    bool live;
    in.init(kp, cur, live);
    if (STAGED) {                                        // prologue: first unit's bytes
        if constexpr (COLS) { if (live) in.template async_tile<N, T * G, G>(tid, g, raw); }
        else in.template async_stage<N, T>(t, raw);
        fsea_cp_async_commit();
    }

    // (cur - gw is the same for all groups of a warp: they leave the loop together)
    while (TILED ? (cur - g < n_claims) : (cur - gw < n_units)) {
        const bool last_of_ticket = (kb == KB - 1);
        if (leader && kb == 0) pend = fsea_take_ticket(ticket);   // the ticket after next; read at the bottom
        const long unit = cur;
        const long nxt = last_of_ticket ? nxt_base : cur + STEP;
        typename fsea_out_sel<OK, N, NTOT, V>::type out;
        out.init(kp, live ? unit : 0, live, out_esz);
        fsea_in<FMT, V> in_next;
        bool live_next;
        in_next.init(kp, nxt, live_next);
        const bool more = live_next;

        const float probe = fsea_transform<SH, V, FMT, OK, NTOT, G, WIN>(kp, in, in_next, more, out, xp, raw, tw1, tw2,
                                                                         t, g, tid, dc_patch, out_mode, fsea_smem,
                                                                         WREG ? wreg : nullptr,
                                                                         kb * G + g, kb == 0, last_of_ticket);
        if (NPASS > 1) fsea_group_sync<T>(g);                      // the next unit rewrites the exchange plane
        in = in_next;
        live = live_next;
        cur = nxt;
        if (last_of_ticket) {
            // (never true: no finite transform produces this NaN payload in its last register)
            const int late = (__float_as_uint(probe) == 0x7fc0deadu) ? 1 : 0;
            nxt_base = share(pend_prev, late);
            pend_prev = pend;
            kb = 0;
        } else {
            kb++;
        }
    }
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

// ---- four-step, step B: twiddle + radix-N2 across the N2 sub-spectra, all in registers ----
// N2 <= 32: a "pair" is two neighbouring bins (k, k+1), k < N1, of one frame -- the two packed
// lanes.  2 N2 float2 loads of Y[n2][k..k+1] (a warp covers 256 contiguous bytes of every
// row), the four-step twiddles W_N^{n2 k} built from the log2 N2 table rows n2 = 1, 2, 4, ..
// (at most log2 N2 - 1 complex products deep), one packed radix-N2 codelet, and N2 stores of
// spectrum positions k + N1 (m ^ N2/2) (a warp covers 256 contiguous bytes each; the xor is the
// reference's (-1)^n, src/nrf.c:611-612).  A thread works on U = 32 / N2 pairs at a time
// (pair index stride = `pstride`, so that every load and store instruction stays coalesced).
// tw4: planes [2r][k] = cos, [2r+1][k] = sin of W_N^{2^r k}, r < log2 N2, k < N1.
template <int OUT, int N1, int N2> struct fsea_comb {
    static const int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;
    static const int LG = fsea_log2<N2>::v;
    static const int U = 32 / N2;
    static const int PTS = 64;                                        // points per thread and call

    // y = frame's Yre[0][0]; k even
    static __device__ __forceinline__ void load(const float* yf, int k, f2* re, f2* im) {
        const float2* y = (const float2*)(yf + k);
#pragma unroll
        for (int j = 0; j < N2; j++) {
            const float2 a = __ldcg(y + (size_t)j * (N1 / 2));                    // L2 only: written by step A
            const float2 b = __ldcg(y + (size_t)(N2 + j) * (N1 / 2));
            re[j] = mk2(a.x, a.y); im[j] = mk2(b.x, b.y);
        }
    }
    // the log2 N2 table rows of a bin pair (loads only: can be issued long before the sub-spectra are there) ...
    static __device__ __forceinline__ void twiddle_load(const fsea_kparams& kp, int k, f2* br, f2* bi) {
#pragma unroll
        for (int r = 0; r < LG; r++) {
            const float2 wa = fsea_ldg_stream((const float2*)((const float*)kp.tw4 + (size_t)(2 * r) * N1 + k));       // re plane
            const float2 wb = fsea_ldg_stream((const float2*)((const float*)kp.tw4 + (size_t)(2 * r + 1) * N1 + k));   // im plane
            br[r] = mk2(wa.x, wa.y); bi[r] = mk2(wb.x, wb.y);
        }
    }
    // ... and the chain of products W^(n2 k) built from them, applied
    static __device__ __forceinline__ void twiddle_apply(const f2* br, const f2* bi, f2* re, f2* im) {
        f2 wr[N2], wi[N2];                                                       // (only a chain of <= LG is live at a time)
#pragma unroll
        for (int j = 1; j < N2; j++) {
            const int low = j & -j;
            const int r = (low == 1) ? 0 : (low == 2) ? 1 : (low == 4) ? 2 : (low == 8) ? 3 : 4;
            if (j == low) { wr[j] = br[r]; wi[j] = bi[r]; }
            else {
                const int q = j - low;
                wr[j] = vsub(vmul(wr[q], br[r]), vmul(wi[q], bi[r]));
                wi[j] = vfma2(wr[q], bi[r], vmul(wi[q], br[r]));
            }
            const f2 a = vsub(vmul(re[j], wr[j]), vmul(im[j], wi[j]));
            im[j] = vfma2(re[j], wi[j], vmul(im[j], wr[j]));
            re[j] = a;
        }
    }
    static __device__ __forceinline__ void twiddle(const fsea_kparams& kp, int k, f2* re, f2* im) {
        f2 br[LG], bi[LG];                                                       // W^(2^r k), r < LG
        twiddle_load(kp, k, br, bi);
        twiddle_apply(br, bi, re, im);
    }
    static __device__ __forceinline__ void finish(const fsea_kparams& kp, int k, uint8_t* row, bool dc_patch, f2* re, f2* im) {
        fsea_dft<N2, f2>::run(re, im);
        if (kp.dc_term && k == 0) { re[0].v.x += (float)(N2 / 2) * (float)N1; im[0].v.x += (float)(N2 / 2) * (float)N1; }   // N/2 (1+i)
#pragma unroll
        for (int m = 0; m < N2; m++) {
            const f2 pw = vfma2(re[m], re[m], vmul(im[m], im[m]));
            uint8_t* dst = row + ((size_t)k + (size_t)N1 * (m ^ (N2 / 2))) * ESZ;
            __align__(8) unsigned char tmp[8];
            fsea_tile_put<OUT>(tmp, pw, kp.pix_scale, kp.lg);
            if (m == 0 && dc_patch && k == 0) {                              // bin 0 is not stored (src/nrf.c:621-622)
                if (ESZ == 4) ((float*)dst)[1] = ((const float*)tmp)[1]; else dst[1] = tmp[1];
            } else if (ESZ == 4) {
                fsea_st_stream((float2*)dst, *(const float2*)tmp);
            } else {
                *(uchar2*)dst = *(const uchar2*)tmp;
            }
            if (m == N2 - 1 && dc_patch && k == N1 - 2) {                    // bin N-1 also fills the DC slot N/2
                uint8_t* dc = row + (size_t)(N2 / 2) * N1 * ESZ;
                if (ESZ == 4) ((float*)dc)[0] = ((const float*)tmp)[1]; else dc[0] = tmp[1];
            }
        }
    }
    // the U pairs p0 + u pstride, u < U, of one frame; ARRIVE: the fused kernel's "intermediate
    // consumed" signal, given as soon as the loads have delivered their values
    template <int ARRIVE>
    static __device__ __forceinline__ void run(const fsea_kparams& kp, const float* yf, int p0, int pstride, uint8_t* row, bool dc_patch) {
        f2 re[U][N2], im[U][N2];
#pragma unroll
        for (int u = 0; u < U; u++) load(yf, 2 * (p0 + u * pstride), re[u], im[u]);
#pragma unroll
        for (int u = 0; u < U; u++) twiddle(kp, 2 * (p0 + u * pstride), re[u], im[u]);
        if (ARRIVE == 1) fsea_cluster_arrive_relaxed();
#pragma unroll
        for (int u = 0; u < U; u++) finish(kp, 2 * (p0 + u * pstride), row, dc_patch, re[u], im[u]);
    }
};

// N2 = 64: a thread combs ONE bin k of one frame; the packed lanes are the even and the odd
// sub-spectra (n2 = 2j, 2j+1 -- the two lanes step A produced together).  Decimation in time
// once more: A = DFT32 over the even, B = DFT32 over the odd ones (one packed codelet), then
// X[m] = A_m + W64^m B_m, X[m+32] = A_m - W64^m B_m across the lanes (fsea_cross64, scalar).
// Twiddle pairs w[j] = (W^{2j k}, W^{(2j+1) k}) = w[j - low] * W^{2 low k}: the chain runs on
// packed registers with a broadcast scalar factor.
// (twiddle and finish as free functions: fsea_wsc.cuh feeds them from shared memory)
__device__ __forceinline__ void fsea_comb64_twiddle(const fsea_kparams& kp, int N1, int k, f2* re, f2* im) {
    float bc[6], bs[6];                                                      // W^(2^r k), r < 6
#pragma unroll
    for (int r = 0; r < 6; r++) {
        bc[r] = __ldg((const float*)kp.tw4 + (size_t)(2 * r) * N1 + k);
        bs[r] = __ldg((const float*)kp.tw4 + (size_t)(2 * r + 1) * N1 + k);
    }
    f2 wr[32], wi[32];
    wr[0] = mk2(1.0f, bc[0]); wi[0] = mk2(0.0f, bs[0]);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if (j > 0) {
            const int low = j & -j;
            const int r = (low == 1) ? 1 : (low == 2) ? 2 : (low == 4) ? 3 : (low == 8) ? 4 : 5;   // W^{2 low k}
            const int q = j - low;
            wr[j] = vfnmas(wi[q], bs[r], vmuls(wr[q], bc[r]));               // wr c - wi s
            wi[j] = vfmas(wr[q], bs[r], vmuls(wi[q], bc[r]));                // wr s + wi c
        }
        const f2 a = vsub(vmul(re[j], wr[j]), vmul(im[j], wi[j]));
        im[j] = vfma2(re[j], wi[j], vmul(im[j], wr[j]));
        re[j] = a;
    }
}
template <int OUT, int N1>
__device__ __forceinline__ void fsea_comb64_finish(const fsea_kparams& kp, int k, uint8_t* row, bool dc_patch, f2* re, f2* im) {
    constexpr int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;
    fsea_dft32<f2>(re, im);
    fsea_cross<32>(re, im);                                                   // lanes: (X[m], X[m+32])
    if (kp.dc_term && k == 0) { re[0].v.x += 32.0f * (float)N1; im[0].v.x += 32.0f * (float)N1; }     // N/2 (1+i)
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const f2 pw = vfma2(re[m], re[m], vmul(im[m], im[m]));
        __align__(8) unsigned char tmp[8];
        fsea_tile_put<OUT>(tmp, pw, kp.pix_scale, kp.lg);
        uint8_t* d0 = row + ((size_t)k + (size_t)N1 * (m + 32)) * ESZ;       // bin k + N1 m        (lane x)
        uint8_t* d1 = row + ((size_t)k + (size_t)N1 * m) * ESZ;              // bin k + N1 (m + 32) (lane y)
        if (!(m == 0 && dc_patch && k == 0)) {                               // bin 0 is not stored (src/nrf.c:621-622)
            if (ESZ == 4) fsea_st_stream((float*)d0, ((const float*)tmp)[0]); else d0[0] = tmp[0];
        }
        if (ESZ == 4) fsea_st_stream((float*)d1, ((const float*)tmp)[1]); else d1[0] = tmp[1];
        if (m == 31 && dc_patch && k == N1 - 1) {                            // bin N-1 also fills the DC slot N/2
            uint8_t* dc = row + (size_t)32 * N1 * ESZ;
            if (ESZ == 4) *(float*)dc = ((const float*)tmp)[1]; else dc[0] = tmp[1];
        }
    }
}

template <int OUT, int N1> struct fsea_comb<OUT, N1, 64> {
    static const int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;
    static const int U = 1;
    static const int PTS = 64;
    template <int ARRIVE>
    static __device__ __forceinline__ void run(const fsea_kparams& kp, const float* yf, int k, int, uint8_t* row, bool dc_patch) {
        f2 re[32], im[32];
        const float2* y = (const float2*)yf + k;                                  // rows (2j, 2j+1) interleaved (fsea_out_cplx PAIRED)
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float2 a = __ldcg(y + (size_t)j * N1);
            const float2 b = __ldcg(y + (size_t)(32 + j) * N1);
            re[j] = mk2(a.x, a.y); im[j] = mk2(b.x, b.y);
        }
        fsea_comb64_twiddle(kp, N1, k, re, im);
        if (ARRIVE == 1) fsea_cluster_arrive_relaxed();
        fsea_comb64_finish<OUT, N1>(kp, k, row, dc_patch, re, im);
    }
};

// ---- four-step, both steps in one kernel: a thread-block CLUSTER per frame -------------------
// The CS blocks of a cluster transform the sub-sequences of one frame (one tile each, step A),
// meet at the cluster barrier, and each combs its share of the frame's bins (step B).  The
// intermediate of a frame (8 N bytes) is written and read back within microseconds by blocks
// of the same GPC: it lives in L2 (resident clusters x 8 N bytes) and never reaches HBM, which
// leaves the algorithmic 2 + 4 bytes per point as the only DRAM traffic.
// Frames are dealt to clusters round-robin.  Two barrier phases per frame: "sub-spectra
// published" (release/acquire, exposed) and "intermediate consumed" (relaxed arrive as soon as
// step B's loads have landed, waited for at the top of the next frame: hidden).  CS = 1 (one
// block per frame, N = 8192): the block barrier plays both roles.
template <class SH, int N2, int G, bool TWS, int MINB, bool WIN, int OUT, int CS>
__global__ void __launch_bounds__(SH::T* G, MINB)
fsea_fused4_kernel(const fsea_kparams kp) {
    typedef float2 S2;
    constexpr int N = SH::N, T = SH::T, NPASS = SH::NPASS;            // N: sub-spectrum length N1
    constexpr int UPT = (N2 / 2) / CS;                                // units per tile = per block and frame
    constexpr int RND = UPT / G;
    static_assert(UPT * CS == N2 / 2 && RND * G == UPT, "the cluster's tiles must cover the N2/2 unit pairs of a frame");
    constexpr int FMT = FSEA_IN_COLS_FMT(N2, RND);
    typedef fsea_in<FMT, f2> IN;
    typedef fsea_comb<OUT, N, N2> COMB;
    constexpr int XB = SH::UNITS * (int)sizeof(f2);
    constexpr int ESZ = (OUT == FSEA_OUT_PIX) ? 1 : 4;
    constexpr int TPB = T * G;
    constexpr int IPB = (N * N2 / COMB::PTS) / CS;                    // step-B thread-calls per block and frame
    constexpr int BIT = IPB / TPB;                                    // step-B passes of the block per frame
    static_assert(IPB % TPB == 0 && BIT >= 1, "step B must divide over the block");

    FSEA_DYN_SMEM(fsea_smem);
    const int tid = threadIdx.x;
    const int g = tid / T;
    const int t = tid - g * T;
    f2* xp = reinterpret_cast<f2*>(fsea_smem + (size_t)g * XB);
    unsigned char* ring = fsea_smem + (size_t)G * XB;
    const S2* tw1 = (const S2*)kp.tw1;
    const S2* tw2 = (const S2*)kp.tw2;
    if (TWS) {
        S2* stw = reinterpret_cast<S2*>(fsea_smem + (size_t)G * XB + IN::template ring<N, G>::BYTES);
        for (int q = tid; q < SH::TW1 + SH::TW2; q += TPB) stw[q] = (q < SH::TW1) ? tw1[q] : tw2[q - SH::TW1];
        __syncthreads();
        tw1 = stw;
        tw2 = stw + SH::TW1;
    }
    const bool dc_patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
    const unsigned crank = CS > 1 ? fsea_cluster_ctarank() : 0u;
    const long cid = CS > 1 ? (long)fsea_cluster_id() : (long)blockIdx.x;
    const long step = CS > 1 ? (long)fsea_cluster_count() : (long)gridDim.x;
    const long frames = kp.n_seq;
    float* scr = (float*)kp.scratch + (size_t)cid * 2 * N2 * N;        // the cluster's slot: Yre[N2][N], Yim[N2][N]
    fsea_kparams kc = kp;                                            // the loader counts sub-sequences
    kc.n_seq = frames * N2;

    long f = cid;
    {
        IN in;
        bool live;
        in.init(kc, f * (N2 / 2) + crank * UPT + g, live);
        if (f < frames) in.template async_tile<N, TPB, G>(tid, g, ring);
        fsea_cp_async_commit();
    }
    // step A of frame f: this block's tile -> the sub-spectra rows of slot `dst`
    auto step_a = [&](long fa, float* dst) {
        const bool more = fa + step < frames;
#pragma unroll 1
        for (int r = 0; r < RND; r++) {
            const int u = crank * UPT + r * G + g;                   // unit pair of the frame
            IN in_r, in_next;
            bool lv;
            in_r.init(kc, fa * (N2 / 2) + u, lv);
            in_next.init(kc, (fa + step) * (N2 / 2) + crank * UPT + g, lv);
            fsea_out_cplx<N, N2> out;
            out.set(dst, u);
            fsea_transform<SH, f2, FMT, FSEA_OK_CPLX, N2, G, WIN>(kc, in_r, in_next, more, out, xp, ring, tw1, tw2, t, g, tid,
                                                                 false, 0, fsea_smem, nullptr, r * G + g, r == 0, r == RND - 1);
            if (NPASS > 1) fsea_group_sync<T>(g);
        }
    };
    bool pending = false;                                        // an arrive of the "intermediate consumed" phase
    for (; f < frames; f += step) {
        if (CS > 1) { if (pending) fsea_cluster_wait_acquire(); }
        else if (pending) __syncthreads();
        step_a(f, scr);
        if (CS > 1) {
            fsea_cluster_arrive_release();
            fsea_cluster_wait_acquire();
        } else {
            __syncthreads();
        }
        uint8_t* row = (uint8_t*)kp.out + fsea_out_row(kp, f);
#pragma unroll 1
        for (int it = 0; it < BIT; it++) {
            // the block's thread-calls of this pass: first index and stride between a thread's U items
            const int first = (int)crank * (IPB * COMB::U) + it * TPB + tid;
            if (CS > 1 && it == BIT - 1) COMB::template run<1>(kp, scr, first, BIT * TPB, row, dc_patch);
            else COMB::template run<0>(kp, scr, first, BIT * TPB, row, dc_patch);
        }
        pending = true;
    }
    if (CS > 1 && pending) fsea_cluster_wait_acquire();
}

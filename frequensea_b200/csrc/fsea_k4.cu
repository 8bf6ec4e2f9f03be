// fsea_k4.cu -- four-step instantiations (N = N1 x N2) for N = 8192 .. 65536, 8-bit input.
#include "fsea_instantiate.cuh"

// Both steps in one kernel, a cluster of CS blocks per frame (fsea_fused4_kernel): step A =
// N1-point transforms of the N2 interleaved sub-sequences (COLS loader, complex rows into the
// L2-resident intermediate), step B = twiddle + radix-N2 comb in registers.
template <class SHA, int N2, int CS>
static bool fill_fused(fsea_kernel_info4* o, bool window, int out_mode, const char* name) {
    constexpr int G = 128 / SHA::T;                                   // 128 threads per block
    constexpr int RND = (N2 / 2) / CS / G;
    static_assert(RND >= 1 && RND * G * CS == N2 / 2, "tiles of the cluster must cover the frame");
    typedef fsea_cfg<SHA, f2, FSEA_IN_COLS_FMT(N2, RND), G> C;
#define FSEA_FUSED(W, O) (const void*)&fsea_fused4_kernel<SHA, N2, C::G, C::TWS, C::MINB, W, O, CS>
    o->fn = window ? (out_mode == FSEA_OUT_MAG ? FSEA_FUSED(true, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_FUSED(true, FSEA_OUT_LOGMAG) : FSEA_FUSED(true, FSEA_OUT_PIX))
                   : (out_mode == FSEA_OUT_MAG ? FSEA_FUSED(false, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_FUSED(false, FSEA_OUT_LOGMAG) : FSEA_FUSED(false, FSEA_OUT_PIX));
#undef FSEA_FUSED
    o->fn_sums = nullptr;
    o->cs = CS;
    o->ws = 0;
    o->tma = 0;
    o->threads = SHA::T * C::G;
    o->smem = C::SMEM;
    o->n1 = SHA::N; o->n2 = N2;
    o->r0 = SHA::R0; o->r1 = SHA::R1; o->r2 = SHA::R2;
    o->tw1 = SHA::TW1; o->tw2 = SHA::TW2;
    o->name = name;
    return true;
}

bool fsea_get_kernel4(int n, bool window, int out_mode, int alt, fsea_kernel_info4* out) {
    typedef fsea_shape<1024, 32, 32, 1> S1024;          // warp-synchronous units, one exchange
    typedef fsea_shape<512, 32, 16, 1> S512;
    switch (n) {
        case 8192:
            return fill_fused<S1024, 8, 1>(out, window, out_mode, "fft4<8192=1024x8: sub 32x32 + comb8, block per frame, f32x2>");
        case 16384:
            if (alt == 1) return fill_fused<S512, 32, 2>(out, window, out_mode, "fft4<16384=512x32: sub 32x16 + comb32, cluster 2, f32x2>");
            return fill_fused<S1024, 16, 2>(out, window, out_mode, "fft4<16384=1024x16: sub 32x32 + comb16, cluster 2, f32x2>");
        case 32768:
            if (alt == 1) return fill_fused<S1024, 32, 2>(out, window, out_mode, "fft4<32768=1024x32: sub 32x32 + comb32, cluster 2, f32x2>");
            return fill_fused<S1024, 32, 4>(out, window, out_mode, "fft4<32768=1024x32: sub 32x32 + comb32, cluster 4, f32x2>");
        case 65536:
            if (alt == 1) return fill_fused<S1024, 64, 8>(out, window, out_mode, "fft4<65536=1024x64: sub 32x32 + comb64, cluster 8, f32x2>");
            return fill_fused<S1024, 64, 4>(out, window, out_mode, "fft4<65536=1024x64: sub 32x32 + comb64, cluster 4, f32x2>");
    }
    return false;
}

// fsea_k4.cu -- four-step instantiations for N = 16384, 32768, 65536 (N = N1 x 32).
#include "fsea_instantiate.cuh"

// ---- decimation in time: step A = N1-point transforms of the 32 interleaved sub-sequences
// (fsea_fft_kernel, COLS loader, complex rows out), step B = fsea_comb32_kernel ----
// TWS: twiddle tables in shared memory (-1 = when they fit 16 KB).  At N1 = 2048 they cost the
// third resident block per SM and are still worth it (3.26 vs 3.06 M frames/s at N = 65536).
#ifndef FSEA_COLS2048_TWS
#define FSEA_COLS2048_TWS -1
#endif
template <class SHA, int TWS = -1>
static bool fill4dit(fsea_kernel_info4* o, const char* na, const char* nb, bool window, int out_mode) {
    if (window) fsea_fill<SHA, f2, FSEA_IN_COLS, FSEA_OK_CPLX, SHA::N, true, 0, TWS>(&o->a, na);
    else fsea_fill<SHA, f2, FSEA_IN_COLS, FSEA_OK_CPLX, SHA::N, false, 0, TWS>(&o->a, na);
    o->b.fn = out_mode == FSEA_OUT_MAG ? (const void*)&fsea_comb32_kernel<FSEA_OUT_MAG, SHA::N>
            : out_mode == FSEA_OUT_LOGMAG ? (const void*)&fsea_comb32_kernel<FSEA_OUT_LOGMAG, SHA::N>
                                          : (const void*)&fsea_comb32_kernel<FSEA_OUT_PIX, SHA::N>;
    o->b.threads = 128; o->b.groups = 128; o->b.lanes = 2; o->b.smem = 0;
    o->b.n = 32; o->b.r0 = 32; o->b.r1 = 1; o->b.r2 = 1; o->b.tw1 = 0; o->b.tw2 = 0; o->b.scalar_bytes = 4;
    o->b.name = nb;
    o->n1 = SHA::N;          // sub-spectrum length
    o->n2 = 32;
    return true;
}

// both steps in one kernel, a cluster of CS blocks per frame (fsea_fused4_kernel)
template <class SHA, int CS>
static void fill_fused(fsea_kernel_info4* o, bool window, int out_mode, const char* name) {
    typedef fsea_cfg<SHA, f2, FSEA_IN_COLS> C;
#define FSEA_FUSED(W, O) (const void*)&fsea_fused4_kernel<SHA, C::G, C::TWS, C::MINB, W, O, CS>
    o->fused = window ? (out_mode == FSEA_OUT_MAG ? FSEA_FUSED(true, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_FUSED(true, FSEA_OUT_LOGMAG) : FSEA_FUSED(true, FSEA_OUT_PIX))
                      : (out_mode == FSEA_OUT_MAG ? FSEA_FUSED(false, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_FUSED(false, FSEA_OUT_LOGMAG) : FSEA_FUSED(false, FSEA_OUT_PIX));
#undef FSEA_FUSED
    o->cs = CS;
    o->fused_name = name;
}

bool fsea_get_kernel4(int n, bool window, int out_mode, fsea_kernel_info4* out) {
    switch (n) {
        case 16384:
            fill_fused<fsea_shape<512, 32, 16, 1>, 2>(out, window, out_mode, "fft4<16384=512x32: sub 32x16 + comb32, cluster 2, f32x2>");
            return fill4dit<fsea_shape<512, 32, 16, 1>>(out, "fft4<16384=512x32: sub 32x16 + comb32,f32x2>", "fft4B<16384: comb32>", window, out_mode);
        case 32768:
            fill_fused<fsea_shape<1024, 32, 32, 1>, 4>(out, window, out_mode, "fft4<32768=1024x32: sub 32x32 + comb32, cluster 4, f32x2>");
            return fill4dit<fsea_shape<1024, 32, 32, 1>>(out, "fft4<32768=1024x32: sub 32x32 + comb32,f32x2>", "fft4B<32768: comb32>", window, out_mode);
        case 65536:
            // (a cluster-of-4 fused kernel was measured at 3.38 vs 3.46 M frames/s for the two launches:
            //  2 resident blocks per SM leave the barrier phases exposed)
            out->fused = nullptr; out->cs = 0; out->fused_name = "";
            return fill4dit<fsea_shape<2048, 32, 8, 8>, FSEA_COLS2048_TWS>(out, "fft4<65536=2048x32: sub 32x8x8 + comb32,f32x2>", "fft4B<65536: comb32>", window, out_mode);
    }
    return false;
}

// fsea_k4.cu -- four-step instantiations for N = 16384, 32768, 65536 (N = 32 x N2).
#include "fsea_instantiate.cuh"

template <int NTOT, class SHB>
static bool fill4(fsea_kernel_info4* o, const char* na, const char* nb, bool window) {
    // step A: the in-register column kernel (its own launch geometry)
    o->a.fn = window ? (const void*)&fsea_cols32_kernel<true> : (const void*)&fsea_cols32_kernel<false>;
    o->a.threads = 128; o->a.groups = 128; o->a.lanes = 2; o->a.smem = 0;
    o->a.n = 32; o->a.r0 = 32; o->a.r1 = 1; o->a.r2 = 1; o->a.tw1 = 0; o->a.tw2 = 0; o->a.scalar_bytes = 4;
    o->a.name = na;
    fsea_fill<SHB, f2, FSEA_IN_ROWS, FSEA_OK_STRIDED, NTOT>(&o->b, nb);
    o->n1 = 32;
    o->n2 = SHB::N;
    return true;
}

bool fsea_get_kernel4_old(int n, bool window, fsea_kernel_info4* out) {
    switch (n) {
        case 16384: return fill4<16384, fsea_shape<512, 32, 16, 1>>(out, "fft4<16384=32x512: cols32 + rows 32x16,f32x2>", "fft4B<16384: rows 512=32x16,f32x2>", window);
        case 32768: return fill4<32768, fsea_shape<1024, 32, 32, 1>>(out, "fft4<32768=32x1024: cols32 + rows 32x32,f32x2>", "fft4B<32768: rows 1024=32x32,f32x2>", window);
        case 65536: return fill4<65536, fsea_shape<2048, 32, 8, 8>>(out, "fft4<65536=32x2048: cols32 + rows 32x8x8,f32x2>", "fft4B<65536: rows 2048=32x8x8,f32x2>", window);
    }
    return false;
}

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
    o->b.fn = out_mode == FSEA_OUT_MAG ? (const void*)&fsea_comb32_kernel<FSEA_OUT_MAG>
            : out_mode == FSEA_OUT_LOGMAG ? (const void*)&fsea_comb32_kernel<FSEA_OUT_LOGMAG>
                                          : (const void*)&fsea_comb32_kernel<FSEA_OUT_PIX>;
    o->b.threads = 128; o->b.groups = 128; o->b.lanes = 2; o->b.smem = 0;
    o->b.n = 32; o->b.r0 = 32; o->b.r1 = 1; o->b.r2 = 1; o->b.tw1 = 0; o->b.tw2 = 0; o->b.scalar_bytes = 4;
    o->b.name = nb;
    o->n1 = SHA::N;          // sub-spectrum length
    o->n2 = 32;
    return true;
}

bool fsea_get_kernel4(int n, bool window, int out_mode, fsea_kernel_info4* out) {
    switch (n) {
        case 16384: return fill4dit<fsea_shape<512, 32, 16, 1>>(out, "fft4<16384=512x32: sub 32x16 + comb32,f32x2>", "fft4B<16384: comb32>", window, out_mode);
        case 32768: return fill4dit<fsea_shape<1024, 32, 32, 1>>(out, "fft4<32768=1024x32: sub 32x32 + comb32,f32x2>", "fft4B<32768: comb32>", window, out_mode);
        case 65536: return fill4dit<fsea_shape<2048, 32, 8, 8>, FSEA_COLS2048_TWS>(out, "fft4<65536=2048x32: sub 32x8x8 + comb32,f32x2>", "fft4B<65536: comb32>", window, out_mode);
    }
    return false;
}

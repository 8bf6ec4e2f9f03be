// fsea_k4.cu -- four-step instantiations for N = 16384, 32768, 65536 (N = N1 x N2).
#include "fsea_instantiate.cuh"

template <int NTOT, class SHA, class SHB>
static bool fill4(fsea_kernel_info4* o, const char* na, const char* nb, bool window) {
    if (window) fsea_fill<SHA, f2, FSEA_IN_COLS8, FSEA_OK_SCRATCH, NTOT, true>(&o->a, na);
    else fsea_fill<SHA, f2, FSEA_IN_COLS8, FSEA_OK_SCRATCH, NTOT, false>(&o->a, na);
    fsea_fill<SHB, f2, FSEA_IN_ROWS, FSEA_OK_STRIDED, NTOT>(&o->b, nb);
    o->n1 = SHA::N;
    o->n2 = SHB::N;
    return true;
}

bool fsea_get_kernel4(int n, bool window, fsea_kernel_info4* out) {
    typedef fsea_shape<128, 16, 8, 1> S128;
    typedef fsea_shape<256, 16, 16, 1> S256;
    switch (n) {
        case 16384: return fill4<16384, S128, S128>(out, "fft4A<16384=128x128:cols 16x8,f32x2>", "fft4B<16384:rows 16x8,f32x2>", window);
        case 32768: return fill4<32768, S128, S256>(out, "fft4A<32768=128x256:cols 16x8,f32x2>", "fft4B<32768:rows 16x16,f32x2>", window);
        case 65536: return fill4<65536, S256, S256>(out, "fft4A<65536=256x256:cols 16x16,f32x2>", "fft4B<65536:rows 16x16,f32x2>", window);
    }
    return false;
}

// Host-side harness for the generated register codelets (CPU unit test):
// compiles frequensea_b200/csrc/fsea_codelets.cuh with g++ for float/double.
#include "../frequensea_b200/csrc/fsea_codelets.cuh"
template <typename T> static void run(int R, double* re, double* im) {
    T r[32], i[32];
    for (int k = 0; k < R; k++) { r[k] = (T)re[k]; i[k] = (T)im[k]; }
    switch (R) {
        case 1: break;
        case 2: fsea_dft2<T>(r, i); break;
        case 4: fsea_dft4<T>(r, i); break;
        case 8: fsea_dft8<T>(r, i); break;
        case 16: fsea_dft16<T>(r, i); break;
        case 32: fsea_dft32<T>(r, i); break;
    }
    for (int k = 0; k < R; k++) { re[k] = (double)r[k]; im[k] = (double)i[k]; }
}
extern "C" void codelet_run(int R, int is_double, double* re, double* im) {
    if (is_double) run<double>(R, re, im); else run<float>(R, re, im);
}

// raw variants: inputs are the magic floats 32768 + byte; result is the DFT of (byte - 128)
template <typename T> static void run_raw(int R, double* re, double* im) {
    T r[32], i[32];
    for (int k = 0; k < R; k++) { r[k] = (T)re[k]; i[k] = (T)im[k]; }
    switch (R) {
        case 8: fsea_dft8_raw<T>(r, i); break;
        case 16: fsea_dft16_raw<T>(r, i); break;
        case 32: fsea_dft32_raw<T>(r, i); break;
    }
    for (int k = 0; k < R; k++) { re[k] = (double)r[k]; im[k] = (double)i[k]; }
}
extern "C" void codelet_run_raw(int R, int is_double, double* re, double* im) {
    if (is_double) run_raw<double>(R, re, im); else run_raw<float>(R, re, im);
}

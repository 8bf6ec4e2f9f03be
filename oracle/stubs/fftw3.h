/*
 * fftw3.h -- stand-in for the ABSENT third-party FFTW3 library, used only to
 * build oracle/_ref/libnrf_ref.so from the reference's own sources
 * (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 *
 * Upstream links -lfftw3 (CMakeLists.txt:16, no version pinned); the image has
 * no FFTW.  This header declares just the API subset src/nrf.c calls
 * (src/nrf.c:562-564,615,638-640); stubs_impl.c backs it with a float64
 * radix-2 forward DFT, i.e. FFTW's published definition of FFTW_FORWARD:
 *   Y[k] = sum_j X[j] exp(-2 pi i j k / n), unnormalised.
 */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stddef.h>
typedef double fftw_complex[2];
typedef struct orc_fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#endif

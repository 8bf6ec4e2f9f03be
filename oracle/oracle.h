/*
 * oracle.h -- CPU restatement of frequensea's IQ -> spectrum path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (frequensea_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the upstream tree).  The DFT arithmetic itself lives upstream in FFTW3
 * (libfftw3, double precision, un-vendored, NO version pinned:
 * CMakeLists.txt:16, src/nrf.h:6); FFTW is absent from this image, so the
 * transform is restated here as the textbook forward DFT
 *     X[k] = sum_n x[n] exp(-2 pi i n k / N)        (unnormalised)
 * in float64, which agrees with any FFTW build to ~1e-15 relative.
 *
 * Parity pin: the reference ships no tests or golden vectors ("parity
 * unpinned" by the reference itself).  This oracle is pinned instead against
 *   (1) oracle/_ref/libnrf_ref.so -- the reference's own src/nrf.c + src/nut.c
 *       compiled from where they lie, with only the absent third-party
 *       libraries (FFTW3, libhackrf, librtlsdr, OpenAL) stubbed
 *       (oracle/stubs/), see oracle/Makefile; and
 *   (2) the float64 golden values recorded in SURVEY.md section 8(c),
 *       committed under tests/golden/.
 */
#ifndef FSEA_ORACLE_H
#define FSEA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- formats / modes (mirror include/fsea.h numerically) ---- */
enum { ORC_IN_U8 = 0, ORC_IN_S8 = 1, ORC_IN_F32C = 2, ORC_IN_F64C = 3 };
enum { ORC_WIN_NONE = 0, ORC_WIN_HANN = 1 };
enum { ORC_OUT_MAG = 0, ORC_OUT_LOGMAG = 1, ORC_OUT_PIX = 2 };
enum { ORC_FLAG_DC_PATCH = 1 };

/* src/nrf.c:95-110 -- HackRF / file-replay bytes are two's-complement int8;
 * the device layer turns them into offset-binary u8 with (b + 128) % 256.
 * RTL-SDR bytes are copied unchanged (convert == 0). */
void orc_device_unpack(const uint8_t *raw, uint8_t *samples, size_t nbytes, int convert);

/* Forward complex DFT, float64, interleaved re/im, n a power of two (radix-2)
 * -- the role of fftw_plan_dft_1d(n, in, out, FFTW_FORWARD, ..) + fftw_execute
 * (src/nrf.c:564,615; c/fft-batch.c:144,69). */
void orc_dft_forward(int n, const double *in, double *out);
/* O(n^2) direct evaluation of the same sum: cross-check of orc_dft_forward. */
void orc_dft_naive(int n, const double *in, double *out);

/* src/nrf.c:598-631, one frame: u8 offset-binary samples -> x = u8/256.0 with
 * the (-1)^n pre-multiply (powf(-1, ii)), forward DFT, linear magnitude,
 * mag[n/2] = mag[n/2-1] ("DC compensation").  out: n doubles. */
void orc_nrf_fft_frame_u8(int n, const uint8_t *samples, double *mag);
/* same for NUT_BUFFER_F64 input (src/nrf.c:607-609): interleaved doubles used as is. */
void orc_nrf_fft_frame_f64(int n, const double *samples, double *mag);

/* src/nrf.c:617 -- history scroll (memmove intent) and row-0 write. */
void orc_history_push(double *history, int n, int rows, const double *row0);
/* src/nrf.c:569-596 -- nrf_fft_shift. */
void orc_fft_shift(double *history, int n, int rows, double d);

/* c/fft-batch.c:61-92 (dc_patch=0, scale=10) and c/fft-batch-broad.c:63-118
 * (dc_patch=1, scale=5): raw HackRF bytes -> (b+128)%256 -> (-1)^n * v/256.0
 * -> DFT -> pix = clamp_u8((int)(10*log10(re^2+im^2+1e-20)*scale)).
 * cplx (may be NULL) receives the n complex bins (2n doubles) that the
 * reference keeps in fft_history. */
void orc_batch_row(int n, const uint8_t *raw, double scale, int dc_patch,
                   uint8_t *pix, double *cplx);
/* c/fft-batch-broad.c:80-98 -- mean |X| over `rows` complex rows of width n. */
double orc_broad_gate_mean(const double *cplx_rows, int n, int rows);

/* c/fft-stitch-broad.c:28-36,66-88 -- max-blend n_images gray images (rows x width) side by
 * side at column offsets i*width_step; out is rows x (width + (n_images-1)*width_step). */
void orc_stitch(const uint8_t *images, int n_images, int rows, int width, int width_step, uint8_t *out);

/* c/fft.c:45-56,181-197 -- 512-pt scope: bytes/255.0 (no offset conversion),
 * half-swap fftshift; "as intended" variant (1/N on re and im, power from the
 * same bin); the upstream precedence bug (*pt[1] == pt[1][0]) is described in
 * DESIGN.md and deliberately not reproduced. out: n doubles
 * intensity*log10(power+1e-20). */
void orc_scope_row(int n, const uint8_t *raw, double intensity, double *out);

/* General engine definition (SURVEY.md section 8(a) net definition + 8(d)
 * extensions), one frame, float64:
 *   in_format U8:  v = b            S8: v = (b+128)%256      (then)
 *   window NONE:   x[n] = (-1)^n * v/256                      (reference, uncentred)
 *   window HANN:   x[n] = (-1)^n * w[n]*(v-128)/256, w[n] = 0.5*(1-cos(2 pi n/N))
 *   F32C/F64C:     x[n] = (-1)^n * w[n] * sample[n]           (w = 1 if NONE)
 *   X = DFT(x);  MAG: |X|;  LOGMAG: 10*log10(|X|^2+1e-20);
 *   PIX: clamp_u8((int)(LOGMAG*scale));  DC_PATCH: out[N/2] = out[N/2-1].
 * out is double[n] for MAG/LOGMAG and uint8_t[n] for PIX. */
void orc_engine_frame(int n, int in_format, int window, int out_mode, double pix_scale,
                      int flags, const void *frame, void *out);

/* Batched driver used by bench.py's CPU baseline: frames at stride `hop`
 * samples, out rows of n, split over `threads` pthreads.  Output is float32
 * (MAG/LOGMAG) or uint8 (PIX), like the engine's device output. */
void orc_engine_batch(int n, int hop, int in_format, int window, int out_mode,
                      double pix_scale, int flags, const void *in, long n_frames,
                      void *out, int threads);

#ifdef __cplusplus
}
#endif
#endif

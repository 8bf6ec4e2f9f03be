/*
 * stubs_impl.c -- link-time stand-ins for the third-party libraries the
 * reference's src/nrf.c needs but this image lacks (FFTW3, libhackrf,
 * librtlsdr, OpenAL).  TEST INFRASTRUCTURE ONLY: linked into
 * oracle/_ref/libnrf_ref.so together with the reference's own, unmodified
 * src/nrf.c + src/nut.c (compiled from /root/reference where they lie).
 *
 * - SDR opens fail  -> the reference takes its own file-replay path.
 * - OpenAL is inert -> the audio player is never used by the FFT path.
 * - FFTW's forward DFT is computed in float64 by a radix-2 FFT (power-of-two
 *   n) or by direct summation (other n).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fftw3.h"
#include "libhackrf/hackrf.h"
#include "rtl-sdr.h"
#include "AL/al.h"
#include "AL/alc.h"

/* ---------------- FFTW role ---------------- */
struct orc_fftw_plan_s {
    int n, log2n, sign;
    double *in, *out;
    double *tw;        /* n/2 complex twiddles exp(sign * 2 pi i k / n) */
    unsigned *rev;     /* bit-reversal table */
};

void *fftw_malloc(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
    return p;
}
void fftw_free(void *p) { free(p); }

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
    (void)flags;
    struct orc_fftw_plan_s *p = (struct orc_fftw_plan_s *)calloc(1, sizeof(*p));
    p->n = n; p->sign = sign; p->in = (double *)in; p->out = (double *)out;
    int l = 0;
    while ((1 << l) < n) l++;
    p->log2n = ((1 << l) == n) ? l : -1;
    if (p->log2n >= 0 && n > 1) {
        p->tw = (double *)malloc(sizeof(double) * (size_t)n);
        for (int k = 0; k < n / 2; k++) {
            long double a = (long double)sign * 2.0L * 3.14159265358979323846264338327950288L * k / n;
            p->tw[2 * k] = (double)cosl(a);
            p->tw[2 * k + 1] = (double)sinl(a);
        }
        p->rev = (unsigned *)malloc(sizeof(unsigned) * (size_t)n);
        for (int i = 0; i < n; i++) {
            unsigned r = 0, v = (unsigned)i;
            for (int b = 0; b < l; b++) { r = (r << 1) | (v & 1u); v >>= 1; }
            p->rev[i] = r;
        }
    }
    return p;
}

void fftw_execute(const fftw_plan p) {
    int n = p->n;
    const double *in = p->in;
    double *out = p->out;
    if (p->log2n < 0) {
        for (int k = 0; k < n; k++) {
            long double sr = 0, si = 0;
            for (int j = 0; j < n; j++) {
                long long m = ((long long)j * k) % n;
                long double a = (long double)p->sign * 2.0L * 3.14159265358979323846264338327950288L * m / n;
                long double c = cosl(a), s = sinl(a);
                sr += in[2 * j] * c - in[2 * j + 1] * s;
                si += in[2 * j] * s + in[2 * j + 1] * c;
            }
            out[2 * k] = (double)sr; out[2 * k + 1] = (double)si;
        }
        return;
    }
    if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
    for (int i = 0; i < n; i++) {
        unsigned r = p->rev[i];
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int base = 0; base < n; base += len) {
            double *a = out + 2 * base, *b = a + 2 * half;
            for (int j = 0; j < half; j++) {
                double wr = p->tw[2 * j * step], wi = p->tw[2 * j * step + 1];
                double tr = b[2 * j] * wr - b[2 * j + 1] * wi;
                double ti = b[2 * j] * wi + b[2 * j + 1] * wr;
                b[2 * j] = a[2 * j] - tr; b[2 * j + 1] = a[2 * j + 1] - ti;
                a[2 * j] += tr;           a[2 * j + 1] += ti;
            }
        }
    }
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    free(p->tw); free(p->rev); free(p);
}

/* ---------------- libhackrf role: no hardware ---------------- */
int hackrf_init(void) { return -1; }
int hackrf_exit(void) { return 0; }
int hackrf_open(hackrf_device **d) { if (d) *d = NULL; return -1; }
int hackrf_close(hackrf_device *d) { (void)d; return 0; }
int hackrf_set_freq(hackrf_device *d, uint64_t f) { (void)d; (void)f; return -1; }
int hackrf_set_sample_rate(hackrf_device *d, double f) { (void)d; (void)f; return -1; }
int hackrf_set_amp_enable(hackrf_device *d, uint8_t v) { (void)d; (void)v; return -1; }
int hackrf_set_lna_gain(hackrf_device *d, uint32_t v) { (void)d; (void)v; return -1; }
int hackrf_set_vga_gain(hackrf_device *d, uint32_t v) { (void)d; (void)v; return -1; }
int hackrf_start_rx(hackrf_device *d, hackrf_sample_block_cb_fn cb, void *ctx) { (void)d; (void)cb; (void)ctx; return -1; }
int hackrf_stop_rx(hackrf_device *d) { (void)d; return 0; }

/* ---------------- librtlsdr role: no hardware ---------------- */
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index) { (void)index; if (dev) *dev = NULL; return -1; }
int rtlsdr_close(rtlsdr_dev_t *dev) { (void)dev; return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t f) { (void)dev; (void)f; return -1; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t r) { (void)dev; (void)r; return -1; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int m) { (void)dev; (void)m; return -1; }
int rtlsdr_set_agc_mode(rtlsdr_dev_t *dev, int on) { (void)dev; (void)on; return -1; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev) { (void)dev; return -1; }
int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read) { (void)dev; (void)buf; (void)len; if (n_read) *n_read = 0; return -1; }

/* ---------------- OpenAL role: inert ---------------- */
ALenum alGetError(void) { return AL_NO_ERROR; }
void alGenSources(ALsizei n, ALuint *s) { for (int i = 0; i < n; i++) s[i] = (ALuint)(i + 1); }
void alGenBuffers(ALsizei n, ALuint *b) { for (int i = 0; i < n; i++) b[i] = (ALuint)(i + 1); }
void alDeleteBuffers(ALsizei n, const ALuint *b) { (void)n; (void)b; }
void alBufferData(ALuint b, ALenum f, const ALvoid *d, ALsizei s, ALsizei fr) { (void)b; (void)f; (void)d; (void)s; (void)fr; }
void alSourcei(ALuint s, ALenum p, ALint v) { (void)s; (void)p; (void)v; }
void alSourcef(ALuint s, ALenum p, ALfloat v) { (void)s; (void)p; (void)v; }
void alGetSourcei(ALuint s, ALenum p, ALint *v) { (void)s; (void)p; if (v) *v = 0; }
void alGetSourceiv(ALuint s, ALenum p, ALint *v) { (void)s; (void)p; if (v) *v = 0; }
void alSourceQueueBuffers(ALuint s, ALsizei n, const ALuint *b) { (void)s; (void)n; (void)b; }
void alSourceUnqueueBuffers(ALuint s, ALsizei n, ALuint *b) { (void)s; (void)n; (void)b; }
void alSourcePlay(ALuint s) { (void)s; }
ALCdevice *alcOpenDevice(const ALCchar *name) { (void)name; return NULL; }
int alcCloseDevice(ALCdevice *d) { (void)d; return 1; }
ALCcontext *alcCreateContext(ALCdevice *d, const ALCint *a) { (void)d; (void)a; return NULL; }
int alcMakeContextCurrent(ALCcontext *c) { (void)c; return 1; }
void alcDestroyContext(ALCcontext *c) { (void)c; }

/*
 * oracle.c -- CPU restatement of frequensea's IQ -> spectrum path (float64).
 * TEST INFRASTRUCTURE ONLY -- see oracle.h for scope, citations and the pin.
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ */
/* src/nrf.c:95-110                                                    */
void orc_device_unpack(const uint8_t *raw, uint8_t *samples, size_t nbytes, int convert) {
    for (size_t i = 0; i < nbytes; i++) {
        uint8_t v = raw[i];
        if (convert) v = (uint8_t)((v + 128) % 256);
        samples[i] = v;
    }
}

/* ------------------------------------------------------------------ */
/* Forward DFT (the FFTW3 role).  Iterative radix-2, decimation in time,
 * twiddles evaluated once per size in long double.                     */
#define ORC_MAX_LOG2 26
static double *g_tw[ORC_MAX_LOG2 + 1];
static pthread_mutex_t g_tw_lock = PTHREAD_MUTEX_INITIALIZER;

static const double *orc_twiddles(int log2n) {
    pthread_mutex_lock(&g_tw_lock);
    if (!g_tw[log2n]) {
        int n = 1 << log2n;
        double *t = (double *)malloc(sizeof(double) * (size_t)n); /* n/2 complex */
        for (int k = 0; k < n / 2; k++) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
            t[2 * k] = (double)cosl(a);
            t[2 * k + 1] = (double)sinl(a);
        }
        g_tw[log2n] = t;
    }
    const double *r = g_tw[log2n];
    pthread_mutex_unlock(&g_tw_lock);
    return r;
}

static int orc_log2(int n) {
    int l = 0;
    while ((1 << l) < n) l++;
    return ((1 << l) == n) ? l : -1;
}

void orc_dft_naive(int n, const double *in, double *out) {
    for (int k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (int j = 0; j < n; j++) {
            /* reduce j*k mod n before the trig call: keeps the angle exact */
            long long m = ((long long)j * k) % n;
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
            long double c = cosl(a), s = sinl(a);
            sr += in[2 * j] * c - in[2 * j + 1] * s;
            si += in[2 * j] * s + in[2 * j + 1] * c;
        }
        out[2 * k] = (double)sr;
        out[2 * k + 1] = (double)si;
    }
}

void orc_dft_forward(int n, const double *in, double *out) {
    int l2 = orc_log2(n);
    if (l2 < 0) { orc_dft_naive(n, in, out); return; }
    if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
    const double *tw = orc_twiddles(l2);
    /* bit-reversal permutation */
    for (int i = 0; i < n; i++) {
        unsigned r = 0, v = (unsigned)i;
        for (int b = 0; b < l2; b++) { r = (r << 1) | (v & 1u); v >>= 1; }
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int j = 0; j < half; j++) {
                double wr = tw[2 * j * step], wi = tw[2 * j * step + 1];
                double *a = out + 2 * (base + j), *b = out + 2 * (base + j + half);
                double tr = b[0] * wr - b[1] * wi;
                double ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr; b[1] = a[1] - ti;
                a[0] += tr;       a[1] += ti;
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* src/nrf.c:598-631                                                   */
static void magnitude_dc_patch(int n, const double *X, double *mag) {
    for (int i = 0; i < n; i++) {
        if (i == n / 2) {
            mag[i] = mag[i - 1];                      /* src/nrf.c:621-622 */
        } else {
            double fi = X[2 * i], fq = X[2 * i + 1];
            mag[i] = sqrt(fi * fi + fq * fq);         /* src/nrf.c:624-628 */
        }
    }
}

void orc_nrf_fft_frame_u8(int n, const uint8_t *samples, double *mag) {
    double *x = (double *)malloc(sizeof(double) * 4 * (size_t)n), *X = x + 2 * n;
    for (int ii = 0; ii < n; ii++) {
        double di = samples[2 * ii] / 256.0;          /* src/nrf.c:605-606 */
        double dq = samples[2 * ii + 1] / 256.0;
        double sgn = (ii & 1) ? -1.0 : 1.0;           /* powf(-1, ii), :611-612 */
        x[2 * ii] = sgn * di;
        x[2 * ii + 1] = sgn * dq;
    }
    orc_dft_forward(n, x, X);                         /* fftw_execute, :615 */
    magnitude_dc_patch(n, X, mag);
    free(x);
}

void orc_nrf_fft_frame_f64(int n, const double *samples, double *mag) {
    double *x = (double *)malloc(sizeof(double) * 4 * (size_t)n), *X = x + 2 * n;
    for (int ii = 0; ii < n; ii++) {
        double sgn = (ii & 1) ? -1.0 : 1.0;
        x[2 * ii] = sgn * samples[2 * ii];            /* src/nrf.c:608-609 */
        x[2 * ii + 1] = sgn * samples[2 * ii + 1];
    }
    orc_dft_forward(n, x, X);
    magnitude_dc_patch(n, X, mag);
    free(x);
}

/* src/nrf.c:617 (overlapping memcpy; memmove is the evident intent) */
void orc_history_push(double *history, int n, int rows, const double *row0) {
    if (rows > 1) memmove(history + n, history, sizeof(double) * (size_t)n * (size_t)(rows - 1));
    memcpy(history, row0, sizeof(double) * (size_t)n);
}

/* src/nrf.c:569-596 */
void orc_fft_shift(double *history, int n, int rows, double d) {
    int shift_pixels = (int)round(n / d);
    if (shift_pixels == 0) return;
    if (abs(shift_pixels) >= n) {
        memset(history, 0, sizeof(double) * (size_t)n * (size_t)rows);
        return;
    }
    for (int y = 0; y < rows; y++) {
        double *row = history + (size_t)y * n;
        if (shift_pixels > 0) {
            for (int x = 0; x < n - shift_pixels; x++) row[x] = row[x + shift_pixels];
            for (int x = n - shift_pixels; x < n; x++) row[x] = 0;
        } else {
            for (int x = n - 1; x >= -shift_pixels; x--) row[x] = row[x + shift_pixels];
            for (int x = 0; x < -shift_pixels; x++) row[x] = 0;
        }
    }
}

/* ------------------------------------------------------------------ */
/* c/fft-batch.c:33-35 */
static uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* c/fft-batch.c:61-92, c/fft-batch-broad.c:63-118 */
void orc_batch_row(int n, const uint8_t *raw, double scale, int dc_patch,
                   uint8_t *pix, double *cplx) {
    double *x = (double *)malloc(sizeof(double) * 4 * (size_t)n), *X = x + 2 * n;
    for (int ii = 0; ii < n; ii++) {
        int vi = (raw[2 * ii] + 128) % 256;           /* c/fft-batch.c:63-64 */
        int vq = (raw[2 * ii + 1] + 128) % 256;
        double sgn = (ii & 1) ? -1.0 : 1.0;
        x[2 * ii] = sgn * vi / 256.0;                 /* :65-66 */
        x[2 * ii + 1] = sgn * vq / 256.0;
    }
    orc_dft_forward(n, x, X);
    if (cplx) memcpy(cplx, X, sizeof(double) * 2 * (size_t)n);
    for (int k = 0; k < n; k++) {
        double ci = X[2 * k], cq = X[2 * k + 1];
        double pwr = ci * ci + cq * cq;
        double pwr_dbfs = 10.0 * log10(pwr + 1.0e-20);  /* c/fft-batch.c:89 */
        pwr_dbfs = pwr_dbfs * scale;                    /* :90 (x10) / broad :113 (x5) */
        uint8_t v = clamp_u8((int)pwr_dbfs);
        if (dc_patch && k == n / 2) v = pix[k - 1];     /* c/fft-batch-broad.c:115-117 */
        pix[k] = v;
    }
    free(x);
}

/* c/fft-batch-broad.c:80-98 */
double orc_broad_gate_mean(const double *cplx_rows, int n, int rows) {
    double total = 0;
    long size = (long)n * rows;
    for (long i = 0; i < size; i++) {
        double ci = cplx_rows[2 * i], cq = cplx_rows[2 * i + 1];
        total += sqrt(ci * ci + cq * cq);
    }
    return total / (double)size;
}

/* c/fft-stitch-broad.c:28-36 (img_gray_copy) and :66-88 (placement loop) */
void orc_stitch(const uint8_t *images, int n_images, int rows, int width, int width_step, uint8_t *out) {
    int out_width = width + (n_images - 1) * width_step;
    memset(out, 0, (size_t)rows * out_width);                 /* calloc, :63 */
    int x = 0;
    for (int i = 0; i < n_images; i++) {
        const uint8_t *src = images + (size_t)i * rows * width;
        for (int y = 0; y < rows; y++)
            for (int j = 0; j < width; j++) {
                uint8_t *d = out + (size_t)y * out_width + x + j;
                uint8_t s = src[(size_t)y * width + j];
                *d = *d > s ? *d : s;                          /* max_u8, :22-24 */
            }
        x += width_step;
    }
}

/* c/fft.c:45-56,181-197 (as intended, see oracle.h) */
void orc_scope_row(int n, const uint8_t *raw, double intensity, double *out) {
    double *x = (double *)malloc(sizeof(double) * 4 * (size_t)n), *X = x + 2 * n;
    for (int ii = 0; ii < n; ii++) {
        x[2 * ii] = raw[2 * ii] / 255.0;              /* c/fft.c:50-51 */
        x[2 * ii + 1] = raw[2 * ii + 1] / 255.0;
    }
    orc_dft_forward(n, x, X);
    for (int i = 0; i < n; i++) {
        int src = (i < n / 2) ? (n / 2 + i) : (i - n / 2);   /* c/fft.c:185-190 */
        double re = X[2 * src] / (double)n, im = X[2 * src + 1] / (double)n;
        out[i] = intensity * log10(re * re + im * im + 1.0e-20);
    }
    free(x);
}

/* ------------------------------------------------------------------ */
/* General engine definition                                           */
void orc_engine_frame(int n, int in_format, int window, int out_mode, double pix_scale,
                      int flags, const void *frame, void *out) {
    double *x = (double *)malloc(sizeof(double) * 5 * (size_t)n), *X = x + 2 * n, *val = x + 4 * n;
    for (int ii = 0; ii < n; ii++) {
        double di, dq;
        double w = 1.0;
        if (window == ORC_WIN_HANN) w = 0.5 * (1.0 - cos(2.0 * M_PI * (double)ii / (double)n));
        if (in_format == ORC_IN_U8 || in_format == ORC_IN_S8) {
            const uint8_t *b = (const uint8_t *)frame;
            int vi = b[2 * ii], vq = b[2 * ii + 1];
            if (in_format == ORC_IN_S8) { vi = (vi + 128) % 256; vq = (vq + 128) % 256; }
            if (window == ORC_WIN_NONE) { di = vi / 256.0; dq = vq / 256.0; }
            else { di = w * ((vi - 128) / 256.0); dq = w * ((vq - 128) / 256.0); }
        } else if (in_format == ORC_IN_F32C) {
            const float *f = (const float *)frame;
            di = w * (double)f[2 * ii]; dq = w * (double)f[2 * ii + 1];
        } else {
            const double *f = (const double *)frame;
            di = w * f[2 * ii]; dq = w * f[2 * ii + 1];
        }
        double sgn = (ii & 1) ? -1.0 : 1.0;
        x[2 * ii] = sgn * di;
        x[2 * ii + 1] = sgn * dq;
    }
    orc_dft_forward(n, x, X);
    for (int k = 0; k < n; k++) {
        double pwr = X[2 * k] * X[2 * k] + X[2 * k + 1] * X[2 * k + 1];
        if (out_mode == ORC_OUT_MAG) val[k] = sqrt(pwr);
        else val[k] = 10.0 * log10(pwr + 1.0e-20);
    }
    if (out_mode == ORC_OUT_PIX) {
        uint8_t *p = (uint8_t *)out;
        for (int k = 0; k < n; k++) {
            uint8_t v = clamp_u8((int)(val[k] * pix_scale));
            if ((flags & ORC_FLAG_DC_PATCH) && k == n / 2 && k > 0) v = p[k - 1];
            p[k] = v;
        }
    } else {
        double *o = (double *)out;
        for (int k = 0; k < n; k++) o[k] = val[k];
        if ((flags & ORC_FLAG_DC_PATCH) && n >= 2) o[n / 2] = o[n / 2 - 1];
    }
    free(x);
}

/* ------------------------------------------------------------------ */
typedef struct {
    int n, hop, in_format, window, out_mode, flags;
    double pix_scale;
    const uint8_t *in;
    uint8_t *out;
    long f0, f1;
} orc_job;

static size_t sample_bytes(int in_format) {
    return in_format == ORC_IN_F64C ? 16 : in_format == ORC_IN_F32C ? 8 : 2;
}

static void *orc_worker(void *p) {
    orc_job *j = (orc_job *)p;
    int n = j->n;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    size_t sb = sample_bytes(j->in_format);
    for (long f = j->f0; f < j->f1; f++) {
        const uint8_t *frame = j->in + (size_t)f * (size_t)j->hop * sb;
        if (j->out_mode == ORC_OUT_PIX) {
            orc_engine_frame(n, j->in_format, j->window, j->out_mode, j->pix_scale, j->flags,
                             frame, j->out + (size_t)f * n);
        } else {
            orc_engine_frame(n, j->in_format, j->window, j->out_mode, j->pix_scale, j->flags,
                             frame, tmp);
            float *o = (float *)j->out + (size_t)f * n;
            for (int k = 0; k < n; k++) o[k] = (float)tmp[k];
        }
    }
    free(tmp);
    return NULL;
}

void orc_engine_batch(int n, int hop, int in_format, int window, int out_mode,
                      double pix_scale, int flags, const void *in, long n_frames,
                      void *out, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    if ((long)threads > n_frames) threads = n_frames > 0 ? (int)n_frames : 1;
    (void)orc_twiddles(orc_log2(n) >= 0 ? orc_log2(n) : 0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    orc_job *jobs = (orc_job *)malloc(sizeof(orc_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        orc_job j = { n, hop, in_format, window, out_mode, flags, pix_scale,
                      (const uint8_t *)in, (uint8_t *)out,
                      n_frames * t / threads, n_frames * (t + 1) / threads };
        jobs[t] = j;
        pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

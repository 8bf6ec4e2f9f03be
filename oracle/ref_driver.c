/*
 * ref_driver.c -- drives the REFERENCE's own public API (src/nrf.h:138-142,
 * src/nut.h:32-45) over a batch of frames, for the CPU baseline / reference arm
 * of bench.py and for cross-checks.  TEST INFRASTRUCTURE ONLY; linked into
 * oracle/_ref/libnrf_ref.so next to the reference's unmodified src/nrf.c and
 * src/nut.c.  Contains no transform code of its own: every frame goes
 *   nut_buffer_new_u8 -> nrf_fft_process -> nrf_fft_get_buffer -> nut_buffer_free
 * exactly as lua/fft.lua:42-45 drives it (one nrf_fft per worker thread,
 * history of one row so that the per-frame 8 MiB scroll of src/nrf.c:617 is
 * not what is being timed).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "nrf.h"
#include "nut.h"

typedef struct {
    const uint8_t *raw;   /* capture bytes, two's-complement IQ (convert=1) or offset-binary (0) */
    int n, convert;
    long hop, f0, f1;
    float *out;           /* rows of n float32 (the consumer's precision, src/ngl.c:229-239) */
} ref_job;

static void *ref_worker(void *arg) {
    ref_job *j = (ref_job *)arg;
    const int n = j->n;
    nrf_fft *fft = nrf_fft_new(n, 1);
    uint8_t *tmp = (uint8_t *)malloc(2 * (size_t)n);
    for (long f = j->f0; f < j->f1; f++) {
        const uint8_t *src = j->raw + 2 * (size_t)f * (size_t)j->hop;
        /* the device layer's byte convention, src/nrf.c:99-107 */
        for (int i = 0; i < 2 * n; i++) tmp[i] = j->convert ? (uint8_t)((src[i] + 128) % 256) : src[i];
        nut_buffer *in = nut_buffer_new_u8(n, 2, tmp);
        nrf_fft_process(fft, in);
        nut_buffer *res = nrf_fft_get_buffer(fft);
        float *o = j->out + (size_t)f * n;
        for (int k = 0; k < n; k++) o[k] = (float)res->data.f64[k];
        nut_buffer_free(res);
        nut_buffer_free(in);
    }
    free(tmp);
    nrf_fft_free(fft);
    return NULL;
}

void ref_run_frames(const uint8_t *raw, int n, long hop, long frames, int convert, float *out, int threads) {
    if (threads < 1) threads = 1;
    if ((long)threads > frames) threads = frames > 0 ? (int)frames : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    ref_job *jobs = (ref_job *)malloc(sizeof(ref_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        ref_job j = { raw, n, convert, hop, frames * t / threads, frames * (t + 1) / threads, out };
        jobs[t] = j;
        pthread_create(&th[t], NULL, ref_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

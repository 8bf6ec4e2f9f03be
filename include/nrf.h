/*
 * nrf.h -- the reference's SDR operator API for the IQ -> spectrum path,
 * backed by the B200 engine (include/fsea.h) instead of FFTW3.
 *
 * Mirror of the hot-path subset of the reference's src/nrf.h (SURVEY.md
 * section 8(a),(b)): the block graph header (src/nrf.h:29-52), the
 * file-replay device (src/nrf.h:56-108, src/nrf.c:95-131,162-170,256-357),
 * the FFT block (src/nrf.h:128-142, src/nrf.c:557-642) and the frequency
 * shifter that lua/fft-shifted.lua chains in front of it (src/nrf.h:206-221,
 * src/nrf.c:814-874).  Names, argument meaning, ownership (every returned
 * nut_buffer* is a fresh caller-owned copy) and fail-stop error behaviour
 * follow the reference, so its Lua binding (src/main.cpp:616-717,775-811) and
 * its lua/fft*.lua scripts work against this header unchanged.
 *
 * Out of scope (not declared): live RTL-SDR / HackRF USB drivers, downsampler,
 * interpolator, IQ drawing, signal detector, demodulators, OpenAL player.
 */
#ifndef NRF_H
#define NRF_H

#include <pthread.h>
#include <stdint.h>

#include "nut.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NRF_BUFFER_SIZE_BYTES (16 * 16384)   /* one device block, src/nrf.h:19 */
#define NRF_SAMPLES_LENGTH 131072            /* complex samples per block, src/nrf.h:20 */
#define DEFAULT_FFT_SIZE 128
#define DEFAULT_FFT_HISTORY_SIZE 128

/* ---- block graph (src/nrf.h:25-52, src/nrf.c:24-50) -------------------- */
#define NRF_BLOCK_MAX_OUTPUTS 10

typedef enum {
    NRF_BLOCK_SOURCE = 1,
    NRF_BLOCK_GENERIC = 2,
    NRF_BLOCK_SINK = 3
} nrf_block_type;

typedef struct nrf_block nrf_block;
typedef void (*nrf_block_process_fn)(nrf_block *block, nut_buffer *buffer);
typedef nut_buffer *(*nrf_block_result_fn)(void *block);

/* Must stay the FIRST member of every block object: the Lua binding reads
 * block->type through the object's pointer (src/main.cpp:616-624). */
struct nrf_block {
    nrf_block_type type;
    nrf_block_process_fn process_fn;
    nrf_block_result_fn result_fn;
    int n_outputs;
    void *outputs[NRF_BLOCK_MAX_OUTPUTS];
};
#define NRF_BLOCK nrf_block block

void nrf_block_init(nrf_block *block, nrf_block_type type,
                    nrf_block_process_fn process_fn, nrf_block_result_fn result_fn);
void nrf_block_connect(nrf_block *input, nrf_block *output);
void nrf_block_process(nrf_block *block, nut_buffer *buffer);

/* ---- device: file replay (src/nrf.h:54-108) ----------------------------- */
typedef struct {
    int sample_rate;
    double freq_mhz;
    const char *data_file;
} nrf_device_config;

typedef enum {
    NRF_DEVICE_DUMMY = 0,
    NRF_DEVICE_RTLSDR = 1,   /* value kept for ABI; no USB driver in this build */
    NRF_DEVICE_HACKRF = 2
} nrf_device_type;

typedef struct nrf_device nrf_device;
typedef void (*nrf_device_decode_cb_fn)(nrf_device *device, void *ctx);

struct nrf_device {
    NRF_BLOCK;
    nrf_device_type device_type;
    void *device;
    int sample_rate;                 /* read by the Lua binding, src/main.cpp:642-644 */
    nrf_device_decode_cb_fn decode_cb_fn;
    void *decode_cb_ctx;
    pthread_t receive_thread;        /* 60 Hz block replay, src/nrf.c:162-170 */
    pthread_mutex_t data_mutex;
    int receiving;
    int paused;
    uint8_t *receive_buffer;         /* whole capture file */
    int dummy_block_length;          /* file size / NRF_BUFFER_SIZE_BYTES */
    int dummy_block_index;
    uint8_t samples[NRF_BUFFER_SIZE_BYTES];   /* current block, offset-binary u8 */
    /* (not in the reference; appended, the layout above is the reference's) */
    pthread_mutex_t graph_mutex;     /* one thread at a time publishes a block and walks the graph behind it */
    int manual;                      /* set by nrf_device_publish_now: the replay thread stops publishing */
};

nrf_device *nrf_device_new(double freq_mhz, const char *data_file);
nrf_device *nrf_device_new_with_config(const nrf_device_config config);
double nrf_device_set_frequency(nrf_device *device, double freq_mhz);
void nrf_device_set_decode_handler(nrf_device *device, nrf_device_decode_cb_fn fn, void *ctx);
void nrf_device_set_paused(nrf_device *device, int paused);
void nrf_device_step(nrf_device *device);
nut_buffer *nrf_device_get_samples_buffer(nrf_device *device);
void nrf_device_free(nrf_device *device);
/* extension: publish the current block now, on the caller's thread; from the first call on the 60 Hz replay
 * thread no longer publishes (deterministic replay of a paused device: nrf_device_set_paused + nrf_device_step) */
void nrf_device_publish_now(nrf_device *device);

/* ---- FFT analysis (src/nrf.h:126-142) ----------------------------------- */
typedef struct {
    NRF_BLOCK;
    int fft_size;
    int fft_history_size;
    double *buffer;      /* OPAQUE here: the host mirror of a device-resident RING of fft_history_size rows (not the
                            reference's "row 0 newest" image, and only as fresh as the last nrf_fft_get_buffer).
                            Read the history through nrf_fft_get_buffer, which returns the reference's layout. */
    void *engine;        /* opaque; stands where the reference kept
                            fftw_complex *fft_in, *fft_out; fftw_plan fft_plan */
} nrf_fft;

nrf_fft *nrf_fft_new(int fft_size, int fft_history_size);
void nrf_fft_shift(nrf_fft *fft, double d);
void nrf_fft_process(nrf_fft *fft, nut_buffer *buffer);
nut_buffer *nrf_fft_get_buffer(nrf_fft *fft);
void nrf_fft_free(nrf_fft *fft);
/* extension: the history where it lives -- fft_history_size x fft_size float64 rows in DEVICE memory, ring order;
 * *head = the row of the newest spectrum.  Valid until the next nrf_fft_shift / nrf_fft_free. */
void nrf_fft_get_history_device(nrf_fft *fft, const double **d_ring, int *head);

/* ---- frequency shifter (src/nrf.h:206-221): host pass-through ------------ */
typedef struct {
    NRF_BLOCK;
    int freq_offset;
    int sample_rate;
    double cosine;
    double sine;
    nut_buffer *buffer;
} nrf_freq_shifter;

nrf_freq_shifter *nrf_freq_shifter_new(int freq_offset, int sample_rate);
void nrf_freq_shifter_process(nrf_freq_shifter *shifter, nut_buffer *buffer);
nut_buffer *nrf_freq_shifter_get_buffer(nrf_freq_shifter *shifter);
void nrf_freq_shifter_free(nrf_freq_shifter *shifter);

/* ---- FIR low-pass + IQ filter (src/nrf.h:144-177, src/nrf.c:644-776) ------
 * the operator in front of the path in lua/iq-tex-filtered.lua:43-46.
 * Coefficients, the carried-over sample line and nrf_fir_filter_get() are host code as in the
 * reference; nrf_iq_filter_get_buffer() -- length x taps multiply-adds per block -- runs on the
 * device (fsea_fir_iq_host) and is bit-identical to the reference's loop. */
typedef struct {
    int length;            /* taps */
    double *coefficients;
    int offset;            /* length - 1 samples carried over between blocks */
    int center;
    int samples_length;
    double *samples;       /* offset carried-over samples, then the current block */
} nrf_fir_filter;

double *nrf_fir_get_low_pass_coefficients(int sample_rate, int half_ampl_freq, int length);
nrf_fir_filter *nrf_fir_filter_new(int sample_rate, int half_ampl_freq, int length);
void nrf_fir_filter_load(nrf_fir_filter *filter, double *samples, int length);
double nrf_fir_filter_get(nrf_fir_filter *filter, int index);
void nrf_fir_filter_free(nrf_fir_filter *filter);

typedef struct {
    NRF_BLOCK;
    nrf_fir_filter *filter_i;
    nrf_fir_filter *filter_q;
    int samples_length;
    double *samples_i;
    double *samples_q;
} nrf_iq_filter;

nrf_iq_filter *nrf_iq_filter_new(int sample_rate, int half_ampl_freq, int kernel_length);
void nrf_iq_filter_process(nrf_iq_filter *filter, nut_buffer *buffer);
nut_buffer *nrf_iq_filter_get_buffer(nrf_iq_filter *f);
void nrf_iq_filter_free(nrf_iq_filter *filter);

#ifdef __cplusplus
}
#endif
#endif /* NRF_H */

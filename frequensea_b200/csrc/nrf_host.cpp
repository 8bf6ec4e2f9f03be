// nrf_host.cpp -- the reference's nrf_* operator API for the IQ -> spectrum
// path (include/nrf.h), as a one-frame client of the B200 engine.
//
// What is kept from the reference (behaviour, not code): the push-style block
// graph (src/nrf.c:24-50), the file-replay device with its 60 Hz producer
// thread and (b+128)%256 byte convention (src/nrf.c:95-131,162-170,256-357),
// nrf_fft's observable semantics -- first N samples of the buffer, x = u8/256
// or raw f64, (-1)^n shift, linear magnitude, row[N/2] = row[N/2-1], N x H
// float64 history with row 0 newest, nrf_fft_shift (src/nrf.c:557-642) -- and
// fail-stop errors (message on stderr + exit, src/nrf.c:54-78).
//
// What is replaced: fftw_plan_dft_1d / fftw_execute (src/nrf.c:564,615) and
// the per-sample powf loop (src/nrf.c:601-614) by fsea_exec_host(), i.e. the
// sm_100a kernels.  There is no CPU FFT here: without a CUDA device
// nrf_fft_new() fails loudly.
#include "nrf.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include <mutex>

#include "fsea.h"

// ---------------------------------------------------------------- fail-stop
static void nrf_die(const char* what, int code) {
    fprintf(stderr, "NRF fatal error: %s: %s", what, fsea_strerror(code));
    if (code == FSEA_ERR_CUDA) fprintf(stderr, " (%s)", fsea_last_cuda_error());
    fprintf(stderr, "\n");
    exit(EXIT_FAILURE);
}

// ---------------------------------------------------------------- block graph
extern "C" void nrf_block_init(nrf_block* block, nrf_block_type type,
                               nrf_block_process_fn process_fn, nrf_block_result_fn result_fn) {
    assert(block->n_outputs == 0);          // objects are calloc'ed (src/nrf.c:28)
    block->type = type;
    block->process_fn = process_fn;
    block->result_fn = result_fn;
}

extern "C" void nrf_block_connect(nrf_block* input, nrf_block* output) {
    assert(input->n_outputs < NRF_BLOCK_MAX_OUTPUTS);
    input->outputs[input->n_outputs++] = output;
}

// process this block, then hand its result to every connected output; the
// intermediate result belongs to this call and is released here (src/nrf.c:37-50)
extern "C" void nrf_block_process(nrf_block* block, nut_buffer* buffer) {
    if (block->process_fn) block->process_fn(block, buffer);
    if (block->n_outputs <= 0) return;
    nut_buffer* result = block->result_fn(block);
    for (int o = 0; o < block->n_outputs; o++) nrf_block_process((nrf_block*)block->outputs[o], result);
    nut_buffer_free(result);
}

// ---------------------------------------------------------------- device (file replay)
static const int kReplaySampleRate = 5000000;     // what the reference reports for replay, src/nrf.c:254

// One device block arrives: convert two's-complement capture bytes to the
// offset-binary u8 convention of nut buffers, publish, run the graph (src/nrf.c:95-131).
static void device_ingest_block(nrf_device* dev, const uint8_t* raw) {
    if (!dev->receiving) return;
    // one thread at a time publishes a block and walks the graph behind it: the 60 Hz replay thread and a
    // caller of nrf_device_publish_now would otherwise both drive the same nrf_fft / filter objects
    pthread_mutex_lock(&dev->graph_mutex);
    pthread_mutex_lock(&dev->data_mutex);
    const bool flip = dev->device_type != NRF_DEVICE_RTLSDR;
    for (int i = 0; i < NRF_BUFFER_SIZE_BYTES; i++) dev->samples[i] = flip ? (uint8_t)(raw[i] ^ 0x80u) : raw[i];
    pthread_mutex_unlock(&dev->data_mutex);
    if (dev->decode_cb_fn) dev->decode_cb_fn(dev, dev->decode_cb_ctx);
    if (dev->receiving) nrf_block_process(&dev->block, NULL);
    pthread_mutex_unlock(&dev->graph_mutex);
}

static void* device_replay_thread(void* arg) {
    nrf_device* dev = (nrf_device*)arg;
    while (dev->receiving) {
        // (manual: the caller publishes blocks itself with nrf_device_publish_now -- deterministic replay)
        if (!dev->manual) device_ingest_block(dev, dev->receive_buffer + (size_t)dev->dummy_block_index * NRF_BUFFER_SIZE_BYTES);
        if (!dev->paused) {                       // src/nrf.c:153-160
            if (++dev->dummy_block_index >= dev->dummy_block_length) dev->dummy_block_index = 0;
        }
        nut_sleep_milliseconds(1000 / 60);
    }
    return NULL;
}

extern "C" nrf_device* nrf_device_new_with_config(const nrf_device_config config) {
    nrf_device* dev = (nrf_device*)calloc(1, sizeof(nrf_device));
    nrf_block_init(&dev->block, NRF_BLOCK_SOURCE, NULL, (nrf_block_result_fn)nrf_device_get_samples_buffer);
    pthread_mutex_init(&dev->data_mutex, NULL);
    pthread_mutex_init(&dev->graph_mutex, NULL);
    dev->device_type = NRF_DEVICE_DUMMY;          // no USB SDR drivers in this build
    dev->sample_rate = kReplaySampleRate;
    const char* path = config.data_file;
    fprintf(stderr, "WARN nrf_device_new: no SDR driver in this build. Replaying data file %s\n", path ? path : "(none)");
    FILE* fp = path ? fopen(path, "rb") : NULL;
    long size = 0;
    if (fp) {
        fseek(fp, 0L, SEEK_END);
        size = ftell(fp);
        rewind(fp);
    }
    if (fp && size >= NRF_BUFFER_SIZE_BYTES) {
        dev->receive_buffer = (uint8_t*)calloc((size_t)size, 1);
        dev->dummy_block_length = (int)(size / NRF_BUFFER_SIZE_BYTES);
        if (fread(dev->receive_buffer, (size_t)size, 1, fp) != 1) fprintf(stderr, "WARN nrf_device_new: short read on %s\n", path);
    } else {
        // Missing file -> one silent block, as the reference (src/nrf.c:273-277).  A file
        // shorter than one block would give the reference a zero-length ring and an
        // out-of-bounds read (src/nrf.c:268); it is padded to one block here instead.
        if (path && !fp) fprintf(stderr, "WARN nrf_device_new: Couldn't open %s. Using empty buffer.\n", path);
        dev->receive_buffer = (uint8_t*)calloc(NRF_BUFFER_SIZE_BYTES, 1);
        dev->dummy_block_length = 1;
        if (fp && size > 0 && fread(dev->receive_buffer, (size_t)size, 1, fp) != 1) { /* keep zeros */ }
    }
    if (fp) fclose(fp);
    dev->dummy_block_index = 0;
    dev->receiving = 1;
    pthread_create(&dev->receive_thread, NULL, device_replay_thread, dev);
    return dev;
}

extern "C" nrf_device* nrf_device_new(double freq_mhz, const char* data_file) {
    nrf_device_config config;
    memset(&config, 0, sizeof(config));
    config.freq_mhz = freq_mhz;
    config.data_file = data_file;
    return nrf_device_new_with_config(config);
}

// replay has no tuner: the value is returned unclamped (src/nrf.c:81-90,324-334)
extern "C" double nrf_device_set_frequency(nrf_device* device, double freq_mhz) {
    (void)device;
    return freq_mhz;
}

extern "C" void nrf_device_set_decode_handler(nrf_device* device, nrf_device_decode_cb_fn fn, void* ctx) {
    device->decode_cb_fn = fn;
    device->decode_cb_ctx = ctx;
}

extern "C" void nrf_device_set_paused(nrf_device* device, int paused) { device->paused = paused; }

extern "C" void nrf_device_step(nrf_device* device) {
    if (++device->dummy_block_index >= device->dummy_block_length) device->dummy_block_index = 0;
}

// Extension (not in the reference): publish the current capture block synchronously, so a
// paused device (nrf_device_set_paused + nrf_device_step, src/nrf.c:341-350) can be replayed
// block by block without waiting for the 60 Hz producer thread.  Used by tools/fsea-lua --step.
extern "C" void nrf_device_publish_now(nrf_device* device) {
    device->manual = 1;                               // from now on only the caller publishes: the replay thread idles
    device_ingest_block(device, device->receive_buffer + (size_t)device->dummy_block_index * NRF_BUFFER_SIZE_BYTES);
}

extern "C" nut_buffer* nrf_device_get_samples_buffer(nrf_device* device) {
    pthread_mutex_lock(&device->data_mutex);
    nut_buffer* snapshot = nut_buffer_new_u8(NRF_SAMPLES_LENGTH, 2, device->samples);
    pthread_mutex_unlock(&device->data_mutex);
    return snapshot;
}

extern "C" void nrf_device_free(nrf_device* device) {
    device->receiving = 0;
    pthread_join(device->receive_thread, NULL);
    free(device->receive_buffer);
    pthread_mutex_destroy(&device->data_mutex);
    pthread_mutex_destroy(&device->graph_mutex);
    free(device);
}

// ---------------------------------------------------------------- FFT block
// Engine state behind nrf_fft::engine.
//
// History.  The reference scrolls its whole N x H image by one row on every
// frame (an 8 MiB memmove at 1024 x 1024, src/nrf.c:617) and on a retune scrolls every row sideways
// (src/nrf.c:569-596).  Here the image is a RING of float64 rows in DEVICE memory: nrf_fft_process
// uploads the N samples that matter, the transform writes its row on the device and one small kernel
// puts it in place of the oldest row -- nothing comes back and nothing is waited for; nrf_fft_shift
// is one kernel over the ring.  The host keeps a mirror of the ring (fft->buffer) that
// nrf_fft_get_buffer brings up to date by copying only the rows written since the last call, before
// it unrolls the ring into the reference's layout (row 0 newest) for the caller's fresh copy.
//
// The staging arrays play the role of the reference's persistent fft_in: a buffer shorter than N
// samples leaves the tail of the previous call in place (src/nrf.c:599-614 only writes size/2
// entries); buffers longer than the reference's 131072-sample fft_in (lua/fft-shifted.lua produces
// one) are clamped instead of overflowing.
struct nrf_fft_engine {
    std::mutex mu;                   // process / shift / get_buffer may come from the device's producer thread and the script's thread
    int head = 0;                    // ring row that holds the newest spectrum
    int dirty = 0;                   // rows written on the device since the host mirror was last brought up to date
    fsea_plan* plan_u8 = nullptr;
    fsea_plan* plan_f64 = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t staged = nullptr;    // the last upload has left the pinned staging arrays
    uint8_t* stage_u8 = nullptr;     // 2N bytes, pinned
    double* stage_f64 = nullptr;     // 2N doubles, pinned
    void* d_in = nullptr;            // 2N doubles' worth
    void* d_row = nullptr;           // N doubles' worth: the transform's output row
    double* d_ring = nullptr;        // H x N, the history
    double* d_ring_alt = nullptr;    // second buffer of the retune scroll
};

static void cu_die(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return;
    fprintf(stderr, "NRF fatal error: %s: %s\n", what, cudaGetErrorString(e));
    exit(EXIT_FAILURE);
}

static fsea_plan* make_plan(int n, int in_format) {
    fsea_config cfg;
    fsea_config_default(&cfg, n);
    cfg.in_format = in_format;
    cfg.out_mode = FSEA_OUT_MAG;
    cfg.flags = FSEA_FLAG_DC_PATCH;              // row[N/2] = row[N/2-1], src/nrf.c:621-622
    fsea_plan* plan = nullptr;
    int rc = fsea_plan_create(&cfg, &plan);
    if (rc != FSEA_OK) nrf_die("nrf_fft: fsea_plan_create", rc);
    return plan;
}

extern "C" nrf_fft* nrf_fft_new(int fft_size, int fft_history_size) {
    nrf_fft* fft = (nrf_fft*)calloc(1, sizeof(nrf_fft));
    nrf_block_init(&fft->block, NRF_BLOCK_GENERIC, (nrf_block_process_fn)nrf_fft_process,
                   (nrf_block_result_fn)nrf_fft_get_buffer);
    fft->fft_size = fft_size;
    fft->fft_history_size = fft_history_size;
    const size_t n = (size_t)fft_size, rows = (size_t)(fft_history_size > 0 ? fft_history_size : 1);
    fft->buffer = (double*)calloc(n * rows, sizeof(double));
    nrf_fft_engine* eng = new nrf_fft_engine();
    eng->plan_u8 = make_plan(fft_size, FSEA_IN_U8);          // (fails loudly here for a size the engine does not have)
    cu_die(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking), "nrf_fft_new: stream");
    cu_die(cudaEventCreateWithFlags(&eng->staged, cudaEventDisableTiming), "nrf_fft_new: event");
    cu_die(cudaMallocHost((void**)&eng->stage_u8, 2 * n), "nrf_fft_new: staging");
    cu_die(cudaMallocHost((void**)&eng->stage_f64, 2 * n * sizeof(double)), "nrf_fft_new: staging");
    memset(eng->stage_u8, 0, 2 * n);
    memset(eng->stage_f64, 0, 2 * n * sizeof(double));
    cu_die(cudaMalloc(&eng->d_in, 2 * n * sizeof(double)), "nrf_fft_new: device input");
    cu_die(cudaMalloc(&eng->d_row, n * sizeof(double)), "nrf_fft_new: device row");
    cu_die(cudaMalloc((void**)&eng->d_ring, n * rows * sizeof(double)), "nrf_fft_new: history");
    cu_die(cudaMalloc((void**)&eng->d_ring_alt, n * rows * sizeof(double)), "nrf_fft_new: history");
    cu_die(cudaMemset(eng->d_ring, 0, n * rows * sizeof(double)), "nrf_fft_new: history");
    fft->engine = eng;
    return fft;
}

extern "C" void nrf_fft_process(nrf_fft* fft, nut_buffer* buffer) {
    nrf_fft_engine* eng = (nrf_fft_engine*)fft->engine;
    std::lock_guard<std::mutex> lock(eng->mu);
    const int n = fft->fft_size;
    const size_t have = (size_t)buffer->length * (size_t)buffer->channels / 2;   // complex samples offered
    const size_t take = have < (size_t)n ? have : (size_t)n;                      // only the first N matter
    const int rows = fft->fft_history_size > 0 ? fft->fft_history_size : 1;
    eng->head = (eng->head + rows - 1) % rows;                   // newest row moves "up" by one
    double* ring_row = eng->d_ring + (size_t)eng->head * n;
    cu_die(cudaEventSynchronize(eng->staged), "nrf_fft_process");             // (long since true: the previous upload has left the staging arrays)
    int rc;
    if (buffer->type == NUT_BUFFER_U8) {
        memcpy(eng->stage_u8, buffer->data.u8, 2 * take);
        cu_die(cudaMemcpyAsync(eng->d_in, eng->stage_u8, 2 * (size_t)n, cudaMemcpyHostToDevice, eng->stream), "nrf_fft_process: upload");
        cu_die(cudaEventRecord(eng->staged, eng->stream), "nrf_fft_process");
        rc = fsea_exec_device(eng->plan_u8, eng->d_in, eng->d_row, 1, eng->stream);
        if (rc == FSEA_OK) rc = fsea_history_put_device(eng->d_row, 0, ring_row, n, eng->stream);
    } else {
        if (!eng->plan_f64) eng->plan_f64 = make_plan(n, FSEA_IN_F64C);
        memcpy(eng->stage_f64, buffer->data.f64, 2 * take * sizeof(double));
        cu_die(cudaMemcpyAsync(eng->d_in, eng->stage_f64, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, eng->stream), "nrf_fft_process: upload");
        cu_die(cudaEventRecord(eng->staged, eng->stream), "nrf_fft_process");
        rc = fsea_exec_device(eng->plan_f64, eng->d_in, eng->d_row, 1, eng->stream);
        if (rc == FSEA_OK) rc = fsea_history_put_device(eng->d_row, 1, ring_row, n, eng->stream);
    }
    if (rc != FSEA_OK) nrf_die("nrf_fft_process", rc);
    if (eng->dirty < rows) eng->dirty++;
}

// the rows written on the device since the last call -> the host mirror (ring order, like the device's)
static void fft_flush_to_host(nrf_fft* fft, nrf_fft_engine* eng) {
    const size_t n = (size_t)fft->fft_size, rows = (size_t)(fft->fft_history_size > 0 ? fft->fft_history_size : 1);
    if (eng->dirty > 0) {
        // dirty rows are head, head+1, .. (mod rows): at most two contiguous pieces of the ring
        const size_t first = (size_t)eng->head, count = (size_t)eng->dirty;
        const size_t a = count < rows - first ? count : rows - first;
        cu_die(cudaMemcpyAsync(fft->buffer + first * n, eng->d_ring + first * n, a * n * sizeof(double), cudaMemcpyDeviceToHost, eng->stream), "nrf_fft: history download");
        if (count > a)
            cu_die(cudaMemcpyAsync(fft->buffer, eng->d_ring, (count - a) * n * sizeof(double), cudaMemcpyDeviceToHost, eng->stream), "nrf_fft: history download");
        eng->dirty = 0;
    }
    cu_die(cudaStreamSynchronize(eng->stream), "nrf_fft: history download");
}

// horizontal scroll of every history row after a retune (src/nrf.c:569-596): one kernel over the device
// ring, and the same scroll on the host mirror (which is brought up to date first)
extern "C" void nrf_fft_shift(nrf_fft* fft, double d) {
    nrf_fft_engine* eng = (nrf_fft_engine*)fft->engine;
    std::lock_guard<std::mutex> lock(eng->mu);
    const int n = fft->fft_size, rows = fft->fft_history_size > 0 ? fft->fft_history_size : 1;
    const int s = (int)round(n / d);
    if (s == 0) return;
    fft_flush_to_host(fft, eng);
    const int rc = fsea_history_shift_device(eng->d_ring, eng->d_ring_alt, n, rows, s, eng->stream);
    if (rc != FSEA_OK) nrf_die("nrf_fft_shift", rc);
    double* t = eng->d_ring; eng->d_ring = eng->d_ring_alt; eng->d_ring_alt = t;
    if (abs(s) >= n) {
        memset(fft->buffer, 0, (size_t)n * rows * sizeof(double));
        return;
    }
    for (int y = 0; y < rows; y++) {
        double* row = fft->buffer + (size_t)y * n;
        if (s > 0) {
            memmove(row, row + s, (size_t)(n - s) * sizeof(double));
            memset(row + (n - s), 0, (size_t)s * sizeof(double));
        } else {
            memmove(row - s, row, (size_t)(n + s) * sizeof(double));
            memset(row, 0, (size_t)(-s) * sizeof(double));
        }
    }
}

extern "C" nut_buffer* nrf_fft_get_buffer(nrf_fft* fft) {
    nrf_fft_engine* eng = (nrf_fft_engine*)fft->engine;
    const size_t n = (size_t)fft->fft_size, rows = (size_t)(fft->fft_history_size > 0 ? fft->fft_history_size : 1);
    nut_buffer* out = nut_buffer_new_f64((int)(n * rows), 1, NULL);
    std::lock_guard<std::mutex> lock(eng->mu);
    fft_flush_to_host(fft, eng);
    const size_t head = (size_t)eng->head;
    memcpy(out->data.f64, fft->buffer + head * n, (rows - head) * n * sizeof(double));          // rows head .. H-1
    if (head) memcpy(out->data.f64 + (rows - head) * n, fft->buffer, head * n * sizeof(double)); // rows 0 .. head-1
    return out;
}

// Extension (not in the reference): the history where it lives.  *d_ring = H x N float64 rows in device memory,
// ring order; *head = the row holding the newest spectrum (row (head + k) mod H is k frames old).  Valid until the
// next nrf_fft_shift / nrf_fft_free; work queued by nrf_fft_process is complete when this returns.
extern "C" void nrf_fft_get_history_device(nrf_fft* fft, const double** d_ring, int* head) {
    nrf_fft_engine* eng = (nrf_fft_engine*)fft->engine;
    std::lock_guard<std::mutex> lock(eng->mu);
    cu_die(cudaStreamSynchronize(eng->stream), "nrf_fft_get_history_device");
    if (d_ring) *d_ring = eng->d_ring;
    if (head) *head = eng->head;
}

extern "C" void nrf_fft_free(nrf_fft* fft) {
    nrf_fft_engine* eng = (nrf_fft_engine*)fft->engine;
    if (eng) {
        if (eng->stream) cudaStreamSynchronize(eng->stream);
        fsea_plan_destroy(eng->plan_u8);
        fsea_plan_destroy(eng->plan_f64);
        if (eng->stage_u8) cudaFreeHost(eng->stage_u8);
        if (eng->stage_f64) cudaFreeHost(eng->stage_f64);
        if (eng->d_in) cudaFree(eng->d_in);
        if (eng->d_row) cudaFree(eng->d_row);
        if (eng->d_ring) cudaFree(eng->d_ring);
        if (eng->d_ring_alt) cudaFree(eng->d_ring_alt);
        if (eng->staged) cudaEventDestroy(eng->staged);
        if (eng->stream) cudaStreamDestroy(eng->stream);
        delete eng;
    }
    free(fft->buffer);          // the reference leaks this (src/nrf.c:637-642)
    free(fft);
}

// ---------------------------------------------------------------- frequency shifter
// Complex rotator in front of the FFT in lua/fft-shifted.lua (src/nrf.c:814-874):
// out = in * e^{i phi_n} + 0.5, phase advanced by recurrence, state kept across calls.
extern "C" nrf_freq_shifter* nrf_freq_shifter_new(int freq_offset, int sample_rate) {
    nrf_freq_shifter* sh = (nrf_freq_shifter*)calloc(1, sizeof(nrf_freq_shifter));
    nrf_block_init(&sh->block, NRF_BLOCK_GENERIC, (nrf_block_process_fn)nrf_freq_shifter_process,
                   (nrf_block_result_fn)nrf_freq_shifter_get_buffer);
    sh->freq_offset = freq_offset;
    sh->sample_rate = sample_rate;
    sh->cosine = 1.0;
    sh->sine = 0.0;
    return sh;
}

extern "C" void nrf_freq_shifter_process(nrf_freq_shifter* sh, nut_buffer* buffer) {
    assert(buffer->channels == 2);
    const double step = 2.0 * M_PI * sh->freq_offset / (double)sh->sample_rate;
    const double dc = cos(step), ds = sin(step);
    const int count = buffer->length * buffer->channels;
    // The reference sizes its result as (count, 2), i.e. twice the payload, and
    // fills only the first `count` values (src/nrf.c:849-852); kept, because
    // nrf_fft_process() clamps what it reads.
    if (!sh->buffer) sh->buffer = nut_buffer_new_f64(count, 2, NULL);
    double* out = sh->buffer->data.f64;
    double c = sh->cosine, s = sh->sine;
    for (int i = 0; i + 1 < count; i += 2) {
        const double vi = nut_buffer_get_f64(buffer, i), vq = nut_buffer_get_f64(buffer, i + 1);
        out[i] = vi * c - vq * s + 0.5;
        out[i + 1] = vi * s + vq * c + 0.5;
        const double ns = c * ds + s * dc, nc = c * dc - s * ds;
        s = ns;
        c = nc;
    }
    sh->cosine = c;
    sh->sine = s;
}

extern "C" nut_buffer* nrf_freq_shifter_get_buffer(nrf_freq_shifter* sh) { return nut_buffer_copy(sh->buffer); }

extern "C" void nrf_freq_shifter_free(nrf_freq_shifter* sh) {
    if (sh->buffer) nut_buffer_free(sh->buffer);
    free(sh);
}

// ---------------------------------------------------------------- FIR low-pass / IQ filter
// Host side of src/nrf.c:644-776.  Windowed-sinc taps (Blackman-style window over length + 1
// points, odd length enforced, unit DC gain), a sample line that keeps the last `length - 1`
// samples of the previous block in front of the new one -- and the convolution itself on the device.
static const double kTau = 2.0 * M_PI;

extern "C" double* nrf_fir_get_low_pass_coefficients(int sample_rate, int half_ampl_freq, int length) {
    if (length % 2 == 0) length++;                                   // src/nrf.c:655
    const double f = half_ampl_freq / (double)sample_rate;
    double* taps = (double*)calloc((size_t)length, sizeof(double));
    const int mid = length / 2;
    double total = 0.0;
    for (int i = 0; i < length; i++) {
        double v;
        if (i == mid) {
            v = kTau * f;
        } else {
            const double a = kTau * (i + 1) / (double)(length + 1);
            v = sin(kTau * f * (i - mid)) / (double)(i - mid);
            v *= 0.42 - 0.5 * cos(a) + 0.08 * cos(2 * a);
        }
        total += v;
        taps[i] = v;
    }
    for (int i = 0; i < length; i++) taps[i] /= total;
    return taps;
}

extern "C" nrf_fir_filter* nrf_fir_filter_new(int sample_rate, int half_ampl_freq, int length) {
    nrf_fir_filter* f = (nrf_fir_filter*)calloc(1, sizeof(nrf_fir_filter));
    f->length = length;                                              // (the even -> odd bump stays inside the
    f->coefficients = nrf_fir_get_low_pass_coefficients(sample_rate, half_ampl_freq, length);   //  coefficient call, src/nrf.c:680-681)
    f->offset = length - 1;
    f->center = length / 2;
    f->samples_length = f->offset;
    f->samples = (double*)calloc((size_t)f->samples_length, sizeof(double));
    return f;
}

extern "C" void nrf_fir_filter_load(nrf_fir_filter* f, double* samples, int length) {
    const int want = length + f->offset;
    double* line = f->samples;
    const bool fresh = f->samples_length != want;
    if (fresh) line = (double*)calloc((size_t)want, sizeof(double));
    // tail of the old line first (memmove: source and destination are the same array when not fresh)
    memmove(line, f->samples + f->samples_length - f->offset, (size_t)f->offset * sizeof(double));
    memcpy(line + f->offset, samples, (size_t)length * sizeof(double));
    if (fresh && f->samples_length != 0) free(f->samples);
    f->samples = line;
    f->samples_length = want;
}

extern "C" double nrf_fir_filter_get(nrf_fir_filter* f, int index) {
    double v = 0;
    for (int k = 0; k < f->length; k++) v += f->coefficients[k] * f->samples[index + k];
    return v;
}

extern "C" void nrf_fir_filter_free(nrf_fir_filter* f) {
    free(f->coefficients);
    free(f->samples);
    free(f);
}

extern "C" nrf_iq_filter* nrf_iq_filter_new(int sample_rate, int half_ampl_freq, int kernel_length) {
    nrf_iq_filter* f = (nrf_iq_filter*)calloc(1, sizeof(nrf_iq_filter));
    nrf_block_init(&f->block, NRF_BLOCK_GENERIC, (nrf_block_process_fn)nrf_iq_filter_process,
                   (nrf_block_result_fn)nrf_iq_filter_get_buffer);
    f->filter_i = nrf_fir_filter_new(sample_rate, half_ampl_freq, kernel_length);
    f->filter_q = nrf_fir_filter_new(sample_rate, half_ampl_freq, kernel_length);
    return f;
}

extern "C" void nrf_iq_filter_process(nrf_iq_filter* f, nut_buffer* buffer) {
    const int length = buffer->length;
    if (length != f->samples_length) {
        free(f->samples_i);
        free(f->samples_q);
        f->samples_i = (double*)calloc((size_t)length, sizeof(double));
        f->samples_q = (double*)calloc((size_t)length, sizeof(double));
    }
    f->samples_length = length;
    for (int j = 0; j < length; j++) {                               // u8 -> /256, f64 as is (src/nut.c:129-135)
        f->samples_i[j] = nut_buffer_get_f64(buffer, 2 * j);
        f->samples_q[j] = nut_buffer_get_f64(buffer, 2 * j + 1);
    }
    nrf_fir_filter_load(f->filter_i, f->samples_i, length);
    nrf_fir_filter_load(f->filter_q, f->samples_q, length);
}

extern "C" nut_buffer* nrf_iq_filter_get_buffer(nrf_iq_filter* f) {
    const int length = f->samples_length;
    nut_buffer* result = nut_buffer_new_f64(length, 2, NULL);
    if (length <= 0) return result;
    const int rc = fsea_fir_iq_host(0, f->filter_i->coefficients, f->filter_i->length, f->filter_i->samples,
                                    f->filter_q->samples, length, result->data.f64);
    if (rc != FSEA_OK) nrf_die("nrf_iq_filter_get_buffer", rc);
    return result;
}

extern "C" void nrf_iq_filter_free(nrf_iq_filter* f) {
    nrf_fir_filter_free(f->filter_i);
    nrf_fir_filter_free(f->filter_q);
    free(f->samples_i);
    free(f->samples_q);
    free(f);
}

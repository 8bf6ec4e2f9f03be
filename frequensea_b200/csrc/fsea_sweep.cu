// fsea_sweep.cu -- the multi-centre-frequency sweep (BASELINE cfg4) as a device-resident engine.
//
// c/fft-batch-broad.c:56-129,189-209 per centre frequency: rows of an N-point spectrum; after
// GATE rows the mean |X| decides whether anything is there (:80-98); if so HEIGHT rows of u8
// pixels (S = 5, pixel N/2 := pixel N/2-1, :108-118), one image per frequency;
// c/fft-stitch-broad.c:28-36,66-88 then places the images side by side.  Here the raw IQ of a
// frequency becomes device-resident once, the gate sum is reduced on the device in the fixed
// order of fsea_sum_kernel (the same bits on every run, rank and sharding), the keep flags cross
// to the host in ONE small copy per job, and the pixel rows are written where the caller wants
// them: the engine's own images or, row by row, a stitched canvas -- which may be a peer GPU's
// memory, so that the transform kernel's stores are the gather.  No allocation after create.
#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/fsea.h"
#include "fsea_internal.h"

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fsea_fail_cuda_internal(e__, #call); \
    } while (0)

struct fsea_sweep {
    fsea_sweep_config cfg;
    fsea_plan* mag = nullptr;                   // |X| of the gate rows, DC bin included (c/fft-batch-broad.c:84-90 sums every bin)
    fsea_plan* pix = nullptr;                   // pixel rows, S = pix_scale, DC patch
    cudaStream_t stream = nullptr;              // the engine's stream: submits, decisions, what callers order against
    // A frequency is a small launch (gate: 100 frames, pixels: 512 frames on 148 SMs): slots are dealt to NS
    // side streams so that the tails of neighbouring launches overlap; events fork them from / join them to `stream`.
    static const int NS = 4;
    cudaStream_t side[NS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[NS] = {nullptr, nullptr, nullptr, nullptr};
    size_t raw_bytes = 0, img_bytes = 0;        // per frequency
    uint8_t* d_raw = nullptr;                   // [capacity][raw_bytes]   (own_raw)
    std::vector<const void*> raw_ptr;           // where each slot's raw IQ lives
    float* d_gate_mag = nullptr;                // NS x gate_rows x N
    float* d_gate_rowmag = nullptr;             // NS x capacity x N: |X| of ONE gate row of every slot (own_raw, sweep_gate_by_rows)
    double* d_partials = nullptr;               // [capacity][FSEA_SUM_BLOCKS]
    double* d_warpsums = nullptr;               // [capacity][gate_rows][4]: the SUMS kernel's per-warp sums (N = 16384, own_raw)
    bool by_sums = false;                       // the last gate used them
    double* d_means = nullptr;
    uint8_t* d_keep = nullptr;
    double* h_means = nullptr;                  // pinned
    uint8_t* h_keep = nullptr;
    uint8_t* d_images = nullptr;                // [capacity][rows][N]     (own_images)
    std::vector<uint8_t> keep;                  // the last gate's decisions
    int gated = 0;                              // slots covered by `keep`
    long launches = 0;
};

// mean over the gate rows from the FSEA_SUM_BLOCKS partial sums of a slot, added in block order as
// fsea_gate_mean_host does on the host (same bits), and the "interesting?" decision (:93-97)
__global__ void fsea_gate_decide_kernel(const double* __restrict__ partials, int entries, int n_slots, double count, double threshold,
                                        double* __restrict__ means, uint8_t* __restrict__ keep) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    double total = 0.0;
    for (int b = 0; b < entries; b++) total += partials[(size_t)s * entries + b];
    const double m = total / count;
    means[s] = m;
    keep[s] = (m < threshold) ? 0 : 1;
}

// SURVEY.md 8(d) cfg4 synthetic source, drawn on the device (measurement tooling: 16 GiB of raw IQ
// for the stated job cannot come from a host generator in reasonable time).  Sample s of centre
// frequency f: gaussian noise from a counter-based hash of (seed = 1000 + f, s), sigma 12 (0.4 for
// the quiet frequencies the gate must reject), plus a tone of amplitude 40 at shifted bin
// (37 f + 11) mod N; two's-complement int8.  Parity tests read the bytes back and run the oracle on them.
__device__ __forceinline__ unsigned long long fsea_mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void fsea_synth_sweep_kernel(uint8_t* __restrict__ out, long count, int n, unsigned seed, int quiet, int tone_bin) {
    const float sigma = quiet ? 0.4f : 12.0f;
    for (long s = (long)blockIdx.x * blockDim.x + threadIdx.x; s < count; s += (long)gridDim.x * blockDim.x) {
        const unsigned long long h = fsea_mix64(((unsigned long long)seed << 40) ^ (unsigned long long)s);
        const float u1 = ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(unsigned)((h >> 16) & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sigma * sqrtf(-2.0f * logf(u1));
        float vi = r * cosf(6.28318530718f * u2), vq = r * sinf(6.28318530718f * u2);
        if (!quiet) {
            const long ph = ((long)(tone_bin - n / 2) * (s % n)) % n;            // exact phase index, |.| < n
            const float a = 6.28318530718f * (float)ph / (float)n;
            vi += 40.0f * cosf(a);
            vq += 40.0f * sinf(a);
        }
        const float ci = fminf(fmaxf(rintf(vi), -128.0f), 127.0f), cq = fminf(fmaxf(rintf(vq), -128.0f), 127.0f);
        out[2 * s] = (uint8_t)(int8_t)(int)ci;
        out[2 * s + 1] = (uint8_t)(int8_t)(int)cq;
    }
}

extern "C" int fsea_sweep_destroy(fsea_sweep* s) {
    if (!s) return FSEA_OK;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    fsea_plan_destroy(s->mag);
    fsea_plan_destroy(s->pix);
    if (s->d_raw) cudaFree(s->d_raw);
    if (s->d_gate_mag) cudaFree(s->d_gate_mag);
    if (s->d_gate_rowmag) cudaFree(s->d_gate_rowmag);
    if (s->d_partials) cudaFree(s->d_partials);
    if (s->d_warpsums) cudaFree(s->d_warpsums);
    if (s->d_means) cudaFree(s->d_means);
    if (s->d_keep) cudaFree(s->d_keep);
    if (s->h_means) cudaFreeHost(s->h_means);
    if (s->h_keep) cudaFreeHost(s->h_keep);
    if (s->d_images) cudaFree(s->d_images);
    for (int k = 0; k < fsea_sweep::NS; k++) {
        if (s->side[k]) { cudaStreamSynchronize(s->side[k]); cudaStreamDestroy(s->side[k]); }
        if (s->join[k]) cudaEventDestroy(s->join[k]);
    }
    if (s->fork) cudaEventDestroy(s->fork);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return FSEA_OK;
}

static int sweep_setup(fsea_sweep* s) {
    const fsea_sweep_config& c = s->cfg;
    fsea_config pc;
    fsea_config_default(&pc, c.fft_size);
    pc.in_format = c.in_format;
    pc.device = c.device;
    pc.out_mode = FSEA_OUT_MAG;
    int rc = fsea_plan_create(&pc, &s->mag);
    if (rc != FSEA_OK) return rc;
    pc.out_mode = FSEA_OUT_PIX;
    pc.arith = (c.pixel_arith == FSEA_ARITH_F64) ? FSEA_ARITH_F64 : FSEA_ARITH_AUTO;
    pc.flags = FSEA_FLAG_DC_PATCH;
    pc.pix_scale = c.pix_scale;
    if ((rc = fsea_plan_create(&pc, &s->pix)) != FSEA_OK) return rc;
    CU(cudaSetDevice(c.device));
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    for (int k = 0; k < fsea_sweep::NS; k++) {
        CU(cudaStreamCreateWithFlags(&s->side[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->join[k], cudaEventDisableTiming));
    }
    s->raw_bytes = (size_t)c.rows * c.fft_size * 2;
    s->img_bytes = (size_t)c.rows * c.fft_size;
    if (c.own_raw) CU(cudaMalloc(&s->d_raw, s->raw_bytes * c.capacity));
    if (c.own_images) CU(cudaMalloc(&s->d_images, s->img_bytes * c.capacity));
    CU(cudaMalloc(&s->d_gate_mag, sizeof(float) * fsea_sweep::NS * (size_t)c.gate_rows * c.fft_size));
    if (c.own_raw && c.gate_rows <= FSEA_SUM_BLOCKS && !fsea_plan_has_sums_internal(s->mag)) CU(cudaMalloc(&s->d_gate_rowmag, sizeof(float) * fsea_sweep::NS * (size_t)c.capacity * c.fft_size));
    CU(cudaMalloc(&s->d_partials, sizeof(double) * FSEA_SUM_BLOCKS * (size_t)c.capacity));
    if (c.own_raw && fsea_plan_has_sums_internal(s->mag)) CU(cudaMalloc(&s->d_warpsums, sizeof(double) * 4 * (size_t)c.gate_rows * c.capacity));
    CU(cudaMalloc(&s->d_means, sizeof(double) * c.capacity));
    CU(cudaMalloc(&s->d_keep, (size_t)c.capacity));
    CU(cudaMallocHost((void**)&s->h_means, sizeof(double) * c.capacity));
    CU(cudaMallocHost((void**)&s->h_keep, (size_t)c.capacity));
    s->raw_ptr.assign((size_t)c.capacity, nullptr);
    s->keep.assign((size_t)c.capacity, 0);
    return FSEA_OK;
}

extern "C" int fsea_sweep_create(const fsea_sweep_config* cfg, fsea_sweep** out) {
    if (!cfg || !out) return FSEA_ERR_INVALID;
    *out = nullptr;
    if (cfg->rows <= 0 || cfg->capacity <= 0 || cfg->gate_rows <= 0) return FSEA_ERR_INVALID;
    if (cfg->in_format != FSEA_IN_S8 && cfg->in_format != FSEA_IN_U8) return FSEA_ERR_INVALID;
    fsea_sweep* s = new (std::nothrow) fsea_sweep();
    if (!s) return FSEA_ERR_NOMEM;
    s->cfg = *cfg;
    if (s->cfg.gate_rows > s->cfg.rows) s->cfg.gate_rows = s->cfg.rows;
    const int rc = sweep_setup(s);
    if (rc != FSEA_OK) { fsea_sweep_destroy(s); return rc; }
    *out = s;
    return FSEA_OK;
}

// the gate rows of a slot whose raw IQ is in place: |X| of the first gate_rows rows, summed per block
static int sweep_fork(fsea_sweep* s) {                       // side streams start behind everything queued on `stream`
    CU(cudaEventRecord(s->fork, s->stream));
    for (int k = 0; k < fsea_sweep::NS; k++) CU(cudaStreamWaitEvent(s->side[k], s->fork, 0));
    return FSEA_OK;
}
static int sweep_join(fsea_sweep* s) {                       // ... and `stream` continues behind the side streams
    for (int k = 0; k < fsea_sweep::NS; k++) {
        CU(cudaEventRecord(s->join[k], s->side[k]));
        CU(cudaStreamWaitEvent(s->stream, s->join[k], 0));
    }
    return FSEA_OK;
}

static int sweep_gate_rows(fsea_sweep* s, int slot) {
    const int k = slot % fsea_sweep::NS;
    const long n = (long)s->cfg.gate_rows * s->cfg.fft_size;
    float* mag = s->d_gate_mag + (size_t)k * n;
    int rc = fsea_launch_internal(s->mag, s->raw_ptr[slot], mag, s->cfg.gate_rows, s->side[k], 0, k);
    if (rc != FSEA_OK) return rc;
    FSEA_LAUNCH(fsea_sum_kernel, FSEA_SUM_BLOCKS, 256, s->side[k], (const float*)mag, n, s->d_partials + (size_t)slot * FSEA_SUM_BLOCKS);
    CU(cudaGetLastError());
    s->launches += 2;
    return FSEA_OK;
}

// The same sums, taken ROW by row over all slots at once when the raw IQ lies in the engine's own store: gate row r of
// slots 0 .. n_slots-1 are frames rows * N samples apart, so one launch transforms row r of every slot (n_slots frames:
// a full grid, where a slot's gate_rows frames fill two thirds of one wave) and one block per slot sums its N magnitudes.
// partials[slot][r] = that row's sum (block order = row order: the same bits wherever the slot sits and however many
// slots there are), entries r >= gate_rows zero; fsea_gate_decide_kernel adds them as before.
__global__ void fsea_rowsum_kernel(const float* __restrict__ x, int n, double* __restrict__ partials, int r) {
    FSEA_STATIC_SMEM(double, warp_sum, 8);
    const float* row = x + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)row[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += warp_sum[w];
        partials[(size_t)blockIdx.x * FSEA_SUM_BLOCKS + r] = t;
    }
}

static bool sweep_in_own_store(const fsea_sweep* s, int n_slots) {
    if (!s->d_raw || (!s->d_gate_rowmag && !s->d_warpsums)) return false;
    for (int j = 0; j < n_slots; j++) if (s->raw_ptr[j] != s->d_raw + (size_t)j * s->raw_bytes) return false;
    return true;
}

static int sweep_gate_by_rows(fsea_sweep* s, int n_slots) {          // (between sweep_fork and sweep_join)
    const int n = s->cfg.fft_size;
    s->by_sums = s->d_warpsums != nullptr;
    if (s->by_sums) {
        // the kernel's SUMS form: no magnitude row is written at all, comb warp w of (slot, row r) leaves its share at
        // warpsums[slot][r][w]; the decision kernel adds a slot's 4 gate_rows values in that order
        int rc = sweep_fork(s);
        for (int r = 0; r < s->cfg.gate_rows && rc == FSEA_OK; r++) {
            rc = fsea_launch_sums_internal(s->mag, s->d_raw + (size_t)r * n * 2, n_slots, s->side[r % fsea_sweep::NS], (long)s->cfg.rows * n,
                                           s->d_warpsums, 4L * s->cfg.gate_rows, 4L * r);
            s->launches++;
        }
        return rc;
    }
    CU(cudaMemsetAsync(s->d_partials, 0, sizeof(double) * FSEA_SUM_BLOCKS * (size_t)n_slots, s->stream));
    int rc = sweep_fork(s);
    for (int r = 0; r < s->cfg.gate_rows && rc == FSEA_OK; r++) {
        const int k = r % fsea_sweep::NS;
        float* mag = s->d_gate_rowmag + (size_t)k * s->cfg.capacity * n;
        rc = fsea_launch_internal(s->mag, s->d_raw + (size_t)r * n * 2, mag, n_slots, s->side[k], 0, k, 0, 0, (long)s->cfg.rows * n);
        if (rc != FSEA_OK) break;
        FSEA_LAUNCH(fsea_rowsum_kernel, n_slots, 256, s->side[k], (const float*)mag, n, s->d_partials, r);
        CU(cudaGetLastError());
        s->launches += 2;
    }
    return rc;
}

extern "C" int fsea_sweep_submit_host(fsea_sweep* s, int slot, const void* h_raw) {
    if (!s || !h_raw || slot < 0 || slot >= s->cfg.capacity || !s->d_raw) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    uint8_t* dst = s->d_raw + (size_t)slot * s->raw_bytes;
    CU(cudaMemcpyAsync(dst, h_raw, s->raw_bytes, cudaMemcpyHostToDevice, s->stream));
    s->raw_ptr[slot] = dst;
    return FSEA_OK;
}

extern "C" int fsea_sweep_submit_device(fsea_sweep* s, int slot, const void* d_raw) {
    if (!s || !d_raw || slot < 0 || slot >= s->cfg.capacity || ((uintptr_t)d_raw & 15)) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    s->raw_ptr[slot] = d_raw;
    return FSEA_OK;
}

extern "C" int fsea_sweep_submit_synth(fsea_sweep* s, int slot, int freq_idx, int quiet_every) {
    if (!s || slot < 0 || slot >= s->cfg.capacity || !s->d_raw || freq_idx < 0) return FSEA_ERR_INVALID;
    if (s->cfg.in_format != FSEA_IN_S8) return FSEA_ERR_INVALID;                 // the source draws two's-complement int8
    CU(cudaSetDevice(s->cfg.device));
    uint8_t* dst = s->d_raw + (size_t)slot * s->raw_bytes;
    const int n = s->cfg.fft_size;
    const int quiet = (quiet_every > 0 && (freq_idx % quiet_every) == quiet_every - 1) ? 1 : 0;
    const int tone_bin = (int)((37L * freq_idx + 11) % n);
    FSEA_LAUNCH(fsea_synth_sweep_kernel, 148 * 8, 256, s->stream, dst, (long)s->cfg.rows * n, n, (unsigned)(1000 + freq_idx), quiet, tone_bin);
    CU(cudaGetLastError());
    s->launches++;
    s->raw_ptr[slot] = dst;
    return FSEA_OK;
}

// gate in two halves, so that one host thread can drive several devices at once: begin queues the gate
// rows, the decision and the small D2H on the engine's stream; end waits for them
extern "C" int fsea_sweep_gate_begin(fsea_sweep* s, int n_slots) {
    if (!s || n_slots <= 0 || n_slots > s->cfg.capacity) return FSEA_ERR_INVALID;
    for (int j = 0; j < n_slots; j++) if (!s->raw_ptr[j]) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    s->by_sums = false;
    if (sweep_in_own_store(s, n_slots)) {
        rc = sweep_gate_by_rows(s, n_slots);
    } else {
        rc = sweep_fork(s);
        for (int j = 0; j < n_slots && rc == FSEA_OK; j++) rc = sweep_gate_rows(s, j);
    }
    if (rc == FSEA_OK) rc = sweep_join(s);
    if (rc != FSEA_OK) return rc;
    const double count = (double)s->cfg.gate_rows * (double)s->cfg.fft_size;
    FSEA_LAUNCH(fsea_gate_decide_kernel, (n_slots + 127) / 128, 128, s->stream, (const double*)(s->by_sums ? s->d_warpsums : s->d_partials),
                s->by_sums ? 4 * s->cfg.gate_rows : FSEA_SUM_BLOCKS, n_slots, count, s->cfg.gate_threshold, s->d_means, s->d_keep);
    CU(cudaGetLastError());
    s->launches++;
    CU(cudaMemcpyAsync(s->h_means, s->d_means, sizeof(double) * n_slots, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(s->h_keep, s->d_keep, (size_t)n_slots, cudaMemcpyDeviceToHost, s->stream));
    return FSEA_OK;
}

extern "C" int fsea_sweep_gate_end(fsea_sweep* s, int n_slots, double* means, uint8_t* keep) {
    if (!s || n_slots <= 0 || n_slots > s->cfg.capacity) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    for (int j = 0; j < n_slots; j++) s->keep[j] = s->h_keep[j];
    s->gated = n_slots;
    if (means) memcpy(means, s->h_means, sizeof(double) * n_slots);
    if (keep) memcpy(keep, s->h_keep, (size_t)n_slots);
    return FSEA_OK;
}

extern "C" int fsea_sweep_gate(fsea_sweep* s, int n_slots, double* means, uint8_t* keep) {
    const int rc = fsea_sweep_gate_begin(s, n_slots);
    return rc != FSEA_OK ? rc : fsea_sweep_gate_end(s, n_slots, means, keep);
}

extern "C" int fsea_sweep_render(fsea_sweep* s, int n_slots, uint8_t* d_canvas, long canvas_stride, long canvas_col0, long canvas_col_step) {
    if (!s || n_slots <= 0 || n_slots > s->gated) return FSEA_ERR_INVALID;
    const int n = s->cfg.fft_size;
    if (!d_canvas && !s->d_images) return FSEA_ERR_INVALID;
    if (d_canvas && (canvas_stride < n || (canvas_stride & 3) || (canvas_col0 & 3) || (canvas_col_step & 3) || ((uintptr_t)d_canvas & 3))) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    int rc = sweep_fork(s);
    if (rc != FSEA_OK) return rc;
    // A run of consecutive kept slots whose raw IQ is contiguous is ONE launch of run x rows frames: the rows of
    // slot j + q start q * col_step bytes further on (images one after the other, or side by side in a canvas).
    const long col_step = d_canvas ? canvas_col_step : (long)s->img_bytes;
    int launches = 0;
    for (int j = 0; j < n_slots;) {
        uint8_t* dst = d_canvas ? d_canvas + canvas_col0 + (size_t)j * canvas_col_step : s->d_images + (size_t)j * s->img_bytes;
        const long stride = d_canvas ? canvas_stride : (long)n;
        cudaStream_t st = s->side[launches % fsea_sweep::NS];
        if (!s->keep[j]) {
            CU(cudaMemset2DAsync(dst, (size_t)stride, 0, (size_t)n, (size_t)s->cfg.rows, st));
            j++;
            continue;
        }
        int run = 1;
        while (j + run < n_slots && s->keep[j + run] &&
               (const uint8_t*)s->raw_ptr[j + run] == (const uint8_t*)s->raw_ptr[j] + (size_t)run * s->raw_bytes) run++;
        rc = fsea_launch_internal(s->pix, s->raw_ptr[j], dst, (long)run * s->cfg.rows, st, stride, launches % fsea_sweep::NS,
                                  run > 1 ? (long)s->cfg.rows : 0, col_step);
        if (rc != FSEA_OK) return rc;
        s->launches++;
        launches++;
        j += run;
    }
    return sweep_join(s);
}

extern "C" int fsea_sweep_sync(fsea_sweep* s) {
    if (!s) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    return FSEA_OK;
}
// results / inputs back on the host (tests, small jobs): images of slots [first, first + count), raw IQ of a slot
extern "C" int fsea_sweep_read_images(fsea_sweep* s, int first, int count, uint8_t* h_out) {
    if (!s || !h_out || !s->d_images || first < 0 || count <= 0 || first + count > s->cfg.capacity) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(h_out, s->d_images + (size_t)first * s->img_bytes, (size_t)count * s->img_bytes, cudaMemcpyDeviceToHost));
    return FSEA_OK;
}
extern "C" int fsea_sweep_read_raw(fsea_sweep* s, int slot, uint8_t* h_out) {
    if (!s || !h_out || slot < 0 || slot >= s->cfg.capacity || !s->raw_ptr[slot]) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(h_out, s->raw_ptr[slot], s->raw_bytes, cudaMemcpyDeviceToHost));
    return FSEA_OK;
}
extern "C" uint8_t* fsea_sweep_images(fsea_sweep* s) { return s ? s->d_images : nullptr; }
extern "C" const void* fsea_sweep_raw(fsea_sweep* s, int slot) { return (s && slot >= 0 && slot < s->cfg.capacity) ? s->raw_ptr[slot] : nullptr; }
extern "C" void* fsea_sweep_stream(fsea_sweep* s) { return s ? (void*)s->stream : nullptr; }
extern "C" long fsea_sweep_launch_count(const fsea_sweep* s) { return s ? s->launches : 0; }

// fsea_engine.cu -- C ABI of the B200 IQ -> spectrum engine (include/fsea.h).
//
// Host side of the hot path: plans (twiddle / window tables in HBM, kernel
// selection, launch geometry), the device-resident batch call and the
// host-buffer call with a pinned, double-buffered H2D -> kernel -> D2H
// pipeline.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fcntl.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/fsea.h"
#include "fsea_dispatch.h"
#include "fsea_kparams.h"
#include "fsea_ptx.cuh"
#include "fsea_internal.h"


// ------------------------------------------------------------------ errors
static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return FSEA_ERR_CUDA;
}
#define CU(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

extern "C" const char* fsea_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char* fsea_strerror(int code) {
    switch (code) {
        case FSEA_OK: return "ok";
        case FSEA_ERR_INVALID: return "invalid argument";
        case FSEA_ERR_UNSUPPORTED: return "unsupported fft size (power of two, 16..65536)";
        case FSEA_ERR_NO_DEVICE: return "no CUDA device (this engine has no CPU fallback)";
        case FSEA_ERR_CUDA: return "CUDA runtime error";
        case FSEA_ERR_NOMEM: return "out of memory";
    }
    return "unknown error";
}

extern "C" int fsea_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------ plan
struct fsea_stage {                 // one pinned staging slot of the host pipeline
    void* h_in = nullptr;
    void* h_out = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    long frames = 0;                // frames in flight in this slot
    long first = 0;
};

struct fsea_plan {
    fsea_config cfg;
    int hop;
    int variant;
    fsea_kernel_info ki;
    int sample_bytes;               // bytes per complex input sample
    int out_elem_bytes;
    void* d_tw1 = nullptr;
    void* d_tw2 = nullptr;
    void* d_twh = nullptr;                       // "half" kernels: W_N^i of the output stage
    float* d_window = nullptr;
    int max_blocks = 0;             // resident blocks on the device
    long launches = 0;
    // host pipeline
    static const int NSTAGE = 3;
    fsea_stage stage[NSTAGE];
    long stage_frames = 0;
    double* d_gate = nullptr;
    unsigned* d_tickets = nullptr;             // TICKET_SLOTS x {ticket, done}, zero between launches
    static const int TICKET_SLOTS = 64;
    void* d_dump = nullptr;
    // four-step (N = N1 x N2 >= 8192, 8-bit input): both steps in one cluster kernel
    fsea_bigf* bigf = nullptr;                   // float frames longer than 8192 points (fsea_bigf.cu)
    bool four = false;
    int max_clusters = 0;                        // clusters (frames in flight) resident at once
    fsea_kernel_info4 k4;
    void* d_tw4 = nullptr;                       // comb twiddles: W_N^{2^r k}, r < log2 N2, as cos / sin planes
    static const int SCRATCH_SLOTS = 4;           // slot 0: fsea_exec_device, 1..3: host pipeline stages
    void* d_cscratch[SCRATCH_SLOTS] = {nullptr, nullptr, nullptr, nullptr};   // max_clusters x N float2 each (L2-resident)
    // A scratch slot is written and read back by every launch that uses it: launches of one slot
    // that go to DIFFERENT streams are ordered by an event (same stream: stream order does it).
    cudaEvent_t scratch_done[SCRATCH_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t scratch_stream[SCRATCH_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    bool scratch_used[SCRATCH_SLOTS] = {false, false, false, false};
    // gate (fsea_gate_mean_host): grow-only device buffers, no allocation per call
    void* d_gate_in = nullptr; size_t gate_in_bytes = 0;
    void* d_gate_out = nullptr; size_t gate_out_bytes = 0;
    // FSEA_ARITH_F64 on 8-bit input: samples widened to double (window applied in double) into `d_wide`, chunk by
    // chunk, and transformed by a float64 plan -- the arithmetic of the reference itself (src/nrf.c:599-615)
    fsea_plan* wide = nullptr;
    void* d_wide = nullptr;
    long wide_frames = 0;                        // frames per chunk
    double* d_window64 = nullptr;
    cudaEvent_t wide_done = nullptr;             // launches on different streams share d_wide: ordered by this event
    cudaStream_t wide_stream = nullptr;
    bool wide_used = false;
    char name_store[160] = "";
};

static fsea_get_kernel_fn size_table(int n) {
    switch (n) {
        case 16: return fsea_get_kernel_16;
        case 32: return fsea_get_kernel_32;
        case 64: return fsea_get_kernel_64;
        case 128: return fsea_get_kernel_128;
        case 256: return fsea_get_kernel_256;
        case 512: return fsea_get_kernel_512;
        case 1024: return fsea_get_kernel_1024;
        case 2048: return fsea_get_kernel_2048;
        case 4096: return fsea_get_kernel_4096;
        case 8192: return fsea_get_kernel_8192;
    }
    return nullptr;
}

extern "C" void fsea_config_default(fsea_config* cfg, int fft_size) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->fft_size = fft_size;
    cfg->hop = fft_size;
    cfg->in_format = FSEA_IN_U8;
    cfg->window = FSEA_WIN_NONE;
    cfg->out_mode = FSEA_OUT_MAG;
    cfg->flags = 0;
    cfg->pix_scale = 10.0f;
    cfg->device = 0;
    cfg->arith = FSEA_ARITH_AUTO;
    cfg->log_mult = 10.0f;
    cfg->power_gain = 1.0f;
}

template <typename S2>
static void fill_twiddles(std::vector<S2>& t1, std::vector<S2>& t2, const fsea_kernel_info& ki) {
    const long double PI2 = 6.283185307179586476925286766559005768L;
    const int R0 = ki.r0, R1 = ki.r1, R2 = ki.r2;
    t1.resize(std::max(ki.tw1, 1));
    t2.resize(std::max(ki.tw2, 1));
    for (int j = 1; j < R1; j++)
        for (int k = 0; k < R0; k++) {
            long double a = -PI2 * (long double)(j * k) / (long double)(R0 * R1);
            t1[(j - 1) * R0 + k].x = cosl(a);
            t1[(j - 1) * R0 + k].y = sinl(a);
        }
    const int M = R0 * R1;
    for (int j = 1; j < R2; j++)
        for (int i = 0; i < M; i++) {
            long double a = -PI2 * (long double)((long)j * i) / (long double)ki.n;
            t2[(j - 1) * M + i].x = cosl(a);
            t2[(j - 1) * M + i].y = sinl(a);
        }
}

extern "C" int fsea_plan_destroy(fsea_plan* p) {
    if (!p) return FSEA_OK;
    cudaSetDevice(p->cfg.device);
    for (int s = 0; s < fsea_plan::NSTAGE; s++) {
        fsea_stage& st = p->stage[s];
        if (st.stream) cudaStreamSynchronize(st.stream);
        if (st.h_in) cudaFreeHost(st.h_in);
        if (st.h_out) cudaFreeHost(st.h_out);
        if (st.d_in) cudaFree(st.d_in);
        if (st.d_out) cudaFree(st.d_out);
        if (st.done) cudaEventDestroy(st.done);
        if (st.stream) cudaStreamDestroy(st.stream);
    }
    if (p->d_tw1) cudaFree(p->d_tw1);
    if (p->d_tw2) cudaFree(p->d_tw2);
    if (p->d_twh) cudaFree(p->d_twh);
    if (p->d_window) cudaFree(p->d_window);
    if (p->d_gate) cudaFree(p->d_gate);
    if (p->d_tickets) cudaFree(p->d_tickets);
    if (p->d_dump) cudaFree(p->d_dump);
    if (p->d_tw4) cudaFree(p->d_tw4);
    fsea_bigf_destroy(p->bigf);
    for (int q = 0; q < fsea_plan::SCRATCH_SLOTS; q++) {
        if (p->d_cscratch[q]) cudaFree(p->d_cscratch[q]);
        if (p->scratch_done[q]) cudaEventDestroy(p->scratch_done[q]);
    }
    if (p->d_gate_in) cudaFree(p->d_gate_in);
    if (p->d_gate_out) cudaFree(p->d_gate_out);
    if (p->wide) fsea_plan_destroy(p->wide);
    if (p->d_wide) cudaFree(p->d_wide);
    if (p->d_window64) cudaFree(p->d_window64);
    if (p->wide_done) cudaEventDestroy(p->wide_done);
    delete p;
    return FSEA_OK;
}

// shared-memory opt-in + resident-block count of one kernel
static int setup_kernel(const fsea_kernel_info& ki, int device, int* max_blocks) {
    CU(cudaFuncSetAttribute(ki.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    int per_sm = 0, sms = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ki.fn, ki.threads, ki.smem));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) per_sm = 1;
    *max_blocks = per_sm * sms;
    return FSEA_OK;
}

// inter-pass twiddle tables of one kernel: evaluated in long double, stored in its precision
static int upload_twiddles(const fsea_kernel_info& ki, void** d1, void** d2, float scale1 = 1.0f) {
    if (ki.scalar_bytes == 4) {
        std::vector<float2> t1, t2;
        fill_twiddles(t1, t2, ki);
        for (auto& w : t1) { w.x *= scale1; w.y *= scale1; }   // exact: a power of two
        CU(cudaMalloc(d1, t1.size() * sizeof(float2)));
        CU(cudaMalloc(d2, t2.size() * sizeof(float2)));
        CU(cudaMemcpy(*d1, t1.data(), t1.size() * sizeof(float2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(*d2, t2.data(), t2.size() * sizeof(float2), cudaMemcpyHostToDevice));
    } else {
        std::vector<double2> t1, t2;
        fill_twiddles(t1, t2, ki);
        CU(cudaMalloc(d1, t1.size() * sizeof(double2)));
        CU(cudaMalloc(d2, t2.size() * sizeof(double2)));
        CU(cudaMemcpy(*d1, t1.data(), t1.size() * sizeof(double2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(*d2, t2.data(), t2.size() * sizeof(double2), cudaMemcpyHostToDevice));
    }
    return FSEA_OK;
}

// ------------------------------------------------------------------ FSEA_ARITH_F64 on 8-bit input
// x[n] = w[n] * sample[n] exactly as the reference forms it in double (src/nrf.c:599-612, c/fft-batch.c:61-66):
// b / 256 (after the HackRF byte flip) without a window, w (b - 128) / 256 with one; dense frames of N double2.
__global__ void fsea_widen_kernel(const uint8_t* __restrict__ in, long hop, int n, unsigned flip, const double* __restrict__ win,
                                  double2* __restrict__ out, long total) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long f = idx / n;
        const int i = (int)(idx - f * n);
        const uchar2 b = ((const uchar2*)in)[f * hop + i];
        const int vi = (int)(b.x ^ flip), vq = (int)(b.y ^ flip);            // (v + 128) % 256 == v ^ 0x80
        double re, im;
        if (win) { const double w = win[i]; re = w * ((double)(vi - 128) / 256.0); im = w * ((double)(vq - 128) / 256.0); }
        else { re = (double)vi / 256.0; im = (double)vq / 256.0; }
        out[idx] = make_double2(re, im);
    }
}

static int create_wide_plan(const fsea_config* cfg, fsea_plan** out) {
    const int n = cfg->fft_size;
    if (n < 16 || n > 65536 || (n & (n - 1))) return FSEA_ERR_UNSUPPORTED;
    if (cfg->window != FSEA_WIN_NONE && cfg->window != FSEA_WIN_HANN) return FSEA_ERR_INVALID;
    if (cfg->hop < 0 || (cfg->hop % 8) != 0) return FSEA_ERR_INVALID;
    fsea_config sub = *cfg;
    sub.in_format = FSEA_IN_F64C; sub.window = FSEA_WIN_NONE; sub.hop = n; sub.arith = FSEA_ARITH_AUTO;
    fsea_plan* inner = nullptr;
    int rc = fsea_plan_create(&sub, &inner);
    if (rc != FSEA_OK) return rc;
    fsea_plan* p = new (std::nothrow) fsea_plan();
    if (!p) { fsea_plan_destroy(inner); return FSEA_ERR_NOMEM; }
    p->cfg = *cfg;
    p->hop = cfg->hop > 0 ? cfg->hop : n;
    p->variant = FSEA_VAR_F64;
    p->sample_bytes = 2;
    p->out_elem_bytes = inner->out_elem_bytes;
    p->wide = inner;
    memset(&p->ki, 0, sizeof(p->ki));
    p->ki.n = n; p->ki.lanes = 1; p->ki.groups = 1; p->ki.threads = 256; p->ki.scalar_bytes = 8;
    snprintf(p->name_store, sizeof(p->name_store), "f64 <- 8-bit (widened in double): %s", inner->ki.name);
    p->ki.name = p->name_store;
    p->wide_frames = std::max<long>(2, (64L << 20) / ((long)n * 16));
    cudaError_t e;
    if ((e = cudaMalloc(&p->d_wide, (size_t)p->wide_frames * n * sizeof(double2))) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&p->wide_done, cudaEventDisableTiming)) != cudaSuccess) {
        rc = cuda_fail(e, "float64 sample buffer");
    }
    if (rc == FSEA_OK && cfg->window == FSEA_WIN_HANN) {
        std::vector<double> w(n);
        const long double PI2 = 6.283185307179586476925286766559005768L;
        for (int i = 0; i < n; i++) w[i] = (double)(0.5L * (1.0L - cosl(PI2 * i / n)));
        if ((e = cudaMalloc(&p->d_window64, n * sizeof(double))) != cudaSuccess ||
            (e = cudaMemcpy(p->d_window64, w.data(), n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) rc = cuda_fail(e, "float64 window");
    }
    if (rc != FSEA_OK) { fsea_plan_destroy(p); return rc; }
    *out = p;
    return FSEA_OK;
}

extern "C" int fsea_plan_create(const fsea_config* cfg, fsea_plan** out) {
    if (!cfg || !out) return FSEA_ERR_INVALID;
    *out = nullptr;
    const int n = cfg->fft_size;
    if (n < 16 || n > 65536 || (n & (n - 1))) return FSEA_ERR_UNSUPPORTED;
    const bool eight = cfg->in_format == FSEA_IN_U8 || cfg->in_format == FSEA_IN_S8;
    if (cfg->arith == FSEA_ARITH_F64) return eight ? create_wide_plan(cfg, out) : FSEA_ERR_INVALID;   // (float64 samples are always transformed in fp64)
    fsea_get_kernel_fn get = size_table(n);
    fsea_kernel_info4 k4;
    int alt = 0;
    if (const char* ev = getenv("FSEA_FOUR_ALT")) alt = atoi(ev);                 // A/B runs
    // N >= 16384 always runs four-step; N = 8192 can (FSEA_8192_FOUR, A/B runs: measured slower
    // than the single-kernel transform)
    bool want4 = !get || (n == 8192 && eight && cfg->arith == FSEA_ARITH_AUTO && getenv("FSEA_8192_FOUR"));
    // alt 0: the warp-specialised block-per-frame kernel where it exists (fsea_ws4.cuh), else the cluster kernel
    bool four = false;
    int ws_env = -1;
    if (const char* ev = getenv("FSEA_WS")) ws_env = atoi(ev);                       // A/B runs: 0 = never, 1 = wherever it exists
    const bool ws_default = (n == 16384 || n == 32768);      // measured: DESIGN.md 3.3 (65536: the L2 cluster kernel is still ahead, 3.68 vs 3.40 M frames/s)
    if (eight && cfg->arith == FSEA_ARITH_AUTO && alt == 0 && (ws_env < 0 ? ws_default : ws_env != 0) && n >= 4096) {
        four = n <= 16384 ? fsea_get_kernel_ws4(n, cfg->window == FSEA_WIN_HANN, cfg->out_mode, &k4)
                          : fsea_get_kernel_wsc(n, cfg->window == FSEA_WIN_HANN, cfg->out_mode, &k4);
        if (four) want4 = true;
    }
    if (!four) four = want4 && eight && fsea_get_kernel4(n, cfg->window == FSEA_WIN_HANN, cfg->out_mode, alt == 9 ? 0 : alt, &k4);
    const bool bigf = !get && !four && !eight && (cfg->in_format == FSEA_IN_F32C || cfg->in_format == FSEA_IN_F64C);   // float frames > 8192 points: fsea_bigf.cu
    if (!get && !four && !bigf) return FSEA_ERR_UNSUPPORTED;
    if (cfg->in_format < FSEA_IN_U8 || cfg->in_format > FSEA_IN_F64C) return FSEA_ERR_INVALID;
    if (cfg->window != FSEA_WIN_NONE && cfg->window != FSEA_WIN_HANN) return FSEA_ERR_INVALID;
    if (cfg->out_mode < FSEA_OUT_MAG || cfg->out_mode > FSEA_OUT_PIX) return FSEA_ERR_INVALID;
    if (cfg->hop < 0) return FSEA_ERR_INVALID;
    if (!(cfg->power_gain >= 0.0f) || !(cfg->log_mult == cfg->log_mult)) return FSEA_ERR_INVALID;   // (also rejects NaN)
    if ((cfg->in_format == FSEA_IN_U8 || cfg->in_format == FSEA_IN_S8) && (cfg->hop % 8) != 0)
        return FSEA_ERR_INVALID;          /* 8-bit frames are fetched in 16-byte pieces */
    if (fsea_device_count() <= 0) return FSEA_ERR_NO_DEVICE;
    if (cfg->device < 0 || cfg->device >= fsea_device_count()) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(cfg->device));

    fsea_plan* p = new (std::nothrow) fsea_plan();
    if (!p) return FSEA_ERR_NOMEM;
    p->cfg = *cfg;
    p->hop = cfg->hop > 0 ? cfg->hop : n;
    switch (cfg->in_format) {
        case FSEA_IN_U8:
        case FSEA_IN_S8:
            p->variant = (cfg->arith == FSEA_ARITH_F32) ? FSEA_VAR_F32 : FSEA_VAR_F32X2;
            // one frame per unit where that is the faster arrangement (measured)
            if (cfg->arith == FSEA_ARITH_F32X2_FRAME || (cfg->arith == FSEA_ARITH_AUTO && (n == 2048 || n == 4096 || n == 8192)))
                p->variant = FSEA_VAR_F32X2H;
            // N = 4096 without a window: 64 x 64 with one exchange (137 vs 129 M frames/s)
            if ((cfg->arith == FSEA_ARITH_F32X2_RADIX64 || cfg->arith == FSEA_ARITH_AUTO) && n == 4096 && cfg->window == FSEA_WIN_NONE)
                p->variant = FSEA_VAR_F32X2X;
            if (cfg->window == FSEA_WIN_HANN)
                p->variant = (p->variant == FSEA_VAR_F32) ? FSEA_VAR_F32_WIN : (p->variant == FSEA_VAR_F32X2H ? FSEA_VAR_F32X2H_WIN : FSEA_VAR_F32X2_WIN);
            p->sample_bytes = 2;
            break;
        case FSEA_IN_F32C: p->variant = FSEA_VAR_F32C; p->sample_bytes = 8; break;
        default: p->variant = FSEA_VAR_F64; p->sample_bytes = 16; break;
    }
    if (bigf) {
        memset(&p->ki, 0, sizeof(p->ki));
        p->ki.n = n; p->ki.lanes = 1; p->ki.groups = 1; p->ki.threads = 256;
        p->ki.scalar_bytes = cfg->in_format == FSEA_IN_F64C ? 8 : 4;
        p->ki.name = cfg->in_format == FSEA_IN_F64C ? "bigf<radix-16 Stockham, fp64, one launch per stage, f64 samples>" : "bigf<radix-16 Stockham, fp64, one launch per stage, f32 samples>";
    } else if (four) {
        p->four = true;
        p->k4 = k4;
        memset(&p->ki, 0, sizeof(p->ki));   // the step-A transform's shape: tables, lanes, names
        p->ki.fn = k4.fn; p->ki.threads = k4.threads; p->ki.groups = 1; p->ki.lanes = 2; p->ki.smem = k4.smem;
        p->ki.n = k4.n1; p->ki.r0 = k4.r0; p->ki.r1 = k4.r1; p->ki.r2 = k4.r2;
        p->ki.tw1 = k4.tw1; p->ki.tw2 = k4.tw2; p->ki.scalar_bytes = 4; p->ki.name = k4.name;
        p->variant = FSEA_VAR_F32X2;
    } else {
        memset(&p->ki, 0, sizeof(p->ki));
        if (!get(p->variant, &p->ki)) { delete p; return FSEA_ERR_UNSUPPORTED; }
    }
    p->out_elem_bytes = (cfg->out_mode == FSEA_OUT_PIX) ? 1 : p->ki.scalar_bytes;

    int rc = FSEA_OK;
    do {
        cudaError_t e;
        if (bigf) {
            if ((rc = fsea_bigf_create(n, cfg->in_format, &p->bigf)) != FSEA_OK) break;
        } else {
        if ((rc = setup_kernel(p->ki, cfg->device, &p->max_blocks)) != FSEA_OK) break;
        if (p->four && p->k4.fn_sums &&
            (e = cudaFuncSetAttribute(p->k4.fn_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->k4.smem)) != cudaSuccess) { rc = cuda_fail(e, "sums kernel setup"); break; }
        // staged 8-bit kernels transform integer-scaled samples (b - 128): the 1/256 of
        // src/nrf.c:605-606 rides on the pass-1 twiddles
        const bool staged = p->four || ((p->variant == FSEA_VAR_F32 || p->variant == FSEA_VAR_F32X2 ||
                                         p->variant == FSEA_VAR_F32_WIN || p->variant == FSEA_VAR_F32X2_WIN ||
                                         p->variant == FSEA_VAR_F32X2H || p->variant == FSEA_VAR_F32X2H_WIN) && p->ki.r1 > 1);
        if (!p->ki.x64 && (rc = upload_twiddles(p->ki, &p->d_tw1, &p->d_tw2, staged ? 1.0f / 256.0f : 1.0f)) != FSEA_OK) break;
        if (p->ki.x64) {
            // fsea_fft4096x_kernel: W_4096^{k0 t} / 256 as packed pairs (t = 2e, 2e+1), [e][k0], cos plane then sin plane
            std::vector<float2> tx((size_t)2 * 32 * 64);
            const long double PI2 = 6.283185307179586476925286766559005768L;
            for (int e = 0; e < 32; e++)
                for (int k0 = 0; k0 < 64; k0++) {
                    const long double a0 = -PI2 * (long double)(k0 * (2 * e)) / 4096.0L, a1 = -PI2 * (long double)(k0 * (2 * e + 1)) / 4096.0L;
                    tx[(size_t)e * 64 + k0] = make_float2((float)(cosl(a0) / 256.0L), (float)(cosl(a1) / 256.0L));
                    tx[(size_t)(32 + e) * 64 + k0] = make_float2((float)(sinl(a0) / 256.0L), (float)(sinl(a1) / 256.0L));
                }
            if ((e = cudaMalloc(&p->d_twh, tx.size() * sizeof(float2))) != cudaSuccess ||
                (e = cudaMemcpy(p->d_twh, tx.data(), tx.size() * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
                rc = cuda_fail(e, "radix-64 table"); break;
            }
        }
        if (p->ki.half) {
            // output stage of the "half" kernels: W_N^i for the butterflies i < M of the last pass
            const int rl = p->ki.r2 > 1 ? p->ki.r2 : (p->ki.r1 > 1 ? p->ki.r1 : p->ki.r0);
            const int m = p->ki.n / rl;
            std::vector<float2> th((size_t)m);
            const long double PI2 = 6.283185307179586476925286766559005768L;
            for (int i = 0; i < m; i++) {
                const long double a = -PI2 * (long double)i / (long double)n;
                th[i].x = (float)cosl(a); th[i].y = (float)sinl(a);
            }
            if ((e = cudaMalloc(&p->d_twh, th.size() * sizeof(float2))) != cudaSuccess ||
                (e = cudaMemcpy(p->d_twh, th.data(), th.size() * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
                rc = cuda_fail(e, "half-kernel table"); break;
            }
        }
        if (p->four) {
            // step B builds W_N^{n2 k} from the rows n2 = 1, 2, 4, .. of this table
            const int n1 = p->k4.n1;
            int lg = 0;
            while ((1 << lg) < p->k4.n2) lg++;
            std::vector<float> t4((size_t)2 * lg * n1);                    // planes: row 2r = cos, row 2r+1 = sin
            const long double PI2 = 6.283185307179586476925286766559005768L;
            for (int r = 0; r < lg; r++)
                for (int k = 0; k < n1; k++) {
                    const long double a = -PI2 * (long double)(((long)k) << r) / (long double)n;
                    t4[(size_t)(2 * r) * n1 + k] = (float)cosl(a);
                    t4[(size_t)(2 * r + 1) * n1 + k] = (float)sinl(a);
                }
            if ((e = cudaMalloc(&p->d_tw4, t4.size() * sizeof(float))) != cudaSuccess ||
                (e = cudaMemcpy(p->d_tw4, t4.data(), t4.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
                rc = cuda_fail(e, "four-step tables"); break;
            }
            // launch geometry: how many clusters of `cs` blocks (= frames in flight) are resident at once
            if (p->k4.cs > 1) {
                cudaLaunchConfig_t lc;
                memset(&lc, 0, sizeof(lc));
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = p->k4.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.gridDim = dim3((unsigned)(p->k4.cs * 64)); lc.blockDim = dim3((unsigned)p->k4.threads);
                lc.dynamicSmemBytes = p->k4.smem; lc.attrs = at; lc.numAttrs = 1;
                int nc = 0;
                if ((e = cudaOccupancyMaxActiveClusters(&nc, p->k4.fn, &lc)) != cudaSuccess) { rc = cuda_fail(e, "cluster kernel setup"); break; }
                if (nc < 1) { rc = FSEA_ERR_UNSUPPORTED; break; }
                p->max_clusters = nc;
            } else {
                p->max_clusters = p->max_blocks;
            }
            // the intermediate of every slot is allocated here, not at the first launch: a lazy
            // cudaMalloc would synchronise the device and break stream capture
            bool bad = false;
            for (int q = 0; q < fsea_plan::SCRATCH_SLOTS && !bad && !p->k4.ws; q++) {       // (ws: the intermediate lives in shared memory)
                if ((e = cudaMalloc(&p->d_cscratch[q], (size_t)p->max_clusters * n * sizeof(float2))) != cudaSuccess ||
                    (e = cudaEventCreateWithFlags(&p->scratch_done[q], cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(e, "four-step intermediate"); bad = true; }
            }
            if (bad) break;
        }
        }
        if ((e = cudaMalloc(&p->d_tickets, sizeof(unsigned) * 2 * fsea_plan::TICKET_SLOTS)) != cudaSuccess ||
            (e = cudaMemset(p->d_tickets, 0, sizeof(unsigned) * 2 * fsea_plan::TICKET_SLOTS)) != cudaSuccess) {
            rc = cuda_fail(e, "ticket counters"); break;
        }
        if ((e = cudaMalloc(&p->d_dump, (size_t)n * 8)) != cudaSuccess) { rc = cuda_fail(e, "dump row"); break; }
        if (cfg->window == FSEA_WIN_HANN) {
            std::vector<float> w(n);
            const long double PI2 = 6.283185307179586476925286766559005768L;
            for (int i = 0; i < n; i++) w[i] = (float)(0.5L * (1.0L - cosl(PI2 * i / n)));
            if ((e = cudaMalloc(&p->d_window, n * sizeof(float))) != cudaSuccess ||
                (e = cudaMemcpy(p->d_window, w.data(), n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
                rc = cuda_fail(e, "window upload"); break;
            }
        }
    } while (0);
    if (rc != FSEA_OK) { fsea_plan_destroy(p); return rc; }
    *out = p;
    return FSEA_OK;
}

extern "C" size_t fsea_input_bytes(const fsea_plan* p, long n_frames) {
    if (!p || n_frames <= 0) return 0;
    return ((size_t)(n_frames - 1) * p->hop + p->cfg.fft_size) * p->sample_bytes;
}
extern "C" size_t fsea_output_bytes(const fsea_plan* p, long n_frames) {
    if (!p || n_frames <= 0) return 0;
    return (size_t)n_frames * p->cfg.fft_size * p->out_elem_bytes;
}
extern "C" long fsea_frames_in(const fsea_plan* p, size_t n_samples) {
    if (!p || n_samples < (size_t)p->cfg.fft_size) return 0;
    return (long)((n_samples - p->cfg.fft_size) / p->hop) + 1;
}
extern "C" long fsea_plan_launch_count(const fsea_plan* p) { return p ? p->launches : 0; }
extern "C" const char* fsea_plan_kernel_name(const fsea_plan* p) { return p ? p->ki.name : ""; }


// The input of a launch as a 3-d TMA tensor of bytes: [frame][row n1 of 2 N2 bytes][byte], box = 16 bytes x 256 rows
// (fsea_wsc.cuh fetches a block's 32-byte column slice of a frame with 8 such boxes).  cuTensorMapEncodeTiled comes
// from the driver through the runtime's entry-point query: libfsea.so does not link libcuda.
#if !defined(FSEA_EMU)
typedef CUresult (*fsea_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#endif
static int make_input_tmap(fsea_tmap* tm, const void* d_in, int n2, long hop, long n_frames) {
#if defined(FSEA_EMU)
    tm->base = (const unsigned char*)d_in;
    tm->dim[0] = 2ull * n2; tm->dim[1] = 1024; tm->dim[2] = (unsigned long long)n_frames;
    tm->stride[0] = 1; tm->stride[1] = 2ull * n2; tm->stride[2] = 2ull * hop;
    tm->box[0] = 16; tm->box[1] = 256; tm->box[2] = 1;
    return FSEA_OK;
#else
    static fsea_encode_tiled_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) return FSEA_ERR_UNSUPPORTED;
        encode = (fsea_encode_tiled_fn)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)(2 * n2), 1024, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)(2 * n2), (cuuint64_t)(2 * hop)};          // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {16, 256, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(d_in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return FSEA_ERR_CUDA;
    }
    return FSEA_OK;
#endif
}

static int launch(fsea_plan* p, const void* d_in, void* d_out, long n_frames, cudaStream_t stream, int slot = 0, long out_stride = 0,
                  long out_group = 0, long out_group_stride = 0, long hop = 0, double* sums = nullptr, long sums_stride = 0,
                  long sums_offset = 0) {
    if (n_frames <= 0) return FSEA_OK;
    if (sums && !(p->four && p->k4.fn_sums)) return FSEA_ERR_UNSUPPORTED;     // (sums: fsea_kparams::frame_sums, no spectra stored)
    if (hop <= 0) hop = p->hop;                          // (hop > 0: this launch's frames lie `hop` samples apart, whatever the plan says)
    if (p->wide) {
        if (hop != p->hop) return FSEA_ERR_INVALID;
        const int n = p->cfg.fft_size;
        const long ostride = out_stride > 0 ? out_stride : (long)n * p->out_elem_bytes;
        if (p->wide_used && p->wide_stream != stream) CU(cudaStreamWaitEvent(stream, p->wide_done, 0));
        // (rows in groups, fsea_out_row(): a chunk never straddles two groups, so the inner launch sees plain rows)
        const long group = out_group > 0 ? out_group : n_frames;
        for (long g0 = 0; g0 < n_frames; g0 += group) {
            const long ng = std::min(group, n_frames - g0);
            uint8_t* gout = (uint8_t*)d_out + (out_group > 0 ? (size_t)(g0 / group) * out_group_stride : 0);
            for (long f0 = 0; f0 < ng; f0 += p->wide_frames) {
                const long nf = std::min(p->wide_frames, ng - f0);
                const long total = nf * n;
                const unsigned blocks = (unsigned)std::min<long>((total + 255) / 256, 148L * 16);
                FSEA_LAUNCH(fsea_widen_kernel, blocks, 256, stream, (const uint8_t*)d_in + (size_t)(g0 + f0) * p->hop * 2, (long)p->hop, n,
                            p->cfg.in_format == FSEA_IN_S8 ? 0x80u : 0u, (const double*)p->d_window64, (double2*)p->d_wide, total);
                CU(cudaGetLastError());
                p->launches++;
                const long before = p->wide->launches;
                const int rc = launch(p->wide, p->d_wide, gout + (size_t)f0 * ostride, nf, stream, slot, ostride);
                if (rc != FSEA_OK) return rc;
                p->launches += p->wide->launches - before;
            }
        }
        CU(cudaEventRecord(p->wide_done, stream));
        p->wide_stream = stream;
        p->wide_used = true;
        return FSEA_OK;
    }
    fsea_kparams kp;
    kp.in = d_in;
    kp.out = d_out;
    kp.n_seq = n_frames;
    kp.seq_stride = hop;
    kp.tw1 = p->d_tw1;
    kp.tw2 = p->d_tw2;
    kp.window = p->d_window;
    kp.pix_scale = p->cfg.pix_scale;
    {   // out = log_mult log10(gain P + 1e-20) = log_mult log10(2) log2(P + 1e-20 / gain) + log_mult log10(gain)
        const double mult = p->cfg.log_mult != 0.0f ? (double)p->cfg.log_mult : 10.0;
        const double gain = p->cfg.power_gain > 0.0f ? (double)p->cfg.power_gain : 1.0;
        kp.lg.dmul = mult * 0.30102999566398119521; kp.lg.deps = 1.0e-20 / gain; kp.lg.dadd = gain == 1.0 ? 0.0 : mult * log10(gain);
        kp.lg.mul = (float)kp.lg.dmul; kp.lg.eps = (float)kp.lg.deps; kp.lg.add = (float)kp.lg.dadd;
    }
    kp.flags = p->cfg.flags;
    const bool eight_bit = p->cfg.in_format == FSEA_IN_U8 || p->cfg.in_format == FSEA_IN_S8;
    kp.dc_term = (eight_bit && p->cfg.window == FSEA_WIN_NONE && !(p->cfg.flags & FSEA_FLAG_DC_PATCH)) ? 1 : 0;
    kp.out_mode = p->cfg.out_mode;
    kp.xor_mask = p->cfg.in_format == FSEA_IN_S8 ? 0x8080u : 0u;
    kp.dump = p->d_dump;
    kp.tw4 = p->d_tw4;
    kp.twh = p->d_twh;
    kp.n1 = kp.n2 = 0;
    kp.scratch = nullptr;
    kp.out_stride = out_stride > 0 ? out_stride : (long)p->cfg.fft_size * p->out_elem_bytes;
    kp.out_group = out_group;
    kp.out_group_stride = out_group_stride;
    kp.tmap_arg = 0;
    kp.frame_sums = sums; kp.sums_stride = sums_stride; kp.sums_offset = sums_offset;
    if (p->bigf) return fsea_bigf_launch(p->bigf, kp, n_frames, stream, &p->launches);
    if (p->four) {
        // one launch: a cluster per frame, the intermediate lives in the clusters' L2-resident slots
        const int n = p->cfg.fft_size;
        kp.n1 = p->k4.n1; kp.n2 = p->k4.n2;
        (void)n;
        if (!p->k4.ws && p->scratch_used[slot] && p->scratch_stream[slot] != stream) CU(cudaStreamWaitEvent(stream, p->scratch_done[slot], 0));
        kp.scratch = p->d_cscratch[slot];
        kp.ticket = nullptr;
        const long nc = p->k4.ws ? std::min<long>((n_frames + p->k4.ws - 1) / p->k4.ws, p->max_clusters) : std::min<long>(n_frames, p->max_clusters);
        fsea_tmap tmap;
        memset(&tmap, 0, sizeof(tmap));
        void* af[] = {&kp, &tmap};
        kp.tmap_arg = p->k4.tma ? 1 : 0;
        if (p->k4.tma == 1) {
            const int trc = make_input_tmap(&tmap, d_in, p->k4.n2, hop, n_frames);
            if (trc != FSEA_OK) return trc;
        }
        if (p->k4.cs > 1) {
            cudaLaunchConfig_t lc;
            memset(&lc, 0, sizeof(lc));
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = p->k4.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.gridDim = dim3((unsigned)(nc * p->k4.cs)); lc.blockDim = dim3((unsigned)p->k4.threads);
            lc.dynamicSmemBytes = p->k4.smem; lc.stream = stream; lc.attrs = at; lc.numAttrs = 1;
            CU(cudaLaunchKernelExC(&lc, p->k4.fn, af));
        } else {
            CU(cudaLaunchKernel(sums ? p->k4.fn_sums : p->k4.fn, dim3((unsigned)nc), dim3((unsigned)p->k4.threads), af, p->k4.smem, stream));
        }
        if (!p->k4.ws) {
            CU(cudaEventRecord(p->scratch_done[slot], stream));
            p->scratch_stream[slot] = stream;
            p->scratch_used[slot] = true;
        }
        p->launches++;
        return FSEA_OK;
    }
    kp.ticket = p->d_tickets + 2 * (p->launches % fsea_plan::TICKET_SLOTS);
    const long units = (n_frames + p->ki.lanes - 1) / p->ki.lanes;
    long blocks = (units + p->ki.groups - 1) / p->ki.groups;
    if (blocks > p->max_blocks) blocks = p->max_blocks;
    void* args[] = {&kp};
    CU(cudaLaunchKernel(p->ki.fn, dim3((unsigned)blocks), dim3((unsigned)p->ki.threads), args, p->ki.smem, stream));
    p->launches++;
    return FSEA_OK;
}

extern "C" int fsea_exec_device(fsea_plan* p, const void* d_in, void* d_out, long n_frames, void* stream) {
    if (!p || n_frames < 0 || (n_frames > 0 && (!d_in || !d_out))) return FSEA_ERR_INVALID;
    if (((uintptr_t)d_in & 15) != 0) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(p->cfg.device));
    return launch(p, d_in, d_out, n_frames, (cudaStream_t)stream);
}

extern "C" int fsea_exec_device_strided(fsea_plan* p, const void* d_in, void* d_out, long out_row_stride_bytes, long n_frames, void* stream) {
    if (!p || n_frames < 0 || (n_frames > 0 && (!d_in || !d_out))) return FSEA_ERR_INVALID;
    if (((uintptr_t)d_in & 15) != 0) return FSEA_ERR_INVALID;
    const long packed = (long)p->cfg.fft_size * p->out_elem_bytes;
    if (out_row_stride_bytes < packed || (out_row_stride_bytes % (p->out_elem_bytes == 1 ? 4 : p->out_elem_bytes)) != 0 ||
        ((uintptr_t)d_out % (p->out_elem_bytes == 1 ? 4 : 8)) != 0) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(p->cfg.device));
    return launch(p, d_in, d_out, n_frames, (cudaStream_t)stream, 0, out_row_stride_bytes);
}

// the sweep engine (fsea_sweep.cu) launches plans on its own stream
int fsea_launch_internal(fsea_plan* p, const void* d_in, void* d_out, long n_frames, cudaStream_t stream, long out_stride, int scratch_slot,
                         long out_group, long out_group_stride, long hop) {
    return launch(p, d_in, d_out, n_frames, stream, scratch_slot, out_stride, out_group, out_group_stride, hop);
}
bool fsea_plan_has_sums_internal(const fsea_plan* p) { return p->four && p->k4.fn_sums != nullptr && p->k4.cs == 1; }
int fsea_launch_sums_internal(fsea_plan* p, const void* d_in, long n_frames, cudaStream_t stream, long hop, double* sums, long sums_stride,
                              long sums_offset) {
    return launch(p, d_in, p->d_dump, n_frames, stream, 0, 0, 0, 0, hop, sums, sums_stride, sums_offset);
}
int fsea_fail_cuda_internal(cudaError_t e, const char* what) { return cuda_fail(e, what); }
int fsea_plan_elem_bytes_internal(const fsea_plan* p) { return p->out_elem_bytes; }
void fsea_count_launch_internal(fsea_plan* p) { p->launches++; }

// ------------------------------------------------------------------ host placement
// A rank's pinned buffers and the thread that feeds its copy engines belong on the NUMA node
// its GPU hangs off: sysfs names the CPUs local to the device's PCI function.
extern "C" int fsea_host_bind_near_device(int device) {
#if defined(FSEA_EMU)
    (void)device; return 0;
#else
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return FSEA_ERR_INVALID; }
    for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* fp = fopen(path, "r");
    if (!fp) return 0;
    char list[1024] = "";
    const bool ok = fgets(list, sizeof(list), fp) != nullptr;
    fclose(fp);
    if (!ok) return 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_cpus = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int got = sscanf(tok, "%d-%d", &a, &b);
        if (got == 1) b = a;
        if (got >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n_cpus++; }
    }
    if (n_cpus == 0) return 0;
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return 0;           // (a cgroup may forbid the CPUs: leave things as they are)
    return n_cpus;
#endif
}

// ------------------------------------------------------------------ host pipeline
// Three slots, each with its own stream: while slot k computes, slot k+1's
// input crosses PCIe and slot k-1's spectra return.  Caller memory that is
// already page-locked (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch
// tensor) is copied from / to directly; pageable memory goes through the
// plan's pinned staging buffers.
static bool is_pinned_host(const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_stages(fsea_plan* p, long want_frames, bool staging) {
    // slot size: ~32 MiB of output or the whole job if smaller
    const size_t out_per = (size_t)p->cfg.fft_size * p->out_elem_bytes;
    long sf = (long)std::max<size_t>(2, (32u << 20) / out_per);
    if (sf > want_frames) sf = want_frames;
    if (sf & 1) sf++;                                   // keep frame pairs together
    const bool have_staging = p->stage[0].h_in != nullptr;
    if (p->stage_frames >= sf && (!staging || have_staging)) return FSEA_OK;
    if (sf < p->stage_frames) sf = p->stage_frames;
    for (int s = 0; s < fsea_plan::NSTAGE; s++) {
        fsea_stage& st = p->stage[s];
        if (st.stream) CU(cudaStreamSynchronize(st.stream));
        if (st.h_in) cudaFreeHost(st.h_in);
        if (st.h_out) cudaFreeHost(st.h_out);
        if (st.d_in) cudaFree(st.d_in);
        if (st.d_out) cudaFree(st.d_out);
        st.h_in = st.h_out = st.d_in = st.d_out = nullptr;
        const size_t ib = fsea_input_bytes(p, sf), ob = fsea_output_bytes(p, sf);
        if (staging || have_staging) {
            CU(cudaMallocHost(&st.h_in, ib));
            CU(cudaMallocHost(&st.h_out, ob));
        }
        CU(cudaMalloc(&st.d_in, ib));
        CU(cudaMalloc(&st.d_out, ob));
        if (!st.stream) CU(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        if (!st.done) CU(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    }
    p->stage_frames = sf;
    return FSEA_OK;
}

// `fill` (may be empty) loads the raw bytes of frames [first, first + nf) into a pinned slot;
// without it the bytes come from h_in (directly when that is page-locked).
template <class FILL>
static int exec_pipeline(fsea_plan* p, const void* h_in, void* h_out, long n_frames, bool from_source, FILL fill, bool copy_only = false) {
    CU(cudaSetDevice(p->cfg.device));
    const bool out_pinned = is_pinned_host(h_out);
    const bool direct = !from_source && is_pinned_host(h_in) && out_pinned;     // no staging copy on either side
    const bool stage_out = from_source ? !out_pinned : !direct;                 // spectra return through the slot's pinned buffer
    int rc = ensure_stages(p, n_frames, !direct);
    if (rc != FSEA_OK) return rc;
    const size_t in_frame_stride = (size_t)p->hop * p->sample_bytes;
    const size_t out_per = (size_t)p->cfg.fft_size * p->out_elem_bytes;
    const long sf = p->stage_frames;
    long issued = 0;
    int slot = 0;
    long pending[fsea_plan::NSTAGE] = {0, 0, 0};
    long pending_first[fsea_plan::NSTAGE] = {0, 0, 0};
    auto drain = [&](int s) -> int {
        if (!pending[s]) return FSEA_OK;
        fsea_stage& st = p->stage[s];
        CU(cudaEventSynchronize(st.done));
        if (stage_out)
            memcpy((uint8_t*)h_out + (size_t)pending_first[s] * out_per, st.h_out, (size_t)pending[s] * out_per);
        pending[s] = 0;
        return FSEA_OK;
    };
    while (issued < n_frames) {
        fsea_stage& st = p->stage[slot];
        if ((rc = drain(slot)) != FSEA_OK) break;
        const long nf = std::min(sf, n_frames - issued);
        const size_t ib = fsea_input_bytes(p, nf);
        const uint8_t* src = (const uint8_t*)h_in + (size_t)issued * in_frame_stride;
        uint8_t* dst = (uint8_t*)h_out + (size_t)issued * out_per;
        if (from_source) {
            if ((rc = fill(st.h_in, issued, nf, ib)) != FSEA_OK) break;
            src = (const uint8_t*)st.h_in;
            if (stage_out) dst = (uint8_t*)st.h_out;
        } else if (!direct) {
            memcpy(st.h_in, src, ib); src = (const uint8_t*)st.h_in; dst = (uint8_t*)st.h_out;
        }
        cudaError_t ce;
        if ((ce = cudaMemcpyAsync(st.d_in, src, ib, cudaMemcpyHostToDevice, st.stream)) != cudaSuccess) { rc = cuda_fail(ce, "H2D"); break; }
        if (!copy_only && (rc = launch(p, st.d_in, st.d_out, nf, st.stream, 1 + slot)) != FSEA_OK) break;
        if ((ce = cudaMemcpyAsync(dst, st.d_out, (size_t)nf * out_per, cudaMemcpyDeviceToHost, st.stream)) != cudaSuccess ||
            (ce = cudaEventRecord(st.done, st.stream)) != cudaSuccess) { rc = cuda_fail(ce, "D2H"); break; }
        pending[slot] = nf;
        pending_first[slot] = issued;
        issued += nf;
        slot = (slot + 1) % fsea_plan::NSTAGE;
    }
    // (also on an error: copies already in flight land in caller memory before this returns)
    for (int s = 0; s < fsea_plan::NSTAGE; s++) {
        const int rd = drain((slot + s) % fsea_plan::NSTAGE);
        if (rc == FSEA_OK) rc = rd;
    }
    return rc;
}

// The same chunks over the same three streams with the kernel left out: what the box's PCIe
// links allow for this job (bench.py reports e2e as a fraction of it).  h_out receives the
// (stale) contents of the device slots.
extern "C" int fsea_copy_only_host(fsea_plan* p, const void* h_in, void* h_out, long n_frames) {
    if (!p || n_frames < 0 || (n_frames > 0 && (!h_in || !h_out))) return FSEA_ERR_INVALID;
    if (n_frames == 0) return FSEA_OK;
    return exec_pipeline(p, h_in, h_out, n_frames, false, [](void*, long, long, size_t) { return (int)FSEA_OK; }, true);
}

extern "C" int fsea_exec_host(fsea_plan* p, const void* h_in, void* h_out, long n_frames) {
    if (!p || n_frames < 0 || (n_frames > 0 && (!h_in || !h_out))) return FSEA_ERR_INVALID;
    if (n_frames == 0) return FSEA_OK;
    return exec_pipeline(p, h_in, h_out, n_frames, false, [](void*, long, long, size_t) { return (int)FSEA_OK; });
}

// Capture files are the bytes the SDR delivered, block after block (c/rfcap.c:26-41 one file per
// block, c/reader.c:25-49 one file per run; src/nrf.c:256-284 replays them): no header, interleaved
// 8-bit I,Q.  pread() fills the pinned slot of the pipeline while the other slots are in flight.
extern "C" int fsea_exec_file(fsea_plan* p, const char* path, long byte_offset, long n_frames, void* h_out) {
    if (!p || !path || byte_offset < 0 || n_frames < 0 || (n_frames > 0 && !h_out)) return FSEA_ERR_INVALID;
    if (p->cfg.in_format != FSEA_IN_U8 && p->cfg.in_format != FSEA_IN_S8) return FSEA_ERR_INVALID;
    if (byte_offset & 15) return FSEA_ERR_INVALID;                  // frames are fetched in 16-byte pieces
    if (n_frames == 0) return FSEA_OK;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return FSEA_ERR_INVALID;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < (size_t)byte_offset + fsea_input_bytes(p, n_frames)) { close(fd); return FSEA_ERR_INVALID; }
    const size_t stride = (size_t)p->hop * p->sample_bytes;
    const int rc = exec_pipeline(p, nullptr, h_out, n_frames, true, [&](void* dst, long first, long, size_t bytes) {
        size_t got = 0;
        while (got < bytes) {
            const ssize_t r = pread(fd, (char*)dst + got, bytes - got, (off_t)(byte_offset + first * stride + got));
            if (r <= 0) return (int)FSEA_ERR_INVALID;
            got += (size_t)r;
        }
        return (int)FSEA_OK;
    });
    close(fd);
    return rc;
}

// ------------------------------------------------------------------ sweep gate
// c/fft-batch-broad.c:80-98: mean |X| over the first `rows` rows.  The rows'
// magnitudes are produced by the MAG kernel and summed in fp64 on the device.
// Deterministic: every block leaves ONE partial sum (lanes by shuffle, warps in index order), the
// host adds the 256 partials in block order -- the same bits on every run and on every rank.
__global__ void fsea_sum_kernel(const float* __restrict__ x, long n, double* __restrict__ partial) {
    FSEA_STATIC_SMEM(double, warp_sum, 8);
    double s = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += warp_sum[w];
        partial[blockIdx.x] = t;
    }
}

extern "C" int fsea_gate_mean_host(fsea_plan* p, const void* h_in, long rows, double* mean_out) {
    if (!p || !h_in || !mean_out || rows <= 0) return FSEA_ERR_INVALID;
    if (p->cfg.out_mode != FSEA_OUT_MAG || p->ki.scalar_bytes != 4) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(p->cfg.device));
    const size_t ib = fsea_input_bytes(p, rows), ob = fsea_output_bytes(p, rows);
    if (!p->d_gate) CU(cudaMalloc(&p->d_gate, sizeof(double) * FSEA_SUM_BLOCKS));
    if (p->gate_in_bytes < ib) {                                   // grow-only: nothing is allocated per call once warm
        if (p->d_gate_in) cudaFree(p->d_gate_in);
        p->d_gate_in = nullptr; p->gate_in_bytes = 0;
        CU(cudaMalloc(&p->d_gate_in, ib));
        p->gate_in_bytes = ib;
    }
    if (p->gate_out_bytes < ob) {
        if (p->d_gate_out) cudaFree(p->d_gate_out);
        p->d_gate_out = nullptr; p->gate_out_bytes = 0;
        CU(cudaMalloc(&p->d_gate_out, ob));
        p->gate_out_bytes = ob;
    }
    void *d_in = p->d_gate_in, *d_out = p->d_gate_out;
    cudaError_t e;
    int rc = FSEA_OK;
    do {
        if ((e = cudaMemcpy(d_in, h_in, ib, cudaMemcpyHostToDevice)) != cudaSuccess) { rc = cuda_fail(e, "H2D"); break; }
        if ((rc = launch(p, d_in, d_out, rows, 0)) != FSEA_OK) break;
        const long n = rows * p->cfg.fft_size;
        FSEA_LAUNCH(fsea_sum_kernel, FSEA_SUM_BLOCKS, 256, (cudaStream_t)0, (const float*)d_out, n, p->d_gate);
        p->launches++;
        double part[FSEA_SUM_BLOCKS], total = 0;
        if ((e = cudaMemcpy(part, p->d_gate, sizeof(part), cudaMemcpyDeviceToHost)) != cudaSuccess) { rc = cuda_fail(e, "D2H"); break; }
        for (int b = 0; b < FSEA_SUM_BLOCKS; b++) total += part[b];
        *mean_out = total / (double)n;
    } while (0);
    return rc;
}

// ------------------------------------------------------------------ sweep stitch
// c/fft-stitch-broad.c:28-36,66-88 (and c/fft-stitch.c:46-54,167-189): the per-frequency
// pixel images of a sweep are max-blended side by side, image i at column i * width_step,
// into one picture of width  width + (n_images - 1) * width_step  (zero where nothing lands).
// One thread per 4 output pixels; every source pixel is read once per output it covers.
// (world, per: the images as a sharded sweep gathers them -- frequency i is image i / world of
// rank i % world, ranks `per` images apart; world = 1: plain frequency order)
__global__ void fsea_stitch_kernel(const uint8_t* __restrict__ img, int n_images, int rows, int width, int step,
                                   int out_width, uint8_t* __restrict__ out, int world, int per) {
    const long quads_per_row = (out_width + 3) / 4;
    const long total = (long)rows * quads_per_row;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x) {
        const long y = q / quads_per_row;
        const int x0 = (int)(q - y * quads_per_row) * 4;
        uint8_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = x0 + k;
            if (x >= out_width) break;
            // images i with 0 <= x - i*step < width
            int hi = x / step;
            if (hi > n_images - 1) hi = n_images - 1;
            int lo = (x - width + step) / step;          // ceil((x - width + 1) / step)
            if (x - width + 1 <= 0) lo = 0;
            uint8_t m = 0;
            for (int i = lo; i <= hi; i++) {
                const size_t at = (world == 1) ? (size_t)i : (size_t)(i % world) * per + (size_t)(i / world);
                const uint8_t s = img[(at * rows + y) * width + (x - i * step)];
                m = s > m ? s : m;
            }
            v[k] = m;
        }
        uint8_t* o = out + (size_t)y * out_width + x0;
        for (int k = 0; k < 4 && x0 + k < out_width; k++) o[k] = v[k];
    }
}

// 16 pixels per thread (one 128-bit load per image that covers them, bytewise max in 4 VMAX.U8x4): the form
// every real sweep has -- width, step and both buffers multiples of 16
__global__ void fsea_stitch16_kernel(const uint4* __restrict__ img, int n_images, int rows, int width16, int step16,
                                     int out_width16, uint4* __restrict__ out, int world, int per) {
    const long total = (long)rows * out_width16;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x) {
        const long y = q / out_width16;
        const int x = (int)(q - y * out_width16);
        int hi = x / step16;
        if (hi > n_images - 1) hi = n_images - 1;
        const int lo = (x - width16 + 1 <= 0) ? 0 : (x - width16 + step16) / step16;
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (int i = lo; i <= hi; i++) {
            const size_t at = (world == 1) ? (size_t)i : (size_t)(i % world) * per + (size_t)(i / world);
            const uint4 v = img[(at * rows + y) * width16 + (x - i * step16)];
            acc.x = __vmaxu4(acc.x, v.x); acc.y = __vmaxu4(acc.y, v.y); acc.z = __vmaxu4(acc.z, v.z); acc.w = __vmaxu4(acc.w, v.w);
        }
        out[q] = acc;
    }
}

static int stitch_launch(const uint8_t* d_images, int n_images, int world, int per, int rows, int width, int width_step,
                         uint8_t* d_out, cudaStream_t stream) {
    const int out_width = width + (n_images - 1) * width_step;
    if ((width % 16) == 0 && (width_step % 16) == 0 && (((uintptr_t)d_images | (uintptr_t)d_out) & 15) == 0) {
        const long items = (long)rows * (out_width / 16);
        const int blocks = (int)std::min<long>((items + 255) / 256, 148 * 32);
        FSEA_LAUNCH(fsea_stitch16_kernel, blocks, 256, stream, (const uint4*)d_images, n_images, rows, width / 16, width_step / 16,
                    out_width / 16, (uint4*)d_out, world, per);
    } else {
        const long quads = (long)rows * ((out_width + 3) / 4);
        const int blocks = (int)std::min<long>((quads + 255) / 256, 148 * 16);
        FSEA_LAUNCH(fsea_stitch_kernel, blocks, 256, stream, d_images, n_images, rows, width, width_step, out_width, d_out, world, per);
    }
    CU(cudaGetLastError());
    return FSEA_OK;
}

extern "C" int fsea_stitch_device(const uint8_t* d_images, int n_images, int rows, int width, int width_step,
                                  uint8_t* d_out, void* stream) {
    if (!d_images || !d_out || n_images <= 0 || rows <= 0 || width <= 0 || width_step <= 0) return FSEA_ERR_INVALID;
    return stitch_launch(d_images, n_images, 1, n_images, rows, width, width_step, d_out, (cudaStream_t)stream);
}

extern "C" int fsea_stitch_sharded_device(const uint8_t* d_images, int n_images, int world, int per_rank, int rows, int width, int width_step,
                                          uint8_t* d_out, void* stream) {
    if (!d_images || !d_out || n_images <= 0 || rows <= 0 || width <= 0 || width_step <= 0 || world <= 0 || per_rank * world < n_images) return FSEA_ERR_INVALID;
    return stitch_launch(d_images, n_images, world, per_rank, rows, width, width_step, d_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ history image on the device
// nrf_fft keeps H rows of N magnitudes, newest first; the reference scrolls the whole image by one row per frame
// (src/nrf.c:617, 8 MiB at 1024 x 1024) and, on a retune, every row sideways (src/nrf.c:569-596).  Here the image
// is a ring of float64 rows in device memory: a new spectrum is ONE row written in place (widened from the fp32
// kernels' output), the retune scroll is one kernel.
__global__ void fsea_history_put_kernel(const void* __restrict__ row, int row_is_f64, double* __restrict__ dst, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[i] = row_is_f64 ? ((const double*)row)[i] : (double)((const float*)row)[i];
}
// every row scrolled by s bins: s > 0 left (row[i] = old[i + s]), s < 0 right, vacated bins zero (src/nrf.c:580-595)
__global__ void fsea_history_shift_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int rows, int s) {
    const long total = (long)n * rows;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x) {
        const int i = (int)(q % n);
        const int from = i + s;
        dst[q] = (from >= 0 && from < n) ? src[q + s] : 0.0;
    }
}

extern "C" int fsea_history_put_device(const void* d_row, int row_is_f64, double* d_ring_row, int n, void* stream) {
    if (!d_row || !d_ring_row || n <= 0) return FSEA_ERR_INVALID;
    FSEA_LAUNCH(fsea_history_put_kernel, (n + 255) / 256 < 64 ? (n + 255) / 256 : 64, 256, (cudaStream_t)stream, d_row, row_is_f64, d_ring_row, n);
    CU(cudaGetLastError());
    return FSEA_OK;
}

extern "C" int fsea_history_shift_device(const double* d_src, double* d_dst, int n, int rows, int s, void* stream) {
    if (!d_src || !d_dst || d_src == d_dst || n <= 0 || rows <= 0) return FSEA_ERR_INVALID;
    if (s >= n || s <= -n) {                                            // everything scrolls out: clear (src/nrf.c:574-577)
        CU(cudaMemsetAsync(d_dst, 0, sizeof(double) * (size_t)n * rows, (cudaStream_t)stream));
        return FSEA_OK;
    }
    const long total = (long)n * rows;
    const int blocks = (int)std::min<long>((total + 255) / 256, 148 * 8);
    FSEA_LAUNCH(fsea_history_shift_kernel, blocks, 256, (cudaStream_t)stream, d_src, d_dst, n, rows, s);
    CU(cudaGetLastError());
    return FSEA_OK;
}

// ------------------------------------------------------------------ device memory another process of the box can store into
// A sweep's stitched canvas lives on one GPU; with one process per GPU (torchrun) the other ranks reach it through
// CUDA IPC: the owner allocates and exports, the others open the 64-byte handle and get a pointer their kernels can
// store through (peer access over NVLink).
extern "C" int fsea_ipc_alloc(int device, size_t bytes, void** d_ptr, unsigned char handle[64]) {
    if (!d_ptr || !handle || bytes == 0) return FSEA_ERR_INVALID;
#if defined(FSEA_EMU)
    (void)device; return FSEA_ERR_UNSUPPORTED;
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CU(cudaSetDevice(device));
    CU(cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    return FSEA_OK;
#endif
}
extern "C" int fsea_ipc_open(int device, const unsigned char handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return FSEA_ERR_INVALID;
#if defined(FSEA_EMU)
    (void)device; return FSEA_ERR_UNSUPPORTED;
#else
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FSEA_OK;
#endif
}
extern "C" int fsea_ipc_close(void* d_ptr) {
#if defined(FSEA_EMU)
    (void)d_ptr; return FSEA_OK;
#else
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr));
    return FSEA_OK;
#endif
}
extern "C" int fsea_ipc_free(void* d_ptr) {
    if (d_ptr) CU(cudaFree(d_ptr));
    return FSEA_OK;
}

// Device scratch of the *_host convenience calls (stitch, FIR): one grow-only buffer per device
// and purpose, so that a call allocates nothing once warm.  The lock is held for the whole call.
static std::mutex g_scratch_mu;
static void* g_scratch[64][2];
static size_t g_scratch_bytes[64][2];
static int scratch_get(int device, int which, size_t bytes, void** out) {
    if (device < 0 || device >= 64) return FSEA_ERR_INVALID;
    if (g_scratch_bytes[device][which] < bytes) {
        if (g_scratch[device][which]) cudaFree(g_scratch[device][which]);
        g_scratch[device][which] = nullptr; g_scratch_bytes[device][which] = 0;
        CU(cudaMalloc(&g_scratch[device][which], bytes));
        g_scratch_bytes[device][which] = bytes;
    }
    *out = g_scratch[device][which];
    return FSEA_OK;
}

extern "C" int fsea_stitch_host(int device, const uint8_t* h_images, int n_images, int rows, int width, int width_step,
                                uint8_t* h_out) {
    if (!h_images || !h_out || n_images <= 0 || rows <= 0 || width <= 0 || width_step <= 0) return FSEA_ERR_INVALID;
    if (fsea_device_count() <= 0) return FSEA_ERR_NO_DEVICE;
    CU(cudaSetDevice(device));
    const size_t ib = (size_t)n_images * rows * width, ob = (size_t)rows * (width + (size_t)(n_images - 1) * width_step);
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    uint8_t *d_in = nullptr, *d_out = nullptr;
    int rc = scratch_get(device, 0, ib, (void**)&d_in);
    if (rc == FSEA_OK) rc = scratch_get(device, 1, ob, (void**)&d_out);
    if (rc != FSEA_OK) return rc;
    cudaError_t e;
    if ((e = cudaMemcpy(d_in, h_images, ib, cudaMemcpyHostToDevice)) != cudaSuccess) rc = cuda_fail(e, "H2D");
    if (rc == FSEA_OK) rc = fsea_stitch_device(d_in, n_images, rows, width, width_step, d_out, nullptr);
    if (rc == FSEA_OK && (e = cudaMemcpy(h_out, d_out, ob, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "D2H");
    return rc;
}

// ------------------------------------------------------------------ IQ low-pass (FIR)
// src/nrf.c:718-724,759-768 (nrf_fir_filter_get per output of nrf_iq_filter_get_buffer):
//   out[2i] = sum_k c[k] si[i+k],  out[2i+1] = sum_k c[k] sq[i+k],  k = 0 .. taps-1 in this order,
// si / sq = the filter's sample line (taps-1 carried-over samples, then the new block).
// float64, every product and every sum rounded on its own (__dmul_rn / __dadd_rn: no contraction),
// accumulated from 0 in the reference's order: the result is bit-identical to the reference.
#if defined(FSEA_EMU)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
#endif
// A block of 256 threads produces 1024 consecutive outputs, a thread four neighbouring ones: every
// sample is read from shared memory once per thread and feeds up to four accumulators (k = j - r),
// each of which still sees its taps in ascending order.  The sample lines are staged
// de-interleaved by (index mod 4), so that the 32 lanes of a load hit consecutive doubles; the
// finished tile goes out through shared memory as coalesced 16-byte (I, Q) stores.
#define FSEA_FIR_TPB 256
#define FSEA_FIR_R 4
static inline int fir_quarter(int taps) {                      // doubles per residue class, = 8 mod 16 (bank spread of the staging stores)
    const int q = (FSEA_FIR_TPB * FSEA_FIR_R + taps - 1 + 3) / 4;
    return ((q + 7) / 16) * 16 + 8;
}
__global__ void __launch_bounds__(FSEA_FIR_TPB)
fsea_fir_iq_kernel(const double* __restrict__ coef, int taps, int quarter, const double* __restrict__ si,
                   const double* __restrict__ sq, long length, double* __restrict__ out) {
    FSEA_DYN_SMEM(fir_smem);
    constexpr int TILE = FSEA_FIR_TPB * FSEA_FIR_R;
    const int span = TILE + taps - 1;
    double* c = (double*)fir_smem;
    double* ti = c + ((taps + 1) & ~1);
    double* tq = ti + 4 * quarter;
    const int t = threadIdx.x;
    for (int k = t; k < taps; k += FSEA_FIR_TPB) c[k] = coef[k];
    for (long i0 = (long)blockIdx.x * TILE; i0 < length; i0 += (long)gridDim.x * TILE) {
        __syncthreads();                                        // previous tile fully written out
        for (int s = t; s < span; s += FSEA_FIR_TPB) {
            const bool in = i0 + s < length + taps - 1;
            const int q = (s & 3) * quarter + (s >> 2);
            ti[q] = in ? si[i0 + s] : 0.0;
            tq[q] = in ? sq[i0 + s] : 0.0;
        }
        __syncthreads();
        double ai[FSEA_FIR_R], aq[FSEA_FIR_R];
#pragma unroll
        for (int r = 0; r < FSEA_FIR_R; r++) ai[r] = aq[r] = 0.0;
        // sample 4t + j sits at (j & 3) quarter + t + (j >> 2); output 4t + r takes it with tap k = j - r
        auto edge = [&](int j) {                                // first / last three samples: not every r has a tap
            const int q = (j & 3) * quarter + t + (j >> 2);
            const double xi = ti[q], xq = tq[q];
#pragma unroll
            for (int r = 0; r < FSEA_FIR_R; r++) {
                const int k = j - r;
                if (k >= 0 && k < taps) {
                    ai[r] = __dadd_rn(ai[r], __dmul_rn(c[k], xi));
                    aq[r] = __dadd_rn(aq[r], __dmul_rn(c[k], xq));
                }
            }
        };
        const int head = taps < 3 ? taps : 3;
        for (int j = 0; j < head; j++) edge(j);
        if (taps > 3) {
            double c1 = c[2], c2 = c[1], c3 = c[0];
#pragma unroll 4
            for (int j = 3; j < taps; j++) {
                const int q = (j & 3) * quarter + t + (j >> 2);
                const double xi = ti[q], xq = tq[q], c0 = c[j];
                ai[0] = __dadd_rn(ai[0], __dmul_rn(c0, xi)); aq[0] = __dadd_rn(aq[0], __dmul_rn(c0, xq));
                ai[1] = __dadd_rn(ai[1], __dmul_rn(c1, xi)); aq[1] = __dadd_rn(aq[1], __dmul_rn(c1, xq));
                ai[2] = __dadd_rn(ai[2], __dmul_rn(c2, xi)); aq[2] = __dadd_rn(aq[2], __dmul_rn(c2, xq));
                ai[3] = __dadd_rn(ai[3], __dmul_rn(c3, xi)); aq[3] = __dadd_rn(aq[3], __dmul_rn(c3, xq));
                c3 = c2; c2 = c1; c1 = c0;
            }
        }
        for (int j = (taps > 3 ? taps : head); j < taps + FSEA_FIR_R - 1; j++) edge(j);
        __syncthreads();                                        // everyone is done reading the lines
        double2* tile = reinterpret_cast<double2*>(ti);         // 1024 x 16 bytes <= 4 quarter doubles x 2 lines
#pragma unroll
        for (int r = 0; r < FSEA_FIR_R; r++) tile[FSEA_FIR_R * t + r] = make_double2(ai[r], aq[r]);
        __syncthreads();
        for (int e = t; e < TILE; e += FSEA_FIR_TPB)
            if (i0 + e < length) reinterpret_cast<double2*>(out)[i0 + e] = tile[e];
    }
}

static size_t fir_smem_bytes(int taps) { return sizeof(double) * ((size_t)((taps + 1) & ~1) + 8 * (size_t)fir_quarter(taps)); }

extern "C" int fsea_fir_iq_device(const double* d_coef, int taps, const double* d_si, const double* d_sq, long length,
                                  double* d_out, void* stream) {
    if (!d_coef || !d_si || !d_sq || !d_out || taps <= 0 || length < 0) return FSEA_ERR_INVALID;
    if (((uintptr_t)d_out & 15) != 0) return FSEA_ERR_INVALID;
    const size_t smem = fir_smem_bytes(taps);
    if (smem > 200 * 1024) return FSEA_ERR_UNSUPPORTED;                  // taps <= ~10000
    if (length == 0) return FSEA_OK;
    const int blocks = (int)std::min<long>((length + 1023) / 1024, 148 * 8);
    const int quarter = fir_quarter(taps);
#if defined(FSEA_EMU)
    fsea_emu::launch(dim3(blocks), dim3(FSEA_FIR_TPB), smem, 1, [=] { fsea_fir_iq_kernel(d_coef, taps, quarter, d_si, d_sq, length, d_out); });
#else
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fsea_fir_iq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fsea_fir_iq_kernel<<<blocks, FSEA_FIR_TPB, smem, (cudaStream_t)stream>>>(d_coef, taps, quarter, d_si, d_sq, length, d_out);
#endif
    CU(cudaGetLastError());
    return FSEA_OK;
}

extern "C" int fsea_fir_iq_host(int device, const double* coef, int taps, const double* si, const double* sq, long length,
                                double* out) {
    if (!coef || !si || !sq || !out || taps <= 0 || length < 0) return FSEA_ERR_INVALID;
    if (fsea_device_count() <= 0) return FSEA_ERR_NO_DEVICE;
    if (length == 0) return FSEA_OK;
    CU(cudaSetDevice(device));
    const size_t line = sizeof(double) * (size_t)(length + taps - 1), ob = sizeof(double) * 2 * (size_t)length;
    double* d = nullptr;                                           // coef | si | sq | out (out 16-byte aligned)
    const size_t o_coef = 0, o_si = ((size_t)taps * 8 + 15) & ~(size_t)15, o_sq = o_si + ((line + 15) & ~(size_t)15),
                 o_out = o_sq + ((line + 15) & ~(size_t)15);
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    int rc = scratch_get(device, 0, o_out + ob, (void**)&d);
    if (rc != FSEA_OK) return rc;
    char* b = (char*)d;
    cudaError_t e;
    if ((e = cudaMemcpy(b + o_coef, coef, (size_t)taps * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(b + o_si, si, line, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(b + o_sq, sq, line, cudaMemcpyHostToDevice)) != cudaSuccess) rc = cuda_fail(e, "H2D");
    if (rc == FSEA_OK) rc = fsea_fir_iq_device((const double*)(b + o_coef), taps, (const double*)(b + o_si), (const double*)(b + o_sq),
                                               length, (double*)(b + o_out), nullptr);
    if (rc == FSEA_OK && (e = cudaMemcpy(out, b + o_out, ob, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "D2H");
    return rc;
}

// fsea_chain.cu -- the operators in FRONT of the path fused with it.
//
// lua/fft-shifted.lua:42-54 and lua/iq-tex-filtered.lua:43-46 put nrf_freq_shifter (src/nrf.c:843-866) or
// nrf_iq_filter (src/nrf.c:738-768) between the device and nrf_fft.  Through the reference's API every hop is a
// fresh float64 nut_buffer of the whole 131072-sample block on the host -- 2 MiB out, 2 MiB in, and 131072 x taps
// multiply-adds of which nrf_fft then reads the first N samples (src/nrf.c:598-615).  A chain object does the same
// arithmetic for one device block in one call, on the device, and only for what the transform consumes:
//   u8 block -> [rotator: v * (c + i s) + 0.5]  ->  [FIR: N outputs from N + taps - 1 line samples]  ->  N-point FFT (fp64),
//   |X|, DC patch  ->  one row of N doubles back on the host.
// State carried from block to block, as the reference carries it: the rotator's phasor and the last taps - 1 samples
// of the previous block in front of the filter's line.  The FIR sums are the reference's (ascending taps, every
// product and sum rounded: bit-identical, fsea_fir_iq_kernel); so are the rotator's products.  The rotator's
// PHASOR is a 131072-step recurrence with a rounding at every step (src/nrf.c:858-863): serial by nature, it stays
// on the host (0.4 ms per block) and only the phasors the chain needs (N + taps of them) cross to the device.
#include <cuda_runtime.h>

#include <math.h>
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

#if defined(FSEA_EMU)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
#endif

struct fsea_chain {
    fsea_chain_config cfg;
    int n = 0, L = 0, taps = 0, carry = 0;       // carry = taps - 1 (0 without a filter)
    fsea_plan* plan = nullptr;                   // N-point, float64 samples, |X| + DC patch
    cudaStream_t stream = nullptr;
    uint8_t* d_block = nullptr;                  // 2 L bytes
    double* d_coef = nullptr;
    double* d_line_i = nullptr;                  // carry + N doubles
    double* d_line_q = nullptr;
    double* d_carry_i[2] = {nullptr, nullptr};   // the previous block's tail (double-buffered)
    double* d_carry_q[2] = {nullptr, nullptr};
    int carry_cur = 0;
    double* d_fft_in = nullptr;                  // N interleaved (I, Q) doubles
    double* d_row = nullptr;                     // N doubles
    double2* d_phasor = nullptr;                 // head: N phasors, tail: carry phasors
    double2* h_phasor = nullptr;                 // pinned
    uint8_t* h_block = nullptr;                  // pinned staging of the caller's block
    double* h_row = nullptr;                     // pinned
    double rot_c = 1.0, rot_s = 0.0;             // rotator state (src/nrf.c:820-821)
    long blocks = 0, launches = 0;
};

// value of block sample j: u8 / 256 (nut_buffer_get_f64, src/nut.c:129-135), optionally through the rotator:
// (vi c - vq s) + 0.5, (vi s + vq c) + 0.5, every operation rounded (src/nrf.c:855-857)
__device__ __forceinline__ void fsea_chain_value(const uint8_t* __restrict__ block, long j, const double2* ph, double& oi, double& oq) {
    const double vi = (double)block[2 * j] / 256.0, vq = (double)block[2 * j + 1] / 256.0;
    if (ph) {
        const double c = ph->x, s = ph->y;
        oi = __dadd_rn(__dadd_rn(__dmul_rn(vi, c), -__dmul_rn(vq, s)), 0.5);
        oq = __dadd_rn(__dadd_rn(__dmul_rn(vi, s), __dmul_rn(vq, c)), 0.5);
    } else {
        oi = vi; oq = vq;
    }
}

// line = [carry from the previous block | first N samples of this block]; new carry = the last `carry` samples of this block
__global__ void fsea_chain_prepare_kernel(const uint8_t* __restrict__ block, int n, long L, int carry, const double2* __restrict__ phasor,
                                          const double* __restrict__ old_i, const double* __restrict__ old_q,
                                          double* __restrict__ line_i, double* __restrict__ line_q,
                                          double* __restrict__ new_i, double* __restrict__ new_q, double* __restrict__ fft_in) {
    const int total = carry + n + carry;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        if (q < carry) {
            line_i[q] = old_i[q]; line_q[q] = old_q[q];
        } else if (q < carry + n) {
            const int j = q - carry;
            double oi, oq;
            fsea_chain_value(block, j, phasor ? phasor + j : nullptr, oi, oq);
            if (line_i) { line_i[q] = oi; line_q[q] = oq; }
            else { fft_in[2 * j] = oi; fft_in[2 * j + 1] = oq; }           // no filter: straight into the transform
        } else {
            const int t = q - carry - n;                                    // tail sample L - carry + t
            double oi, oq;
            fsea_chain_value(block, L - carry + t, phasor ? phasor + n + t : nullptr, oi, oq);
            new_i[t] = oi; new_q[t] = oq;
        }
    }
}

extern "C" int fsea_chain_destroy(fsea_chain* c) {
    if (!c) return FSEA_OK;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    fsea_plan_destroy(c->plan);
    void* dev[] = {c->d_block, c->d_coef, c->d_line_i, c->d_line_q, c->d_carry_i[0], c->d_carry_i[1], c->d_carry_q[0], c->d_carry_q[1],
                   c->d_fft_in, c->d_row, c->d_phasor};
    for (void* p : dev) if (p) cudaFree(p);
    if (c->h_phasor) cudaFreeHost(c->h_phasor);
    if (c->h_block) cudaFreeHost(c->h_block);
    if (c->h_row) cudaFreeHost(c->h_row);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return FSEA_OK;
}

static int chain_setup(fsea_chain* c) {
    const fsea_chain_config& g = c->cfg;
    fsea_config pc;
    fsea_config_default(&pc, g.fft_size);
    pc.in_format = FSEA_IN_F64C;
    pc.out_mode = FSEA_OUT_MAG;
    pc.flags = FSEA_FLAG_DC_PATCH;
    pc.device = g.device;
    int rc = fsea_plan_create(&pc, &c->plan);
    if (rc != FSEA_OK) return rc;
    CU(cudaSetDevice(g.device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t n = (size_t)c->n, carry = (size_t)c->carry;
    CU(cudaMalloc((void**)&c->d_block, 2 * (size_t)c->L));
    CU(cudaMalloc((void**)&c->d_fft_in, 2 * n * sizeof(double)));
    CU(cudaMalloc((void**)&c->d_row, n * sizeof(double)));
    CU(cudaMallocHost((void**)&c->h_block, 2 * (size_t)c->L));
    CU(cudaMallocHost((void**)&c->h_row, n * sizeof(double)));
    if (g.filter) {
        CU(cudaMalloc((void**)&c->d_coef, (size_t)c->taps * sizeof(double)));
        CU(cudaMemcpy(c->d_coef, g.coefficients, (size_t)c->taps * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMalloc((void**)&c->d_line_i, (carry + n) * sizeof(double)));
        CU(cudaMalloc((void**)&c->d_line_q, (carry + n) * sizeof(double)));
        for (int k = 0; k < 2; k++) {
            CU(cudaMalloc((void**)&c->d_carry_i[k], (carry + 1) * sizeof(double)));
            CU(cudaMalloc((void**)&c->d_carry_q[k], (carry + 1) * sizeof(double)));
            CU(cudaMemset(c->d_carry_i[k], 0, (carry + 1) * sizeof(double)));      // the reference's line starts out zero (calloc, src/nrf.c:687)
            CU(cudaMemset(c->d_carry_q[k], 0, (carry + 1) * sizeof(double)));
        }
    }
    if (g.shifter) {
        CU(cudaMalloc((void**)&c->d_phasor, (n + carry + 1) * sizeof(double2)));
        CU(cudaMallocHost((void**)&c->h_phasor, (n + carry + 1) * sizeof(double2)));
    }
    return FSEA_OK;
}

extern "C" int fsea_chain_create(const fsea_chain_config* cfg, fsea_chain** out) {
    if (!cfg || !out) return FSEA_ERR_INVALID;
    *out = nullptr;
    if (cfg->block_samples < cfg->fft_size || cfg->fft_size < 16) return FSEA_ERR_INVALID;
    if (cfg->filter && (!cfg->coefficients || cfg->taps <= 0 || cfg->taps - 1 > cfg->block_samples - cfg->fft_size)) return FSEA_ERR_INVALID;
    if (cfg->shifter && cfg->shifter_sample_rate <= 0) return FSEA_ERR_INVALID;
    fsea_chain* c = new (std::nothrow) fsea_chain();
    if (!c) return FSEA_ERR_NOMEM;
    c->cfg = *cfg;
    c->n = cfg->fft_size;
    c->L = cfg->block_samples;
    c->taps = cfg->filter ? cfg->taps : 0;
    c->carry = cfg->filter ? cfg->taps - 1 : 0;
    c->cfg.coefficients = cfg->coefficients;     // (setup copies them to the device; the caller's array is not kept)
    const int rc = chain_setup(c);
    c->cfg.coefficients = nullptr;
    if (rc != FSEA_OK) { fsea_chain_destroy(c); return rc; }
    *out = c;
    return FSEA_OK;
}

extern "C" int fsea_chain_process_u8(fsea_chain* c, const uint8_t* h_block, double* h_row) {
    if (!c || !h_block || !h_row) return FSEA_ERR_INVALID;
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n, carry = c->carry;
    const long L = c->L;
    memcpy(c->h_block, h_block, 2 * (size_t)L);
    CU(cudaMemcpyAsync(c->d_block, c->h_block, 2 * (size_t)L, cudaMemcpyHostToDevice, c->stream));
    if (c->cfg.shifter) {
        // the rotator's recurrence over the whole block, exactly as src/nrf.c:843-866 steps it; the phasors of the
        // samples the chain consumes are kept
        const double step = 2.0 * M_PI * c->cfg.freq_offset / (double)c->cfg.shifter_sample_rate;
        const double dc = cos(step), ds = sin(step);
        double cc = c->rot_c, ss = c->rot_s;
        for (long j = 0; j < L; j++) {
            if (j < n) c->h_phasor[j] = make_double2(cc, ss);
            else if (j >= L - carry) c->h_phasor[n + (j - (L - carry))] = make_double2(cc, ss);
            const double ns = cc * ds + ss * dc, nc = cc * dc - ss * ds;
            ss = ns; cc = nc;
        }
        c->rot_c = cc; c->rot_s = ss;
        CU(cudaMemcpyAsync(c->d_phasor, c->h_phasor, (size_t)(n + carry) * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    }
    const int nxt = c->carry_cur ^ 1;
    const int total = carry + n + carry;
    FSEA_LAUNCH(fsea_chain_prepare_kernel, (total + 255) / 256, 256, c->stream, (const uint8_t*)c->d_block, n, L, carry,
                (const double2*)(c->cfg.shifter ? c->d_phasor : nullptr),
                (const double*)c->d_carry_i[c->carry_cur], (const double*)c->d_carry_q[c->carry_cur],
                c->cfg.filter ? c->d_line_i : nullptr, c->cfg.filter ? c->d_line_q : nullptr,
                c->d_carry_i[nxt], c->d_carry_q[nxt], c->d_fft_in);
    CU(cudaGetLastError());
    c->launches++;
    c->carry_cur = nxt;
    int rc;
    if (c->cfg.filter) {
        rc = fsea_fir_iq_device(c->d_coef, c->taps, c->d_line_i, c->d_line_q, n, c->d_fft_in, c->stream);
        if (rc != FSEA_OK) return rc;
        c->launches++;
    }
    const long l0 = fsea_plan_launch_count(c->plan);
    rc = fsea_launch_internal(c->plan, c->d_fft_in, c->d_row, 1, c->stream, 0);
    if (rc != FSEA_OK) return rc;
    c->launches += fsea_plan_launch_count(c->plan) - l0;
    CU(cudaMemcpyAsync(c->h_row, c->d_row, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(h_row, c->h_row, (size_t)n * sizeof(double));
    c->blocks++;
    return FSEA_OK;
}

extern "C" long fsea_chain_launch_count(const fsea_chain* c) { return c ? c->launches : 0; }

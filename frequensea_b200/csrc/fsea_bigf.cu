// fsea_bigf.cu -- float-sample frames longer than 8192 points (FSEA_IN_F32C / FSEA_IN_F64C, N = 16384 .. 65536);
// also what FSEA_ARITH_F64 runs above 8192 points once fsea_engine.cu has widened the 8-bit samples.
//
// The reference's float64 branch (src/nrf.c:607-609: raw doubles from a frequency shifter or an IQ filter)
// has no length limit below its 131072-sample block.  The fast kernels of this engine take 8-bit IQ above
// 8192 points; a float frame that long is one spectrum row per device block for a script such as
// lua/fft-shifted.lua, not a throughput case -- it gets a plain, exact path: a radix-16 Stockham autosort
// transform in fp64, one launch per stage through two work buffers in global memory (L2-resident for the
// one-frame use), samples widened on load, |X| / log / pixel and the fftshift rotation in a last launch.
#include <cuda_runtime.h>

#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/fsea.h"
#include "fsea_internal.h"
#include "fsea_kparams.h"

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fsea_fail_cuda_internal(e__, #call); \
    } while (0)

struct fsea_bigf {
    int n = 0, in_format = 0, out_mode = 0, flags = 0;
    long chunk = 1;                     // frames per pass through the work buffers
    double2* d_tw = nullptr;            // exp(-2 pi i j / N), j < N
    double2* d_work[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;         // the work buffers are shared: launches on different streams are ordered by this event
    cudaStream_t last = nullptr;
    bool used = false;
};

// x[n] = w[n] * sample[n]  (src/nrf.c:607-612; the (-1)^n is the output rotation of the last launch)
__global__ void fsea_bigf_load_kernel(const void* __restrict__ in, int f32, long seq_stride, int n, const float* __restrict__ win,
                                      double2* __restrict__ out, int gx) {
    const long f = blockIdx.x / gx;                            // (one-dimensional grid: gx blocks per frame)
    for (int i = (blockIdx.x % gx) * blockDim.x + threadIdx.x; i < n; i += gx * blockDim.x) {
        double re, im;
        if (f32) { const float2 v = ((const float2*)in)[f * seq_stride + i]; re = v.x; im = v.y; }
        else { const double2 v = ((const double2*)in)[f * seq_stride + i]; re = v.x; im = v.y; }
        if (win) { const double w = (double)win[i]; re *= w; im *= w; }
        out[f * n + i] = make_double2(re, im);
    }
}

// One radix-R Stockham autosort stage (R = 2, 4, 8, 16), ns = product of the radices already done: item j < N/R takes
// in[j + t N/R], t < R, turns input t by W_{ns R}^{t k} (k = j mod ns), transforms the R values in registers and writes
// out[(j - k) R + k + m ns], m < R.  tw[q] = exp(-2 pi i q / N), q < N.
__constant__ double2 fsea_bigf_w16[8] = {
    {1.0, -0.0},
    {0.92387953251128675613, -0.38268343236508977173},
    {0.70710678118654752440, -0.70710678118654752440},
    {0.38268343236508977173, -0.92387953251128675613},
    {0.0, -1.0},
    {-0.38268343236508977173, -0.92387953251128675613},
    {-0.70710678118654752440, -0.70710678118654752440},
    {-0.92387953251128675613, -0.38268343236508977173}};

template <int R> __device__ __forceinline__ void fsea_bigf_dft(double2* v) {       // in place, natural order in and out
    constexpr int LG = (R == 2) ? 1 : (R == 4) ? 2 : (R == 8) ? 3 : 4;
#pragma unroll
    for (int i = 0; i < R; i++) {                                                    // bit reversal
        int r = 0;
#pragma unroll
        for (int b = 0; b < LG; b++) r |= ((i >> b) & 1) << (LG - 1 - b);
        if (r > i) { const double2 x = v[i]; v[i] = v[r]; v[r] = x; }
    }
#pragma unroll
    for (int len = 2; len <= R; len <<= 1) {
#pragma unroll
        for (int base = 0; base < R; base += len) {
#pragma unroll
            for (int q = 0; q < len / 2; q++) {
                const double2 w = fsea_bigf_w16[q * (16 / len)];                    // W_len^q
                const double2 a = v[base + q], b0 = v[base + q + len / 2];
                const double2 b = (q == 0) ? b0 : make_double2(b0.x * w.x - b0.y * w.y, b0.x * w.y + b0.y * w.x);
                v[base + q] = make_double2(a.x + b.x, a.y + b.y);
                v[base + q + len / 2] = make_double2(a.x - b.x, a.y - b.y);
            }
        }
    }
}

template <int R>
__global__ void fsea_bigf_stage_kernel(const double2* __restrict__ in, double2* __restrict__ out, const double2* __restrict__ tw,
                                       int n, int ns, int gx) {
    const long f = blockIdx.x / gx;
    const int items = n / R;
    const int tstep = n / (ns * R);                                     // table stride of W_{ns R}
    for (int j = (blockIdx.x % gx) * blockDim.x + threadIdx.x; j < items; j += gx * blockDim.x) {
        const int k = j & (ns - 1);
        double2 v[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            const double2 x = in[f * n + j + (size_t)t * items];
            if (t == 0 || ns == 1) { v[t] = x; continue; }
            const double2 w = tw[(size_t)t * k * tstep];
            v[t] = make_double2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
        }
        fsea_bigf_dft<R>(v);
        const size_t o = (size_t)f * n + (size_t)(j - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; m++) out[o + (size_t)m * ns] = v[m];
    }
}

// bin k -> position (k + N/2) mod N; magnitude / log / pixel; out[N/2] = out[N/2 - 1] with the DC patch
__global__ void fsea_bigf_finish_kernel(const double2* __restrict__ x, int n, fsea_kparams kp, int out_f64, long frame0, int gx) {
    const long f = blockIdx.x / gx;
    uint8_t* row = (uint8_t*)kp.out + fsea_out_row(kp, frame0 + f);
    const bool patch = (kp.flags & FSEA_FLAG_DC_PATCH) != 0;
    for (int p = (blockIdx.x % gx) * blockDim.x + threadIdx.x; p < n; p += gx * blockDim.x) {
        const int src_pos = (patch && p == n / 2) ? n / 2 - 1 : p;     // src/nrf.c:621-622
        const int k = (src_pos + n / 2) & (n - 1);
        const double2 v = x[f * n + k];
        const double pw = v.x * v.x + v.y * v.y;
        double val;
        if (kp.out_mode == FSEA_OUT_MAG) val = sqrt(pw);
        else val = kp.lg.dmul * log2(pw + kp.lg.deps) + kp.lg.dadd;
        if (kp.out_mode == FSEA_OUT_PIX) {
            int q = (int)(val * (double)kp.pix_scale);
            q = q < 0 ? 0 : (q > 255 ? 255 : q);
            row[p] = (uint8_t)q;
        } else if (out_f64) {
            ((double*)row)[p] = val;
        } else {
            ((float*)row)[p] = (float)val;
        }
    }
}

int fsea_bigf_destroy(fsea_bigf* b) {
    if (!b) return FSEA_OK;
    if (b->d_tw) cudaFree(b->d_tw);
    if (b->d_work[0]) cudaFree(b->d_work[0]);
    if (b->d_work[1]) cudaFree(b->d_work[1]);
    if (b->done) cudaEventDestroy(b->done);
    delete b;
    return FSEA_OK;
}

int fsea_bigf_create(int n, int in_format, fsea_bigf** out) {
    fsea_bigf* b = new fsea_bigf();
    b->n = n;
    b->in_format = in_format;
    b->chunk = std::max<long>(1, (1L << 21) / n);
    std::vector<double2> tw((size_t)n);
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (int j = 0; j < n; j++) {
        const long double a = -PI2 * (long double)j / (long double)n;
        tw[j].x = (double)cosl(a); tw[j].y = (double)sinl(a);
    }
    cudaError_t e;
    if ((e = cudaMalloc(&b->d_tw, tw.size() * sizeof(double2))) != cudaSuccess ||
        (e = cudaMemcpy(b->d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&b->d_work[0], (size_t)b->chunk * n * sizeof(double2))) != cudaSuccess ||
        (e = cudaMalloc(&b->d_work[1], (size_t)b->chunk * n * sizeof(double2))) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming)) != cudaSuccess) {
        fsea_bigf_destroy(b);
        return fsea_fail_cuda_internal(e, "float-frame work buffers");
    }
    *out = b;
    return FSEA_OK;
}

// kp: the launch's argument block as fsea_engine.cu fills it (in / out / strides / log constants / flags)
int fsea_bigf_launch(fsea_bigf* b, const fsea_kparams& kp, long n_frames, cudaStream_t stream, long* launches) {
    const int n = b->n;
    const int f32 = b->in_format == FSEA_IN_F32C;
    const int gx = std::min(n / 2 / 256, 32);
    if (b->used && b->last != stream) CU(cudaStreamWaitEvent(stream, b->done, 0));
    for (long f0 = 0; f0 < n_frames; f0 += b->chunk) {
        const long nf = std::min(b->chunk, n_frames - f0);
        const void* in = f32 ? (const void*)((const float2*)kp.in + f0 * kp.seq_stride) : (const void*)((const double2*)kp.in + f0 * kp.seq_stride);
        FSEA_LAUNCH(fsea_bigf_load_kernel, (unsigned)(gx * nf), 256, stream, in, f32, kp.seq_stride, n, kp.window, b->d_work[0], gx);
        int cur = 0;
        for (int ns = 1; ns < n;) {                                      // radix 16 while it divides what is left, then the rest
            const int left = n / ns;
            const int r = left >= 16 ? 16 : left;
            const double2* src = b->d_work[cur];
            double2* dst = b->d_work[cur ^ 1];
            const int gs = std::max(1, std::min(n / r / 128, 32));     // blocks per frame
            const unsigned grid = (unsigned)(gs * nf);
            if (r == 16) FSEA_LAUNCH(fsea_bigf_stage_kernel<16>, grid, 128, stream, src, dst, (const double2*)b->d_tw, n, ns, gs);
            else if (r == 8) FSEA_LAUNCH(fsea_bigf_stage_kernel<8>, grid, 128, stream, src, dst, (const double2*)b->d_tw, n, ns, gs);
            else if (r == 4) FSEA_LAUNCH(fsea_bigf_stage_kernel<4>, grid, 128, stream, src, dst, (const double2*)b->d_tw, n, ns, gs);
            else FSEA_LAUNCH(fsea_bigf_stage_kernel<2>, grid, 128, stream, src, dst, (const double2*)b->d_tw, n, ns, gs);
            cur ^= 1;
            ns *= r;
            (*launches)++;
        }
        FSEA_LAUNCH(fsea_bigf_finish_kernel, (unsigned)(gx * nf), 256, stream, (const double2*)b->d_work[cur], n, kp, f32 ? 0 : 1, f0, gx);
        (*launches) += 2;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(b->done, stream));
    b->last = stream;
    b->used = true;
    return FSEA_OK;
}

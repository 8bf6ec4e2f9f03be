// fma_pipe.cu -- microbenchmark: FP32 issue rate of scalar FFMA vs packed FFMA2
// on sm_100a (B200).  Decides which arithmetic variant of the FFT kernels can
// reach the HBM roofline (DESIGN.md "FP32 co-limit").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_pipe fma_pipe.cu && ./fma_pipe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: FFMA2 stream with an ALU (PRMT-like LOP3) instruction interleaved 1:2
template <int ILP>
__global__ void k_mix(float* out, int iters, float a, float b, unsigned m) {
    float2 x[ILP];
    unsigned y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i); y[i] = threadIdx.x + i; }
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) { x[i] = __ffma2_rn(x[i], aa, bb); y[i] = __byte_perm(y[i], m, 0x7250) ^ m; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 4096;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const int ILP = 8;
    const double inst = (double)blocks * threads * iters * 8 * ILP;      // per-thread instructions
    printf("device %s, %d SMs, max clock %.0f MHz\n", p.name, sms, clk_khz / 1e3);
    float t1 = time_ms([&] { k_ffma<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_ffma2<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float t3 = time_ms([&] { k_mix<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f, 0x01020304u); });
    printf("FFMA : %.3f ms  %.2f Tinst/s  %.2f TFMA/s  (%.1f FMA lanes/clk/SM at max clock)\n", t1, inst / t1 / 1e9, inst / t1 / 1e9,
           inst / (t1 * 1e-3) / sms / (clk_khz * 1e3));
    printf("FFMA2: %.3f ms  %.2f Tinst/s  %.2f TFMA/s  (%.1f FMA lanes/clk/SM at max clock)\n", t2, inst / t2 / 1e9, 2 * inst / t2 / 1e9,
           2 * inst / (t2 * 1e-3) / sms / (clk_khz * 1e3));
    printf("FFMA2+2ALU: %.3f ms  %.2f T FFMA2/s (%.1f FMA lanes/clk/SM) while co-issuing 2 ALU ops per FFMA2\n", t3, inst / t3 / 1e9,
           2 * inst / (t3 * 1e-3) / sms / (clk_khz * 1e3));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

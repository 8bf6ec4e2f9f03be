// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE, not product code.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the product's CUDA sources
// (frequensea_b200/csrc/*.cu, -DFSEA_EMU) into tests/emu/libfsea_emu.so: the SAME kernel
// source runs on the CPU, every CUDA thread a cooperative fiber (ucontext), so the index
// arithmetic, barrier placement, shared-memory layout and cp.async discipline of a kernel
// can be checked against the oracle in the GPU-less authoring container.  It is loaded only
// by tests/test_emu_kernels.py (through an explicit path); the product library
// (libfsea.so, nvcc, sm_100a) never sees this header and has no CPU path.
//
// Semantics worth knowing:
//   * blocks of a launch run as fibers of ONE OS thread, round-robin, switching only at
//     barriers / collectives / spin points: deterministic, no data races;
//   * cp.async copies are DEFERRED until the issuing thread executes cp.async.wait_group 0,
//     so a kernel that reads staged bytes without waiting reads stale memory and fails;
//   * the "device" has 2 SMs; occupancy queries answer from shared-memory / thread limits.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define FSEA_EMU_ACTIVE 1

// ---- vector types ---------------------------------------------------------------------
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uchar2 { unsigned char x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline unsigned __vmaxu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) { const unsigned x = (a >> (8 * i)) & 255u, y = (b >> (8 * i)) & 255u; r |= (x > y ? x : y) << (8 * i); }
    return r;
}
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline uchar2 make_uchar2(unsigned char a, unsigned char b) { uchar2 r; r.x = a; r.y = b; return r; }

// ---- function / variable qualifiers -----------------------------------------------------
#define __global__
#define __device__
#define __constant__ static const
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

// ---- the fiber machine (tests/emu/emu.cpp) ------------------------------------------------
namespace fsea_emu {
struct Thread {
    uint3 tid, bid, bdim, gdim;
    int cluster_size;
    void* block;          // Block*
};
extern Thread* cur;                                   // the running CUDA thread
unsigned char* dyn_smem();                            // this block's dynamic shared memory
void* static_smem(int key, size_t bytes);             // this block's named static shared variable
void sync_block();                                    // __syncthreads
void sync_named(int id, int count);                   // bar.sync id, count
void arrive_named(int id, int count);                 // bar.arrive id, count
void sync_warp();                                     // __syncwarp (full mask)
unsigned shfl_idx(unsigned v, int src_lane);          // shfl.sync.idx, full mask
double shfl_down_f64(double v, int delta);
void cluster_arrive();
unsigned char* map_cluster(unsigned char* local_smem, unsigned cta_rank);   // mapa: the same offset in another block of the cluster
void mbar_init(void* bar, unsigned count);
void mbar_arrive(void* bar, unsigned expect_tx_bytes);
void mbar_complete_tx(void* bar, unsigned bytes);
void mbar_wait(void* bar, unsigned parity);
void cluster_wait();
void cp_async(void* smem_dst, const void* gsrc, int bytes);
void cp_async_wait_all();
void yield();
void launch(dim3 grid, dim3 block, size_t smem, int cluster_size, const std::function<void()>& body);
}  // namespace fsea_emu

struct fsea_tmap;
namespace fsea_emu { void tma_load_3d(unsigned char* dst, const fsea_tmap* tm, int c0, int c1, int c2, void* bar); }

// threadIdx.x etc.: objects whose .x reads the running fiber's coordinates when converted
namespace fsea_emu {
struct TidX { operator unsigned() const { return cur->tid.x; } };
struct BidX { operator unsigned() const { return cur->bid.x; } };
struct BdimX { operator unsigned() const { return cur->bdim.x; } };
struct GdimX { operator unsigned() const { return cur->gdim.x; } };
}  // namespace fsea_emu
static const struct { fsea_emu::TidX x; } threadIdx = {};
static const struct { fsea_emu::BidX x; } blockIdx = {};
static const struct { fsea_emu::BdimX x; } blockDim = {};
static const struct { fsea_emu::GdimX x; } gridDim = {};

static inline void __syncthreads() { fsea_emu::sync_block(); }
static inline void __syncwarp() { fsea_emu::sync_warp(); }
static inline void __threadfence() {}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const uint64_t w = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((w >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#define __log2f(x) log2f(x)
static inline void sincospif(float x, float* s, float* c) { *s = (float)sin(3.14159265358979323846 * (double)x); *c = (float)cos(3.14159265358979323846 * (double)x); }
static inline int __float2int_rz(float x) { return (int)x; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline double __shfl_down_sync(unsigned, double v, int d) { return fsea_emu::shfl_down_f64(v, d); }

// ---- runtime API subset -------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { int type; };
enum { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttribute {
    int id;
    union { struct { unsigned x, y, z; } clusterDim; char pad[64]; } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};

static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorInvalidValue; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t y = 0; y < h; y++) memset((char*)d + y * pitch, v, w);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
static inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
// resident blocks per SM: what 227 KB of shared memory and 2048 threads allow, at most 3
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int threads, size_t smem) {
    int a = smem ? (int)((227 * 1024) / (smem + 1024)) : 3, b = 2048 / threads;
    a = a < b ? a : b;
    *n = a < 1 ? 0 : (a > 3 ? 3 : a);
    return cudaSuccess;
}
static inline cudaError_t cudaOccupancyMaxActiveClusters(int* n, const void*, const cudaLaunchConfig_t*) { *n = 3; return cudaSuccess; }

// every kernel of the engine's table takes one by-value argument block
struct fsea_kparams;
cudaError_t fsea_emu_launch_kparams(const void* fn, dim3 grid, dim3 block, void** args, size_t smem, int cluster);
static inline cudaError_t cudaLaunchKernel(const void* fn, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t) {
    return fsea_emu_launch_kparams(fn, grid, block, args, smem, 1);
}
static inline cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t* lc, const void* fn, void** args) {
    int cs = 1;
    for (unsigned i = 0; i < lc->numAttrs; i++)
        if (lc->attrs[i].id == cudaLaunchAttributeClusterDimension) cs = (int)lc->attrs[i].val.clusterDim.x;
    return fsea_emu_launch_kparams(fn, lc->gridDim, lc->blockDim, args, lc->dynamicSmemBytes, cs);       // (args[1], if the kernel takes one: a tensor map)
}

// fsea_ptx.cuh -- every inline-PTX instruction the kernels use, one wrapper each.
//
// The product build (nvcc, sm_100a) gets the PTX.  A -DFSEA_EMU build (g++, tests/emu/: the
// kernel SOURCE executed on the CPU by a fiber machine, test infrastructure only) gets the
// same operations from tests/emu/cuda_runtime.h, so that kernel logic can be checked against
// the oracle where there is no GPU.  Nothing here is a CPU fallback of the product.
#pragma once
#include <stdint.h>

#if defined(FSEA_EMU)
#include <cuda_runtime.h>          // tests/emu/cuda_runtime.h
#define FSEA_DYN_SMEM(name) unsigned char* name = fsea_emu::dyn_smem()
#define FSEA_STATIC_SMEM(type, name, count) type* name = (type*)fsea_emu::static_smem(__LINE__, sizeof(type) * (count))
#else
#define FSEA_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define FSEA_STATIC_SMEM(type, name, count) __shared__ type name[count]
#endif

__device__ __forceinline__ float fsea_sqrt_approx(float x) {
#if defined(FSEA_EMU)
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

__device__ __forceinline__ void fsea_prefetch_l2(const void* p) {
#if defined(FSEA_EMU)
    (void)p;
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// ---- streaming (evict-first) stores: spectra are written once and never read again by the
// kernel, the four-step intermediate next to them in L2 is -- keep the spectra from pushing it out
#ifndef FSEA_COMB_STREAM
#define FSEA_COMB_STREAM 1
#endif
__device__ __forceinline__ void fsea_st_stream(float* p, float v) {
#if defined(FSEA_EMU) || !FSEA_COMB_STREAM
    *p = v;
#else
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#endif
}
__device__ __forceinline__ void fsea_st_stream(float2* p, float2 v) {
#if defined(FSEA_EMU) || !FSEA_COMB_STREAM
    *p = v;
#else
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}

// ---- asynchronous global -> shared copies (LDGSTS) --------------------------------------
__device__ __forceinline__ void fsea_cp_async16(void* smem_dst, const void* gsrc) {        // 16 bytes, L2 only
#if defined(FSEA_EMU)
    fsea_emu::cp_async(smem_dst, gsrc, 16);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void fsea_cp_async8(void* smem_dst, const void* gsrc) {         // 8 bytes
#if defined(FSEA_EMU)
    fsea_emu::cp_async(smem_dst, gsrc, 8);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void fsea_cp_async4(void* smem_dst, const void* gsrc) {         // 4 bytes
#if defined(FSEA_EMU)
    fsea_emu::cp_async(smem_dst, gsrc, 4);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void fsea_cp_async_commit() {
#if !defined(FSEA_EMU)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void fsea_cp_async_wait_all() {
#if defined(FSEA_EMU)
    fsea_emu::cp_async_wait_all();
#else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- barriers ---------------------------------------------------------------------------
template <int COUNT> __device__ __forceinline__ void fsea_bar_sync(int id) {                // named barrier 1..15
#if defined(FSEA_EMU)
    fsea_emu::sync_named(id, COUNT);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
#endif
}
// producer / consumer form of the named barriers (warp-specialised kernels): the barrier completes
// when `count` threads have arrived, by either instruction; only bar.sync waits
__device__ __forceinline__ void fsea_named_sync(int id, int count) {
#if defined(FSEA_EMU)
    fsea_emu::sync_named(id, count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void fsea_named_arrive(int id, int count) {
#if defined(FSEA_EMU)
    fsea_emu::arrive_named(id, count);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void fsea_cluster_arrive_release() {
#if defined(FSEA_EMU)
    fsea_emu::cluster_arrive();
#else
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void fsea_cluster_arrive_relaxed() {
#if defined(FSEA_EMU)
    fsea_emu::cluster_arrive();
#else
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void fsea_cluster_wait_acquire() {
#if defined(FSEA_EMU)
    fsea_emu::cluster_wait();
#else
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ unsigned fsea_cluster_ctarank() {
#if defined(FSEA_EMU)
    return blockIdx.x % (unsigned)fsea_emu::cur->cluster_size;
#else
    unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
#endif
}
__device__ __forceinline__ unsigned fsea_cluster_id() {
#if defined(FSEA_EMU)
    return blockIdx.x / (unsigned)fsea_emu::cur->cluster_size;
#else
    unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r;
#endif
}
__device__ __forceinline__ unsigned fsea_cluster_count() {
#if defined(FSEA_EMU)
    return gridDim.x / (unsigned)fsea_emu::cur->cluster_size;
#else
    unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r;
#endif
}

// ---- atomics / warp exchange ----------------------------------------------------------------
// atom.inc as opaque PTX: nvcc turns atomicAdd() in a one-lane branch into its
// warp-aggregated form (elect + atom + SHFL of the result), and that compiler-made SHFL
// waits for the atomic's round trip right where it is issued.
__device__ __forceinline__ unsigned fsea_atom_inc(unsigned* counter) {
#if defined(FSEA_EMU)
    const unsigned old = *counter;
    *counter = (old >= 0xfffffffeu) ? 0u : old + 1u;
    return old;
#else
    unsigned old;
    asm volatile("atom.global.inc.u32 %0, [%1], 0xfffffffe;" : "=r"(old) : "l"(counter) : "memory");
    return old;
#endif
}
__device__ __forceinline__ unsigned fsea_shfl_idx(unsigned v, int src_lane) {               // full warp
#if defined(FSEA_EMU)
    return fsea_emu::shfl_idx(v, src_lane);
#else
    unsigned r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(src_lane) : "memory");
    return r;
#endif
}

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

// hides a value's provenance from the optimiser (keeps loop-invariant arithmetic INSIDE a loop whose registers are all spoken for)
__device__ __forceinline__ void fsea_opaque(float& x) {
#if !defined(FSEA_EMU)
    asm volatile("" : "+f"(x));
#endif
}
__device__ __forceinline__ void fsea_prefetch_l1(const void* p) {
#if defined(FSEA_EMU)
    (void)p;
#else
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
// read-only table loads that do not take L1 space (the comb's four-step twiddle rows: 32 KB per frame that would push
// the raw IQ lines the transform warps are about to gather out of the ~50 KB of L1 left beside 206 KB of shared memory)
__device__ __forceinline__ float2 fsea_ldg_stream(const float2* p) {
#if defined(FSEA_EMU)
    return *p;
#else
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
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

__device__ __forceinline__ void fsea_st_stream(float4* p, float4 v) {
#if defined(FSEA_EMU) || !FSEA_COMB_STREAM
    *p = v;
#else
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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


// ---- distributed shared memory, mbarriers, TMA (the cluster kernel of fsea_wsc.cuh) -------------------
// fsea_cptr: an address in the shared memory of a block of this cluster (shared::cluster window).
#if defined(FSEA_EMU)
typedef unsigned char* fsea_cptr;
struct fsea_tmap { const unsigned char* base; unsigned long long dim[3], stride[3]; unsigned box[3]; };   // stand-in for CUtensorMap
#else
#include <cuda.h>
typedef uint32_t fsea_cptr;
typedef CUtensorMap fsea_tmap;
#endif

__device__ __forceinline__ fsea_cptr fsea_mapa(const void* local_smem, unsigned cta_rank) {
#if defined(FSEA_EMU)
    return fsea_emu::map_cluster((unsigned char*)local_smem, cta_rank);
#else
    uint32_t r;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem);
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta_rank));
    return r;
#endif
}
__device__ __forceinline__ void fsea_st_cluster(fsea_cptr p, int byte_off, float v) {
#if defined(FSEA_EMU)
    *(float*)(p + byte_off) = v;
#else
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(p + (uint32_t)byte_off), "f"(v) : "memory");
#endif
}
__device__ __forceinline__ void fsea_st_cluster(fsea_cptr p, int byte_off, float2 v) {
#if defined(FSEA_EMU)
    *(float2*)(p + byte_off) = v;
#else
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(p + (uint32_t)byte_off), "f"(v.x), "f"(v.y) : "memory");
#endif
}
// asynchronous store into the shared memory of a block of the cluster; the bytes are counted in on `bar`, an mbarrier of
// the SAME destination block (complete_tx): the consumer's wait on that barrier is the only synchronisation -- no
// release fence (a release at cluster scope is a MEMBAR.ALL.GPU + ERRBAR per thread)
__device__ __forceinline__ void fsea_st_async(fsea_cptr p, int byte_off, float v, fsea_cptr bar) {
#if defined(FSEA_EMU)
    *(float*)(p + byte_off) = v;
    fsea_emu::mbar_complete_tx(bar, 4);
#else
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 ::"r"(p + (uint32_t)byte_off), "f"(v), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void fsea_st_async(fsea_cptr p, int byte_off, float2 v, fsea_cptr bar) {
#if defined(FSEA_EMU)
    *(float2*)(p + byte_off) = v;
    fsea_emu::mbar_complete_tx(bar, 8);
#else
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(p + (uint32_t)byte_off), "f"(v.x), "f"(v.y), "r"(bar) : "memory");
#endif
}
// the same with a compile-time byte offset folded into the instruction's address (one base register for a whole row of stores)
template <int OFF> __device__ __forceinline__ void fsea_st_async_at(fsea_cptr p, float v, fsea_cptr bar) {
#if defined(FSEA_EMU)
    *(float*)(p + OFF) = v;
    fsea_emu::mbar_complete_tx(bar, 4);
#else
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0+%3], %1, [%2];"
                 ::"r"(p), "f"(v), "r"(bar), "n"(OFF) : "memory");
#endif
}
template <int OFF> __device__ __forceinline__ void fsea_st_async_at(fsea_cptr p, float2 v, fsea_cptr bar) {
#if defined(FSEA_EMU)
    *(float2*)(p + OFF) = v;
    fsea_emu::mbar_complete_tx(bar, 8);
#else
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0+%4], {%1, %2}, [%3];"
                 ::"r"(p), "f"(v.x), "f"(v.y), "r"(bar), "n"(OFF) : "memory");
#endif
}
// plain stores into a block of the cluster with the byte offset as an immediate; made visible by fsea_fence_cluster()
// + an arrive on the consumer's barrier
template <int OFF> __device__ __forceinline__ void fsea_st_cluster_at(fsea_cptr p, float v) {
#if defined(FSEA_EMU)
    *(float*)(p + OFF) = v;
#else
    asm volatile("st.shared::cluster.f32 [%0+%2], %1;" ::"r"(p), "f"(v), "n"(OFF) : "memory");
#endif
}
template <int OFF> __device__ __forceinline__ void fsea_st_cluster_at(fsea_cptr p, float2 v) {
#if defined(FSEA_EMU)
    *(float2*)(p + OFF) = v;
#else
    asm volatile("st.shared::cluster.v2.f32 [%0+%3], {%1, %2};" ::"r"(p), "f"(v.x), "f"(v.y), "n"(OFF) : "memory");
#endif
}
__device__ __forceinline__ void fsea_fence_cluster() {
#if !defined(FSEA_EMU)
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_arrive_cluster_relaxed(fsea_cptr bar) {     // a signal that orders nothing but itself
#if defined(FSEA_EMU)
    fsea_emu::mbar_arrive(bar, 0);
#else
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_init(void* bar, unsigned count) {
#if defined(FSEA_EMU)
    fsea_emu::mbar_init(bar, count);
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_init_fence() {          // makes the inits visible to the other blocks of the cluster
#if !defined(FSEA_EMU)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_arrive(void* bar) {     // this block's barrier
#if defined(FSEA_EMU)
    fsea_emu::mbar_arrive(bar, 0);
#else
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_arrive_cluster(fsea_cptr bar) {   // a barrier of any block of the cluster; releases this thread's (remote) stores
#if defined(FSEA_EMU)
    fsea_emu::mbar_arrive(bar, 0);
#else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_expect_tx(void* bar, unsigned bytes) {      // arrive + expect `bytes` of asynchronous copies
#if defined(FSEA_EMU)
    fsea_emu::mbar_arrive(bar, bytes);
#else
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void fsea_mbar_wait(void* bar, unsigned parity) {          // until the phase of that parity has completed
#if defined(FSEA_EMU)
    fsea_emu::mbar_wait(bar, parity);
#else
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FSEA_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FSEA_MBAR_DONE_%=;\n"
        "bra FSEA_MBAR_WAIT_%=;\n"
        "FSEA_MBAR_DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
#endif
}
#ifndef FSEA_MBARC_MODE
#define FSEA_MBARC_MODE 1
#endif
__device__ __forceinline__ void fsea_mbar_wait_cluster(void* bar, unsigned parity) {  // ... acquiring what the arriving threads of OTHER blocks released
#if defined(FSEA_EMU)
    fsea_emu::mbar_wait(bar, parity);
#elif FSEA_MBARC_MODE == 0
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FSEA_MBARC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"      // (invalidates L1 on EVERY poll)
        "@p bra FSEA_MBARC_DONE_%=;\n"
        "bra FSEA_MBARC_WAIT_%=;\n"
        "FSEA_MBARC_DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
#else
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FSEA_MBARC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n"      // spin relaxed, acquire once
        "@p bra FSEA_MBARC_DONE_%=;\n"
        "bra FSEA_MBARC_WAIT_%=;\n"
        "FSEA_MBARC_DONE_%=:\n"
#if FSEA_MBARC_MODE == 1
        "fence.acq_rel.cluster;\n"
#endif
        "}\n" ::"r"(a), "r"(parity) : "memory");
#endif
}
// one box of a 3-d tensor (bytes) -> shared memory, completion counted on `bar` in bytes
__device__ __forceinline__ void fsea_tma_load_3d(void* smem_dst, const fsea_tmap* tm, int c0, int c1, int c2, void* bar) {
#if defined(FSEA_EMU)
    fsea_emu::tma_load_3d((unsigned char*)smem_dst, tm, c0, c1, c2, bar);
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2),
                   "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
__device__ __forceinline__ void fsea_cluster_sync_all() {
    fsea_cluster_arrive_release();
    fsea_cluster_wait_acquire();
}

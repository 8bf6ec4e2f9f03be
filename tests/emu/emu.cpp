// tests/emu/emu.cpp -- TEST INFRASTRUCTURE: the fiber machine behind tests/emu/cuda_runtime.h.
// Every CUDA thread of a launch is a ucontext fiber; all fibers of a launch run on the calling
// OS thread, round-robin, and switch only inside barriers, warp collectives and cluster waits.
#include "cuda_runtime.h"

#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <map>
#include <vector>

#include "../../frequensea_b200/csrc/fsea_kparams.h"
struct fsea_tmap { const unsigned char* base; unsigned long long dim[3], stride[3]; unsigned box[3]; };

namespace fsea_emu {

struct Barrier { int arrived = 0; unsigned gen = 0; };

struct Block {
    std::vector<unsigned char> smem;
    std::map<int, std::vector<unsigned char>> statics;
    Barrier block_bar;
    Barrier named[16];
    std::vector<Barrier> warp_bar;
    std::vector<unsigned> warp_mail;          // 32 words per warp
    std::vector<double> warp_mail64;
    int threads = 0;
    int cluster = 0;                          // index into Launch::clusters
};

struct PendingCopy { void* dst; const void* src; int bytes; };

struct Fiber {
    Thread th;
    ucontext_t ctx;
    void* stack = nullptr;
    bool done = false;
    unsigned cluster_gen = 0;                 // generation seen at this thread's last cluster arrive
    std::vector<PendingCopy> copies;
};

struct Launch {
    std::vector<Block> blocks;
    std::vector<Barrier> clusters;
    std::vector<Fiber> fibers;
    const std::function<void()>* body = nullptr;
    ucontext_t sched;
    int current = -1;
    int cluster_threads = 0;
};

Thread* cur = nullptr;
static Launch* L = nullptr;
static const size_t STACK = 512 * 1024;
static unsigned long progress = 0;            // state changes (arrivals, exits): a round without any is a deadlock

static Fiber& me() { return L->fibers[L->current]; }
static Block& myblock() { return *(Block*)cur->block; }

void yield() { swapcontext(&me().ctx, &L->sched); }

static void barrier_wait(Barrier& b, int count) {
    const unsigned g = b.gen;
    progress++;
    if (++b.arrived == count) { b.arrived = 0; b.gen++; return; }
    while (b.gen == g) yield();
}

unsigned char* dyn_smem() { return myblock().smem.data(); }
void* static_smem(int key, size_t bytes) {
    auto& v = myblock().statics[key];
    if (v.size() < bytes) v.resize(bytes);
    return v.data();
}
void sync_block() { Block& b = myblock(); barrier_wait(b.block_bar, b.threads); }
void sync_named(int id, int count) { barrier_wait(myblock().named[id & 15], count); }
void arrive_named(int id, int count) {
    Barrier& b = myblock().named[id & 15];
    progress++;
    if (++b.arrived == count) { b.arrived = 0; b.gen++; }
}
void sync_warp() { Block& b = myblock(); barrier_wait(b.warp_bar[cur->tid.x >> 5], 32); }
unsigned shfl_idx(unsigned v, int src_lane) {
    Block& b = myblock();
    const int w = cur->tid.x >> 5, lane = cur->tid.x & 31;
    b.warp_mail[w * 32 + lane] = v;
    barrier_wait(b.warp_bar[w], 32);
    const unsigned r = b.warp_mail[w * 32 + (src_lane & 31)];
    barrier_wait(b.warp_bar[w], 32);
    return r;
}
double shfl_down_f64(double v, int delta) {
    Block& b = myblock();
    const int w = cur->tid.x >> 5, lane = cur->tid.x & 31;
    b.warp_mail64[w * 32 + lane] = v;
    barrier_wait(b.warp_bar[w], 32);
    const double r = (lane + delta < 32) ? b.warp_mail64[w * 32 + lane + delta] : v;
    barrier_wait(b.warp_bar[w], 32);
    return r;
}
void cluster_arrive() {
    Barrier& c = L->clusters[myblock().cluster];
    me().cluster_gen = c.gen;
    progress++;
    if (++c.arrived == L->cluster_threads) { c.arrived = 0; c.gen++; }
}
void cluster_wait() {
    Barrier& c = L->clusters[myblock().cluster];
    while (c.gen == me().cluster_gen) yield();
}
// ---- distributed shared memory / mbarriers / TMA (fsea_wsc.cuh) ---------------------------------------
unsigned char* map_cluster(unsigned char* local_smem, unsigned cta_rank) {
    Block& b = myblock();
    const size_t off = (size_t)(local_smem - b.smem.data());
    if (off >= b.smem.size()) { fprintf(stderr, "fsea_emu: mapa of an address outside dynamic shared memory\n"); abort(); }
    const int cs = cur->cluster_size;
    const int first = (int)(cur->bid.x / cs) * cs;
    return L->blocks[first + (int)cta_rank].smem.data() + off;
}
struct MBar { unsigned count = 0, pending = 0; long tx = 0; unsigned phase = 0; };
static std::map<void*, MBar> g_mbar;                     // keyed by the barrier's address (in some block's shared memory)
static void mbar_try_complete(MBar& m) {
    if (m.pending == 0 && m.tx == 0) { m.pending = m.count; m.phase ^= 1u; progress++; }
}
void mbar_init(void* bar, unsigned count) { MBar m; m.count = m.pending = count; g_mbar[bar] = m; }
void mbar_arrive(void* bar, unsigned expect_tx_bytes) {
    auto it = g_mbar.find(bar);
    if (it == g_mbar.end()) { fprintf(stderr, "fsea_emu: arrive on an uninitialised mbarrier\n"); abort(); }
    MBar& m = it->second;
    if (m.pending == 0) { fprintf(stderr, "fsea_emu: more arrivals than the mbarrier's count\n"); abort(); }
    m.tx += expect_tx_bytes;
    m.pending--;
    progress++;
    mbar_try_complete(m);
}
void mbar_complete_tx(void* bar, unsigned bytes) {
    MBar& m = g_mbar.at(bar);
    m.tx -= bytes;
    progress++;
    mbar_try_complete(m);
}
void mbar_wait(void* bar, unsigned parity) {
    auto it = g_mbar.find(bar);
    if (it == g_mbar.end()) { fprintf(stderr, "fsea_emu: wait on an uninitialised mbarrier\n"); abort(); }
    while (it->second.phase == (parity & 1u)) yield();     // the phase of that parity is still in progress
}
void tma_load_3d(unsigned char* dst, const fsea_tmap* tm, int c0, int c1, int c2, void* bar) {
    // box[0] bytes x box[1] rows x box[2] planes, rows packed in shared memory; out-of-bounds elements read as zero
    unsigned bytes = 0;
    for (unsigned z = 0; z < tm->box[2]; z++)
        for (unsigned y = 0; y < tm->box[1]; y++)
            for (unsigned x = 0; x < tm->box[0]; x++, bytes++) {
                const unsigned long long i0 = c0 + x, i1 = c1 + y, i2 = c2 + z;
                unsigned char v = 0;
                if (i0 < tm->dim[0] && i1 < tm->dim[1] && i2 < tm->dim[2]) v = tm->base[i0 + i1 * tm->stride[1] + i2 * tm->stride[2]];
                dst[bytes] = v;
            }
    mbar_complete_tx(bar, bytes);
}

void cp_async(void* dst, const void* src, int bytes) { me().copies.push_back({dst, src, bytes}); }
void cp_async_wait_all() {
    for (auto& c : me().copies) memcpy(c.dst, c.src, (size_t)c.bytes);
    me().copies.clear();
}

static void trampoline() {
    (*L->body)();
    Fiber& f = me();
    // (exiting with copies in flight is legal -- idle groups do it after their prologue -- so only on request)
    if (!f.copies.empty() && getenv("FSEA_EMU_VERBOSE")) fprintf(stderr, "fsea_emu: thread exited with %zu cp.async copies in flight\n", f.copies.size());
    f.done = true;
    swapcontext(&f.ctx, &L->sched);
}

void launch(dim3 grid, dim3 block, size_t smem, int cluster_size, const std::function<void()>& body) {
    if (L) { fprintf(stderr, "fsea_emu: nested launch\n"); abort(); }
    Launch launch_state;
    L = &launch_state;
    const int nb = (int)grid.x, nt = (int)block.x;
    if (nt % 32) { fprintf(stderr, "fsea_emu: block size %d is not a multiple of 32\n", nt); abort(); }
    L->body = &body;
    L->blocks.resize(nb);
    L->clusters.resize((nb + cluster_size - 1) / cluster_size);
    L->cluster_threads = cluster_size * nt;
    L->fibers.resize((size_t)nb * nt);
    for (int b = 0; b < nb; b++) {
        Block& B = L->blocks[b];
        B.smem.assign(smem + 64, 0xAB);                    // poison: uninitialised reads show up
        B.threads = nt;
        B.cluster = b / cluster_size;
        B.warp_bar.resize(nt / 32);
        B.warp_mail.resize(nt);
        B.warp_mail64.resize(nt);
        for (int t = 0; t < nt; t++) {
            Fiber& f = L->fibers[(size_t)b * nt + t];
            f.th.tid = {(unsigned)t, 0, 0};
            f.th.bid = {(unsigned)b, 0, 0};
            f.th.bdim = {(unsigned)nt, 1, 1};
            f.th.gdim = {(unsigned)nb, 1, 1};
            f.th.cluster_size = cluster_size;
            f.th.block = &B;
            f.stack = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (f.stack == MAP_FAILED) { perror("fsea_emu mmap"); abort(); }
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
    }
    size_t live = L->fibers.size();
    while (live) {
        const unsigned long before = progress;
        for (size_t i = 0; i < L->fibers.size(); i++) {
            Fiber& f = L->fibers[i];
            if (f.done) continue;
            L->current = (int)i;
            cur = &f.th;
            swapcontext(&L->sched, &f.ctx);
            if (f.done) { live--; progress++; munmap(f.stack, STACK); f.stack = nullptr; }
        }
        if (progress == before) { fprintf(stderr, "fsea_emu: deadlock (a scheduler round changed nothing)\n"); abort(); }
    }
    cur = nullptr;
    L = nullptr;
    g_mbar.clear();
}

}  // namespace fsea_emu

cudaError_t fsea_emu_launch_kparams(const void* fn, dim3 grid, dim3 block, void** args, size_t smem, int cluster) {
    const fsea_kparams kp = *(const fsea_kparams*)args[0];
    if (kp.tmap_arg) {
        // kernel(kparams, tensor map)
        const fsea_tmap tm = *(const fsea_tmap*)args[1];
        void (*k2)(const fsea_kparams, const fsea_tmap) = (void (*)(const fsea_kparams, const fsea_tmap))fn;
        fsea_emu::launch(grid, block, smem, cluster, [&] { k2(kp, tm); });
        return cudaSuccess;
    }
    void (*k)(const fsea_kparams) = (void (*)(const fsea_kparams))fn;
    fsea_emu::launch(grid, block, smem, cluster, [&] { k(kp); });
    return cudaSuccess;
}

// fsea_k_ws4.cu -- instantiations of the warp-specialised block-per-frame kernel (fsea_ws4.cuh)
#include "fsea_instantiate.cuh"
#include "fsea_ws4.cuh"
#include "fsea_wsc.cuh"

#ifndef FSEA_WS_TMA
#define FSEA_WS_TMA 1          /* N = 16384: raw tiles by TMA (0: gathered by the transform warps) */
#endif
template <int N2, bool TMA>
static bool fill_ws(fsea_kernel_info4* o, bool window, int out_mode, const char* name) {
#define FSEA_WS(W, O) (const void*)&fsea_ws4_kernel<N2, O, W, TMA>
    o->fn = window ? (out_mode == FSEA_OUT_MAG ? FSEA_WS(true, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_WS(true, FSEA_OUT_LOGMAG) : FSEA_WS(true, FSEA_OUT_PIX))
                   : (out_mode == FSEA_OUT_MAG ? FSEA_WS(false, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_WS(false, FSEA_OUT_LOGMAG) : FSEA_WS(false, FSEA_OUT_PIX));
#undef FSEA_WS
    o->fn_sums = nullptr;
    if constexpr (N2 == 16) { if (!window && out_mode == FSEA_OUT_MAG) o->fn_sums = (const void*)&fsea_ws4_kernel<N2, FSEA_OUT_MAG, false, TMA, true>; }
    o->cs = 1;
    o->ws = 16 / N2;                       // frames per block iteration
    o->tma = TMA ? 1 : 2;                  // (2: the kernel takes a tensor-map argument it does not read)
    o->threads = FSEA_WS_THREADS;
    o->smem = TMA ? FSEA_WS_SMEM_TMA : FSEA_WS_SMEM;
    o->n1 = 1024; o->n2 = N2;
    o->r0 = 32; o->r1 = 32; o->r2 = 1;
    o->tw1 = 31 * 32; o->tw2 = 0;
    o->name = name;
    return true;
}

bool fsea_get_kernel_ws4(int n, bool window, int out_mode, fsea_kernel_info4* o) {
    switch (n) {
        case 4096: return fill_ws<4, false>(o, window, out_mode, "ws4<4096=1024x4: 8 transform + 4 comb warps, 4 frames per block, intermediate in shared memory, f32x2>");
        case 8192:
            // (TMA tiles measured at this length too: 50.3 vs 58.5 M frames/s -- rows of 16 bytes make the gathers cheap, 4 lines per
            //  load, and the staged form pays 4-way conflicts and loses the twiddle table's place in shared memory)
            return fill_ws<8, false>(o, window, out_mode, "ws4<8192=1024x8: 8 transform + 4 comb warps, 2 frames per block, intermediate in shared memory, f32x2>");
        case 16384:
            // measured (profiles/r02_ws4/sweep_tma_ab.txt): float32 spectra 25.7 vs 23.8 M frames/s with TMA tiles; u8 pixels,
            // where the comb warps' heavier epilogue is the critical path, 23.9 vs 24.7: the gathered form stays there
            if (FSEA_WS_TMA && out_mode != FSEA_OUT_PIX) return fill_ws<16, true>(o, window, out_mode, "ws4<16384=1024x16: 8 transform + 4 comb warps per frame, intermediate in shared memory, TMA tiles, f32x2>");
            return fill_ws<16, false>(o, window, out_mode, "ws4<16384=1024x16: 8 transform + 4 comb warps per frame, intermediate in shared memory, f32x2>");
    }
    return false;
}

template <int N2>
static bool fill_wsc(fsea_kernel_info4* o, bool window, int out_mode, const char* name) {
#define FSEA_WSC(W, O) (const void*)&fsea_wsc_kernel<N2, O, W>
    o->fn = window ? (out_mode == FSEA_OUT_MAG ? FSEA_WSC(true, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_WSC(true, FSEA_OUT_LOGMAG) : FSEA_WSC(true, FSEA_OUT_PIX))
                   : (out_mode == FSEA_OUT_MAG ? FSEA_WSC(false, FSEA_OUT_MAG) : out_mode == FSEA_OUT_LOGMAG ? FSEA_WSC(false, FSEA_OUT_LOGMAG) : FSEA_WSC(false, FSEA_OUT_PIX));
#undef FSEA_WSC
    o->fn_sums = nullptr;
    o->cs = N2 / 16;
    o->ws = 1;
    o->tma = 1;
    o->threads = FSEA_WS_THREADS;
    o->smem = FSEA_WSC_SMEM;
    o->n1 = 1024; o->n2 = N2;
    o->r0 = 32; o->r1 = 32; o->r2 = 1;
    o->tw1 = 31 * 32; o->tw2 = 0;
    o->name = name;
    return true;
}

// the cluster form: intermediate in distributed shared memory, raw tiles by TMA (fsea_wsc.cuh)
bool fsea_get_kernel_wsc(int n, bool window, int out_mode, fsea_kernel_info4* o) {
    switch (n) {
        case 32768: return fill_wsc<32>(o, window, out_mode, "wsc<32768=1024x32: cluster of 2, 8 transform + 4 comb warps per block, intermediate in distributed shared memory, TMA tiles, f32x2>");
        case 65536: return fill_wsc<64>(o, window, out_mode, "wsc<65536=1024x64: cluster of 4, 8 transform + 4 comb warps per block, intermediate in distributed shared memory, TMA tiles, f32x2>");
    }
    return false;
}

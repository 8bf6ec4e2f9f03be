// fsea_mgpu.cpp -- the sweep job on the GPUs of one box, one process (include/fsea_mgpu.h).
#include "../../include/fsea_mgpu.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "../../include/fsea.h"

static thread_local char g_err[256] = "";
extern "C" const char* fsea_mgpu_last_error(void) { return g_err; }

#define CU(call)                                                                     \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e__)); rc = FSEA_ERR_CUDA; goto done; } \
    } while (0)
#define NC(call)                                                                     \
    do {                                                                             \
        ncclResult_t r__ = (call);                                                   \
        if (r__ != ncclSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #call, ncclGetErrorString(r__)); rc = FSEA_ERR_CUDA; goto done; } \
    } while (0)
#define FS(call)                                                                     \
    do {                                                                             \
        int r__ = (call);                                                            \
        if (r__ != FSEA_OK) { snprintf(g_err, sizeof(g_err), "%s: %s %s", #call, fsea_strerror(r__), fsea_last_cuda_error()); rc = r__; goto done; } \
    } while (0)

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" int fsea_mgpu_sweep(const fsea_mgpu_sweep_config* cfg, uint8_t* h_stitched, double* means, uint8_t* keep,
                               fsea_mgpu_sweep_timing* timing) {
    if (!cfg || cfg->n_gpus <= 0 || cfg->n_freqs <= 0 || cfg->rows <= 0 || cfg->width_step <= 0) return FSEA_ERR_INVALID;
    if (cfg->mode == FSEA_MGPU_P2P && cfg->width_step != cfg->fft_size) return FSEA_ERR_INVALID;
    if (fsea_device_count() < cfg->n_gpus) return FSEA_ERR_NO_DEVICE;
    const int G = cfg->n_gpus, n = cfg->fft_size, rows = cfg->rows, nf = cfg->n_freqs;
    const int per = (nf + G - 1) / G;
    const size_t img = (size_t)rows * n;
    const long out_w = (long)n + (long)(nf - 1) * cfg->width_step;
    int rc = FSEA_OK;
    std::vector<fsea_sweep*> eng((size_t)G, nullptr);
    std::vector<int> count((size_t)G, 0);
    std::vector<ncclComm_t> comm((size_t)G, nullptr);
    std::vector<cudaEvent_t> ev((size_t)G * 4, nullptr);
    std::vector<double> lm((size_t)per);
    std::vector<uint8_t> lk((size_t)per);
    uint8_t* d_gathered = nullptr;       // device 0: [G][per][rows][n]
    uint8_t* d_canvas = nullptr;         // device 0: rows x out_w
    fsea_mgpu_sweep_timing t;
    memset(&t, 0, sizeof(t));
    double t0, t1;
    const bool gather = cfg->mode == FSEA_MGPU_GATHER, p2p = cfg->mode == FSEA_MGPU_P2P;

    for (int g = 0; g < G; g++) {
        fsea_sweep_config sc;
        memset(&sc, 0, sizeof(sc));
        sc.fft_size = n; sc.rows = rows; sc.gate_rows = cfg->gate_rows; sc.gate_threshold = cfg->gate_threshold;
        sc.pix_scale = cfg->pix_scale; sc.in_format = FSEA_IN_S8; sc.device = g; sc.capacity = per;
        sc.own_raw = 1; sc.own_images = p2p ? 0 : 1;
        FS(fsea_sweep_create(&sc, &eng[g]));
        CU(cudaSetDevice(g));
        for (int k = 0; k < 4; k++) CU(cudaEventCreate(&ev[(size_t)g * 4 + k]));
    }
    if (gather && G > 1) {
        std::vector<int> devs((size_t)G);
        for (int g = 0; g < G; g++) devs[g] = g;
        NC(ncclCommInitAll(comm.data(), G, devs.data()));
    }
    CU(cudaSetDevice(0));
    if (gather) CU(cudaMalloc(&d_gathered, (size_t)G * per * img));
    if (gather && G > 1) {
        // first use of a send / recv pair sets up the NVLink connection (100s of ms): do that outside the timed region
        NC(ncclGroupStart());
        for (int g = 1; g < G; g++) {
            NC(ncclSend(fsea_sweep_images(eng[g]), 256, ncclUint8, 0, comm[g], (cudaStream_t)fsea_sweep_stream(eng[g])));
            NC(ncclRecv(d_gathered + (size_t)g * per * img, 256, ncclUint8, g, comm[0], (cudaStream_t)fsea_sweep_stream(eng[0])));
        }
        NC(ncclGroupEnd());
    }
    if (gather || p2p) {
        CU(cudaMalloc(&d_canvas, (size_t)rows * out_w));
        CU(cudaMemset(d_canvas, 0, (size_t)rows * out_w));
    }
    if (p2p)
        for (int g = 1; g < G; g++) {
            CU(cudaSetDevice(g));
            cudaError_t e = cudaDeviceEnablePeerAccess(0, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { snprintf(g_err, sizeof(g_err), "peer access %d -> 0: %s", g, cudaGetErrorString(e)); rc = FSEA_ERR_CUDA; goto done; }
            cudaGetLastError();
        }

    // ---- stage (untimed): the raw IQ of every frequency becomes resident on its device
    t0 = now_s();
    for (int f = 0; f < nf; f++) {
        const int g = f % G;
        CU(cudaSetDevice(g));
        FS(fsea_sweep_submit_synth(eng[g], count[g]++, f, cfg->quiet_every));
    }
    for (int g = 0; g < G; g++) FS(fsea_sweep_sync(eng[g]));
    t.stage_s = now_s() - t0;

    // ---- timed: decisions, pixel rows, exchange, stitch
    t0 = now_s();
    {
        const long l0 = [&] { long s = 0; for (int g = 0; g < G; g++) s += fsea_sweep_launch_count(eng[g]); return s; }();
        for (int g = 0; g < G; g++) {
            if (!count[g]) continue;
            CU(cudaSetDevice(g));
            cudaStream_t st = (cudaStream_t)fsea_sweep_stream(eng[g]);
            CU(cudaEventRecord(ev[(size_t)g * 4 + 0], st));
            FS(fsea_sweep_gate_begin(eng[g], count[g]));
            CU(cudaEventRecord(ev[(size_t)g * 4 + 1], st));
        }
        for (int g = 0; g < G; g++) {
            if (!count[g]) continue;
            FS(fsea_sweep_gate_end(eng[g], count[g], lm.data(), lk.data()));
            for (int j = 0; j < count[g]; j++) {
                const int f = j * G + g;
                if (means) means[f] = lm[j];
                if (keep) keep[f] = lk[j];
                t.kept += lk[j];
            }
        }
        for (int g = 0; g < G; g++) {
            if (!count[g]) continue;
            CU(cudaSetDevice(g));
            cudaStream_t st = (cudaStream_t)fsea_sweep_stream(eng[g]);
            if (p2p) FS(fsea_sweep_render(eng[g], count[g], d_canvas, out_w, (long)g * n, (long)G * n));
            else FS(fsea_sweep_render(eng[g], count[g], nullptr, 0, 0, 0));
            CU(cudaEventRecord(ev[(size_t)g * 4 + 2], st));
        }
        if (gather) {
            if (G > 1) {
                NC(ncclGroupStart());
                for (int g = 0; g < G; g++) {
                    if (!count[g]) continue;
                    cudaStream_t st = (cudaStream_t)fsea_sweep_stream(eng[g]);
                    if (g != 0) {
                        NC(ncclSend(fsea_sweep_images(eng[g]), (size_t)count[g] * img, ncclUint8, 0, comm[g], st));
                        NC(ncclRecv(d_gathered + (size_t)g * per * img, (size_t)count[g] * img, ncclUint8, g, comm[0], (cudaStream_t)fsea_sweep_stream(eng[0])));
                    }
                }
                NC(ncclGroupEnd());
            }
            CU(cudaSetDevice(0));
            cudaStream_t s0 = (cudaStream_t)fsea_sweep_stream(eng[0]);
            CU(cudaMemcpyAsync(d_gathered, fsea_sweep_images(eng[0]), (size_t)count[0] * img, cudaMemcpyDeviceToDevice, s0));
            CU(cudaEventRecord(ev[3], s0));
            FS(fsea_stitch_sharded_device(d_gathered, nf, G, per, rows, n, cfg->width_step, d_canvas, s0));
            t.launches++;
        }
        for (int g = 0; g < G; g++) FS(fsea_sweep_sync(eng[g]));
        t1 = now_s();
        t.total_s = t1 - t0;
        for (int g = 0; g < G; g++) {
            if (!count[g]) continue;
            float ms = 0;
            CU(cudaSetDevice(g));
            CU(cudaEventElapsedTime(&ms, ev[(size_t)g * 4 + 0], ev[(size_t)g * 4 + 1])); t.gate_s = std::max(t.gate_s, ms * 1e-3);
            CU(cudaEventElapsedTime(&ms, ev[(size_t)g * 4 + 1], ev[(size_t)g * 4 + 2])); t.render_s = std::max(t.render_s, ms * 1e-3);
        }
        if (gather) {
            float ms = 0;
            CU(cudaSetDevice(0));
            CU(cudaEventRecord(ev[1], (cudaStream_t)fsea_sweep_stream(eng[0])));      // (end of the stitch: recorded late, synchronised below)
            CU(cudaEventSynchronize(ev[1]));
            CU(cudaEventElapsedTime(&ms, ev[2], ev[3])); t.exchange_s = ms * 1e-3;
            CU(cudaEventElapsedTime(&ms, ev[3], ev[1])); t.stitch_s = ms * 1e-3;
        }
        long l1 = 0;
        for (int g = 0; g < G; g++) l1 += fsea_sweep_launch_count(eng[g]);
        t.launches += l1 - l0;
        t.rows_rendered = t.kept * rows;
    }
    if (h_stitched && d_canvas) {
        CU(cudaSetDevice(0));
        CU(cudaMemcpy(h_stitched, d_canvas, (size_t)rows * out_w, cudaMemcpyDeviceToHost));
    }
    if (timing) *timing = t;

done:
    for (int g = 0; g < G; g++) {
        if (comm[g]) ncclCommDestroy(comm[g]);
        cudaSetDevice(g);
        for (int k = 0; k < 4; k++) if (ev[(size_t)g * 4 + k]) cudaEventDestroy(ev[(size_t)g * 4 + k]);
        fsea_sweep_destroy(eng[g]);
    }
    cudaSetDevice(0);
    if (d_gathered) cudaFree(d_gathered);
    if (d_canvas) cudaFree(d_canvas);
    return rc;
}

extern "C" int fsea_mgpu_gather_bench(int n_gpus, long bytes_per_gpu, int repeats, double* seconds_out) {
    if (n_gpus < 2 || bytes_per_gpu <= 0 || repeats <= 0 || !seconds_out) return FSEA_ERR_INVALID;
    if (fsea_device_count() < n_gpus) return FSEA_ERR_NO_DEVICE;
    const int G = n_gpus;
    int rc = FSEA_OK;
    std::vector<ncclComm_t> comm((size_t)G, nullptr);
    std::vector<cudaStream_t> st((size_t)G, nullptr);
    std::vector<void*> buf((size_t)G, nullptr);
    std::vector<int> devs((size_t)G);
    void* root = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0;
    for (int g = 0; g < G; g++) devs[g] = g;
    NC(ncclCommInitAll(comm.data(), G, devs.data()));
    for (int g = 0; g < G; g++) {
        CU(cudaSetDevice(g));
        CU(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
        if (g) { CU(cudaMalloc(&buf[g], (size_t)bytes_per_gpu)); CU(cudaMemset(buf[g], g, (size_t)bytes_per_gpu)); }
    }
    CU(cudaSetDevice(0));
    CU(cudaMalloc(&root, (size_t)bytes_per_gpu * (G - 1)));
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int it = 0; it <= repeats; it++) {                      // (first round: warm-up, connections)
        if (it == 1) { for (int g = 0; g < G; g++) { CU(cudaSetDevice(g)); CU(cudaStreamSynchronize(st[g])); } CU(cudaSetDevice(0)); CU(cudaEventRecord(e0, st[0])); }
        NC(ncclGroupStart());
        for (int g = 1; g < G; g++) {
            NC(ncclSend(buf[g], (size_t)bytes_per_gpu, ncclUint8, 0, comm[g], st[g]));
            NC(ncclRecv((char*)root + (size_t)(g - 1) * bytes_per_gpu, (size_t)bytes_per_gpu, ncclUint8, g, comm[0], st[0]));
        }
        NC(ncclGroupEnd());
    }
    CU(cudaSetDevice(0));
    CU(cudaEventRecord(e1, st[0]));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    *seconds_out = ms * 1e-3 / repeats;
done:
    for (int g = 0; g < G; g++) {
        cudaSetDevice(g);
        if (st[g]) { cudaStreamSynchronize(st[g]); cudaStreamDestroy(st[g]); }
        if (buf[g]) cudaFree(buf[g]);
        if (comm[g]) ncclCommDestroy(comm[g]);
    }
    cudaSetDevice(0);
    if (root) cudaFree(root);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

// fsea_internal.h -- what the other translation units of libfsea.so need from fsea_engine.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fsea.h"
#include "fsea_ptx.cuh"

// plain (non-table) kernel launches; -DFSEA_EMU (tests/emu/, test infrastructure) runs them on the CPU fiber machine
#if defined(FSEA_EMU)
#define FSEA_LAUNCH(kernel, grid, block, stream, ...) fsea_emu::launch(dim3(grid), dim3(block), 0, 1, [=] { kernel(__VA_ARGS__); })
#else
#define FSEA_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

// launch a plan on `stream` (no device switch, no argument checks); out_stride = bytes between
// output rows (0: packed); scratch_slot < 4: which L2 intermediate a four-step plan uses (launches that
// share a slot on different streams are serialised)
int fsea_launch_internal(fsea_plan* p, const void* d_in, void* d_out, long n_frames, cudaStream_t stream, long out_stride, int scratch_slot = 0,
                         long out_group = 0, long out_group_stride = 0,    // out_group > 0: rows in groups (fsea_kparams.h)
                         long hop = 0);                                     // hop > 0: samples between this launch's frames (0: the plan's)
// per-frame sums of |X| instead of spectra (plans whose kernel has the SUMS form: N = 16384, 8-bit, float32 magnitudes):
// comb warp w of frame f adds into sums[f * sums_stride + sums_offset + w], w < 4
bool fsea_plan_has_sums_internal(const fsea_plan* p);
int fsea_launch_sums_internal(fsea_plan* p, const void* d_in, long n_frames, cudaStream_t stream, long hop, double* sums, long sums_stride,
                              long sums_offset);
int fsea_fail_cuda_internal(cudaError_t e, const char* what);      // records the text for fsea_last_cuda_error(), returns FSEA_ERR_CUDA
int fsea_plan_elem_bytes_internal(const fsea_plan* p);
void fsea_count_launch_internal(fsea_plan* p);

#define FSEA_SUM_BLOCKS 256
__global__ void fsea_sum_kernel(const float* __restrict__ x, long n, double* __restrict__ partial);

// float-sample frames longer than 8192 points (fsea_bigf.cu)
struct fsea_bigf;
struct fsea_kparams;
int fsea_bigf_create(int n, int in_format, fsea_bigf** out);
int fsea_bigf_destroy(fsea_bigf* b);
int fsea_bigf_launch(fsea_bigf* b, const fsea_kparams& kp, long n_frames, cudaStream_t stream, long* launches);

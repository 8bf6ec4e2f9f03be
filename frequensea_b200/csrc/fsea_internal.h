// fsea_internal.h -- what the other translation units of libfsea.so need from fsea_engine.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fsea.h"

// launch a plan on `stream` (no device switch, no argument checks); out_stride = bytes between
// output rows (0: packed)
int fsea_launch_internal(fsea_plan* p, const void* d_in, void* d_out, long n_frames, cudaStream_t stream, long out_stride);
int fsea_fail_cuda_internal(cudaError_t e, const char* what);      // records the text for fsea_last_cuda_error(), returns FSEA_ERR_CUDA
int fsea_plan_elem_bytes_internal(const fsea_plan* p);
void fsea_count_launch_internal(fsea_plan* p);

#define FSEA_SUM_BLOCKS 256
__global__ void fsea_sum_kernel(const float* __restrict__ x, long n, double* __restrict__ partial);

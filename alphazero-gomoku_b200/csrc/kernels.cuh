// Kernel declarations shared between tree.cu (definitions) and engine.cu (launches).
#pragma once
#include "common.cuh"
extern "C" {
__global__ void azg_fill_kernel(azg_dev e);
__global__ void azg_fill_fast_kernel(azg_dev e);
__global__ void azg_scan_kernel(azg_dev e);
__global__ void azg_leaf_planes_kernel(azg_dev e, float* out);
__global__ void azg_commit_kernel(azg_dev e, const float* probs, const double* noise);
__global__ void azg_begin_kernel(azg_dev e, const int32_t* plies, int n_sims, const int32_t* mask);
__global__ void azg_finish_kernel(azg_dev e, float* pi, int32_t* visits);
__global__ void azg_advance_kernel(azg_dev e, const int32_t* actions, int gc, int reserve, int32_t* status);
__global__ void azg_reset_kernel(azg_dev e, const int32_t* mask, const azg_pos* roots, int clear_tree);
__global__ void azg_stats_kernel(azg_dev e, unsigned long long* out);
}

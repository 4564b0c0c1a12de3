// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include "../../include/azgomoku_b200.h"

int azg_fail(int code, const char* msg);          // records the per-thread error text, returns code
int azg_check_launch(const char* what);           // cudaGetLastError -> AZG_E_CUDA

#define AZG_CUDA(x)                                                        \
  do {                                                                     \
    cudaError_t _e = (x);                                                  \
    if (_e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(_e)); \
  } while (0)

// Every entry point that launches work makes the handle's device current first (callers may drive several
// GPUs from one thread); cudaSetDevice on the already-current device is a cheap no-op.
#define AZG_USE_DEVICE(dev) AZG_CUDA(cudaSetDevice(dev))

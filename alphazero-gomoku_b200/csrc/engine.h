// Host-side engine object behind the opaque azg_engine handle.
#pragma once
#include "common.cuh"

struct azg_engine {
  azg_config cfg{};
  azg_dev dev{};
  cudaStream_t stream = nullptr;
  int64_t bytes = 0;
  unsigned long long* stats_dev = nullptr;
  uint64_t* pinned = nullptr;
};

// Host-side engine object behind the opaque azg_engine handle.
#pragma once
#include "common.cuh"

// Per-game example capture for the self-play driver (selfplay.cu).
struct azg_selfplay_buf {
  uint32_t* ex_key = nullptr;     // [G][max_plies][16] stones of the position each move was chosen in
  int32_t* ex_player = nullptr;   // [G][max_plies]     side to move
  float* ex_pi = nullptr;         // [G][max_plies][225] search policy
  int32_t* n_plies = nullptr;     // [G] plies played in the current game
  int32_t* n_done = nullptr;      // [G] games finished so far (part of the RNG counter: streams never repeat)
  int32_t max_plies = 0;
  const int32_t* active = nullptr; // [G] optional caller-owned mask: slots with 0 are retired (azg_selfplay_set_active)
};

struct azg_engine {
  azg_config cfg{};
  azg_dev dev{};
  cudaStream_t stream = nullptr;
  int64_t bytes = 0;
  unsigned long long* stats_dev = nullptr;
  uint64_t* pinned = nullptr;
  azg_selfplay_buf sp;
};

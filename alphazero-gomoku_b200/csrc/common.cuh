// Shared definitions for the sm_100a self-play engine.
//
// Board geometry and the packed position record follow the reference's game
// objects (games/gomoku.py:20-25, games/pente.py:12-23): 15x15 cells, action
// index a = r*15 + c, players 1/2, optional last move, Pente capture counters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/azgomoku_b200.h"

#define AZG_N 15
#define AZG_A 225
#define AZG_ROW 228              // row stride (elements) of per-node child arrays: 912 B, 16-B aligned
#define AZG_MAX_QUEUE 256        // upper bound for the reference's batch_size
#define AZG_MAX_DEPTH 512        // path stack entries per game
#define AZG_P64_SLOTS 8          // float64 prior rows per game (noised roots)
#define AZG_FULL 0xffffffffu

// Packed position (azg_pos) comes from the public header: stones[colour-1][word], bit a&31 of word a>>5 is cell a.
static_assert(sizeof(azg_pos) == 96, "azg_pos layout is part of the C ABI");

// Per-game search control block (device resident).
enum { AZG_ST_IDLE = 0, AZG_ST_RUN = 1, AZG_ST_NEED_EVAL = 2, AZG_ST_NEED_FINAL = 3, AZG_ST_DONE = 4, AZG_ST_ERROR = 5 };
enum { AZG_ERR_NODES = 1, AZG_ERR_DEPTH = 2, AZG_ERR_ROOT_TERMINAL = 4, AZG_ERR_P64 = 8, AZG_ERR_HASH = 16 };

struct azg_ctl {
  azg_pos root;                 // 96
  azg_pos scratch;              // 96  suspended simulation's position
  int32_t state;
  int32_t err;
  int32_t sims_left;
  int32_t ply;                  // move_number of this run
  int32_t root_node;            // node of the root key in this run, -1 until known
  int32_t susp;                 // 1: a simulation is parked at resume_node
  int32_t depth;                // path length of the parked simulation
  int32_t resume_node;
  int32_t n_pending;
  int32_t n_nodes;              // high-water mark of the slab
  int32_t n_free;               // entries on the free stack
  int32_t n_live;
  int32_t p64_used;             // bit mask of float64 prior rows in use
  int32_t leaf_off;             // offset of this game's queue in the leaf batch
  int32_t resets;               // trees dropped because the slab could not hold another run
  int32_t pad0;
  unsigned long long visits;    // node visits (search() entries in the reference)
  unsigned long long evals;     // rows sent to the evaluator
  unsigned long long sims;      // completed simulations
  unsigned long long pad1;
  int32_t pending[AZG_MAX_QUEUE];
};

// Engine-wide device view handed to kernels by value.
struct azg_dev {
  int32_t G, rule, queue_len, cap, hcap, noise_on, noise_plies, n_sims, game_base;
  int32_t fast_warps, virtual_loss;   // fast (non-parity) mode: warps per game walking the tree concurrently, virtual loss per edge
  int32_t fill_l1;              // Gomoku: the FILL kernel reads the slab through L1 and backs up with plain stores (tree.cu ldx)
  float cpuct;
  double eps, alpha, cpuct64;   // cpuct64: the Python float the reference multiplies with at a float64 root
  unsigned long long seed;
  azg_ctl* ctl;                 // [G]
  float* P;                     // [G][cap][AZG_ROW]
  int32_t* Nv;                  // [G][cap][AZG_ROW]
  int32_t* W;                   // [G][cap][AZG_ROW]
  int32_t* child;               // [G][cap][AZG_ROW] Gomoku only (else null): what playing action a from this node leads to -
                                //   0 unknown, 1 terminal with a winner, 2 terminal draw, n + 4 = node n (tree.cu, fill_body)
  uint32_t* key;                // [G][cap][16]
  uint32_t* meta;               // [G][cap]  bit0 alive, bit1-2 player, bits 4-7 p64 slot+1
  unsigned long long* slots;    // [G][hcap] (tag<<32)|(node+1), 0 empty
  int32_t* freelist;            // [G][cap]
  uint32_t* path;               // [G][AZG_MAX_DEPTH] (node<<8)|action
  double* P64;                  // [G][AZG_P64_SLOTS][AZG_ROW]
  int32_t* leaf_game;           // [G*queue_len]
  int32_t* leaf_node;           // [G*queue_len]
  int32_t* counters;            // [8]: 0 n_leaves, 1 n_active(not done), 2 n_error
};

#define AZG_META_ALIVE 1u

static __host__ __device__ inline size_t azg_node_off(const azg_dev& e, int g, int node) {
  return ((size_t)g * (size_t)e.cap + (size_t)node);
}

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t* p) { return __ldcg(p); }

// board mask for word w: cells 0..224
__device__ __forceinline__ uint32_t board_word_mask(int w) { return w < 7 ? 0xffffffffu : 0x00000001u; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
#endif

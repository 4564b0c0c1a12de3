// Leaf-evaluation network: shared declarations of the CUDA translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "host.h"

#define AZG_NET_FRONT 32          // zero rows in front of the first board of an activation buffer
#define AZG_NET_BACK 32           // zero rows behind the last board
#define AZG_NET_MAX_BLOCKS 40
#define AZG_HEAD_PITCH 704         // floats per board of head features: 450 policy (ch*225+pixel), 225 value, zero pad
#define AZG_HEAD_WP_K 480          // K extent of the padded policy_fc matrix [240][480] (450 real columns)
#define AZG_HEAD_WV_K 256          // K extent of the padded value_fc1 matrix [64][256]: column j is feature 448 + j

struct ConvArgs {
  const int* n_boards;            // device: positions in this batch
  int max_boards;                 // capacity of the activation buffers
  int layer;                      // index into the packed 3x3 weights (tap-major blocks of C x C)
  int relu;
  const float* shift_host;        // HOST [C]: folded BatchNorm shift of this layer (the scale lives in the weights)
  const __nv_bfloat16* residual;  // padded activation buffer added before the ReLU, or null
  __nv_bfloat16* out;             // padded activation buffer (null: do not store, fused-heads layer only)
  const float* head_host;         // HOST [3*C + 6]: fused 1x1 head weights, BN scale[3], shift[3]; null = plain layer
  float* hidden;                  // head features [board][AZG_HEAD_PITCH] (fused-heads layer only)
  int* error;                     // device flag set by the pipeline watchdogs
  unsigned long long* prof;       // optional device counters [16] (cycles spent waiting per role), or null
  int prof_detail;                // also clock the phases of one epilogue warp (slightly intrusive)
  float* stat_partial;            // training forward: [2 * C][AZG_CONV_STAT_SLOTS] per-channel sum / sum of squares of the
                                  // outputs, one slot per epilogue warp (azg_conv3x3_stat_slots of them are written); null = off
  const __nv_bfloat16* bwd_z;     // training, input gradient: with stat_partial, z (activation layout) and the ReLU bit mask
  const uint8_t* bwd_mask;        // ([boards * 256][C / 8]) of the layer whose dL/da this launch produces; the two statistics are
                                  // then sum(dy) and sum(dy * z), dy = out * [a > 0]; null = forward statistics
};
#define AZG_CONV_STAT_SLOTS 1184   // 74 clusters x 2 CTAs x 4 quadrants x 2 epilogue groups
int azg_conv3x3_stat_slots(int max_boards, int n_sm, int C);

// mode: activation staging variant of the kernel (see net_conv.cu); tm_act must have been encoded
// with box rows = azg_conv3x3_rows(mode).
int azg_conv3x3_rows(int mode);
// tm_out: box {32 channels, 32 rows}, SWIZZLE_64B, over the output buffer (epilogue TMA stores).
int azg_conv3x3_launch(int C, int mode, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const CUtensorMap& tm_out,
                       const ConvArgs& args, int n_sm, cudaStream_t stream);

// net_aux.cu
struct StemArgs {
  const int* n_boards; int max_boards;
  const uint32_t* keys;           // leaf stones: [n][16] words (player-1 words 0-7, player-2 words 8-15) ...
  const uint32_t* meta;           // ... and the side to move in bits 1-2 of meta (tree node layout), or
  const int32_t* leaf_game;       // leaf -> (game, node) indirection into the engine slabs (null: keys are dense [n][16], meta [n])
  const int32_t* leaf_node;
  long long slab_stride;          // nodes per game in the slab
  const float* w;                 // [27][C] folded stem weights (tap-major, plane, then channel)
  const float* shift;             // [C]
  __nv_bfloat16* out;
};
int azg_stem_launch(int C, const StemArgs& a, int n_sm, cudaStream_t stream);
int azg_planes_to_keys_launch(const float* planes, int n, uint32_t* keys, uint32_t* meta, cudaStream_t stream);

struct HeadArgs {
  const int* n_boards; int max_boards;
  const __nv_bfloat16* act;       // final trunk activations (padded layout)
  const float* w1;                // [3][C]: policy conv rows 0-1, value conv row 2 (1x1)
  const float* scale1;            // [3] folded BN of the two head convs
  const float* shift1;            // [3]
  float* hidden;                  // [boards][AZG_HEAD_PITCH]: 450 policy features (ch*225+pixel), 225 value features, zeros
  const float* pol_b;             // [240] policy_fc.bias, zero padded
  const float* v1_b;              // [64]
  const float* v2_w;              // [64]
  const float* v2_b;              // [1]
  float* probs;                   // [n][225]
  float* values;                  // [n] (may be null)
  float* logits;                  // [n][225] optional raw logits (may be null)
};
// 1x1 head convs (unless already fused into the last trunk layer) + the dense part on tensor cores (net_heads.cu).
// tm_hid: fp32 [rows][AZG_HEAD_PITCH], box {32, 128}; tm_wp: fp32 [240][AZG_HEAD_WP_K], box {32, 240};
// tm_wv: fp32 [64][AZG_HEAD_WV_K], box {32, 64}; all SWIZZLE_128B.
int azg_heads_launch(int C, const HeadArgs& a, const CUtensorMap& tm_hid, const CUtensorMap& tm_wp, const CUtensorMap& tm_wv,
                     int* error, int n_sm, cudaStream_t stream, bool skip_head1);
int azg_heads_gemm_launch(const CUtensorMap& tm_hid, const CUtensorMap& tm_wp, const CUtensorMap& tm_wv, const HeadArgs& a,
                          int* error, int n_sm, cudaStream_t stream);

struct PackArgs {
  int C, n_blocks;
};

// net_wgrad.cu: weight gradient of one 3x3 layer, dW[tap][co][ci] (fp32, accumulated with reductions: zero it first).
// tm_dz: box {64 channels, 64 rows} over the output-gradient buffer; tm_a: box {64 channels, azg_wgrad3x3_a_rows()}
// over the layer's input activations; both in the padded activation layout, SWIZZLE_128B.
struct WgradArgs {
  int n_boards;
  float* dw;                      // [9][C][C]
  int* error;                     // watchdog flag
  int desc_variant;               // experiment switch for the MN-major descriptor fields (0 = documented layout)
  int ct;                         // channels of the TENSORS when wider than the kernel's tile (256: four 128 x 128 quadrants); 0 = tile width
};
int azg_wgrad3x3_a_rows();
int azg_wgrad3x3_launch(int C, const CUtensorMap& tm_dz, const CUtensorMap& tm_a, const WgradArgs& a, int n_sm, cudaStream_t stream);
// cluster-of-three variant (multicast operands): tm = {dz box 24 rows, dz box 16 rows, a box 32 rows, a box 34 rows}
int azg_wgrad3x3_cluster_launch(int C, const CUtensorMap* tm, const WgradArgs& a, int n_sm, cudaStream_t stream);

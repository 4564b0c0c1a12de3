// Training step (PyTorchModel.train_batch, network.py:199-235): declarations shared by train_kernels.cu
// (BatchNorm / heads / stem / optimiser kernels) and train_engine.cu (buffers, schedule, C ABI azg_train_*).
#pragma once
#include "net.h"

#define AZG_TRAIN_PARTIALS 320        // capacity of partial[output][block]; the reduction kernels launch azg_train_red_blocks() <= this
int azg_train_red_blocks();           // default: one block per SM; AZG_TRAIN_RED_BLOCKS overrides (experiment switch)

// ---- trunk BatchNorm, training mode (batch statistics over boards x 225 pixels) -------------------------------------
struct BnStatsArgs {
  const __nv_bfloat16* z;             // conv output, padded layout (pad rows are zero)
  int n_boards;
  float* partial;                     // [2][C][AZG_TRAIN_PARTIALS] scratch
  unsigned* counter;                  // last-block ticket (zero before the launch, zero again after it)
  float* stats;                       // out [2][C]: batch mean, 1/sqrt(biased var + eps)
  float* running_mean;                // module buffers, updated as nn.BatchNorm2d does (momentum, unbiased variance)
  float* running_var;
  float momentum, eps;
};
int azg_bn_stats_launch(int C, const BnStatsArgs& a, cudaStream_t s);
// statistics already summed per epilogue warp by the convolution (ConvArgs::stat_partial = a.partial, n_slots valid slots)
int azg_bn_finalize_launch(int C, const BnStatsArgs& a, int n_slots, cudaStream_t s);

struct BnApplyArgs {
  const __nv_bfloat16* z;
  const float* stats;                 // [2][C]
  const float* gamma;                 // bn.weight
  const float* beta;                  // bn.bias
  const __nv_bfloat16* residual;      // block input added before the ReLU (second layer of a block), or null
  __nv_bfloat16* out;                 // a = relu(bn(z) (+ residual)), pad rows zero
  int n_boards;
  uint8_t* mask;                      // [n_boards * 256][C / 8]: bit i of a byte = [a > 0] of channel 8 * group + i (what the backward pass reads
                                      // instead of a: 1/16 of the bytes), or null
};
int azg_bn_apply_launch(int C, const BnApplyArgs& a, int n_sm, cudaStream_t s);

struct BnBwdArgs {
  const __nv_bfloat16* g;             // dL/da (gradient of the layer output, after the ReLU)
  const uint8_t* mask;                // ReLU mask of the layer output, one bit per element (bn_apply wrote it)
  const __nv_bfloat16* z;             // conv output (for x_hat)
  const float* stats;                 // [2][C] batch mean, rstd
  const float* gamma;
  int n_boards;
  float* partial;                     // [2 * C][AZG_TRAIN_PARTIALS]
  unsigned* counter;
  float* sums;                        // [2][C]: sum dy, sum dy * x_hat   (reduce writes, apply reads)
  float* dgamma;                      // gradient outputs (reduce writes)
  float* dbeta;
  __nv_bfloat16* dz;                  // apply: dL/dz, pad rows zero
  __nv_bfloat16* gskip;               // apply: dy = g * [a > 0] (gradient of the skip connection), or null
};
int azg_bn_bwd_reduce_launch(int C, const BnBwdArgs& a, cudaStream_t s);
// the same sums from the statistics the input-gradient convolution left in conv_partial (ConvArgs.bwd_z): no pass over g and z
int azg_bn_bwd_finalize_launch(int C, const BnBwdArgs& a, const float* conv_partial, int n_slots, cudaStream_t s);
int azg_bn_bwd_apply_launch(int C, const BnBwdArgs& a, int n_sm, cudaStream_t s);

// ---- stem (conv 3 -> C on the input planes) ---------------------------------------------------------------------------
struct StemTrainArgs {
  const float* planes;                // [n][3][225] encoded states (train.py:287-293)
  const float* w;                     // conv.weight [C][3][3][3]
  int n_boards;
  __nv_bfloat16* z;                   // forward out: conv output, padded layout
  const __nv_bfloat16* dz;            // backward in
  float* partial;                     // [n_boards][27][C] scratch
  float* dw;                          // out [27][C] (plane*9 + tap major, channel minor), accumulated with atomics: zero it first
};
int azg_stem_train_fwd_launch(int C, const StemTrainArgs& a, cudaStream_t s);
int azg_stem_train_wgrad_launch(int C, const StemTrainArgs& a, cudaStream_t s);

// ---- heads + loss (network.py:102-117, 217-222) -----------------------------------------------------------------------
struct HeadTrainArgs {
  int C, n_boards;
  const __nv_bfloat16* act;           // final trunk activations a_L
  const float* w1p;                   // policy_conv.weight [2][C]
  const float* w1v;                   // value_conv.weight [1][C]
  const float* bn_gamma[2];           // policy_bn.weight [2], value_bn.weight [1]
  const float* bn_beta[2];
  float* bn_rmean[2];
  float* bn_rvar[2];
  float momentum, eps;
  const float* wp_t;                  // packed policy_fc.weight^T [450][225]
  const float* bp;                    // [225]
  const float* wv1_t;                 // packed value_fc1.weight^T [225][64]
  const float* bv1;                   // [64]
  const float* w2;                    // value_fc2.weight [64]
  const float* b2;                    // [1]
  const float* wp;                    // policy_fc.weight [225][450]  (backward data)
  const float* wv1;                   // value_fc1.weight [64][225]
  const float* pi;                    // targets [n][225]
  const float* zt;                    // targets [n]
  // activations / scratch
  float* zh;                          // [n][3][225] 1x1 conv outputs
  float* hstats;                      // [3][2] mean, rstd
  float* hidden;                      // [n][675] relu(bn(zh)): 450 policy features (ch*225+pix), 225 value features
  float* h1;                          // [n][64]
  float* value;                       // [n]
  float* dlogits;                     // [n][225]
  float* dvpre;                       // [n]
  float* loss_parts;                  // [n][2]: KL row sum, squared value error (caller sums / n)
  float* dhid;                        // [n][675] gradient at the head BatchNorm outputs, ReLU mask applied
  float* hsums;                       // [3][2]: sum dy, sum dy * x_hat of the head BatchNorms
  float* partial;                     // [3 * C][AZG_TRAIN_PARTIALS]
  unsigned* counter;
  __nv_bfloat16* g;                   // out: dL/da_L, padded layout
  // gradients (flat-buffer slices)
  float *d_w1p, *d_w1v, *d_bn_gamma[2], *d_bn_beta[2], *d_wp, *d_bp, *d_wv1, *d_bv1, *d_w2, *d_b2;
};
int azg_head_train_fwd_launch(const HeadTrainArgs& a, int n_sm, cudaStream_t s);
int azg_head_train_bwd_launch(const HeadTrainArgs& a, int n_sm, cudaStream_t s);

// ---- optimiser: clip_grad_norm_(3.0) + Adam (network.py:141, 224-225) + bf16 repack -------------------------------------
enum { SEG_PLAIN = 0, SEG_CONV3 = 1, SEG_STEM = 2, SEG_WP = 3, SEG_WV1 = 4 };
struct AdamSeg {
  long long off;                      // first element in the flat parameter (and gradient, and moment) vectors
  int count;
  int kind;                           // SEG_*: how parameter index maps to gradient index / packed copies
  int layer;                          // SEG_CONV3: trunk layer
};
struct AdamArgs {
  float* params; const float* grads; float* m; float* v;
  long long n;
  const AdamSeg* segs;                // device
  const int2* block_seg;              // device: per block {segment, first element within the segment}
  int n_blocks;
  int C;
  float lr, wd, b1, b2, eps, clip;
  float inv_world;                    // gradients hold a SUM over ranks: scale by 1/world
  float* norm_partial;                // [blocks]
  unsigned* counter;
  float* scal;                        // device [4]: clip coefficient (incl. 1/world), gradient norm, -, -
  long long* step;                    // device step counter (incremented by the norm kernel)
  __nv_bfloat16* wf;                  // packed forward conv weights  [(layer*9+tap)*C + co][ci]
  __nv_bfloat16* wb;                  // packed backward conv weights [(layer*9+(8-tap))*C + ci][co]
  float* wp_t;                        // [450][225]
  float* wv1_t;                       // [225][64]
};
int azg_grad_norm_launch(const AdamArgs& a, cudaStream_t s);
int azg_adam_launch(const AdamArgs& a, bool update, cudaStream_t s);     // update = false: only (re)pack the bf16 / transposed copies

/* azgomoku_b200 - C ABI of the B200-native AlphaZero Gomoku/Pente self-play engine.
 *
 * This is the drop-in boundary for the hot path of shirongcan/AlphaZero-Gomoku.
 * The reference is pure Python and has no FFI of its own; each entry point below
 * names the reference interface (file:line under the reference tree) whose work
 * it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer of the
 * reference would add.
 *
 * Conventions: every function returns 0 on success and a negative code on error
 * (azg_last_error() gives the text, per host thread); all `*_dev` / unmarked data
 * pointers are DEVICE memory owned by the caller unless the name ends in `_host`;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream); no hidden
 * global state; one host thread per engine.  There is no CPU fallback: every call
 * fails with AZG_E_CUDA when no sm_100 device is usable.
 */
#ifndef AZGOMOKU_B200_H
#define AZGOMOKU_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AZG_ABI_VERSION 3
#define AZG_BOARD 15
#define AZG_ACTIONS 225

enum { AZG_OK = 0, AZG_E_ARG = -1, AZG_E_CUDA = -2, AZG_E_NOMEM = -3, AZG_E_STATE = -4, AZG_E_SEARCH = -5 };
enum { AZG_RULE_GOMOKU = 0, AZG_RULE_PENTE = 1 };

/* Packed position, 96 bytes.  Bit (a & 31) of stones[c][a >> 5] is cell a = r*15+c of
 * colour c+1.  Mirrors the state of games/gomoku.py:20-25 and games/pente.py:12-23:
 * board, current_player, last_move (-1 == None), captures, len(move_history). */
typedef struct azg_pos {
  uint32_t stones[2][8];
  int32_t player;
  int32_t last;
  int32_t caps[2];
  int32_t plies;
  int32_t pad[3];
} azg_pos;

const char* azg_last_error(void);
int azg_abi_version(void);
/* Number of usable sm_100 devices (0 when none: every other call then fails). */
int azg_device_count(void);

/* ------------------------------------------------------------------ rules (batched, n positions)
 * status bits written by azg_rules_play / azg_rules_status:
 *   bits 0-1 winner (0/1/2), bit 2 game over, bit 3 move rejected (off board / occupied). */
#define AZG_STATUS_WINNER(s) ((s) & 3)
#define AZG_STATUS_OVER(s) (((s) >> 2) & 1)
#define AZG_STATUS_REJECTED(s) (((s) >> 3) & 1)

/* boards int8[n][225] (0/1/2) + scalars -> packed positions.  lasts/caps/plies may be NULL. */
int azg_rules_pack(const int8_t* boards, const int32_t* players, const int32_t* lasts, const int32_t* caps,
                   const int32_t* plies, azg_pos* out, int n, void* stream);
/* packed -> boards int8[n][225]; players/lasts/caps/plies outputs may be NULL. */
int azg_rules_unpack(const azg_pos* pos, int8_t* boards, int32_t* players, int32_t* lasts, int32_t* caps,
                     int32_t* plies, int n, void* stream);
/* Gomoku.do_move / Pente.do_move + _handle_captures (games/gomoku.py:60-78,
 * games/pente.py:57-79,114-152), then check_winner / is_game_over
 * (gomoku.py:155-197, pente.py:199-236).  actions[i] = r*15+c, or any value outside
 * 0..224 for an off-board move; rejected moves leave the position unchanged. */
int azg_rules_play(int rule, azg_pos* pos, const int32_t* actions, int32_t* status, int n, void* stream);
int azg_rules_status(int rule, const azg_pos* pos, int32_t* status, int n, void* stream);
/* get_valid_moves (gomoku.py:109-121): float32[n][225]. */
int azg_rules_legal(const azg_pos* pos, float* mask, int n, void* stream);
/* get_encoded_state (gomoku.py:130-150): float32[n][3][15][15]. */
int azg_rules_encode(const azg_pos* pos, float* planes, int n, void* stream);
/* Host-buffer form of azg_rules_play used by the Python game shims: boards int8[n][225],
 * players/lasts/caps[n][2]/plies are updated in place; status_host receives the bits. */
int azg_rules_play_host(int rule, int device, int8_t* boards_host, int32_t* players_host, int32_t* lasts_host,
                        int32_t* caps_host, int32_t* plies_host, const int32_t* actions_host,
                        int32_t* status_host, int n);

/* Host-buffer query (nothing is modified): status bits, and optionally get_valid_moves
 * (legal_host float32[n][225]) and get_encoded_state (planes_host float32[n][3][15][15]). */
int azg_rules_query_host(int rule, int device, const int8_t* boards_host, const int32_t* players_host,
                         const int32_t* lasts_host, const int32_t* caps_host, const int32_t* plies_host,
                         int32_t* status_host, float* legal_host, float* planes_host, int n);

/* ------------------------------------------------------------------ search engine
 * One engine = G concurrent games of one rule on one device, each with its own
 * HBM-resident tree slab.  Replaces MCTS.__init__/run/search/_predict_batch/clear_tree
 * (mcts/new_mcts_alpha.py:12-37, 58-72, 77-185) for G games at once. */
typedef struct azg_engine azg_engine;

typedef struct azg_config {
  int32_t device;          /* CUDA ordinal */
  int32_t rule;            /* AZG_RULE_* (game_class) */
  int32_t n_games;         /* G */
  int32_t queue_len;       /* reference batch_size, 1..256 (default 32) */
  int32_t node_capacity;   /* nodes per game slab */
  int32_t noise_on;        /* add_dirichlet_noise */
  int32_t noise_plies;     /* apply_dirichlet_n_first_moves */
  int32_t game_base;       /* global id of game 0 (rank * G): on-device RNG streams are keyed by the global id, so
                              results do not depend on how games are sharded over GPUs */
  double cpuct;            /* cpuct */
  double alpha;            /* dirichlet_alpha */
  double eps;              /* epsilon */
  uint64_t seed;           /* Philox key for on-device noise / sampling */
  int32_t fast_warps;      /* 0 (default): the reference's strictly sequential simulations, exact visit counts.
                              1..16: NON-PARITY fast mode - that many warps walk one game's tree concurrently, kept apart
                              by a virtual loss (north-star subsystem 1); visit counts differ from the reference's */
  int32_t virtual_loss;    /* fast mode: N += vl, W -= vl on every edge of a simulation in flight (default 1) */
} azg_config;

int azg_create(const azg_config* cfg, azg_engine** out);
int azg_destroy(azg_engine* e);
int azg_set_stream(azg_engine* e, void* stream);
/* Bytes of device memory held by the engine. */
int64_t azg_memory_bytes(const azg_engine* e);

/* Load root positions (roots[g] for every g with mask[g] != 0; mask NULL = all) and, when
 * clear_tree != 0, forget those games' trees (MCTS.clear_tree, new_mcts_alpha.py:58-72). */
int azg_set_roots(azg_engine* e, const azg_pos* roots, const int32_t* mask, int clear_tree);
int azg_get_roots(azg_engine* e, azg_pos* roots_out);

/* MCTS.run, first half (new_mcts_alpha.py:77-83): fix the root keys and the budget.
 * plies[g] is the reference's move_number (NULL = the root's own ply count). */
int azg_search_begin(azg_engine* e, const int32_t* plies, int n_sims);
/* As above, but games with mask[g] == 0 sit this run out (their result rows are zero).  Used by the
 * evaluation arena (train.py:418-486), where two models alternate on the same set of games. */
int azg_search_begin_masked(azg_engine* e, const int32_t* plies, int n_sims, const int32_t* mask);
/* Run simulations on every game until its leaf queue is full or its budget is spent
 * (MCTS.search, new_mcts_alpha.py:102-151), then assemble the leaf batch.
 * The outputs (each may be NULL) receive the batch size, the number of games that need
 * another fill after this batch is committed, and the number of games whose root is
 * in the batch (the only case in which the reference draws Dirichlet noise,
 * new_mcts_alpha.py:171); passing all NULL skips the host synchronisation. */
int azg_search_fill(azg_engine* e, int32_t* n_leaves_host, int32_t* n_active_host, int32_t* n_roots_host);
/* The three outputs of the last azg_search_fill that was issued with NULL outputs (asynchronously):
 * synchronises the engine's stream only, so another engine's work on another stream keeps running. */
int azg_search_read_counters(azg_engine* e, int32_t* n_leaves_host, int32_t* n_active_host, int32_t* n_roots_host);
/* Device address of int32 {n_leaves, n_active, n_errors, n_roots, ...} for on-device consumers. */
const int32_t* azg_search_counters(const azg_engine* e);
/* Encoded planes of the current leaf batch, float32[n_leaves][3][15][15] - what the
 * reference stacks into X at new_mcts_alpha.py:160. */
int azg_search_leaf_planes(azg_engine* e, float* planes);
/* MCTS._predict_batch after the network call (new_mcts_alpha.py:163-185): probs is the
 * unmasked softmax float32[n_leaves][225] in leaf order.  noise (may be NULL) is
 * float64[G][225]: the Dirichlet draw used if game g's root is in its queue. */
int azg_search_commit(azg_engine* e, const float* probs, const double* noise);
/* MCTS.run, second half (new_mcts_alpha.py:88-97): pi float32[G][225] and the raw root
 * visit counts int32[G][225] (either may be NULL). */
int azg_search_result(azg_engine* e, float* pi, int32_t* visits);
/* Play actions[g] (>= 0) on game g's root, report azg_rules status bits, and with
 * gc != 0 reclaim tree nodes that can no longer be reached (DESIGN.md "tree memory").
 * reserve > 0: a game whose slab has fewer than `reserve` free nodes after the sweep drops
 * its whole tree (counted in azg_search_stats) instead of overflowing during the next run. */
int azg_search_advance(azg_engine* e, const int32_t* actions, int gc, int reserve, int32_t* status);
/* out_host[0..7] = completed sims, node visits, leaf evaluations, live nodes (sum),
 * high-water nodes (max over games), games in error, OR of error bits, dropped trees. */
int azg_search_stats(azg_engine* e, uint64_t* out_host);

/* ------------------------------------------------------------------ self-play driver
 * play_game_and_collect for G games at once (train.py:360-412) with on-device Philox. */
/* Allocate example capture for games of up to max_plies moves (train.py max_moves). */
int azg_selfplay_enable(azg_engine* e, int max_plies);
/* Optional slot mask int32[G] (device memory owned by the caller, read by every later azg_selfplay_choose /
 * azg_selfplay_finish; NULL = all slots play).  A slot with active[g] == 0 is retired: it chooses no move
 * (actions[g] = -1), records no example and never reports done.  This is how a fixed number of games is
 * played to completion (train.py:671-694 plays exactly games_per_iteration games). */
int azg_selfplay_set_active(azg_engine* e, const int32_t* active);
/* noise float64[G][225] ~ Dirichlet(alpha) over all 225 actions (new_mcts_alpha.py:172);
 * The Philox stream is keyed by (global game id, action, `draw`, games finished in this slot, ply of
 * the current game); `draw` is an extra caller-chosen index and may stay constant. */
int azg_selfplay_noise(azg_engine* e, uint64_t draw, double* noise);
/* sample_action_from_pi with temperature max(0, 1 - ply/temp_threshold) (train.py:252-266,
 * 647-648), illegal-pick fallback to argmax (train.py:380-382); records (position, pi) as the
 * ply's example (train.py:384).  pi float32[G][225] -> actions int32[G]. */
int azg_selfplay_choose(azg_engine* e, const float* pi, float temp_threshold, uint64_t draw, int32_t* actions);
/* After azg_search_advance: games that are over (status bit 2) or reached max_moves plies get
 * their examples labelled with z and expanded by the 8 symmetries (train.py:392-410,
 * new_mcts_alpha.py:42-56) into out[row][901] = planes[3][225], pi[225], z, rows reserved at
 * *cursor (device counter; rows >= capacity are dropped).  done_mask int32[G] marks the games
 * to restart; winners int32[G] (may be NULL) receives 0/1/2 or -1 for unfinished games. */
int azg_selfplay_finish(azg_engine* e, const int32_t* status, int max_moves, int use_symmetries, float* out,
                        int64_t capacity, uint64_t* cursor, int32_t* done_mask, int32_t* winners);

/* The same, but one PACKED row per ply instead of 8 expanded ones: out uint32[capacity][AZG_PACKED_WORDS] =
 * {stones[2][8], side to move, z (float bits), pi[225] (float bits), 0}.  976 bytes per ply against 8 x 3604: this
 * is what the ranks exchange (train.py:737-742 pickles the expanded rows through a pipe). */
#define AZG_PACKED_WORDS 244
int azg_selfplay_finish_packed(azg_engine* e, const int32_t* status, int max_moves, uint32_t* out, int64_t capacity,
                               uint64_t* cursor, int32_t* done_mask, int32_t* winners);
/* Packed plies -> example rows float32[n * (8 or 1)][901] in the reference's symmetry order
 * (train.py:405-410, new_mcts_alpha.py:42-56), on the current device. */
int azg_examples_expand(const uint32_t* packed, int64_t n, int use_symmetries, float* out, void* stream);

/* ------------------------------------------------------------------ leaf evaluator (policy/value ResNet)
 * Replaces AlphaZeroNet.forward + PyTorchModel.predict (network.py:85-117, 168-183): stem conv,
 * n_blocks residual blocks of two 3x3 convs (tcgen05 implicit GEMM, bf16 in / fp32 accumulate,
 * eval-mode BatchNorm folded into the epilogue), the 1x1 head convolutions fused into the last layer, the
 * dense head layers as tcgen05 TF32 GEMMs with softmax / tanh in their epilogue.  Tolerance against the fp32
 * reference forward is stated and tested in tests/test_net_gpu.py. */
typedef struct azg_net azg_net;
#define AZG_NET_MAX_LAYERS 80

/* Device pointers to float32 tensors in the reference's state_dict layout (network.py:54-71).
 * bn arrays are {weight, bias, running_mean, running_var}. */
typedef struct azg_net_weights {
  const float* conv_w;                          /* conv.weight [C,3,3,3] */
  const float* bn[4];                           /* bn.* */
  const float* res_conv_w[AZG_NET_MAX_LAYERS];  /* res_blocks.i.conv1.weight, res_blocks.i.conv2.weight, ... [C,C,3,3] */
  const float* res_bn[AZG_NET_MAX_LAYERS][4];   /* res_blocks.i.bn1.*, res_blocks.i.bn2.*, ... */
  const float* policy_conv_w;                   /* [2,C,1,1] */
  const float* policy_bn[4];
  const float* policy_fc_w;                     /* [225,450] */
  const float* policy_fc_b;                     /* [225] */
  const float* value_conv_w;                    /* [1,C,1,1] */
  const float* value_bn[4];
  const float* value_fc1_w;                     /* [64,225] */
  const float* value_fc1_b;                     /* [64] */
  const float* value_fc2_w;                     /* [1,64] */
  const float* value_fc2_b;                     /* [1] */
} azg_net_weights;

/* channels: 64, 128 or 256; max_batch: positions per forward pass the activation buffers hold. */
int azg_net_create(int device, int n_blocks, int channels, int max_batch, azg_net** out);
int azg_net_destroy(azg_net* n);
int64_t azg_net_memory_bytes(const azg_net* n);
/* Repack weights (fold BatchNorm, bf16 tap-major conv weights, transposed FC weights). */
int azg_net_load(azg_net* n, const azg_net_weights* w, void* stream);
/* PyTorchModel.predict on device buffers: planes float32[count][3][15][15] ->
 * probs float32[count][225] (softmax over all 225), values float32[count] (may be NULL),
 * logits float32[count][225] (may be NULL). */
int azg_net_forward_planes(azg_net* n, const float* planes, int count, float* probs, float* values, float* logits,
                           void* stream);
/* Evaluate the engine's current leaf batch (azg_search_fill) without leaving the device:
 * probs float32[n_leaves][225] in leaf order, values float32[n_leaves] (may be NULL). */
int azg_net_forward_leaves(azg_net* n, azg_engine* e, float* probs, float* values);
/* Test hook: activations after the stem and the first n_layers 3x3 layers, float32[count][C][15][15]. */
int azg_net_trunk_debug(azg_net* n, const float* planes, int count, int n_layers, float* out, void* stream);
/* Measurement hooks used by bench.py: time the 3x3 trunk of every forward pass with CUDA events
 * on the launch stream; read returns the summed milliseconds and the conv3x3 launches covered. */
int azg_net_profile(azg_net* n, int enable);
int azg_net_profile_read(azg_net* n, double* trunk_ms, int64_t* launches);
/* Cycle counters of the conv3x3 pipeline roles collected while profiling (see net_engine.cu). */
int azg_net_profile_counters(azg_net* n, uint64_t* out32);
/* Synchronise and report the tcgen05 pipeline watchdog (0 = healthy). */
int azg_net_check(azg_net* n, void* stream);

/* ------------------------------------------------------------------ training step
 * Replaces PyTorchModel.train_batch (network.py:199-235): training-mode forward (BatchNorm2d batch statistics,
 * running statistics updated with momentum 0.1), loss = KLDivLoss(batchmean)(log_softmax(logits), pi) +
 * MSELoss(value, z), backward, clip_grad_norm_(3.0), Adam(lr, weight_decay).  3x3 convolutions forward /
 * input gradient / weight gradient run on tcgen05 tensor cores (bf16 operands, fp32 accumulation); master
 * weights, gradients and Adam moments are fp32.  Tolerance against the fp32 reference step is stated and
 * tested in tests/test_train_gpu.py.
 * The parameter vector is FLAT, in net.parameters() order (network.py:47-73): conv.weight, bn.weight, bn.bias,
 * per residual block conv1.weight, bn1.weight, bn1.bias, conv2.weight, bn2.weight, bn2.bias, then
 * policy_conv.weight, policy_bn.{weight,bias}, policy_fc.{weight,bias}, value_conv.weight,
 * value_bn.{weight,bias}, value_fc1.{weight,bias}, value_fc2.{weight,bias}. */
typedef struct azg_train azg_train;
typedef struct azg_train_config {
  int32_t device;
  int32_t n_blocks;
  int32_t channels;        /* 64, 128 or 256 */
  int32_t max_batch;       /* positions per step the activation buffers hold */
  double lr;               /* Adam, network.py:141 defaults: 1e-3 */
  double weight_decay;     /* 1e-4, added to the gradient (torch.optim.Adam) */
  double beta1, beta2, eps;/* 0.9, 0.999, 1e-8 */
  double clip;             /* clip_grad_norm_ max_norm, network.py:224: 3.0 */
  double bn_momentum;      /* 0.1 */
  double bn_eps;           /* 1e-5 */
} azg_train_config;

int azg_train_create(const azg_train_config* cfg, azg_train** out);
int azg_train_destroy(azg_train* t);
/* Number of floats of the flat parameter / gradient / moment vectors. */
int64_t azg_train_param_count(const azg_train* t);
int64_t azg_train_memory_bytes(const azg_train* t);
/* Attach the caller-owned flat device vectors (float32[param_count] each) and the BatchNorm running statistics
 * (only the running_mean / running_var entries of `stats` are read; they are UPDATED by every step).
 * `step` is the number of Adam steps already taken (bias correction). */
int azg_train_bind(azg_train* t, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                   const azg_net_weights* stats, int64_t step, void* stream);
/* Refresh the bf16 / transposed weight copies after the parameters were changed from outside. */
int azg_train_pack(azg_train* t, void* stream);
/* network.py:210-223: planes float32[count][3][15][15], pis float32[count][225], zs float32[count] ->
 * gradients of the summed loss in the bound gradient vector; loss_parts float32[count][2] = per position
 * {KL row sum, squared value error} (policy_loss = sum / count, value_loss = sum / count). */
int azg_train_forward_backward(azg_train* t, const float* planes, const float* pis, const float* zs, int count,
                               float* loss_parts, void* stream);
/* network.py:224-225: clip_grad_norm_ + Adam.step on the bound vectors.  world > 1: the gradient vector holds
 * the SUM over `world` ranks (all-reduced by the caller between the two calls) and is averaged first. */
int azg_train_apply(azg_train* t, int world, void* stream);
/* Synchronise; fails if a tcgen05 pipeline watchdog fired.  out_host (may be NULL) receives
 * {gradient norm before clipping, clip coefficient, Adam steps taken}. */
int azg_train_check(azg_train* t, double* out_host, void* stream);
/* Test hooks: activations of the last forward_backward as float32[count][C][15][15] (what: 0 = after
 * BatchNorm+ReLU, 1 = convolution output; layer 0 = stem ... 2*n_blocks; 2 = gradient at the stem output,
 * 3 = last dL/dz), and the gradient vector rearranged to the parameter layout (what p.grad would hold). */
int azg_train_read_activation(azg_train* t, int what, int layer, float* out, void* stream);
int azg_train_export_grads(azg_train* t, float* out, void* stream);
/* The two tensor-core gradient kernels in isolation (test hook): dz, a float32[count][C][15][15] (rounded to bf16),
 * current weights of trunk layer `layer` -> dw_out float32[C][C][3][3] (weight gradient), da_out
 * float32[count][C][15][15] (input gradient, bf16-rounded).  Overwrites activation scratch only. */
int azg_train_debug_conv_grads(azg_train* t, const float* dz, const float* a, int count, int layer, float* dw_out,
                               float* da_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif

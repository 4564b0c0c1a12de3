// Host side of the search engine: slab allocation, kernel launches, C ABI.
// Replaces the object life-cycle of MCTS (mcts/new_mcts_alpha.py:12-37, 58-97).
#include <new>
#include <string>
#include <vector>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.cuh"
#include "host.h"
#include "engine.h"

static thread_local std::string g_err;

int azg_fail(int code, const char* msg) {
  g_err = msg ? msg : "unknown error";
  return code;
}

int azg_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return AZG_E_CUDA;
  }
  return AZG_OK;
}

extern "C" const char* azg_last_error(void) { return g_err.c_str(); }
extern "C" int azg_abi_version(void) { return AZG_ABI_VERSION; }

extern "C" int azg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

template <typename T>
static int dev_alloc(T** p, size_t count, int64_t* total) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) { cudaGetLastError(); return azg_fail(AZG_E_NOMEM, "cudaMalloc failed (engine slab)"); }
  *total += (int64_t)(count * sizeof(T));
  return AZG_OK;
}

static inline dim3 warp_grid(int n) { return dim3((unsigned)((n + 3) / 4)); }

extern "C" int azg_create(const azg_config* cfg, azg_engine** out) {
  if (!cfg || !out) return azg_fail(AZG_E_ARG, "azg_create: null argument");
  if (cfg->n_games < 1 || cfg->queue_len < 1 || cfg->queue_len > AZG_MAX_QUEUE || cfg->node_capacity < 64 ||
      cfg->node_capacity >= (1 << 24) || (cfg->rule != 0 && cfg->rule != 1))
    return azg_fail(AZG_E_ARG, "azg_create: n_games>=1, 1<=queue_len<=256, 64<=node_capacity<2^24, rule in {0,1}");
  if (cfg->fast_warps < 0 || cfg->fast_warps > 16 || cfg->virtual_loss < 0)
    return azg_fail(AZG_E_ARG, "azg_create: fast_warps must be 0 (exact) .. 16, virtual_loss >= 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return azg_fail(AZG_E_CUDA, "azg_create: no such CUDA device (this engine has no CPU fallback)");
  }
  AZG_CUDA(cudaSetDevice(cfg->device));
  azg_engine* e = new (std::nothrow) azg_engine();
  if (!e) return azg_fail(AZG_E_NOMEM, "azg_create: host allocation failed");
  e->cfg = *cfg;
  azg_dev& d = e->dev;
  d.G = cfg->n_games; d.rule = cfg->rule; d.queue_len = cfg->queue_len; d.cap = cfg->node_capacity;
  int h = 64;
  while (h < 2 * cfg->node_capacity) h <<= 1;
  d.hcap = h;
  d.noise_on = cfg->noise_on; d.noise_plies = cfg->noise_plies; d.n_sims = 0; d.game_base = cfg->game_base;
  d.fast_warps = cfg->fast_warps; d.virtual_loss = cfg->virtual_loss > 0 ? cfg->virtual_loss : 1;
  { const char* v = getenv("AZG_FILL_L1"); d.fill_l1 = v ? atoi(v) : 1; }        // experiment switch, see tree.cu ldx
  d.cpuct = (float)cfg->cpuct; d.cpuct64 = cfg->cpuct; d.eps = cfg->eps; d.alpha = cfg->alpha; d.seed = cfg->seed;
  const size_t G = d.G, C = d.cap;
  int rc = AZG_OK;
  int64_t& tot = e->bytes;
  {
    // Gomoku: per-node child codes (see tree.cu): a re-visited child costs one load, no win test, hash or table probe.
    // A recorded experiment, OFF by default (AZG_CHILD_CODES=1 enables it): exact, but measured SLOWER - single game
    // 160 -> 184 ms per 5 000-simulation move (6x128), batched step -0.4 %, +912 bytes per node (95 -> 126 GB at 2 048
    // x 16 384 nodes).  Most descents end after a few levels at a NEW child (the reference resets N/W on every flush, so
    // placeholder children of equal prior are walked round-robin), where the code is unknown and only costs: one more
    // 32-byte load per lane and node, a shuffle, a zeroed row per new node and a store per new edge.
    const char* v = getenv("AZG_CHILD_CODES");
    const bool want = d.rule == AZG_RULE_GOMOKU && (v ? atoi(v) != 0 : false);
    d.child = nullptr;
    if (want && (rc = dev_alloc(&d.child, G * C * AZG_ROW, &tot))) { azg_destroy(e); return rc; }
  }
  if ((rc = dev_alloc(&d.ctl, G, &tot)) || (rc = dev_alloc(&d.P, G * C * AZG_ROW, &tot)) ||
      (rc = dev_alloc(&d.Nv, G * C * AZG_ROW, &tot)) || (rc = dev_alloc(&d.W, G * C * AZG_ROW, &tot)) ||
      (rc = dev_alloc(&d.key, G * C * 16, &tot)) || (rc = dev_alloc(&d.meta, G * C, &tot)) ||
      (rc = dev_alloc(&d.slots, G * (size_t)d.hcap, &tot)) || (rc = dev_alloc(&d.freelist, G * C, &tot)) ||
      (rc = dev_alloc(&d.path, G * AZG_MAX_DEPTH, &tot)) || (rc = dev_alloc(&d.P64, G * AZG_P64_SLOTS * AZG_ROW, &tot)) ||
      (rc = dev_alloc(&d.leaf_game, G * (size_t)d.queue_len, &tot)) ||
      (rc = dev_alloc(&d.leaf_node, G * (size_t)d.queue_len, &tot)) || (rc = dev_alloc(&d.counters, (size_t)8, &tot)) ||
      (rc = dev_alloc(&e->stats_dev, (size_t)8, &tot))) {
    azg_destroy(e);
    return rc;
  }
  if (cudaMallocHost((void**)&e->pinned, 64 * sizeof(uint64_t)) != cudaSuccess) {
    azg_destroy(e);
    return azg_fail(AZG_E_NOMEM, "azg_create: pinned allocation failed");
  }
  cudaMemset(d.ctl, 0, G * sizeof(azg_ctl));
  cudaMemset(d.meta, 0, G * C * sizeof(uint32_t));
  cudaMemset(d.slots, 0, G * (size_t)d.hcap * sizeof(unsigned long long));
  cudaMemset(d.counters, 0, 8 * sizeof(int32_t));
  // empty boards, player 1 to move
  {
    std::vector<azg_pos> init(G);
    for (size_t g = 0; g < G; ++g) { init[g] = azg_pos{}; init[g].player = 1; init[g].last = -1; }
    azg_pos* tmp = nullptr;
    if (cudaMalloc((void**)&tmp, G * sizeof(azg_pos)) != cudaSuccess) { azg_destroy(e); return azg_fail(AZG_E_NOMEM, "cudaMalloc failed"); }
    cudaMemcpy(tmp, init.data(), G * sizeof(azg_pos), cudaMemcpyHostToDevice);
    azg_reset_kernel<<<warp_grid(d.G), 128>>>(d, nullptr, tmp, 1);
    cudaDeviceSynchronize();
    cudaFree(tmp);
  }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { azg_destroy(e); return azg_fail(AZG_E_CUDA, cudaGetErrorString(ce)); }
  *out = e;
  return AZG_OK;
}

extern "C" int azg_destroy(azg_engine* e) {
  if (!e) return AZG_OK;
  cudaSetDevice(e->cfg.device);
  azg_dev& d = e->dev;
  cudaFree(d.ctl); cudaFree(d.P); cudaFree(d.Nv); cudaFree(d.W); cudaFree(d.child); cudaFree(d.key); cudaFree(d.meta); cudaFree(d.slots);
  cudaFree(d.freelist); cudaFree(d.path); cudaFree(d.P64); cudaFree(d.leaf_game); cudaFree(d.leaf_node);
  cudaFree(d.counters); cudaFree(e->stats_dev);
  cudaFree(e->sp.ex_key); cudaFree(e->sp.ex_player); cudaFree(e->sp.ex_pi); cudaFree(e->sp.n_plies); cudaFree(e->sp.n_done);
  if (e->pinned) cudaFreeHost(e->pinned);
  delete e;
  return AZG_OK;
}

extern "C" int azg_set_stream(azg_engine* e, void* stream) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  e->stream = (cudaStream_t)stream;
  return AZG_OK;
}

extern "C" int64_t azg_memory_bytes(const azg_engine* e) { return e ? e->bytes : 0; }

extern "C" int azg_set_roots(azg_engine* e, const azg_pos* roots, const int32_t* mask, int clear_tree) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  AZG_USE_DEVICE(e->cfg.device);
  azg_reset_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, mask, roots, clear_tree);
  return azg_check_launch("azg_set_roots");
}

extern "C" int azg_get_roots(azg_engine* e, azg_pos* roots_out) {
  if (!e || !roots_out) return azg_fail(AZG_E_ARG, "null argument");
  AZG_USE_DEVICE(e->cfg.device);
  AZG_CUDA(cudaMemcpy2DAsync(roots_out, sizeof(azg_pos), &e->dev.ctl[0].root, sizeof(azg_ctl), sizeof(azg_pos), e->dev.G,
                             cudaMemcpyDeviceToDevice, e->stream));
  return AZG_OK;
}

extern "C" int azg_search_begin_masked(azg_engine* e, const int32_t* plies, int n_sims, const int32_t* mask) {
  if (!e || n_sims < 0) return azg_fail(AZG_E_ARG, "azg_search_begin: bad argument");
  AZG_USE_DEVICE(e->cfg.device);
  e->dev.n_sims = n_sims;
  azg_begin_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, plies, n_sims, mask);
  return azg_check_launch("azg_search_begin");
}

extern "C" int azg_search_begin(azg_engine* e, const int32_t* plies, int n_sims) {
  return azg_search_begin_masked(e, plies, n_sims, nullptr);
}

extern "C" int azg_search_fill(azg_engine* e, int32_t* n_leaves_host, int32_t* n_active_host, int32_t* n_roots_host) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  AZG_USE_DEVICE(e->cfg.device);
  if (e->dev.fast_warps > 0) azg_fill_fast_kernel<<<e->dev.G, 32 * e->dev.fast_warps, 0, e->stream>>>(e->dev);      // non-parity option
  else azg_fill_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev);
  azg_scan_kernel<<<1, 1024, 0, e->stream>>>(e->dev);
  int rc = azg_check_launch("azg_search_fill");
  if (rc) return rc;
  if (n_leaves_host || n_active_host || n_roots_host) {
    int32_t* h = (int32_t*)e->pinned;
    AZG_CUDA(cudaMemcpyAsync(h, e->dev.counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    AZG_CUDA(cudaStreamSynchronize(e->stream));
    if (n_leaves_host) *n_leaves_host = h[0];
    if (n_active_host) *n_active_host = h[1];
    if (n_roots_host) *n_roots_host = h[3];
    if (h[2] > 0) {
      char msg[200];
      snprintf(msg, sizeof msg, "search error on %d game(s): node slab full (raise node_capacity or pass reserve to azg_search_advance), "
               "path deeper than %d, or a root that is already terminal; azg_search_stats reports the error bits", h[2], AZG_MAX_DEPTH);
      return azg_fail(AZG_E_SEARCH, msg);
    }
  }
  return AZG_OK;
}

// Host copy of the counters written by the last azg_search_fill (synchronises the engine's stream only).
extern "C" int azg_search_read_counters(azg_engine* e, int32_t* n_leaves_host, int32_t* n_active_host, int32_t* n_roots_host) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  AZG_USE_DEVICE(e->cfg.device);
  int32_t* h = (int32_t*)e->pinned;
  AZG_CUDA(cudaMemcpyAsync(h, e->dev.counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  AZG_CUDA(cudaStreamSynchronize(e->stream));
  if (n_leaves_host) *n_leaves_host = h[0];
  if (n_active_host) *n_active_host = h[1];
  if (n_roots_host) *n_roots_host = h[3];
  if (h[2] > 0) return azg_fail(AZG_E_SEARCH, "search error on at least one game; azg_search_stats reports the error bits");
  return AZG_OK;
}

extern "C" const int32_t* azg_search_counters(const azg_engine* e) { return e ? e->dev.counters : nullptr; }

extern "C" int azg_search_leaf_planes(azg_engine* e, float* planes) {
  if (!e || !planes) return azg_fail(AZG_E_ARG, "null argument");
  AZG_USE_DEVICE(e->cfg.device);
  azg_leaf_planes_kernel<<<1184, 256, 0, e->stream>>>(e->dev, planes);
  return azg_check_launch("azg_search_leaf_planes");
}

extern "C" int azg_search_commit(azg_engine* e, const float* probs, const double* noise) {
  if (!e || !probs) return azg_fail(AZG_E_ARG, "null argument");
  AZG_USE_DEVICE(e->cfg.device);
  const int warps = e->dev.queue_len < 32 ? e->dev.queue_len : 32;          // one warp per queued leaf
  azg_commit_kernel<<<e->dev.G, 32 * warps, 0, e->stream>>>(e->dev, probs, noise);
  return azg_check_launch("azg_search_commit");
}

extern "C" int azg_search_result(azg_engine* e, float* pi, int32_t* visits) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  AZG_USE_DEVICE(e->cfg.device);
  azg_finish_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, pi, visits);
  return azg_check_launch("azg_search_result");
}

extern "C" int azg_search_advance(azg_engine* e, const int32_t* actions, int gc, int reserve, int32_t* status) {
  if (!e) return azg_fail(AZG_E_ARG, "null engine");
  AZG_USE_DEVICE(e->cfg.device);
  azg_advance_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, actions, gc, reserve, status);
  return azg_check_launch("azg_search_advance");
}

extern "C" int azg_search_stats(azg_engine* e, uint64_t* out_host) {
  if (!e || !out_host) return azg_fail(AZG_E_ARG, "null argument");
  AZG_USE_DEVICE(e->cfg.device);
  azg_stats_kernel<<<1, 256, 0, e->stream>>>(e->dev, e->stats_dev);
  int rc = azg_check_launch("azg_search_stats");
  if (rc) return rc;
  AZG_CUDA(cudaMemcpyAsync(e->pinned, e->stats_dev, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
  AZG_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 8; ++i) out_host[i] = e->pinned[i];
  return AZG_OK;
}

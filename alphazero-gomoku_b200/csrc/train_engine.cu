// Host side of the training step: buffers, TMA descriptors, layer schedule and the C ABI azg_train_*.
//
// Replaces PyTorchModel.train_batch (network.py:199-235) - forward in training mode, KLDiv(batchmean) + MSE,
// backward, clip_grad_norm_(3.0), Adam - with hand-written kernels:
//   * 3x3 convolutions forward (z = conv(a)) and input gradient (da = conv^T(dz)): the tcgen05 implicit-GEMM
//     kernel of the leaf evaluator (net_conv.cu) with linear output, on raw / transposed-flipped bf16 weights;
//   * weight gradient: tcgen05 GEMM over the pixel dimension with MN-major operands (net_wgrad.cu);
//   * BatchNorm (batch statistics), heads, loss, stem, clip + Adam + repack: train_kernels.cu.
// Parameters, gradients and Adam moments live in caller-owned FLAT fp32 vectors in net.parameters() order
// (the Python side makes the module's tensors views into them), so checkpoints keep the reference's format
// and the gradient exchange between ranks is one all-reduce over one buffer.
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "common.cuh"
#include "train.h"

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_fn train_get_encode() {
  static encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (encode_fn)p;
  }
  return fn;
}

// bf16 [rows][cols], box {box_cols, box_rows}; 128-byte swizzle for 64-column boxes, 64-byte for 32-column ones
static int train_make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols = 64) {
  encode_fn enc = train_get_encode();
  if (!enc) return azg_fail(AZG_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return azg_fail(AZG_E_CUDA, "cuTensorMapEncodeTiled failed (training buffers)");
  return AZG_OK;
}

struct ParamLayout {
  long long conv_w, bn_w, bn_b;
  long long res_conv_w[AZG_NET_MAX_LAYERS], res_bn_w[AZG_NET_MAX_LAYERS], res_bn_b[AZG_NET_MAX_LAYERS];
  long long policy_conv_w, policy_bn_w, policy_bn_b, policy_fc_w, policy_fc_b;
  long long value_conv_w, value_bn_w, value_bn_b, value_fc1_w, value_fc1_b, value_fc2_w, value_fc2_b;
  long long total;
};

struct azg_train {
  azg_train_config cfg{};
  int device = 0, C = 0, L = 0, max_batch = 0, n_sm = 0, conv_mode = 1, bound = 0, wgrad_variant = 0;
  int64_t bytes = 0;
  size_t rows = 0;
  ParamLayout lay{};
  std::vector<__nv_bfloat16*> a, z;              // [L+1] activations after / before BatchNorm+ReLU
  std::vector<uint8_t*> relu_mask;               // [L+1] one bit per element of a: [a > 0], read by the BatchNorm backward passes instead of a
  __nv_bfloat16 *g[2] = {nullptr, nullptr}, *dzb[2] = {nullptr, nullptr}, *gskip = nullptr;     // dzb: dL/dz, double buffered (see the backward loop)
  cudaStream_t side = nullptr;                   // the weight-gradient kernels run here, beside the next layer's BatchNorm passes
  std::vector<cudaEvent_t> ev_fork, ev_join;
  int overlap = 1, last_dz = 0;
  std::vector<CUtensorMap> tm_a_in, tm_a_wg, tm_z_st, tm_a_c32, tm_a_c34;      // *_c*: boxes of the cluster weight-gradient kernel
  CUtensorMap tm_dz_c24[2], tm_dz_c16[2];
  int wgrad_cluster = 0;
  CUtensorMap tm_dz_in[2], tm_dz_wg[2], tm_g_st[2], tm_wf, tm_wb;
  __nv_bfloat16 *wf = nullptr, *wb = nullptr;
  float *wp_t = nullptr, *wv1_t = nullptr;
  float *stats = nullptr, *sums = nullptr, *partial = nullptr, *stem_partial = nullptr, *conv_stat = nullptr;
  int fuse_stats = 1, fuse_bwd = 1;
  float *zh = nullptr, *hstats = nullptr, *hidden = nullptr, *h1 = nullptr, *value = nullptr, *dlogits = nullptr, *dvpre = nullptr,
        *dhid = nullptr, *hsums = nullptr, *norm_partial = nullptr, *scal = nullptr;
  unsigned* counters = nullptr;
  long long* step_dev = nullptr;
  int *n_dev = nullptr, *error_dev = nullptr, *pinned = nullptr;
  AdamSeg* segs_dev = nullptr;
  int2* block_seg_dev = nullptr;
  int n_adam_blocks = 0, n_stem_partial = 64;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  azg_net_weights stat_ptrs{};                   // only the running_mean / running_var entries are used
  std::vector<float> zero_shift;
  int last_count = 0, last_g = 0;
  unsigned calls = 0;
};

template <typename T>
static int talloc(azg_train* t, T** p, size_t count) {
  if (cudaMalloc((void**)p, count * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return azg_fail(AZG_E_NOMEM, "cudaMalloc failed (training buffers)"); }
  cudaMemset(*p, 0, count * sizeof(T));
  t->bytes += (int64_t)(count * sizeof(T));
  return AZG_OK;
}

static void layout_params(ParamLayout& l, int C, int L) {
  long long o = 0;
  auto take = [&](long long n) { const long long at = o; o += n; return at; };
  l.conv_w = take((long long)C * 27); l.bn_w = take(C); l.bn_b = take(C);
  for (int i = 0; i < L; ++i) { l.res_conv_w[i] = take((long long)C * C * 9); l.res_bn_w[i] = take(C); l.res_bn_b[i] = take(C); }
  l.policy_conv_w = take(2 * C); l.policy_bn_w = take(2); l.policy_bn_b = take(2);
  l.policy_fc_w = take(225 * 450); l.policy_fc_b = take(225);
  l.value_conv_w = take(C); l.value_bn_w = take(1); l.value_bn_b = take(1);
  l.value_fc1_w = take(64 * 225); l.value_fc1_b = take(64); l.value_fc2_w = take(64); l.value_fc2_b = take(1);
  l.total = o;
}

extern "C" int azg_train_destroy(azg_train* t) {
  if (!t) return AZG_OK;
  cudaSetDevice(t->device);
  for (auto p : t->a) cudaFree(p);
  for (auto p : t->z) cudaFree(p);
  for (auto p : t->relu_mask) cudaFree(p);
  cudaFree(t->g[0]); cudaFree(t->g[1]); cudaFree(t->dzb[0]); cudaFree(t->dzb[1]); cudaFree(t->gskip);
  for (cudaEvent_t e : t->ev_fork) cudaEventDestroy(e);
  for (cudaEvent_t e : t->ev_join) cudaEventDestroy(e);
  if (t->side) cudaStreamDestroy(t->side); cudaFree(t->wf); cudaFree(t->wb);
  cudaFree(t->wp_t); cudaFree(t->wv1_t); cudaFree(t->stats); cudaFree(t->sums); cudaFree(t->partial); cudaFree(t->stem_partial); cudaFree(t->conv_stat);
  cudaFree(t->zh); cudaFree(t->hstats); cudaFree(t->hidden); cudaFree(t->h1); cudaFree(t->value); cudaFree(t->dlogits);
  cudaFree(t->dvpre); cudaFree(t->dhid); cudaFree(t->hsums); cudaFree(t->norm_partial); cudaFree(t->scal); cudaFree(t->counters);
  cudaFree(t->step_dev); cudaFree(t->n_dev); cudaFree(t->error_dev); cudaFree(t->segs_dev); cudaFree(t->block_seg_dev);
  if (t->pinned) cudaFreeHost(t->pinned);
  delete t;
  return AZG_OK;
}

extern "C" int azg_train_create(const azg_train_config* cfg, azg_train** out) {
  if (!cfg || !out) return azg_fail(AZG_E_ARG, "azg_train_create: null argument");
  if (cfg->channels != 64 && cfg->channels != 128 && cfg->channels != 256) return azg_fail(AZG_E_ARG, "azg_train_create: channels must be 64, 128 or 256");
  if (cfg->n_blocks < 0 || cfg->n_blocks > AZG_NET_MAX_BLOCKS || cfg->max_batch < 1) return azg_fail(AZG_E_ARG, "azg_train_create: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return azg_fail(AZG_E_CUDA, "azg_train_create: no such CUDA device (no CPU fallback)");
  }
  AZG_CUDA(cudaSetDevice(cfg->device));
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device);
  if (major != 10) return azg_fail(AZG_E_CUDA, "azg_train_create: tcgen05 kernels need an sm_100 device");
  azg_train* t = new (std::nothrow) azg_train();
  if (!t) return azg_fail(AZG_E_NOMEM, "host allocation failed");
  t->cfg = *cfg; t->device = cfg->device; t->C = cfg->channels; t->L = 2 * cfg->n_blocks; t->max_batch = cfg->max_batch;
  cudaDeviceGetAttribute(&t->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
  t->conv_mode = t->C == 64 ? 3 : 1;
  { const char* v = getenv("AZG_WGRAD_DESC"); t->wgrad_variant = v ? atoi(v) : 0; }
  { const char* v = getenv("AZG_TRAIN_FUSE_BWD"); t->fuse_bwd = v ? atoi(v) : 1; }             // 0: BatchNorm backward sums in their own pass over g and z
  { const char* v = getenv("AZG_TRAIN_FUSE_STATS"); t->fuse_stats = v ? atoi(v) : 1; }     // 0: separate statistics pass over z (experiment switch)
  { const char* v = getenv("AZG_TRAIN_OVERLAP"); t->overlap = v ? atoi(v) : 1; }           // 0: weight gradients on the main stream
  // 1: clusters of three CTAs sharing their operands by TMA multicast (2.4x less L2 -> SMEM traffic).  Measured slower:
  // 71 vs 41 us per launch alone, the same step time when overlapped - the kernel is not bound by L2 traffic but by the
  // 128 x 128 x 16 cta_group::1 UMMA reading 8 KB of shared memory per 64 cycles, and the clusters add lock-step latency.
  { const char* v = getenv("AZG_WGRAD_CLUSTER"); t->wgrad_cluster = v ? atoi(v) : 0; }
  if (t->C == 256) t->wgrad_cluster = 0;                                                   // the cluster experiment is built for one tile per tensor
  const int C = t->C, L = t->L, B = t->max_batch;
  layout_params(t->lay, C, L);
  t->rows = AZG_NET_FRONT + (size_t)B * 256 + AZG_NET_BACK;
  const size_t act = t->rows * C;
  int rc = AZG_OK;
  t->a.assign(L + 1, nullptr); t->z.assign(L + 1, nullptr); t->relu_mask.assign(L + 1, nullptr);
  for (int i = 0; i <= L && !rc; ++i) {
    rc = talloc(t, &t->a[i], act);
    if (!rc) rc = talloc(t, &t->z[i], act);
    if (!rc) rc = talloc(t, &t->relu_mask[i], (size_t)B * 256 * (C / 8));
  }
  if (!rc) rc = talloc(t, &t->g[0], act);
  if (!rc) rc = talloc(t, &t->g[1], act);
  if (!rc) rc = talloc(t, &t->dzb[0], act);
  if (!rc) rc = talloc(t, &t->dzb[1], act);
  if (!rc) rc = talloc(t, &t->gskip, act);
  const size_t wl = (size_t)(L ? L : 1) * 9 * C * C;
  if (!rc) rc = talloc(t, &t->wf, wl);
  if (!rc) rc = talloc(t, &t->wb, wl);
  if (!rc) rc = talloc(t, &t->wp_t, (size_t)450 * 225);
  if (!rc) rc = talloc(t, &t->wv1_t, (size_t)225 * 64);
  if (!rc) rc = talloc(t, &t->stats, (size_t)(L + 1) * 2 * C);
  if (!rc) rc = talloc(t, &t->sums, (size_t)2 * C);
  if (!rc) rc = talloc(t, &t->partial, (size_t)AZG_TRAIN_PARTIALS * 3 * C);
  if (!rc) rc = talloc(t, &t->stem_partial, (size_t)B * 27 * C);
  if (!rc) rc = talloc(t, &t->conv_stat, (size_t)2 * C * AZG_CONV_STAT_SLOTS);
  if (!rc) rc = talloc(t, &t->zh, (size_t)B * 675);
  if (!rc) rc = talloc(t, &t->hstats, (size_t)8);
  if (!rc) rc = talloc(t, &t->hidden, (size_t)B * 675);
  if (!rc) rc = talloc(t, &t->h1, (size_t)B * 64);
  if (!rc) rc = talloc(t, &t->value, (size_t)B);
  if (!rc) rc = talloc(t, &t->dlogits, (size_t)B * 225);
  if (!rc) rc = talloc(t, &t->dvpre, (size_t)B);
  if (!rc) rc = talloc(t, &t->dhid, (size_t)B * 675);
  if (!rc) rc = talloc(t, &t->hsums, (size_t)8);
  if (!rc) rc = talloc(t, &t->norm_partial, (size_t)AZG_TRAIN_PARTIALS);
  if (!rc) rc = talloc(t, &t->scal, (size_t)4);
  if (!rc) rc = talloc(t, &t->counters, (size_t)8);
  if (!rc) rc = talloc(t, &t->step_dev, (size_t)1);
  if (!rc) rc = talloc(t, &t->n_dev, (size_t)4);
  if (!rc) rc = talloc(t, &t->error_dev, (size_t)4);
  if (!rc && cudaMallocHost((void**)&t->pinned, 64) != cudaSuccess) rc = azg_fail(AZG_E_NOMEM, "pinned allocation failed");
  if (rc) { azg_train_destroy(t); return rc; }
  // TMA descriptors
  const uint32_t in_rows = (uint32_t)azg_conv3x3_rows(t->conv_mode), wg_rows = (uint32_t)azg_wgrad3x3_a_rows();
  t->tm_a_in.resize(L + 1); t->tm_a_wg.resize(L + 1); t->tm_z_st.resize(L + 1); t->tm_a_c32.resize(L + 1); t->tm_a_c34.resize(L + 1);
  for (int i = 0; i <= L && !rc; ++i) {
    rc = train_make_map(&t->tm_a_in[i], t->a[i], t->rows, C, in_rows);
    if (!rc) rc = train_make_map(&t->tm_a_wg[i], t->a[i], t->rows, C, wg_rows);
    if (!rc) rc = train_make_map(&t->tm_z_st[i], t->z[i], t->rows, C, 32, 32);
    if (!rc) rc = train_make_map(&t->tm_a_c32[i], t->a[i], t->rows, C, 32);
    if (!rc) rc = train_make_map(&t->tm_a_c34[i], t->a[i], t->rows, C, 34);
  }
  for (int k = 0; k < 2 && !rc; ++k) {
    rc = train_make_map(&t->tm_dz_in[k], t->dzb[k], t->rows, C, in_rows);
    if (!rc) rc = train_make_map(&t->tm_dz_wg[k], t->dzb[k], t->rows, C, 64);
    if (!rc) rc = train_make_map(&t->tm_dz_c24[k], t->dzb[k], t->rows, C, 24);
    if (!rc) rc = train_make_map(&t->tm_dz_c16[k], t->dzb[k], t->rows, C, 16);
  }
  if (!rc) rc = train_make_map(&t->tm_g_st[0], t->g[0], t->rows, C, 32, 32);
  if (!rc) rc = train_make_map(&t->tm_g_st[1], t->g[1], t->rows, C, 32, 32);
  if (!rc) rc = train_make_map(&t->tm_wf, t->wf, (uint64_t)(L ? L : 1) * 9 * C, C, (uint32_t)(C / 2));
  if (!rc) rc = train_make_map(&t->tm_wb, t->wb, (uint64_t)(L ? L : 1) * 9 * C, C, (uint32_t)(C / 2));
  if (rc) { azg_train_destroy(t); return rc; }
  // Adam segments: every parameter tensor is one segment; blocks of 1024 elements
  std::vector<AdamSeg> segs;
  auto add = [&](long long off, long long count, int kind, int layer) { segs.push_back(AdamSeg{off, (int)count, kind, layer}); };
  const ParamLayout& l = t->lay;
  add(l.conv_w, (long long)C * 27, SEG_STEM, 0); add(l.bn_w, C, SEG_PLAIN, 0); add(l.bn_b, C, SEG_PLAIN, 0);
  for (int i = 0; i < L; ++i) { add(l.res_conv_w[i], (long long)C * C * 9, SEG_CONV3, i); add(l.res_bn_w[i], C, SEG_PLAIN, 0); add(l.res_bn_b[i], C, SEG_PLAIN, 0); }
  add(l.policy_conv_w, 2 * C, SEG_PLAIN, 0); add(l.policy_bn_w, 2, SEG_PLAIN, 0); add(l.policy_bn_b, 2, SEG_PLAIN, 0);
  add(l.policy_fc_w, 225 * 450, SEG_WP, 0); add(l.policy_fc_b, 225, SEG_PLAIN, 0);
  add(l.value_conv_w, C, SEG_PLAIN, 0); add(l.value_bn_w, 1, SEG_PLAIN, 0); add(l.value_bn_b, 1, SEG_PLAIN, 0);
  add(l.value_fc1_w, 64 * 225, SEG_WV1, 0); add(l.value_fc1_b, 64, SEG_PLAIN, 0); add(l.value_fc2_w, 64, SEG_PLAIN, 0); add(l.value_fc2_b, 1, SEG_PLAIN, 0);
  std::vector<int2> blocks;
  for (size_t sidx = 0; sidx < segs.size(); ++sidx)
    for (int first = 0; first < segs[sidx].count; first += 1024) blocks.push_back(make_int2((int)sidx, first));
  t->n_adam_blocks = (int)blocks.size();
  if (cudaMalloc((void**)&t->segs_dev, segs.size() * sizeof(AdamSeg)) != cudaSuccess ||
      cudaMalloc((void**)&t->block_seg_dev, blocks.size() * sizeof(int2)) != cudaSuccess) {
    cudaGetLastError(); azg_train_destroy(t); return azg_fail(AZG_E_NOMEM, "cudaMalloc failed (optimiser tables)");
  }
  cudaMemcpy(t->segs_dev, segs.data(), segs.size() * sizeof(AdamSeg), cudaMemcpyHostToDevice);
  cudaMemcpy(t->block_seg_dev, blocks.data(), blocks.size() * sizeof(int2), cudaMemcpyHostToDevice);
  t->zero_shift.assign(C, 0.f);
  if (cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking) != cudaSuccess) { azg_train_destroy(t); return azg_fail(AZG_E_CUDA, "cudaStreamCreate failed"); }
  t->ev_fork.assign(L + 1, nullptr); t->ev_join.assign(L + 1, nullptr);
  for (int i = 0; i <= L; ++i)
    if (cudaEventCreateWithFlags(&t->ev_fork[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->ev_join[i], cudaEventDisableTiming) != cudaSuccess) { azg_train_destroy(t); return azg_fail(AZG_E_CUDA, "cudaEventCreate failed"); }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { azg_train_destroy(t); return azg_fail(AZG_E_CUDA, cudaGetErrorString(ce)); }
  *out = t;
  return AZG_OK;
}

extern "C" int64_t azg_train_param_count(const azg_train* t) { return t ? t->lay.total : 0; }
extern "C" int64_t azg_train_memory_bytes(const azg_train* t) { return t ? t->bytes : 0; }

__global__ void set_count_kernel(int* dst, int v) { *dst = v; }

static AdamArgs adam_args(azg_train* t, int world) {
  AdamArgs a{};
  a.params = t->params; a.grads = t->grads; a.m = t->m; a.v = t->v; a.n = t->lay.total;
  a.segs = t->segs_dev; a.block_seg = t->block_seg_dev; a.n_blocks = t->n_adam_blocks; a.C = t->C;
  a.lr = (float)t->cfg.lr; a.wd = (float)t->cfg.weight_decay; a.b1 = (float)t->cfg.beta1; a.b2 = (float)t->cfg.beta2;
  a.eps = (float)t->cfg.eps; a.clip = (float)t->cfg.clip; a.inv_world = 1.0f / (float)(world < 1 ? 1 : world);
  a.norm_partial = t->norm_partial; a.counter = t->counters + 3; a.scal = t->scal; a.step = t->step_dev;
  a.wf = t->wf; a.wb = t->wb; a.wp_t = t->wp_t; a.wv1_t = t->wv1_t;
  return a;
}

extern "C" int azg_train_bind(azg_train* t, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                              const azg_net_weights* stats, int64_t step, void* stream_) {
  if (!t || !params || !grads || !exp_avg || !exp_avg_sq || !stats) return azg_fail(AZG_E_ARG, "azg_train_bind: null argument");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  t->params = params; t->grads = grads; t->m = exp_avg; t->v = exp_avg_sq; t->stat_ptrs = *stats;
  if (!stats->bn[2] || !stats->bn[3] || !stats->policy_bn[2] || !stats->value_bn[2]) return azg_fail(AZG_E_ARG, "azg_train_bind: missing running statistics");
  for (int i = 0; i < t->L; ++i)
    if (!stats->res_bn[i][2] || !stats->res_bn[i][3]) return azg_fail(AZG_E_ARG, "azg_train_bind: missing running statistics of a residual block");
  long long st = step;
  AZG_CUDA(cudaMemcpyAsync(t->step_dev, &st, sizeof st, cudaMemcpyHostToDevice, s));
  AZG_CUDA(cudaStreamSynchronize(s));
  t->bound = 1;
  return azg_adam_launch(adam_args(t, 1), false, s);          // bf16 / transposed copies of the current weights
}

extern "C" int azg_train_pack(azg_train* t, void* stream_) {
  if (!t || !t->bound) return azg_fail(AZG_E_STATE, "azg_train_pack: not bound");
  AZG_USE_DEVICE(t->device);
  return azg_adam_launch(adam_args(t, 1), false, (cudaStream_t)stream_);
}

static HeadTrainArgs head_args(azg_train* t, int count, const float* pi, const float* zt, float* loss_parts) {
  const ParamLayout& l = t->lay;
  HeadTrainArgs h{};
  h.C = t->C; h.n_boards = count; h.act = t->a[t->L];
  h.w1p = t->params + l.policy_conv_w; h.w1v = t->params + l.value_conv_w;
  h.bn_gamma[0] = t->params + l.policy_bn_w; h.bn_gamma[1] = t->params + l.value_bn_w;
  h.bn_beta[0] = t->params + l.policy_bn_b; h.bn_beta[1] = t->params + l.value_bn_b;
  h.bn_rmean[0] = const_cast<float*>(t->stat_ptrs.policy_bn[2]); h.bn_rvar[0] = const_cast<float*>(t->stat_ptrs.policy_bn[3]);
  h.bn_rmean[1] = const_cast<float*>(t->stat_ptrs.value_bn[2]); h.bn_rvar[1] = const_cast<float*>(t->stat_ptrs.value_bn[3]);
  h.momentum = (float)t->cfg.bn_momentum; h.eps = (float)t->cfg.bn_eps;
  h.wp_t = t->wp_t; h.bp = t->params + l.policy_fc_b; h.wv1_t = t->wv1_t; h.bv1 = t->params + l.value_fc1_b;
  h.w2 = t->params + l.value_fc2_w; h.b2 = t->params + l.value_fc2_b;
  h.wp = t->params + l.policy_fc_w; h.wv1 = t->params + l.value_fc1_w;
  h.pi = pi; h.zt = zt;
  h.zh = t->zh; h.hstats = t->hstats; h.hidden = t->hidden; h.h1 = t->h1; h.value = t->value; h.dlogits = t->dlogits;
  h.dvpre = t->dvpre; h.loss_parts = loss_parts; h.dhid = t->dhid; h.hsums = t->hsums; h.partial = t->partial; h.counter = t->counters + 2;
  h.g = t->g[0];
  h.d_w1p = t->grads + l.policy_conv_w; h.d_w1v = t->grads + l.value_conv_w;
  h.d_bn_gamma[0] = t->grads + l.policy_bn_w; h.d_bn_gamma[1] = t->grads + l.value_bn_w;
  h.d_bn_beta[0] = t->grads + l.policy_bn_b; h.d_bn_beta[1] = t->grads + l.value_bn_b;
  h.d_wp = t->grads + l.policy_fc_w; h.d_bp = t->grads + l.policy_fc_b; h.d_wv1 = t->grads + l.value_fc1_w;
  h.d_bv1 = t->grads + l.value_fc1_b; h.d_w2 = t->grads + l.value_fc2_w; h.d_b2 = t->grads + l.value_fc2_b;
  return h;
}

// weight gradient of trunk layer `layer` from dzb[k] and a[layer] into dw ([9][C][C], accumulated)
static int launch_wgrad(azg_train* t, int k, int layer, int count, float* dw, cudaStream_t s) {
  WgradArgs wa{count, dw, t->error_dev, t->wgrad_variant, 0};
  if (t->wgrad_cluster) {
    const CUtensorMap tm[4] = {t->tm_dz_c24[k], t->tm_dz_c16[k], t->tm_a_c32[layer], t->tm_a_c34[layer]};
    return azg_wgrad3x3_cluster_launch(t->C, tm, wa, t->n_sm, s);
  }
  return azg_wgrad3x3_launch(t->C, t->tm_dz_wg[k], t->tm_a_wg[layer], wa, t->n_sm, s);
}

static ConvArgs conv_args(azg_train* t, int layer, const __nv_bfloat16* residual, __nv_bfloat16* out) {
  ConvArgs a{};
  a.n_boards = t->n_dev; a.max_boards = t->max_batch; a.layer = layer; a.relu = 0; a.shift_host = t->zero_shift.data();
  a.residual = residual; a.out = out; a.head_host = nullptr; a.hidden = nullptr; a.error = t->error_dev; a.prof = nullptr; a.prof_detail = 0;
  return a;
}

// forward (training mode) + loss + backward: gradients of every parameter into the bound flat gradient vector
// (3x3 convolution weights in [tap][cout][cin] order, the stem in [plane*9+tap][c] order - azg_train_apply and
// azg_train_export_grads know).  loss_parts float32[count][2] = {KL row sum, squared value error}.
extern "C" int azg_train_forward_backward(azg_train* t, const float* planes, const float* pis, const float* zs, int count,
                                          float* loss_parts, void* stream_) {
  if (!t || !planes || !pis || !zs || !loss_parts) return azg_fail(AZG_E_ARG, "azg_train_forward_backward: null argument");
  if (!t->bound) return azg_fail(AZG_E_STATE, "azg_train_forward_backward: call azg_train_bind first");
  if (count < 2 || count > t->max_batch) return azg_fail(AZG_E_ARG, "azg_train_forward_backward: count must be 2..max_batch");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  const int C = t->C, L = t->L;
  const ParamLayout& l = t->lay;
  int rc;
  set_count_kernel<<<1, 1, 0, s>>>(t->n_dev, count);      // a kernel argument: also right when the call is replayed from a CUDA graph
  AZG_CUDA(cudaMemsetAsync(t->grads, 0, (size_t)l.total * sizeof(float), s));      // the wgrad kernel accumulates with reductions
  const float mom = (float)t->cfg.bn_momentum, eps = (float)t->cfg.bn_eps;
  auto bn_fwd = [&](int idx, const float* gamma, const float* beta, float* rmean, float* rvar, const __nv_bfloat16* residual) -> int {
    const bool fused = idx > 0 && t->fuse_stats;             // the convolution's epilogue already summed z and z^2 per channel
    BnStatsArgs st{t->z[idx], count, fused ? t->conv_stat : t->partial, t->counters + 0, t->stats + (size_t)idx * 2 * C, rmean, rvar, mom, eps};
    int r = fused ? azg_bn_finalize_launch(C, st, azg_conv3x3_stat_slots(t->max_batch, t->n_sm, C), s) : azg_bn_stats_launch(C, st, s);
    if (r) return r;
    BnApplyArgs ap{t->z[idx], t->stats + (size_t)idx * 2 * C, gamma, beta, residual, t->a[idx], count, t->relu_mask[idx]};
    return azg_bn_apply_launch(C, ap, t->n_sm, s);
  };
  // ---- forward
  StemTrainArgs st{};
  st.planes = planes; st.w = t->params + l.conv_w; st.n_boards = count; st.z = t->z[0];
  if ((rc = azg_stem_train_fwd_launch(C, st, s))) return rc;
  if ((rc = bn_fwd(0, t->params + l.bn_w, t->params + l.bn_b, const_cast<float*>(t->stat_ptrs.bn[2]), const_cast<float*>(t->stat_ptrs.bn[3]), nullptr))) return rc;
  for (int i = 0; i < L; ++i) {
    ConvArgs ca = conv_args(t, i, nullptr, t->z[i + 1]);
    if (t->fuse_stats) ca.stat_partial = t->conv_stat;
    if ((rc = azg_conv3x3_launch(C, t->conv_mode, t->tm_a_in[i], t->tm_wf, t->tm_z_st[i + 1], ca, t->n_sm, s))) return rc;
    if ((rc = bn_fwd(i + 1, t->params + l.res_bn_w[i], t->params + l.res_bn_b[i], const_cast<float*>(t->stat_ptrs.res_bn[i][2]),
                     const_cast<float*>(t->stat_ptrs.res_bn[i][3]), (i & 1) ? t->a[i - 1] : nullptr))) return rc;
  }
  HeadTrainArgs h = head_args(t, count, pis, zs, loss_parts);
  if ((rc = azg_head_train_fwd_launch(h, t->n_sm, s))) return rc;
  // ---- backward
  if ((rc = azg_head_train_bwd_launch(h, t->n_sm, s))) return rc;
  int cur = 0;
  // fused_sums: the input-gradient convolution that produced g already summed dy and dy * z per channel in its epilogue
  auto bn_bwd = [&](int idx, const float* gamma, float* dgamma, float* dbeta, bool want_skip, __nv_bfloat16* dz, bool fused_sums) -> int {
    BnBwdArgs b{};
    b.g = t->g[cur]; b.mask = t->relu_mask[idx]; b.z = t->z[idx]; b.stats = t->stats + (size_t)idx * 2 * C; b.gamma = gamma; b.n_boards = count;
    b.partial = t->partial; b.counter = t->counters + 1; b.sums = t->sums; b.dgamma = dgamma; b.dbeta = dbeta; b.dz = dz;
    b.gskip = want_skip ? t->gskip : nullptr;
    int r = fused_sums ? azg_bn_bwd_finalize_launch(C, b, t->conv_stat, azg_conv3x3_stat_slots(t->max_batch, t->n_sm, C), s)
                       : azg_bn_bwd_reduce_launch(C, b, s);
    if (r) return r;
    return azg_bn_bwd_apply_launch(C, b, t->n_sm, s);
  };
  // Per layer: BatchNorm backward (reduce, apply -> dz), then two independent tensor-core kernels that both read dz:
  // the input gradient (needed by the next layer down, stays on this stream) and the weight gradient (needed only by the
  // optimiser).  The weight gradient runs on a side stream next to the HBM-bound BatchNorm passes of the layer below;
  // dz is double buffered so that those passes do not overwrite what it is still reading.
  const bool ov = t->overlap != 0;
  const bool fuse_bwd = t->fuse_stats != 0 && t->fuse_bwd != 0;
  cudaStream_t ws = ov ? t->side : s;
  for (int i = L - 1; i >= 0; --i) {
    const int k = i & 1;
    if (ov && i + 2 <= L - 1) AZG_CUDA(cudaStreamWaitEvent(s, t->ev_join[i + 2], 0));      // the kernel that last read dzb[k] is done
    if ((rc = bn_bwd(i + 1, t->params + l.res_bn_w[i], t->grads + l.res_bn_w[i], t->grads + l.res_bn_b[i], (i & 1) != 0, t->dzb[k],
                     fuse_bwd && i < L - 1))) return rc;
    if (ov) { AZG_CUDA(cudaEventRecord(t->ev_fork[i], s)); AZG_CUDA(cudaStreamWaitEvent(ws, t->ev_fork[i], 0)); }
    if ((rc = launch_wgrad(t, k, i, count, t->grads + l.res_conv_w[i], ws))) return rc;
    if (ov) AZG_CUDA(cudaEventRecord(t->ev_join[i], ws));
    ConvArgs ca = conv_args(t, i, (i & 1) ? nullptr : t->gskip, t->g[cur ^ 1]);
    if (fuse_bwd) { ca.stat_partial = t->conv_stat; ca.bwd_z = t->z[i]; ca.bwd_mask = t->relu_mask[i]; }   // g is dL/da of BatchNorm idx i
    if ((rc = azg_conv3x3_launch(C, t->conv_mode, t->tm_dz_in[k], t->tm_wb, t->tm_g_st[cur ^ 1], ca, t->n_sm, s))) return rc;
    cur ^= 1;
  }
  const int ks = L >= 2 ? 1 : (L & 1);                     // buffer of the stem's dz: the one layer 1 used (layer 0 may still be read)
  if (ov && L >= 2) AZG_CUDA(cudaStreamWaitEvent(s, t->ev_join[1], 0));
  if (ov && L == 1) AZG_CUDA(cudaStreamWaitEvent(s, t->ev_join[0], 0));
  if ((rc = bn_bwd(0, t->params + l.bn_w, t->grads + l.bn_w, t->grads + l.bn_b, false, t->dzb[L >= 1 ? ks : 0], fuse_bwd && L >= 1))) return rc;
  st.dz = t->dzb[L >= 1 ? ks : 0]; st.partial = t->stem_partial; st.dw = t->grads + l.conv_w;
  if ((rc = azg_stem_train_wgrad_launch(C, st, s))) return rc;
  if (ov && L >= 1) AZG_CUDA(cudaStreamWaitEvent(s, t->ev_join[0], 0));                   // join: every weight gradient has landed
  t->last_dz = L >= 1 ? ks : 0;
  t->last_count = count; t->last_g = cur;
  return AZG_OK;
}

// clip_grad_norm_(clip) + Adam + repack.  `world` > 1: the gradient vector holds the SUM over that many ranks
// (all-reduced by the caller) and is averaged here.
extern "C" int azg_train_apply(azg_train* t, int world, void* stream_) {
  if (!t || !t->bound) return azg_fail(AZG_E_STATE, "azg_train_apply: not bound");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  AdamArgs a = adam_args(t, world);
  int rc = azg_grad_norm_launch(a, s);
  if (rc) return rc;
  return azg_adam_launch(a, true, s);
}

// Synchronise and report the tcgen05 pipeline watchdogs; out_host[0] = gradient norm of the last apply,
// out_host[1] = clip coefficient (including 1/world), out_host[2] = optimiser step count.
extern "C" int azg_train_check(azg_train* t, double* out_host, void* stream_) {
  if (!t) return azg_fail(AZG_E_ARG, "null trainer");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  float sc[4]; long long st = 0; int err = 0;
  AZG_CUDA(cudaMemcpyAsync(sc, t->scal, sizeof sc, cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaMemcpyAsync(&st, t->step_dev, sizeof st, cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaMemcpyAsync(&err, t->error_dev, sizeof err, cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaStreamSynchronize(s));
  if (out_host) { out_host[0] = sc[1]; out_host[1] = sc[0]; out_host[2] = (double)st; }
  if (err != 0) {
    char msg[96];
    snprintf(msg, sizeof msg, "training pipeline watchdog fired (code %d)", err);
    cudaMemsetAsync(t->error_dev, 0, sizeof(int), s);
    return azg_fail(AZG_E_CUDA, msg);
  }
  return AZG_OK;
}

// ---- test hooks -------------------------------------------------------------------------------------------------------
__global__ void train_unpad_kernel(const __nv_bfloat16* act, int C, int count, float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)count * C * 225;
  if (i >= total) return;
  const int pix = (int)(i % 225), c = (int)((i / 225) % C);
  const size_t b = i / ((size_t)225 * C);
  out[i] = __bfloat162float(act[((size_t)AZG_NET_FRONT + b * 256 + (size_t)((pix / 15 + 1) * 16 + pix % 15)) * C + c]);
}

__global__ void train_export_conv_kernel(const float* g, int C, float* out) {       // [tap][co][ci] -> [co][ci][3][3]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * 9) return;
  const int co = i / (C * 9), ci = (i / 9) % C, tap = i % 9;
  out[i] = g[(tap * C + co) * C + ci];
}

__global__ void train_export_stem_kernel(const float* g, int C, float* out) {       // [27][C] -> [C][27]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 27) return;
  out[i] = g[(i % 27) * C + i / 27];
}

// what: 0 = a[layer], 1 = z[layer] (layer 0 = stem .. 2*n_blocks), 2 = dL/da of the stem output (last input gradient),
// 3 = last dz; out float32 [count][C][15][15] of the last forward_backward.
extern "C" int azg_train_read_activation(azg_train* t, int what, int layer, float* out, void* stream_) {
  if (!t || !out || t->last_count < 1) return azg_fail(AZG_E_ARG, "azg_train_read_activation: nothing to read");
  if ((what == 0 || what == 1) && (layer < 0 || layer > t->L)) return azg_fail(AZG_E_ARG, "azg_train_read_activation: bad layer");
  AZG_USE_DEVICE(t->device);
  const __nv_bfloat16* src = what == 0 ? t->a[layer] : what == 1 ? t->z[layer] : what == 2 ? t->g[t->last_g] : t->dzb[t->last_dz];
  const size_t total = (size_t)t->last_count * t->C * 225;
  train_unpad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(src, t->C, t->last_count, out);
  return azg_check_launch("train_unpad_kernel");
}

__global__ void train_pad_kernel(const float* in, int C, int count, __nv_bfloat16* act) {     // [n][C][15][15] fp32 -> padded bf16 rows
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)count * C * 225;
  if (i >= total) return;
  const int pix = (int)(i % 225), c = (int)((i / 225) % C);
  const size_t b = i / ((size_t)225 * C);
  act[((size_t)AZG_NET_FRONT + b * 256 + (size_t)((pix / 15 + 1) * 16 + pix % 15)) * C + c] = __float2bfloat16_rn(in[i]);
}

// Test hook for the two tensor-core gradient kernels in isolation: with dz and a given as float32[count][C][15][15]
// (rounded to bf16 on the way in) and the CURRENT weights of trunk layer `layer`,
//   dw_out float32[C][C][3][3] = weight gradient  sum_p dz[p][co] * a[p + tap][ci]            (net_wgrad.cu)
//   da_out float32[count][C][15][15] = input gradient  conv^T(dz, W[layer])                      (net_conv.cu on transposed weights)
// Overwrites the trainer's activation scratch; the bound gradient vector is left untouched.
extern "C" int azg_train_debug_conv_grads(azg_train* t, const float* dz, const float* a, int count, int layer, float* dw_out,
                                          float* da_out, void* stream_) {
  if (!t || !dz || !a || !dw_out || !da_out || !t->bound) return azg_fail(AZG_E_ARG, "azg_train_debug_conv_grads: bad argument");
  if (count < 1 || count > t->max_batch || layer < 0 || layer >= t->L) return azg_fail(AZG_E_ARG, "azg_train_debug_conv_grads: bad count / layer");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  const int C = t->C;
  const size_t total = (size_t)count * C * 225, act = t->rows * C;
  int rc;
  AZG_CUDA(cudaMemsetAsync(t->dzb[0], 0, act * 2, s));
  AZG_CUDA(cudaMemsetAsync(t->a[layer], 0, act * 2, s));
  train_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dz, C, count, t->dzb[0]);
  train_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, C, count, t->a[layer]);
  set_count_kernel<<<1, 1, 0, s>>>(t->n_dev, count);
  float* scratch = nullptr;                                   // [9][C][C] accumulator, separate from the bound gradients
  AZG_CUDA(cudaMalloc((void**)&scratch, (size_t)9 * C * C * sizeof(float)));
  AZG_CUDA(cudaMemsetAsync(scratch, 0, (size_t)9 * C * C * sizeof(float), s));
  if ((rc = launch_wgrad(t, 0, layer, count, scratch, s))) { cudaFree(scratch); return rc; }
  train_export_conv_kernel<<<(C * C * 9 + 255) / 256, 256, 0, s>>>(scratch, C, dw_out);
  ConvArgs ca = conv_args(t, layer, nullptr, t->g[0]);
  if ((rc = azg_conv3x3_launch(C, t->conv_mode, t->tm_dz_in[0], t->tm_wb, t->tm_g_st[0], ca, t->n_sm, s))) { cudaFree(scratch); return rc; }
  train_unpad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(t->g[0], C, count, da_out);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(scratch);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  return azg_check_launch("azg_train_debug_conv_grads");
}

// The gradient vector in the PARAMETER layout (what p.grad would hold): out float32[param_count].
extern "C" int azg_train_export_grads(azg_train* t, float* out, void* stream_) {
  if (!t || !out || !t->bound) return azg_fail(AZG_E_ARG, "azg_train_export_grads: bad argument");
  AZG_USE_DEVICE(t->device);
  cudaStream_t s = (cudaStream_t)stream_;
  const int C = t->C;
  AZG_CUDA(cudaMemcpyAsync(out, t->grads, (size_t)t->lay.total * sizeof(float), cudaMemcpyDeviceToDevice, s));
  train_export_stem_kernel<<<(C * 27 + 255) / 256, 256, 0, s>>>(t->grads + t->lay.conv_w, C, out + t->lay.conv_w);
  for (int i = 0; i < t->L; ++i)
    train_export_conv_kernel<<<(C * C * 9 + 255) / 256, 256, 0, s>>>(t->grads + t->lay.res_conv_w[i], C, out + t->lay.res_conv_w[i]);
  return azg_check_launch("train_export");
}

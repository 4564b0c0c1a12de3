// Host side of the leaf evaluator: weight repacking, activation buffers, TMA descriptors and the
// layer schedule.  C ABI: azg_net_* (include/azgomoku_b200.h).  Replaces AlphaZeroNet.forward +
// PyTorchModel.predict (network.py:85-117, 168-183).
#include <stdlib.h>
#include <new>
#include <vector>
#include "common.cuh"
#include "engine.h"
#include "net.h"

int azg_pack_launch_fold(const float*, const float*, const float*, const float*, int, float*, float*, cudaStream_t);
int azg_pack_launch_conv3(const float*, const float*, int, int, __nv_bfloat16*, cudaStream_t);
int azg_pack_launch_stem(const float*, const float*, int, float*, cudaStream_t);

struct azg_net {
  int device = 0, n_blocks = 0, C = 0, max_batch = 0, n_sm = 0, loaded = 0, conv_mode = 0, prof_detail = 0;
  int64_t bytes = 0;
  size_t rows = 0;                       // rows of one activation buffer (front pad + boards*256 + back pad)
  __nv_bfloat16* w3 = nullptr;           // [(layer*9+tap)*C + cout][cin]
  float *scale3 = nullptr, *shift3 = nullptr;                         // [2*n_blocks][C]
  float *stem_w = nullptr, *stem_scale = nullptr, *stem_shift = nullptr;
  float *head_w1 = nullptr, *head_scale1 = nullptr, *head_shift1 = nullptr;
  float *pol_w = nullptr, *pol_b = nullptr, *v1_w = nullptr, *v1_b = nullptr, *v2_w = nullptr, *v2_b = nullptr;   // pol_w [240][480], v1_w [64][256]: padded, see net.h
  size_t hid_rows = 0;                   // rows of `hidden` (max_batch rounded up to the 128-board GEMM tile)
  __nv_bfloat16* act[3] = {nullptr, nullptr, nullptr};
  float* hidden = nullptr;
  uint32_t *keys = nullptr, *meta = nullptr;
  int *n_dev = nullptr, *error_dev = nullptr;
  unsigned long long* prof_dev = nullptr;
  int* pinned = nullptr;
  CUtensorMap tm_act[3], tm_w, tm_st[3];      // tm_st: 32x32 SWIZZLE_64B store boxes over the activation buffers
  CUtensorMap tm_hid, tm_wp, tm_wv;          // fp32 operand maps of the heads GEMM
  // optional timing of the 3x3 trunk (one CUDA-event pair per forward pass, on the launch stream)
  int profiling = 0;
  std::vector<cudaEvent_t> ev;          // start/stop pairs
  std::vector<float> shift_host;        // [2*n_blocks][C] folded BatchNorm shifts of the 3x3 layers
  std::vector<float> head_host;         // [3*C + 6] 1x1 head weights + folded BN scale[3], shift[3]
  size_t ev_used = 0;
  long long prof_launches = 0;
};

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_fn get_encode() {
  static encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (encode_fn)p;
  }
  return fn;
}

// 2-D row-major [rows][cols] tensor (bf16, or fp32 when `f32`), box {box_cols columns, box_rows}: 128-byte swizzle
// when the box row is 128 bytes (operand tiles), 64-byte swizzle for the 64-byte rows of the epilogue store tiles.
static int make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols = 64,
                    bool f32 = false) {
  encode_fn enc = get_encode();
  if (!enc) return azg_fail(AZG_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint32_t esz = f32 ? 4u : 2u;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return azg_fail(AZG_E_CUDA, "cuTensorMapEncodeTiled failed");
  return AZG_OK;
}

template <typename T>
static int nalloc(azg_net* n, T** p, size_t count) {
  if (cudaMalloc((void**)p, count * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return azg_fail(AZG_E_NOMEM, "cudaMalloc failed (network buffers)"); }
  n->bytes += (int64_t)(count * sizeof(T));
  return AZG_OK;
}

extern "C" int azg_net_destroy(azg_net* n) {
  if (!n) return AZG_OK;
  cudaSetDevice(n->device);
  cudaFree(n->w3); cudaFree(n->scale3); cudaFree(n->shift3); cudaFree(n->stem_w); cudaFree(n->stem_scale); cudaFree(n->stem_shift);
  cudaFree(n->head_w1); cudaFree(n->head_scale1); cudaFree(n->head_shift1); cudaFree(n->pol_w); cudaFree(n->pol_b);
  cudaFree(n->v1_w); cudaFree(n->v1_b); cudaFree(n->v2_w); cudaFree(n->v2_b);
  for (int i = 0; i < 3; ++i) cudaFree(n->act[i]);
  cudaFree(n->hidden); cudaFree(n->keys); cudaFree(n->meta); cudaFree(n->n_dev); cudaFree(n->error_dev); cudaFree(n->prof_dev);
  if (n->pinned) cudaFreeHost(n->pinned);
  for (cudaEvent_t ev : n->ev) cudaEventDestroy(ev);
  delete n;
  return AZG_OK;
}

extern "C" int azg_net_create(int device, int n_blocks, int channels, int max_batch, azg_net** out) {
  if (!out || n_blocks < 0 || n_blocks > AZG_NET_MAX_BLOCKS || max_batch < 1) return azg_fail(AZG_E_ARG, "azg_net_create: bad argument");
  if (channels != 64 && channels != 128 && channels != 256)
    return azg_fail(AZG_E_ARG, "azg_net_create: channels must be 64, 128 or 256");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return azg_fail(AZG_E_CUDA, "azg_net_create: no such CUDA device (no CPU fallback)");
  }
  AZG_CUDA(cudaSetDevice(device));
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) return azg_fail(AZG_E_CUDA, "azg_net_create: tcgen05 kernels need an sm_100 device");
  azg_net* n = new (std::nothrow) azg_net();
  if (!n) return azg_fail(AZG_E_NOMEM, "host allocation failed");
  n->device = device; n->n_blocks = n_blocks; n->C = channels; n->max_batch = max_batch;
  cudaDeviceGetAttribute(&n->n_sm, cudaDevAttrMultiProcessorCount, device);
  const size_t C = channels, L = 2 * (size_t)n_blocks;
  n->rows = AZG_NET_FRONT + (size_t)max_batch * 256 + AZG_NET_BACK;
  n->hid_rows = ((size_t)max_batch + 127) / 128 * 128;
  int rc = AZG_OK;
  if ((rc = nalloc(n, &n->w3, (L ? L : 1) * 9 * C * C)) || (rc = nalloc(n, &n->scale3, (L ? L : 1) * C)) ||
      (rc = nalloc(n, &n->shift3, (L ? L : 1) * C)) || (rc = nalloc(n, &n->stem_w, 27 * C)) ||
      (rc = nalloc(n, &n->stem_scale, C)) || (rc = nalloc(n, &n->stem_shift, C)) || (rc = nalloc(n, &n->head_w1, 3 * C)) ||
      (rc = nalloc(n, &n->head_scale1, (size_t)4)) || (rc = nalloc(n, &n->head_shift1, (size_t)4)) ||
      (rc = nalloc(n, &n->pol_w, (size_t)240 * AZG_HEAD_WP_K)) || (rc = nalloc(n, &n->pol_b, (size_t)240)) ||
      (rc = nalloc(n, &n->v1_w, (size_t)64 * AZG_HEAD_WV_K)) || (rc = nalloc(n, &n->v1_b, (size_t)64)) ||
      (rc = nalloc(n, &n->v2_w, (size_t)64)) || (rc = nalloc(n, &n->v2_b, (size_t)1)) ||
      (rc = nalloc(n, &n->act[0], n->rows * C)) || (rc = nalloc(n, &n->act[1], n->rows * C)) ||
      (rc = nalloc(n, &n->act[2], n->rows * C)) ||
      (rc = nalloc(n, &n->hidden, n->hid_rows * AZG_HEAD_PITCH)) ||
      (rc = nalloc(n, &n->keys, (size_t)max_batch * 16)) || (rc = nalloc(n, &n->meta, (size_t)max_batch)) ||
      (rc = nalloc(n, &n->n_dev, (size_t)4)) || (rc = nalloc(n, &n->error_dev, (size_t)4)) ||
      (rc = nalloc(n, &n->prof_dev, (size_t)32))) {
    azg_net_destroy(n);
    return rc;
  }
  if (cudaMallocHost((void**)&n->pinned, 64) != cudaSuccess) { azg_net_destroy(n); return azg_fail(AZG_E_NOMEM, "pinned allocation failed"); }
  for (int i = 0; i < 3; ++i) cudaMemset(n->act[i], 0, n->rows * C * 2);      // pad rows must read as zero
  cudaMemset(n->error_dev, 0, 16);
  cudaMemset(n->prof_dev, 0, 256);
  cudaMemset(n->hidden, 0, n->hid_rows * AZG_HEAD_PITCH * 4);        // the pad columns stay zero (finite) for ever
  cudaMemset(n->pol_w, 0, (size_t)240 * AZG_HEAD_WP_K * 4);           // zero rows 225..239 and columns 450..479
  cudaMemset(n->pol_b, 0, 240 * 4);
  cudaMemset(n->v1_w, 0, (size_t)64 * AZG_HEAD_WV_K * 4);
  {
    const char* pd = getenv("AZG_CONV_PHASES");      // clock the epilogue phases too when profiling
    n->prof_detail = pd ? atoi(pd) : 0;
    const char* m = getenv("AZG_CONV_MODE");        // experiment switch for the staging / epilogue variant
    // measured defaults: 128 and 256 channels: single activation copy + staged epilogue with TMA stores (1);
    // 64 channels: single copy + direct stores (3) - the short N = 64 MMAs leave no time for the staging
    // round trip (5.91 M vs 5.41 M sims/s on the 3x64 net)
    n->conv_mode = m ? atoi(m) : (channels == 64 ? 3 : 1);
    if (n->conv_mode != 0 && n->conv_mode != 1 && n->conv_mode != 3 && n->conv_mode != 4 && !(n->conv_mode == 5 && channels == 128))
      n->conv_mode = channels == 64 ? 3 : 1;
    if (channels == 256 && n->conv_mode == 0) n->conv_mode = 1;      // the streaming-weights kernel needs the single-copy layout
  }
  for (int i = 0; i < 3; ++i)
    if ((rc = make_map(&n->tm_act[i], n->act[i], n->rows, C, (uint32_t)azg_conv3x3_rows(n->conv_mode)))) { azg_net_destroy(n); return rc; }
  for (int i = 0; i < 3; ++i)
    if ((rc = make_map(&n->tm_st[i], n->act[i], n->rows, C, 32, 32))) { azg_net_destroy(n); return rc; }
  if ((rc = make_map(&n->tm_w, n->w3, (L ? L : 1) * 9 * C, C, (uint32_t)(C / 2)))) { azg_net_destroy(n); return rc; }
  if ((rc = make_map(&n->tm_hid, n->hidden, n->hid_rows, AZG_HEAD_PITCH, 128, 32, true)) ||
      (rc = make_map(&n->tm_wp, n->pol_w, 240, AZG_HEAD_WP_K, 240, 32, true)) ||
      (rc = make_map(&n->tm_wv, n->v1_w, 64, AZG_HEAD_WV_K, 64, 32, true))) { azg_net_destroy(n); return rc; }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { azg_net_destroy(n); return azg_fail(AZG_E_CUDA, cudaGetErrorString(ce)); }
  *out = n;
  return AZG_OK;
}

extern "C" int64_t azg_net_memory_bytes(const azg_net* n) { return n ? n->bytes : 0; }

extern "C" int azg_net_load(azg_net* n, const azg_net_weights* w, void* stream_) {
  if (!n || !w) return azg_fail(AZG_E_ARG, "azg_net_load: null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  AZG_CUDA(cudaSetDevice(n->device));
  const int C = n->C;
  int rc;
  if ((rc = azg_pack_launch_fold(w->bn[0], w->bn[1], w->bn[2], w->bn[3], C, n->stem_scale, n->stem_shift, s))) return rc;
  if ((rc = azg_pack_launch_stem(w->conv_w, n->stem_scale, C, n->stem_w, s))) return rc;
  for (int l = 0; l < 2 * n->n_blocks; ++l) {
    if (!w->res_conv_w[l]) return azg_fail(AZG_E_ARG, "azg_net_load: missing residual-block weights");
    if ((rc = azg_pack_launch_fold(w->res_bn[l][0], w->res_bn[l][1], w->res_bn[l][2], w->res_bn[l][3], C,
                                   n->scale3 + (size_t)l * C, n->shift3 + (size_t)l * C, s))) return rc;
    if ((rc = azg_pack_launch_conv3(w->res_conv_w[l], n->scale3 + (size_t)l * C, C, l, n->w3, s))) return rc;
  }
  // head 1x1 convs: rows 0-1 policy_conv.weight [2][C], row 2 value_conv.weight [1][C]
  AZG_CUDA(cudaMemcpyAsync(n->head_w1, w->policy_conv_w, 2 * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpyAsync(n->head_w1 + 2 * C, w->value_conv_w, C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if ((rc = azg_pack_launch_fold(w->policy_bn[0], w->policy_bn[1], w->policy_bn[2], w->policy_bn[3], 2, n->head_scale1, n->head_shift1, s))) return rc;
  if ((rc = azg_pack_launch_fold(w->value_bn[0], w->value_bn[1], w->value_bn[2], w->value_bn[3], 1, n->head_scale1 + 2, n->head_shift1 + 2, s))) return rc;
  // dense head layers: the reference's [out][in] matrices are already the K-major B operands of the heads GEMM;
  // value_fc1's feature j sits in hidden column 450 + j = padded column 2 + j (the value GEMM starts at column 448)
  AZG_CUDA(cudaMemcpy2DAsync(n->pol_w, AZG_HEAD_WP_K * sizeof(float), w->policy_fc_w, 450 * sizeof(float), 450 * sizeof(float), 225,
                             cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpy2DAsync(n->v1_w + 2, AZG_HEAD_WV_K * sizeof(float), w->value_fc1_w, 225 * sizeof(float), 225 * sizeof(float), 64,
                             cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpyAsync(n->pol_b, w->policy_fc_b, 225 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpyAsync(n->v1_b, w->value_fc1_b, 64 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpyAsync(n->v2_w, w->value_fc2_w, 64 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  AZG_CUDA(cudaMemcpyAsync(n->v2_b, w->value_fc2_b, sizeof(float), cudaMemcpyDeviceToDevice, s));
  // the epilogue takes the shifts as kernel arguments (constant bank): keep a host copy
  n->shift_host.resize((size_t)2 * n->n_blocks * C + 1);
  if (n->n_blocks > 0)
    AZG_CUDA(cudaMemcpyAsync(n->shift_host.data(), n->shift3, (size_t)2 * n->n_blocks * C * sizeof(float), cudaMemcpyDeviceToHost, s));
  n->head_host.resize((size_t)3 * C + 6);
  AZG_CUDA(cudaMemcpyAsync(n->head_host.data(), n->head_w1, (size_t)3 * C * sizeof(float), cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaMemcpyAsync(n->head_host.data() + 3 * C, n->head_scale1, 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaMemcpyAsync(n->head_host.data() + 3 * C + 3, n->head_shift1, 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaStreamSynchronize(s));
  n->loaded = 1;
  return AZG_OK;
}

// Stem, trunk (first `n_layers` 3x3 layers) and optionally the heads for the batch described by
// `stem` (count on the device).  Returns the index of the activation buffer holding the trunk output.
static int run_network(azg_net* n, StemArgs stem, const int* n_ptr, int max_boards, int n_layers, float* probs, float* values,
                       float* logits, bool heads, cudaStream_t s, int* out_buf) {
  if (!n->loaded) return azg_fail(AZG_E_STATE, "azg_net: weights not loaded");
  const int C = n->C;
  int rc;
  stem.n_boards = n_ptr; stem.max_boards = max_boards; stem.w = n->stem_w; stem.shift = n->stem_shift; stem.out = n->act[0];
  if ((rc = azg_stem_launch(C, stem, n->n_sm, s))) return rc;
  int x = 0, t = 1, y = 2;
  bool fused_heads = false;
  cudaEvent_t ev_stop = nullptr;
  if (n->profiling && n_layers > 0) {
    if (n->ev_used + 2 > n->ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) { n->ev.push_back(a); n->ev.push_back(b); }
    }
    if (n->ev_used + 2 <= n->ev.size()) {
      cudaEventRecord(n->ev[n->ev_used], s);
      ev_stop = n->ev[n->ev_used + 1];
      n->ev_used += 2;
      n->prof_launches += n_layers;
    }
  }
  for (int l = 0; l < n_layers; ++l) {
    ConvArgs a{};
    a.n_boards = n_ptr; a.max_boards = max_boards; a.layer = l; a.relu = 1;
    a.shift_host = n->shift_host.data() + (size_t)l * C; a.error = n->error_dev; a.prof = n->profiling ? n->prof_dev + 16 * (l & 1) : nullptr; a.prof_detail = n->prof_detail;   // [0..15] layers without, [16..31] with residual
    a.head_host = nullptr; a.hidden = nullptr;
    const bool fuse = heads && C <= 128 && l == n_layers - 1 && (l & 1) == 1;     // last conv2: fuse the 1x1 head convs, skip the store
    if (fuse) { a.head_host = n->head_host.data(); a.hidden = n->hidden; fused_heads = true; }
    if ((l & 1) == 0) { a.residual = nullptr; a.out = n->act[t]; rc = azg_conv3x3_launch(C, n->conv_mode, n->tm_act[x], n->tm_w, n->tm_st[t], a, n->n_sm, s); }
    else { a.residual = n->act[x]; a.out = fuse ? nullptr : n->act[y]; rc = azg_conv3x3_launch(C, n->conv_mode, n->tm_act[t], n->tm_w, n->tm_st[y], a, n->n_sm, s); int tmp = x; x = y; y = tmp; }
    if (rc) return rc;
  }
  if (ev_stop) cudaEventRecord(ev_stop, s);
  const int last = (n_layers & 1) ? t : x;
  if (out_buf) *out_buf = last;
  if (heads) {
    HeadArgs h;
    h.n_boards = n_ptr; h.max_boards = max_boards; h.act = n->act[last]; h.w1 = n->head_w1; h.scale1 = n->head_scale1;
    h.shift1 = n->head_shift1; h.hidden = n->hidden; h.pol_b = n->pol_b;
    h.v1_b = n->v1_b; h.v2_w = n->v2_w; h.v2_b = n->v2_b; h.probs = probs; h.values = values; h.logits = logits;
    if ((rc = azg_heads_launch(C, h, n->tm_hid, n->tm_wp, n->tm_wv, n->error_dev, n->n_sm, s, fused_heads))) return rc;
  }
  return AZG_OK;
}

static int check_watchdog(azg_net* n, cudaStream_t s) {
  AZG_CUDA(cudaMemcpyAsync(n->pinned, n->error_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  AZG_CUDA(cudaStreamSynchronize(s));
  if (n->pinned[0] != 0) {
    char msg[96];
    snprintf(msg, sizeof msg, "conv3x3 pipeline watchdog fired (code %d)", n->pinned[0]);
    cudaMemsetAsync(n->error_dev, 0, sizeof(int), s);
    return azg_fail(AZG_E_CUDA, msg);
  }
  return AZG_OK;
}

extern "C" int azg_net_forward_planes(azg_net* n, const float* planes, int count, float* probs, float* values, float* logits,
                                      void* stream_) {
  if (!n || !planes || !probs || count < 0) return azg_fail(AZG_E_ARG, "azg_net_forward_planes: bad argument");
  cudaStream_t s = (cudaStream_t)stream_;
  AZG_CUDA(cudaSetDevice(n->device));
  for (int done = 0; done < count; done += n->max_batch) {
    const int m = count - done < n->max_batch ? count - done : n->max_batch;
    n->pinned[1] = m;
    AZG_CUDA(cudaMemcpyAsync(n->n_dev, &n->pinned[1], sizeof(int), cudaMemcpyHostToDevice, s));
    int rc = azg_planes_to_keys_launch(planes + (size_t)done * 675, m, n->keys, n->meta, s);
    if (rc) return rc;
    StemArgs st = {};
    st.keys = n->keys; st.meta = n->meta; st.leaf_game = nullptr; st.leaf_node = nullptr; st.slab_stride = 0;
    rc = run_network(n, st, n->n_dev, m, 2 * n->n_blocks, probs + (size_t)done * 225, values ? values + done : nullptr,
                     logits ? logits + (size_t)done * 225 : nullptr, true, s, nullptr);
    if (rc) return rc;
    if ((rc = check_watchdog(n, s))) return rc;     // also orders the reuse of pinned[1]
  }
  return AZG_OK;
}

// Debug / test entry: trunk activations after the stem and the first n_layers 3x3 layers,
// float32 [count][C][15][15] (NCHW like the reference's tensors).
__global__ void unpad_kernel(const __nv_bfloat16* act, int C, int count, float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)count * C * 225;
  if (i >= total) return;
  const int pix = (int)(i % 225), c = (int)((i / 225) % C);
  const size_t b = i / ((size_t)225 * C);
  const int r = pix / 15, col = pix % 15;
  out[i] = __bfloat162float(act[((size_t)AZG_NET_FRONT + b * 256 + (size_t)((r + 1) * 16 + col)) * C + c]);
}

extern "C" int azg_net_trunk_debug(azg_net* n, const float* planes, int count, int n_layers, float* out, void* stream_) {
  if (!n || !planes || !out || count < 1 || count > n->max_batch || n_layers < 0 || n_layers > 2 * n->n_blocks)
    return azg_fail(AZG_E_ARG, "azg_net_trunk_debug: bad argument");
  cudaStream_t s = (cudaStream_t)stream_;
  AZG_CUDA(cudaSetDevice(n->device));
  n->pinned[1] = count;
  AZG_CUDA(cudaMemcpyAsync(n->n_dev, &n->pinned[1], sizeof(int), cudaMemcpyHostToDevice, s));
  int rc = azg_planes_to_keys_launch(planes, count, n->keys, n->meta, s);
  if (rc) return rc;
  StemArgs st = {};
  st.keys = n->keys; st.meta = n->meta;
  int buf = 0;
  if ((rc = run_network(n, st, n->n_dev, count, n_layers, nullptr, nullptr, nullptr, false, s, &buf))) return rc;
  const size_t total = (size_t)count * n->C * 225;
  unpad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n->act[buf], n->C, count, out);
  if ((rc = azg_check_launch("unpad"))) return rc;
  return check_watchdog(n, s);
}

// Evaluate the engine's current leaf batch in place (count and leaf list stay on the device).
extern "C" int azg_net_forward_leaves(azg_net* n, azg_engine* e, float* probs, float* values) {
  if (!n || !e || !probs) return azg_fail(AZG_E_ARG, "azg_net_forward_leaves: null argument");
  const int cap = e->dev.G * e->dev.queue_len;
  if (cap > n->max_batch) return azg_fail(AZG_E_ARG, "azg_net_forward_leaves: network max_batch is smaller than games*queue_len");
  if (e->cfg.device != n->device) return azg_fail(AZG_E_ARG, "azg_net_forward_leaves: engine and network live on different devices");
  AZG_CUDA(cudaSetDevice(n->device));
  StemArgs st = {};
  st.keys = e->dev.key; st.meta = e->dev.meta; st.leaf_game = e->dev.leaf_game; st.leaf_node = e->dev.leaf_node;
  st.slab_stride = e->dev.cap;
  return run_network(n, st, e->dev.counters, cap, 2 * n->n_blocks, probs, values, nullptr, true, e->stream, nullptr);
}

// Trunk timing: enable != 0 starts collecting one event pair around the 3x3 layers of every
// forward pass; azg_net_profile_read synchronises the device, returns the summed milliseconds and
// the number of conv3x3 launches they cover, and resets the collection.
extern "C" int azg_net_profile(azg_net* n, int enable) {
  if (!n) return azg_fail(AZG_E_ARG, "null network");
  n->profiling = enable ? 1 : 0;
  return AZG_OK;
}

extern "C" int azg_net_profile_read(azg_net* n, double* trunk_ms, int64_t* launches) {
  if (!n || !trunk_ms || !launches) return azg_fail(AZG_E_ARG, "null argument");
  AZG_CUDA(cudaSetDevice(n->device));
  AZG_CUDA(cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i + 1 < n->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, n->ev[i], n->ev[i + 1]) == cudaSuccess) total += ms;
  }
  *trunk_ms = total;
  *launches = n->prof_launches;
  n->ev_used = 0;
  n->prof_launches = 0;
  return AZG_OK;
}

// Pipeline wait counters of the conv3x3 kernel accumulated while profiling is enabled (cycles,
// summed over clusters): {mma wait-full, mma wait-tmem-empty, mma total, producer wait-empty,
// producer total, epilogue wait-tmem-full, epilogue total, boards, then five epilogue phases of one
// warp: wait-for-store-drain, residual transpose, tcgen05.ld wait, compute + stage, fence + TMA store,
// and three reserved}, once for the layers without residual and once for those with; out must hold 32
// values; reading resets them.
extern "C" int azg_net_profile_counters(azg_net* n, uint64_t* out8) {
  if (!n || !out8) return azg_fail(AZG_E_ARG, "null argument");
  AZG_CUDA(cudaSetDevice(n->device));
  AZG_CUDA(cudaDeviceSynchronize());
  AZG_CUDA(cudaMemcpy(out8, n->prof_dev, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  AZG_CUDA(cudaMemset(n->prof_dev, 0, 256));
  return AZG_OK;
}

// Watchdog status of the tcgen05 pipeline (0 = healthy); synchronises the stream.
extern "C" int azg_net_check(azg_net* n, void* stream_) {
  if (!n) return azg_fail(AZG_E_ARG, "null network");
  AZG_CUDA(cudaSetDevice(n->device));
  return check_watchdog(n, (cudaStream_t)stream_);
}

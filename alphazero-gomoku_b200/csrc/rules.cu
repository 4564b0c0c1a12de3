// Batched rule kernels behind the C ABI (one warp per position).
// Reference semantics: games/gomoku.py:60-78,109-121,130-150,155-197 and
// games/pente.py:57-79,114-152,164-194,199-236.
#include "common.cuh"
#include "rules.cuh"
#include "host.h"

__global__ void __launch_bounds__(128)
rules_pack_kernel(const int8_t* __restrict__ boards, const int32_t* __restrict__ players, const int32_t* __restrict__ lasts,
                  const int32_t* __restrict__ caps, const int32_t* __restrict__ plies, azg_pos* __restrict__ out, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int l = lane_id();
  const int8_t* b = boards + (size_t)i * AZG_A;
  // lane l gathers bit l of each of the 8 words
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int a = w * 32 + l;
    const int v = a < AZG_A ? b[a] : 0;
    const uint32_t m1 = __ballot_sync(AZG_FULL, v == 1), m2 = __ballot_sync(AZG_FULL, v == 2);
    if (l == 0) { out[i].stones[0][w] = m1; out[i].stones[1][w] = m2; }
  }
  if (l == 0) {
    out[i].player = players[i];
    out[i].last = lasts ? lasts[i] : -1;
    out[i].caps[0] = caps ? caps[2 * i] : 0;
    out[i].caps[1] = caps ? caps[2 * i + 1] : 0;
    out[i].plies = plies ? plies[i] : 0;
    out[i].pad[0] = out[i].pad[1] = out[i].pad[2] = 0;
  }
}

__global__ void __launch_bounds__(128)
rules_unpack_kernel(const azg_pos* __restrict__ pos, int8_t* __restrict__ boards, int32_t* __restrict__ players,
                    int32_t* __restrict__ lasts, int32_t* __restrict__ caps, int32_t* __restrict__ plies, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int l = lane_id();
  if (boards)
    for (int a = l; a < AZG_A; a += 32) {
      const uint32_t b1 = (pos[i].stones[0][a >> 5] >> (a & 31)) & 1u, b2 = (pos[i].stones[1][a >> 5] >> (a & 31)) & 1u;
      boards[(size_t)i * AZG_A + a] = (int8_t)(b1 ? 1 : (b2 ? 2 : 0));
    }
  if (l == 0) {
    if (players) players[i] = pos[i].player;
    if (lasts) lasts[i] = pos[i].last;
    if (caps) { caps[2 * i] = pos[i].caps[0]; caps[2 * i + 1] = pos[i].caps[1]; }
    if (plies) plies[i] = pos[i].plies;
  }
}

__device__ __forceinline__ int status_bits(const WPos& p, int rule) {
  const int won = wpos_winner(p, rule);
  const bool over = won != 0 || !wpos_any_empty(p);
  return won | (over ? 4 : 0);
}

__global__ void __launch_bounds__(128)
rules_play_kernel(int rule, azg_pos* __restrict__ pos, const int32_t* __restrict__ actions, int32_t* __restrict__ status, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  WPos p = wpos_load(pos + i);
  int rejected = 0;
  if (actions) {
    const int a = actions[i];
    const bool on = a >= 0 && a < AZG_A;
    if (on && wpos_at(p, a) == 0) { wpos_play(p, rule, a); wpos_store(pos + i, p); }
    else rejected = 8;
  }
  const int s = status_bits(p, rule) | rejected;
  if (status && lane_id() == 0) status[i] = s;
}

__global__ void __launch_bounds__(128)
rules_legal_kernel(const azg_pos* __restrict__ pos, float* __restrict__ mask, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  for (int a = lane_id(); a < AZG_A; a += 32) {
    const uint32_t occ = pos[i].stones[0][a >> 5] | pos[i].stones[1][a >> 5];
    mask[(size_t)i * AZG_A + a] = ((occ >> (a & 31)) & 1u) ? 0.f : 1.f;
  }
}

__global__ void __launch_bounds__(128)
rules_encode_kernel(const azg_pos* __restrict__ pos, float* __restrict__ planes, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int me = pos[i].player - 1;
  float* o = planes + (size_t)i * 3 * AZG_A;
  for (int a = lane_id(); a < AZG_A; a += 32) {
    o[a] = (float)((pos[i].stones[me][a >> 5] >> (a & 31)) & 1u);
    o[AZG_A + a] = (float)((pos[i].stones[1 - me][a >> 5] >> (a & 31)) & 1u);
    o[2 * AZG_A + a] = 1.0f;
  }
}

static inline dim3 warp_grid(int n) { return dim3((unsigned)((n + 3) / 4)); }

extern "C" int azg_rules_pack(const int8_t* boards, const int32_t* players, const int32_t* lasts, const int32_t* caps,
                              const int32_t* plies, azg_pos* out, int n, void* stream) {
  if (!boards || !players || !out || n < 0) return azg_fail(AZG_E_ARG, "azg_rules_pack: null argument");
  if (n == 0) return AZG_OK;
  rules_pack_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(boards, players, lasts, caps, plies, out, n);
  return azg_check_launch("rules_pack");
}

extern "C" int azg_rules_unpack(const azg_pos* pos, int8_t* boards, int32_t* players, int32_t* lasts, int32_t* caps,
                                int32_t* plies, int n, void* stream) {
  if (!pos || n < 0) return azg_fail(AZG_E_ARG, "azg_rules_unpack: null argument");
  if (n == 0) return AZG_OK;
  rules_unpack_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(pos, boards, players, lasts, caps, plies, n);
  return azg_check_launch("rules_unpack");
}

extern "C" int azg_rules_play(int rule, azg_pos* pos, const int32_t* actions, int32_t* status, int n, void* stream) {
  if (!pos || !actions || n < 0 || (rule != 0 && rule != 1)) return azg_fail(AZG_E_ARG, "azg_rules_play: bad argument");
  if (n == 0) return AZG_OK;
  rules_play_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(rule, pos, actions, status, n);
  return azg_check_launch("rules_play");
}

extern "C" int azg_rules_status(int rule, const azg_pos* pos, int32_t* status, int n, void* stream) {
  if (!pos || !status || n < 0 || (rule != 0 && rule != 1)) return azg_fail(AZG_E_ARG, "azg_rules_status: bad argument");
  if (n == 0) return AZG_OK;
  rules_play_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(rule, const_cast<azg_pos*>(pos), nullptr, status, n);
  return azg_check_launch("rules_status");
}

extern "C" int azg_rules_legal(const azg_pos* pos, float* mask, int n, void* stream) {
  if (!pos || !mask || n < 0) return azg_fail(AZG_E_ARG, "azg_rules_legal: null argument");
  if (n == 0) return AZG_OK;
  rules_legal_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(pos, mask, n);
  return azg_check_launch("rules_legal");
}

extern "C" int azg_rules_encode(const azg_pos* pos, float* planes, int n, void* stream) {
  if (!pos || !planes || n < 0) return azg_fail(AZG_E_ARG, "azg_rules_encode: null argument");
  if (n == 0) return AZG_OK;
  rules_encode_kernel<<<warp_grid(n), 128, 0, (cudaStream_t)stream>>>(pos, planes, n);
  return azg_check_launch("rules_encode");
}

extern "C" int azg_rules_play_host(int rule, int device, int8_t* boards, int32_t* players, int32_t* lasts, int32_t* caps,
                                   int32_t* plies, const int32_t* actions, int32_t* status, int n) {
  if (!boards || !players || !lasts || !caps || !plies || !actions || !status || n <= 0)
    return azg_fail(AZG_E_ARG, "azg_rules_play_host: null argument");
  AZG_CUDA(cudaSetDevice(device));
  const size_t nb = (size_t)n * AZG_A, ni = (size_t)n * sizeof(int32_t);
  char* buf = nullptr;
  const size_t total = nb + 8 * ni + (size_t)n * sizeof(azg_pos) + 4096;      // every sub-buffer is padded to 256 B
  AZG_CUDA(cudaMalloc(&buf, total));
  size_t o = 0;
  auto take = [&](size_t bytes) { char* p = buf + o; o += (bytes + 255) & ~(size_t)255; return p; };
  int8_t* d_b = (int8_t*)take(nb);
  int32_t* d_pl = (int32_t*)take(ni); int32_t* d_la = (int32_t*)take(ni); int32_t* d_ca = (int32_t*)take(2 * ni);
  int32_t* d_pi = (int32_t*)take(ni); int32_t* d_ac = (int32_t*)take(ni); int32_t* d_st = (int32_t*)take(ni);
  azg_pos* d_pos = (azg_pos*)take((size_t)n * sizeof(azg_pos));
  int rc = AZG_OK;
  cudaError_t ce;
#define AZG_TRY(x) do { ce = (x); if (ce != cudaSuccess) { rc = azg_fail(AZG_E_CUDA, cudaGetErrorString(ce)); goto done; } } while (0)
  AZG_TRY(cudaMemcpy(d_b, boards, nb, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_pl, players, ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_la, lasts, ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_ca, caps, 2 * ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_pi, plies, ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_ac, actions, ni, cudaMemcpyHostToDevice));
  if ((rc = azg_rules_pack(d_b, d_pl, d_la, d_ca, d_pi, d_pos, n, nullptr))) goto done;
  if ((rc = azg_rules_play(rule, d_pos, d_ac, d_st, n, nullptr))) goto done;
  if ((rc = azg_rules_unpack(d_pos, d_b, d_pl, d_la, d_ca, d_pi, n, nullptr))) goto done;
  AZG_TRY(cudaMemcpy(boards, d_b, nb, cudaMemcpyDeviceToHost));
  AZG_TRY(cudaMemcpy(players, d_pl, ni, cudaMemcpyDeviceToHost));
  AZG_TRY(cudaMemcpy(lasts, d_la, ni, cudaMemcpyDeviceToHost));
  AZG_TRY(cudaMemcpy(caps, d_ca, 2 * ni, cudaMemcpyDeviceToHost));
  AZG_TRY(cudaMemcpy(plies, d_pi, ni, cudaMemcpyDeviceToHost));
  AZG_TRY(cudaMemcpy(status, d_st, ni, cudaMemcpyDeviceToHost));
#undef AZG_TRY
done:
  cudaFree(buf);
  return rc;
}

// Host-buffer query used by the Python game shims: status bits, legal mask and encoded planes of n
// positions given as boards + scalars (nothing is modified).  legal_host / planes_host may be NULL.
extern "C" int azg_rules_query_host(int rule, int device, const int8_t* boards, const int32_t* players, const int32_t* lasts,
                                    const int32_t* caps, const int32_t* plies, int32_t* status, float* legal_host,
                                    float* planes_host, int n) {
  if (!boards || !players || !lasts || !caps || !plies || !status || n <= 0)
    return azg_fail(AZG_E_ARG, "azg_rules_query_host: null argument");
  AZG_CUDA(cudaSetDevice(device));
  const size_t nb = (size_t)n * AZG_A, ni = (size_t)n * sizeof(int32_t);
  char* buf = nullptr;
  const size_t total = nb + 6 * ni + (size_t)n * sizeof(azg_pos) + (size_t)n * AZG_A * 4 * 4 + 4096;
  AZG_CUDA(cudaMalloc(&buf, total));
  size_t o = 0;
  auto take = [&](size_t bytes) { char* p = buf + o; o += (bytes + 255) & ~(size_t)255; return p; };
  int8_t* d_b = (int8_t*)take(nb);
  int32_t* d_pl = (int32_t*)take(ni); int32_t* d_la = (int32_t*)take(ni); int32_t* d_ca = (int32_t*)take(2 * ni);
  int32_t* d_pi = (int32_t*)take(ni); int32_t* d_st = (int32_t*)take(ni);
  azg_pos* d_pos = (azg_pos*)take((size_t)n * sizeof(azg_pos));
  float* d_legal = (float*)take((size_t)n * AZG_A * 4);
  float* d_planes = (float*)take((size_t)n * AZG_A * 12);
  int rc = AZG_OK;
  cudaError_t ce;
#define AZG_TRY(x) do { ce = (x); if (ce != cudaSuccess) { rc = azg_fail(AZG_E_CUDA, cudaGetErrorString(ce)); goto done; } } while (0)
  AZG_TRY(cudaMemcpy(d_b, boards, nb, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_pl, players, ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_la, lasts, ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_ca, caps, 2 * ni, cudaMemcpyHostToDevice));
  AZG_TRY(cudaMemcpy(d_pi, plies, ni, cudaMemcpyHostToDevice));
  if ((rc = azg_rules_pack(d_b, d_pl, d_la, d_ca, d_pi, d_pos, n, nullptr))) goto done;
  if ((rc = azg_rules_status(rule, d_pos, d_st, n, nullptr))) goto done;
  AZG_TRY(cudaMemcpy(status, d_st, ni, cudaMemcpyDeviceToHost));
  if (legal_host) {
    if ((rc = azg_rules_legal(d_pos, d_legal, n, nullptr))) goto done;
    AZG_TRY(cudaMemcpy(legal_host, d_legal, (size_t)n * AZG_A * 4, cudaMemcpyDeviceToHost));
  }
  if (planes_host) {
    if ((rc = azg_rules_encode(d_pos, d_planes, n, nullptr))) goto done;
    AZG_TRY(cudaMemcpy(planes_host, d_planes, (size_t)n * AZG_A * 12, cudaMemcpyDeviceToHost));
  }
#undef AZG_TRY
done:
  cudaFree(buf);
  return rc;
}

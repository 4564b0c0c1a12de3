// Non-GEMM parts of the leaf evaluator: weight packing, fused encode + stem convolution,
// fused policy / value heads.  Reference: network.py:85-117 (forward), :168-183 (predict),
// games/gomoku.py:130-150 (input planes).
#include <cuda_bf16.h>
#include "net.h"

// ------------------------------------------------------------------------------------------------
// weight packing (runs once per weight update)
// ------------------------------------------------------------------------------------------------
// Eval-mode BatchNorm folded to y = x*scale + shift (network.py BatchNorm2d, eps 1e-5).
__global__ void fold_bn_kernel(const float* g, const float* b, const float* mean, const float* var, int n, float* scale,
                               float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / sqrtf(var[i] + 1e-5f);
  scale[i] = s;
  shift[i] = b[i] - mean[i] * s;
}

// conv weight [Cout][Cin][3][3] fp32, times the folded BatchNorm scale of its output channel ->
// bf16 rows (layer*9 + tap)*C + cout, columns cin (K-major B operand).
__global__ void pack_conv3_kernel(const float* w, const float* scale, int C, int layer, __nv_bfloat16* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * C) return;
  const int ci = i % C, co = (i / C) % C, tap = i / (C * C);
  out[((size_t)(layer * 9 + tap) * C + co) * C + ci] = __float2bfloat16(w[((size_t)co * C + ci) * 9 + tap] * scale[co]);
}

// stem weight [C][3][3][3] -> fp32 [(tap*3 + plane)][C], BatchNorm scale folded in.
__global__ void pack_stem_kernel(const float* w, const float* scale, int C, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * C) return;
  const int c = i % C, plane = (i / C) % 3, tap = i / (3 * C);
  out[i] = w[((size_t)c * 3 + plane) * 9 + tap] * scale[c];
}

__global__ void transpose_kernel(const float* in, int rows, int cols, float* out) {   // out[c][r] = in[r][c]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  out[(size_t)c * rows + r] = in[i];
}

int azg_pack_launch_fold(const float* g, const float* b, const float* m, const float* v, int n, float* scale, float* shift,
                         cudaStream_t s) {
  fold_bn_kernel<<<(n + 127) / 128, 128, 0, s>>>(g, b, m, v, n, scale, shift);
  return azg_check_launch("fold_bn");
}
int azg_pack_launch_conv3(const float* w, const float* scale, int C, int layer, __nv_bfloat16* out, cudaStream_t s) {
  pack_conv3_kernel<<<(9 * C * C + 255) / 256, 256, 0, s>>>(w, scale, C, layer, out);
  return azg_check_launch("pack_conv3");
}
int azg_pack_launch_stem(const float* w, const float* scale, int C, float* out, cudaStream_t s) {
  pack_stem_kernel<<<(27 * C + 255) / 256, 256, 0, s>>>(w, scale, C, out);
  return azg_check_launch("pack_stem");
}
int azg_pack_launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t s) {
  transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, s>>>(in, rows, cols, out);
  return azg_check_launch("transpose");
}

// ------------------------------------------------------------------------------------------------
// float planes -> packed stones (PyTorchModel.predict entry, network.py:168-183)
// ------------------------------------------------------------------------------------------------
__global__ void planes_to_keys_kernel(const float* __restrict__ planes, int n, uint32_t* __restrict__ keys,
                                      uint32_t* __restrict__ meta) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int l = threadIdx.x & 31;
  const float* p = planes + (size_t)i * 3 * 225;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int a = w * 32 + l;
    const bool mine = a < 225 && p[a] > 0.5f, theirs = a < 225 && p[225 + a] > 0.5f;
    const uint32_t m1 = __ballot_sync(0xffffffffu, mine), m2 = __ballot_sync(0xffffffffu, theirs);
    if (l == 0) { keys[(size_t)i * 16 + w] = m1; keys[(size_t)i * 16 + 8 + w] = m2; }
  }
  if (l == 0) meta[i] = 1u | (1u << 1);        // "player 1 to move": plane 0 is the mover's stones
}

int azg_planes_to_keys_launch(const float* planes, int n, uint32_t* keys, uint32_t* meta, cudaStream_t stream) {
  if (n <= 0) return AZG_OK;
  planes_to_keys_kernel<<<(n + 3) / 4, 128, 0, stream>>>(planes, n, keys, meta);
  return azg_check_launch("planes_to_keys");
}

// ------------------------------------------------------------------------------------------------
// encode + stem: conv3x3(3 -> C) + BN + ReLU straight from the stone bitboards
// ------------------------------------------------------------------------------------------------
// The three input planes are {mover stones, opponent stones, ones} (gomoku.py:146-150), so each
// tap contributes a sum of at most two weight vectors - no multiplies.  One block per board,
// warp w owns padded rows 32w..32w+31, lane l owns C/32 consecutive channels: every store
// instruction writes one full pixel row (C*2 bytes, coalesced).
// Row-pattern tables: for tap row d (dr = d-1) and the 6-bit pattern of its three neighbour states
// (2 bits each: 0 off-board, 1 empty, 2 mover, 3 opponent) the summed weight vector of that row is
// precomputed in shared memory, so a pixel costs three 16-byte table reads instead of nine.
template <int CT>
__global__ void __launch_bounds__(256) stem_kernel(StemArgs a) {
  // CT channels in total; a block computes a slab of C = min(CT, 128) of them (blockIdx.y selects it)
  constexpr int C = CT > 128 ? 128 : CT;
  constexpr int CPL = C / 32;
  const int cbase = blockIdx.y * C;
  a.w += cbase; a.shift += cbase; a.out += cbase;
  extern __shared__ __align__(16) float s_rows[];   // [3][64][C]
  __shared__ __align__(16) float s_empty[C];         // shift + all nine taps on empty in-board cells
  __shared__ __align__(16) float s_shift[C];
  __shared__ uint8_t s_state[256];                   // 0 off-board / pad, 1 empty, 2 mover stone, 3 opponent stone
  __shared__ uint32_t s_code[256];                   // the nine neighbour states of each padded pixel, 2 bits per tap
  for (int i = threadIdx.x; i < 3 * 64 * C; i += 256) {
    const int c = i % C, combo = (i / C) & 63, d = i / (64 * C);
    float v = 0.f;
#pragma unroll
    for (int dc = 0; dc < 3; ++dc) {
      const int st = (combo >> (2 * dc)) & 3;
      if (st == 0) continue;
      const int tap = d * 3 + dc;
      v += a.w[(tap * 3 + 2) * CT + c];
      if (st >= 2) v += a.w[(tap * 3 + (st - 2)) * CT + c];
    }
    s_rows[i] = v;
  }
  for (int i = threadIdx.x; i < C; i += 256) {
    float e = a.shift[i];
    for (int tap = 0; tap < 9; ++tap) e += a.w[(tap * 3 + 2) * CT + i];
    s_empty[i] = e;
    s_shift[i] = a.shift[i];
  }
  int n = *a.n_boards;
  if (n > a.max_boards) n = a.max_boards;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    size_t node = (size_t)b;
    if (a.leaf_game) node = (size_t)a.leaf_game[b] * (size_t)a.slab_stride + (size_t)a.leaf_node[b];
    const uint32_t* k = a.keys + node * 16;
    const int player = (int)((a.meta[node] >> 1) & 3u);
    __syncthreads();
    {
      const int qi = threadIdx.x, y = qi >> 4, x = qi & 15;
      uint8_t st = 0;
      if (y >= 1 && x < 15) {
        const int cell = (y - 1) * 15 + x;
        const uint32_t b1 = (k[cell >> 5] >> (cell & 31)) & 1u, b2 = (k[8 + (cell >> 5)] >> (cell & 31)) & 1u;
        const uint32_t mine = player == 2 ? b2 : b1, theirs = player == 2 ? b1 : b2;
        st = mine ? 2 : (theirs ? 3 : 1);
      }
      s_state[qi] = st;
    }
    __syncthreads();
    {
      const int qi = threadIdx.x;
      uint32_t code = 0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int q2 = qi + (tap / 3 - 1) * 16 + (tap % 3 - 1);
        const uint32_t st = (q2 >= 0 && q2 < 256) ? s_state[q2] : 0;
        code |= st << (2 * tap);
      }
      s_code[qi] = s_state[qi] ? code : 0xffffffffu;      // pad rows/columns are written as zeros
    }
    __syncthreads();
    __nv_bfloat16* out = a.out + ((size_t)AZG_NET_FRONT + (size_t)b * 256) * CT;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const int qi = warp * 32 + i;
      const uint32_t code = s_code[qi];
      float acc[CPL];
      if (code == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
      } else if (code == 0x15555u) {                       // nine empty in-board neighbours
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = fmaxf(s_empty[lane * CPL + j], 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = s_shift[lane * CPL + j];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float* w = s_rows + (size_t)(d * 64 + ((code >> (6 * d)) & 63u)) * C + lane * CPL;
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[j] += w[j];
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
      __nv_bfloat16* o = out + (size_t)qi * CT + lane * CPL;
      if constexpr (CPL == 2) {
        *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(acc[0], acc[1]);
      } else {
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[j], acc[j + 1]), p1 = __floats2bfloat162_rn(acc[j + 2], acc[j + 3]);
          uint2 v;
          v.x = *reinterpret_cast<const uint32_t*>(&p0);
          v.y = *reinterpret_cast<const uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(o + j) = v;
        }
      }
    }
  }
}

template <int C>
static int stem_launch_t(const StemArgs& a, int n_sm, cudaStream_t stream) {
  constexpr int CB = C > 128 ? 128 : C;
  const int smem = 3 * 64 * CB * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(stem_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  int grid = n_sm * (C == 64 ? 4 : 2) / (C / CB);
  if (grid > a.max_boards) grid = a.max_boards;
  if (grid < 1) grid = 1;
  stem_kernel<C><<<dim3(grid, C / CB), 256, smem, stream>>>(a);
  return azg_check_launch("stem_kernel");
}

int azg_stem_launch(int C, const StemArgs& a, int n_sm, cudaStream_t stream) {
  if (C == 64) return stem_launch_t<64>(a, n_sm, stream);
  if (C == 128) return stem_launch_t<128>(a, n_sm, stream);
  if (C == 256) return stem_launch_t<256>(a, n_sm, stream);
  return azg_fail(AZG_E_ARG, "stem: channels must be 64, 128 or 256");
}

// ------------------------------------------------------------------------------------------------
// heads, part 1: the three 1x1 convolutions (policy 2 channels, value 1) + BN + ReLU
// ------------------------------------------------------------------------------------------------
// Output is tiled for part 2: hidden[(b/32)][k][b%32], k = ch*225 + pixel (policy, the
// reference's view(B, 2*225) order, network.py:105) then 450 + pixel (value).
#define AZG_HID 676      // 675 features + 1 pad
template <int C>
__global__ void __launch_bounds__(256) head1_kernel(HeadArgs a) {
  constexpr int CPL = C / 32;
  int n = *a.n_boards;
  if (n > a.max_boards) n = a.max_boards;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w[3][CPL];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < CPL; ++j) w[r][j] = a.w1[r * C + lane * CPL + j];
  const float sc0 = a.scale1[0], sc1 = a.scale1[1], sc2 = a.scale1[2];
  const float sh0 = a.shift1[0], sh1 = a.shift1[1], sh2 = a.shift1[2];
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    const __nv_bfloat16* act = a.act + ((size_t)AZG_NET_FRONT + (size_t)b * 256) * C;
    float* hid = a.hidden + (size_t)(b >> 5) * (AZG_HID * 32) + (b & 31);
    // four pixels per iteration so that four row loads are in flight per warp
    for (int p0 = warp * 4; p0 < 225; p0 += 32) {
      float x[4][CPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pix = min(p0 + u, 224);
        const int r = pix / 15, c = pix - r * 15;
        const __nv_bfloat16* row = act + (size_t)((r + 1) * 16 + c) * C + lane * CPL;
        if constexpr (CPL == 2) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row));
          x[u][0] = f.x; x[u][1] = f.y;
        } else {
#pragma unroll
          for (int j = 0; j < CPL; j += 4) {
            const uint2 v = *reinterpret_cast<const uint2*>(row + j);
            x[u][j] = __uint_as_float(v.x << 16); x[u][j + 1] = __uint_as_float(v.x & 0xffff0000u);
            x[u][j + 2] = __uint_as_float(v.y << 16); x[u][j + 3] = __uint_as_float(v.y & 0xffff0000u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pix = p0 + u;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) { d0 = fmaf(x[u][j], w[0][j], d0); d1 = fmaf(x[u][j], w[1][j], d1); d2 = fmaf(x[u][j], w[2][j], d2); }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          d0 += __shfl_xor_sync(0xffffffffu, d0, s);
          d1 += __shfl_xor_sync(0xffffffffu, d1, s);
          d2 += __shfl_xor_sync(0xffffffffu, d2, s);
        }
        if (lane == 0 && pix < 225) {
          hid[(size_t)pix * 32] = fmaxf(fmaf(d0, sc0, sh0), 0.f);
          hid[(size_t)(225 + pix) * 32] = fmaxf(fmaf(d1, sc1, sh1), 0.f);
          hid[(size_t)(450 + pix) * 32] = fmaxf(fmaf(d2, sc2, sh2), 0.f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// heads, part 2: policy_fc + softmax, value_fc1 + ReLU + value_fc2 + tanh, 16 boards per block
// ------------------------------------------------------------------------------------------------
// 16 boards (half a feature tile) per block: 62 KB of shared memory, three blocks per SM.
constexpr int kH2B = 16;
constexpr int kPolSlice = 60, kValSlice = 30;      // K slices of the two dense layers: 8 slices each (450 = 7*60+30, 225 = 7*30+15)
constexpr int kHead2Smem = (AZG_HID * kH2B + kH2B * 228 + kH2B * 64) * 4;

__global__ void __launch_bounds__(256) head2_kernel(HeadArgs a) {
  extern __shared__ float sm[];
  float* s_h = sm;                         // [676][16]
  float* s_lg = sm + AZG_HID * kH2B;       // [16][228] logits
  float* s_v = s_lg + kH2B * 228;          // [16][64]
  int n = *a.n_boards;
  if (n > a.max_boards) n = a.max_boards;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_units = (n + kH2B - 1) / kH2B;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int tile = unit >> 1, half = unit & 1, b0 = unit * kH2B;
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(a.hidden + (size_t)tile * (AZG_HID * 32) + half * kH2B);
    float4* dst = reinterpret_cast<float4*>(s_h);
    for (int i = tid; i < AZG_HID * 4; i += 256) dst[i] = src[(i >> 2) * 8 + (i & 3)];     // row k: 16 of the 32 boards
    __syncthreads();
    // policy_fc (network.py:106): logits[b][o] = bias[o] + sum_k W[o][k] h[b][k]
    if (tid < 225) {
      float acc[kH2B];
      const float bias = a.pol_b[tid];
#pragma unroll
      for (int b = 0; b < kH2B; ++b) acc[b] = bias;
      // K is summed in slices of kPolSlice (partial chain from 0, then added to the total in slice order):
      // head2_small_kernel splits the same slices over its warps, so both kernels give identical bits
      for (int s0 = 0; s0 < 450; s0 += kPolSlice) {
        float part[kH2B];
#pragma unroll
        for (int b = 0; b < kH2B; ++b) part[b] = 0.f;
        const int s1 = s0 + kPolSlice < 450 ? s0 + kPolSlice : 450;
        for (int k0 = s0; k0 < s1; k0 += 10) {
          float wv[10];                                   // ten weight loads in flight per thread
#pragma unroll
          for (int u = 0; u < 10; ++u) wv[u] = __ldg(a.pol_wt + (size_t)(k0 + u) * 225 + tid);
#pragma unroll
          for (int u = 0; u < 10; ++u) {
            const float w = wv[u];
            const float4* h = reinterpret_cast<const float4*>(s_h + (k0 + u) * kH2B);
#pragma unroll
            for (int q = 0; q < kH2B / 4; ++q) {
              const float4 hv = h[q];
              part[4 * q] = fmaf(w, hv.x, part[4 * q]); part[4 * q + 1] = fmaf(w, hv.y, part[4 * q + 1]);
              part[4 * q + 2] = fmaf(w, hv.z, part[4 * q + 2]); part[4 * q + 3] = fmaf(w, hv.w, part[4 * q + 3]);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < kH2B; ++b) acc[b] += part[b];
      }
#pragma unroll
      for (int b = 0; b < kH2B; ++b) s_lg[b * 228 + tid] = acc[b];
    }
    // value_fc1 + ReLU (network.py:113)
    if (tid < 64) {
      float acc[kH2B];
      const float bias = a.v1_b[tid];
#pragma unroll
      for (int b = 0; b < kH2B; ++b) acc[b] = bias;
      for (int s0 = 0; s0 < 225; s0 += kValSlice) {
        float part[kH2B];
#pragma unroll
        for (int b = 0; b < kH2B; ++b) part[b] = 0.f;
        const int s1 = s0 + kValSlice < 225 ? s0 + kValSlice : 225;
        for (int k0 = s0; k0 < s1; k0 += 5) {
          float wv[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) wv[u] = __ldg(a.v1_wt + (size_t)(k0 + u) * 64 + tid);
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const float w = wv[u];
            const float4* h = reinterpret_cast<const float4*>(s_h + (450 + k0 + u) * kH2B);
#pragma unroll
            for (int q = 0; q < kH2B / 4; ++q) {
              const float4 hv = h[q];
              part[4 * q] = fmaf(w, hv.x, part[4 * q]); part[4 * q + 1] = fmaf(w, hv.y, part[4 * q + 1]);
              part[4 * q + 2] = fmaf(w, hv.z, part[4 * q + 2]); part[4 * q + 3] = fmaf(w, hv.w, part[4 * q + 3]);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < kH2B; ++b) acc[b] += part[b];
      }
#pragma unroll
      for (int b = 0; b < kH2B; ++b) s_v[b * 64 + tid] = fmaxf(acc[b], 0.f);
    }
    __syncthreads();
    // softmax over all 225 logits (network.py:180), 2 boards per warp
    for (int bb = 0; bb < kH2B / 8; ++bb) {
      const int b = warp * (kH2B / 8) + bb, gb = b0 + b;
      if (gb >= n) break;
      float v[8], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = lane + 32 * j;
        v[j] = o < 225 ? s_lg[b * 228 + o] : -INFINITY;
        mx = fmaxf(mx, v[j]);
      }
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (a.logits && lane + 32 * j < 225) a.logits[(size_t)gb * 225 + lane + 32 * j] = v[j];
        v[j] = lane + 32 * j < 225 ? expf(v[j] - mx) : 0.f;
        sum += v[j];
      }
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < 225) a.probs[(size_t)gb * 225 + lane + 32 * j] = v[j] * inv;
    }
    // value_fc2 + tanh (network.py:114-115)
    if (a.values && tid < kH2B && b0 + tid < n) {
      float acc = a.v2_b[0];
      for (int o = 0; o < 64; ++o) acc = fmaf(s_v[tid * 64 + o], a.v2_w[o], acc);
      a.values[b0 + tid] = tanhf(acc);
    }
  }
}

// Small batches (the single-game path: <= 32 leaves per round): two boards per block, the K slices of the
// dense layers spread over the eight warps, so a round pays one slice of dependent weight loads instead of
// eight.  Same slices, same order of additions as head2_kernel: identical results.
constexpr int kSB = 2;
constexpr int kHead2SmallMax = 296;     // up to one block per SM pair slot: beyond that head2_kernel's 16-board tiles win
__global__ void __launch_bounds__(256) head2_small_kernel(HeadArgs a) {
  __shared__ float s_h[AZG_HID * kSB];             // [k][board]
  __shared__ float s_part[8][kSB][228];
  __shared__ float s_vpart[8][kSB][64];
  __shared__ float s_lg[kSB][228];
  __shared__ float s_v[kSB][64];
  int n = *a.n_boards;
  if (n > a.max_boards) n = a.max_boards;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_units = (n + kSB - 1) / kSB;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int b0 = unit * kSB;
    __syncthreads();
    for (int i = tid; i < AZG_HID * kSB; i += 256) {
      const int k = i / kSB, b = b0 + i % kSB;
      s_h[i] = a.hidden[(size_t)(b >> 5) * (AZG_HID * 32) + (size_t)k * 32 + (b & 31)];
    }
    __syncthreads();
    {   // policy_fc slice `warp`: outputs lane + 32 j
      float part[8][kSB];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int b = 0; b < kSB; ++b) part[j][b] = 0.f;
      const int s0 = warp * kPolSlice, s1 = s0 + kPolSlice < 450 ? s0 + kPolSlice : 450;
#pragma unroll 2
      for (int k = s0; k < s1; ++k) {
        const float* wrow = a.pol_wt + (size_t)k * 225;
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = lane + 32 * j < 225 ? __ldg(wrow + lane + 32 * j) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int b = 0; b < kSB; ++b) part[j][b] = fmaf(wv[j], s_h[k * kSB + b], part[j][b]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < 225)
#pragma unroll
          for (int b = 0; b < kSB; ++b) s_part[warp][b][lane + 32 * j] = part[j][b];
    }
    {   // value_fc1 slice `warp`: outputs lane, lane + 32
      float part[2][kSB];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int b = 0; b < kSB; ++b) part[j][b] = 0.f;
      const int s0 = warp * kValSlice, s1 = s0 + kValSlice < 225 ? s0 + kValSlice : 225;
#pragma unroll 5
      for (int k = s0; k < s1; ++k) {
        const float w0 = __ldg(a.v1_wt + (size_t)k * 64 + lane), w1 = __ldg(a.v1_wt + (size_t)k * 64 + 32 + lane);
#pragma unroll
        for (int b = 0; b < kSB; ++b) {
          part[0][b] = fmaf(w0, s_h[(450 + k) * kSB + b], part[0][b]);
          part[1][b] = fmaf(w1, s_h[(450 + k) * kSB + b], part[1][b]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int b = 0; b < kSB; ++b) s_vpart[warp][b][lane + 32 * j] = part[j][b];
    }
    __syncthreads();
    if (tid < 225) {
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
        float acc = a.pol_b[tid];
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_part[w][b][tid];
        s_lg[b][tid] = acc;
      }
    }
    if (tid < 64) {
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
        float acc = a.v1_b[tid];
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_vpart[w][b][tid];
        s_v[b][tid] = fmaxf(acc, 0.f);
      }
    }
    __syncthreads();
    if (warp < kSB && b0 + warp < n) {          // softmax over all 225 logits (network.py:180), one board per warp
      const int gb = b0 + warp;
      float v[8], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = lane + 32 * j;
        v[j] = o < 225 ? s_lg[warp][o] : -INFINITY;
        mx = fmaxf(mx, v[j]);
      }
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (a.logits && lane + 32 * j < 225) a.logits[(size_t)gb * 225 + lane + 32 * j] = v[j];
        v[j] = lane + 32 * j < 225 ? expf(v[j] - mx) : 0.f;
        sum += v[j];
      }
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < 225) a.probs[(size_t)gb * 225 + lane + 32 * j] = v[j] * inv;
    }
    if (a.values && tid >= 64 && tid < 64 + kSB && b0 + (tid - 64) < n) {     // value_fc2 + tanh (network.py:114-115)
      const int b = tid - 64;
      float acc = a.v2_b[0];
      for (int o = 0; o < 64; ++o) acc = fmaf(s_v[b][o], a.v2_w[o], acc);
      a.values[b0 + b] = tanhf(acc);
    }
  }
}

int azg_heads_launch(int C, const HeadArgs& a, int n_sm, cudaStream_t stream, bool skip_head1) {
  int grid = n_sm * 4;
  if (grid > a.max_boards) grid = a.max_boards;
  if (grid < 1) grid = 1;
  int rc = AZG_OK;
  if (!skip_head1) {          // the 1x1 convs are normally fused into the last trunk layer's epilogue
    if (C == 64) head1_kernel<64><<<grid, 256, 0, stream>>>(a);
    else if (C == 128) head1_kernel<128><<<grid, 256, 0, stream>>>(a);
    else if (C == 256) head1_kernel<256><<<grid, 256, 0, stream>>>(a);
    else return azg_fail(AZG_E_ARG, "heads: channels must be 64, 128 or 256");
    rc = azg_check_launch("head1_kernel");
    if (rc) return rc;
  }
  if (a.max_boards <= kHead2SmallMax) {
    const int units = (a.max_boards + kSB - 1) / kSB;
    head2_small_kernel<<<units, 256, 0, stream>>>(a);
    return azg_check_launch("head2_small_kernel");
  }
  cudaError_t e = cudaFuncSetAttribute(head2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHead2Smem);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  int tiles = (a.max_boards + kH2B - 1) / kH2B;
  int grid2 = tiles < 3 * n_sm ? tiles : 3 * n_sm;
  head2_kernel<<<grid2, 256, kHead2Smem, stream>>>(a);
  return azg_check_launch("head2_kernel");
}

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
// precomputed in shared memory, so a pixel costs at most three 16-byte table reads instead of nine.
// The tables hold the DIFFERENCE to the all-empty in-board row (pattern 0b010101), and each lane keeps
// shift + "nine empty neighbours" in registers: a tap row without stones or border costs no read at all
// (the kernel is bound by shared-memory bandwidth, then by the HBM writes).
constexpr uint32_t kRowEmpty = 0x15u;
template <int CT>
__global__ void __launch_bounds__(512) stem_kernel(StemArgs a) {
  // CT channels in total; a block computes a slab of C = min(CT, 128) of them (blockIdx.y selects it)
  constexpr int C = CT > 128 ? 128 : CT;
  constexpr int CPL = C / 32;
  const int cbase = blockIdx.y * C;
  a.w += cbase; a.shift += cbase; a.out += cbase;
  extern __shared__ __align__(16) float s_rows[];   // [3][64][C]
  __shared__ __align__(16) float s_empty[C];
  // 512 threads: two boards at a time (threads 0-255 and 256-511) share one copy of the tables.  The kernel is bound by
  // instruction issue (ncu: 0.48 IPC per scheduler at 4 warps), and the 96 KB of tables allow only two blocks per SM,
  // so the second board per block is what doubles the resident warps.
  __shared__ uint8_t s_state[2][256];                // 0 off-board / pad, 1 empty, 2 mover stone, 3 opponent stone
  __shared__ uint32_t s_code[2][256];                // the nine neighbour states of each padded pixel, 2 bits per tap
  const int sub = threadIdx.x >> 8, tid = threadIdx.x & 255;
  for (int i = threadIdx.x; i < 3 * 64 * C; i += 512) {
    const int c = i % C, combo = (i / C) & 63, d = i / (64 * C);
    float v = 0.f, e = 0.f;
#pragma unroll
    for (int dc = 0; dc < 3; ++dc) {
      const int tap = d * 3 + dc;
      e += a.w[(tap * 3 + 2) * CT + c];
      const int st = (combo >> (2 * dc)) & 3;
      if (st == 0) continue;
      v += a.w[(tap * 3 + 2) * CT + c];
      if (st >= 2) v += a.w[(tap * 3 + (st - 2)) * CT + c];
    }
    s_rows[i] = v - e;
  }
  for (int i = threadIdx.x; i < C; i += 512) {
    float e = a.shift[i];
    for (int tap = 0; tap < 9; ++tap) e += a.w[(tap * 3 + 2) * CT + i];
    s_empty[i] = e;
  }
  __syncthreads();
  float base[CPL];                                     // shift + all nine taps on empty in-board cells
#pragma unroll
  for (int j = 0; j < CPL; ++j) base[j] = s_empty[(threadIdx.x & 31) * CPL + j];
  int n = *a.n_boards;
  if (n > a.max_boards) n = a.max_boards;
  const int warp = tid >> 5, lane = tid & 31;
  for (int pb = blockIdx.x; pb * 2 < n; pb += gridDim.x) {
    const int b = pb * 2 + sub;
    const bool live = b < n;
    size_t node = (size_t)(live ? b : n - 1);
    if (a.leaf_game) node = (size_t)a.leaf_game[live ? b : n - 1] * (size_t)a.slab_stride + (size_t)a.leaf_node[live ? b : n - 1];
    const uint32_t* k = a.keys + node * 16;
    const int player = (int)((a.meta[node] >> 1) & 3u);
    __syncthreads();
    {
      const int qi = tid, y = qi >> 4, x = qi & 15;
      uint8_t st = 0;
      if (y >= 1 && x < 15) {
        const int cell = (y - 1) * 15 + x;
        const uint32_t b1 = (k[cell >> 5] >> (cell & 31)) & 1u, b2 = (k[8 + (cell >> 5)] >> (cell & 31)) & 1u;
        const uint32_t mine = player == 2 ? b2 : b1, theirs = player == 2 ? b1 : b2;
        st = mine ? 2 : (theirs ? 3 : 1);
      }
      s_state[sub][qi] = st;
    }
    __syncthreads();
    {
      const int qi = tid;
      uint32_t code = 0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int q2 = qi + (tap / 3 - 1) * 16 + (tap % 3 - 1);
        const uint32_t st = (q2 >= 0 && q2 < 256) ? s_state[sub][q2] : 0;
        code |= st << (2 * tap);
      }
      s_code[sub][qi] = (live && s_state[sub][qi]) ? code : 0xffffffffu;      // pad rows/columns (and the odd board out)
    }
    __syncthreads();
    __nv_bfloat16* out = a.out + ((size_t)AZG_NET_FRONT + (size_t)b * 256) * CT;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const int qi = warp * 32 + i;
      const uint32_t code = s_code[sub][qi];
      // pad rows / columns are not written: the buffers are zeroed at creation and every writer (this kernel, the
      // trunk epilogue) keeps the pad positions zero, so they still are (12 % fewer bytes to HBM)
      if (code == 0xffffffffu) continue;
      float acc[CPL];
      {
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = base[j];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const uint32_t pat = (code >> (6 * d)) & 63u;
          if (pat != kRowEmpty) {                          // warp-uniform: the code belongs to the pixel
            const float* w = s_rows + (size_t)(d * 64 + pat) * C + lane * CPL;
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[j] += w[j];
          }
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
  stem_kernel<C><<<dim3(grid, C / CB), 512, smem, stream>>>(a);
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
// Output feeds the dense part (net_heads.cu): hidden[b][k], k = ch*225 + pixel (policy, the
// reference's view(B, 2*225) order, network.py:105) then 450 + pixel (value).
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
    float* hid = a.hidden + (size_t)b * AZG_HEAD_PITCH;
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
          hid[pix] = fmaxf(fmaf(d0, sc0, sh0), 0.f);
          hid[225 + pix] = fmaxf(fmaf(d1, sc1, sh1), 0.f);
          hid[450 + pix] = fmaxf(fmaf(d2, sc2, sh2), 0.f);
        }
      }
    }
  }
}

int azg_heads_launch(int C, const HeadArgs& a, const CUtensorMap& tm_hid, const CUtensorMap& tm_wp, const CUtensorMap& tm_wv,
                     int* error, int n_sm, cudaStream_t stream, bool skip_head1) {
  if (!skip_head1) {          // the 1x1 convs are normally fused into the last trunk layer's epilogue
    int grid = n_sm * 4;
    if (grid > a.max_boards) grid = a.max_boards;
    if (grid < 1) grid = 1;
    if (C == 64) head1_kernel<64><<<grid, 256, 0, stream>>>(a);
    else if (C == 128) head1_kernel<128><<<grid, 256, 0, stream>>>(a);
    else if (C == 256) head1_kernel<256><<<grid, 256, 0, stream>>>(a);
    else return azg_fail(AZG_E_ARG, "heads: channels must be 64, 128 or 256");
    const int rc = azg_check_launch("head1_kernel");
    if (rc) return rc;
  }
  return azg_heads_gemm_launch(tm_hid, tm_wp, tm_wv, a, error, n_sm, stream);
}

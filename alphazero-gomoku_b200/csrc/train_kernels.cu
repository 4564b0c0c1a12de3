// Training-step kernels around the tensor-core convolutions: training-mode BatchNorm (forward and
// backward), stem, policy / value heads with the loss, gradient-norm clip + Adam + bf16 repack.
//
// Replaces what torch autograd does for PyTorchModel.train_batch (network.py:199-235):
//   forward   network.py:94-117 with module.training = True (BatchNorm2d uses batch statistics and
//             updates its running statistics with momentum 0.1 / unbiased variance)
//   loss      KLDivLoss(batchmean)(log_softmax(logits), pi) + MSELoss(value, z)   (network.py:143-144, 217-222)
//   backward  of all of the above
//   update    clip_grad_norm_(3.0) (network.py:224) and Adam(lr, weight_decay) (network.py:141, 225)
//
// These kernels are HBM / L2 bound elementwise and reduction passes over the padded bf16 activation
// layout of net_conv.cu ([board][256 rows][C], pad rows zero); all arithmetic is fp32 (statistics are
// finished in fp64).  Per-channel reductions are deterministic: every block writes its partial sums and
// the last block to finish (ticket counter) adds them in block order.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include "train.h"
#include "ptx.cuh"

namespace {

constexpr int kEwThreads = 256;

__device__ __forceinline__ bool is_pad_row(int qi) { return qi < 16 || (qi & 15) == 15; }

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&p);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Ticket: true in exactly one block - the last one to arrive - after all blocks' global writes are visible.
__device__ __forceinline__ bool last_block_arrives(unsigned* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0;          // ready for the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// ------------------------------------------------------------------------------------------------
// trunk BatchNorm: batch statistics
// ------------------------------------------------------------------------------------------------
// Reduction kernels: thread = (row lane rl, 8-channel group cg); a block of kRedThreads covers RL rows per iteration.
// Every block leaves its per-channel partial sums in partial[output][block] (stride AZG_TRAIN_PARTIALS); the last
// block to finish adds them: one warp per output, lanes over the blocks, shuffle tree - a fixed order.
constexpr int kRedThreads = 512;

template <int C, bool BWD>
__device__ __forceinline__ void channel_reduce(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ g,
                                               const uint8_t* __restrict__ mask, const float* __restrict__ stats,
                                               int n_boards, float* __restrict__ partial) {
  constexpr int CG = C / 8, RL = kRedThreads / CG;
  __shared__ float red[RL][2][C];
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  float s0[8], s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
  float mean[8], rstd[8];
  if (BWD) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mean[i] = stats[cg * 8 + i]; rstd[i] = stats[C + cg * 8 + i]; }
  }
  const long long n_rows = (long long)n_boards * 256;
  const long long step = (long long)gridDim.x * RL;
  constexpr int U = 4;                                     // rows in flight per thread: the passes are latency bound otherwise
  for (long long r0 = (long long)blockIdx.x * RL + rl; r0 < n_rows; r0 += U * step) {
    uint4 zq[U], gq[U];
    unsigned mq[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      live[u] = r < n_rows && !is_pad_row((int)(r & 255));   // pad rows are zero in z and g: nothing to add, nothing to read
      if (live[u]) {
        const size_t off = ((size_t)AZG_NET_FRONT + (size_t)r) * C + (size_t)cg * 8;
        zq[u] = ptx::ldg128(z + off);
        if (BWD) { gq[u] = ptx::ldg128(g + off); mq[u] = mask[(size_t)r * CG + cg]; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      float zf[8];
      unpack8(zq[u], zf);
      if (!BWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += zf[i]; s1[i] = fmaf(zf[i], zf[i], s1[i]); }
      } else {
        float gf[8];
        unpack8(gq[u], gf);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dy = ((mq[u] >> i) & 1u) ? gf[i] : 0.f;
          s0[i] += dy;
          s1[i] = fmaf(dy, (zf[i] - mean[i]) * rstd[i], s1[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[rl][0][cg * 8 + i] = s0[i]; red[rl][1][cg * 8 + i] = s1[i]; }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += kRedThreads) {
    const int k = t / C, c = t % C;
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < RL; ++j) acc += red[j][k][c];
    partial[(size_t)t * AZG_TRAIN_PARTIALS + blockIdx.x] = acc;
  }
}

// Sum of partial[o][0 .. n_blocks) for the calling WARP's output o, in double, identical in every lane.
__device__ __forceinline__ double warp_sum_partials(const float* __restrict__ partial, int o, int n_blocks) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < n_blocks; b += 32) s += (double)partial[(size_t)o * AZG_TRAIN_PARTIALS + b];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  return s;
}
// The same for N outputs o0, o0 + stride, ... at once: all loads are issued before any is used (the tail of the
// last block is a chain of L2 round trips otherwise).
template <int N>
__device__ __forceinline__ void warp_sum_partials_n(const float* __restrict__ partial, int o0, int stride, int n_blocks, double (&out)[N]) {
  const int lane = threadIdx.x & 31;
  constexpr int R = (AZG_TRAIN_PARTIALS + 31) / 32;
  float v[N][R];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int b = lane + 32 * j;
      v[k][j] = b < n_blocks ? partial[(size_t)(o0 + k * stride) * AZG_TRAIN_PARTIALS + b] : 0.f;
    }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) s += (double)v[k][j];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    out[k] = s;
  }
}

template <int C>
__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(BnStatsArgs p) {
  channel_reduce<C, false>(p.z, nullptr, nullptr, nullptr, p.n_boards, p.partial);
  if (!last_block_arrives(p.counter)) return;
  const double n = (double)p.n_boards * 225.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = kRedThreads / 32, PER = C / NW;                 // channels per warp: 8 (C = 128) or 4
  for (int k = 0; k < PER; k += 2) {                                 // two channels (four sums) per round: few registers
    const int c0 = warp + k * NW;
    double sums[4];
    double sq[2];
    warp_sum_partials_n<2>(p.partial, c0, NW, (int)gridDim.x, reinterpret_cast<double(&)[2]>(sums[0]));       // sums of channels c0, c0 + NW
    warp_sum_partials_n<2>(p.partial, C + c0, NW, (int)gridDim.x, sq);                                        // their sums of squares
    if (lane != 0) continue;
    for (int j = 0; j < 2; ++j) {
    const int c = c0 + j * NW;
    const double s = sums[j], q = sq[j];
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    p.stats[c] = (float)mean;
    p.stats[C + c] = (float)(1.0 / sqrt(var + (double)p.eps));
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)mean;
    p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unbiased;
    }
  }
}

// The same finish for statistics that the convolution's epilogue collected (net_conv.cu STATS): partial is
// [2 * C][AZG_CONV_STAT_SLOTS], n_slots of each row are valid; warp = channel.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(BnStatsArgs p, int C, int n_slots) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int j = lane; j < n_slots; j += 32) {
    s += (double)p.partial[(size_t)c * AZG_CONV_STAT_SLOTS + j];
    q += (double)p.partial[(size_t)(C + c) * AZG_CONV_STAT_SLOTS + j];
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  if (lane != 0) return;
  const double n = (double)p.n_boards * 225.0;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  p.stats[c] = (float)mean;
  p.stats[C + c] = (float)(1.0 / sqrt(var + (double)p.eps));
  const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
  p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)mean;
  p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unbiased;
}

// BatchNorm backward sums from what the input-gradient convolution's epilogue collected (net_conv.cu STATS 2):
// partial rows [0, C) hold sum(dy), rows [C, 2C) sum(dy * z) per epilogue warp; warp = channel.
//   sum(dy * x_hat) = rstd * (sum(dy * z) - mean * sum(dy))
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(BnBwdArgs p, const float* __restrict__ conv_partial, int C, int n_slots) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int j = lane; j < n_slots; j += 32) {
    s += (double)conv_partial[(size_t)c * AZG_CONV_STAT_SLOTS + j];
    q += (double)conv_partial[(size_t)(C + c) * AZG_CONV_STAT_SLOTS + j];
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  if (lane != 0) return;
  const double xh = (double)p.stats[C + c] * (q - (double)p.stats[c] * s);
  p.sums[c] = (float)s;
  p.sums[C + c] = (float)xh;
  p.dbeta[c] = (float)s;
  p.dgamma[c] = (float)xh;
}

// a = relu(gamma * (z - mean) * rstd + beta (+ residual)), pad rows zero
template <int C>
__global__ void __launch_bounds__(kEwThreads)
bn_apply_kernel(BnApplyArgs p) {
  // the next kernel in the stream is the convolution that consumes `out`: let it set up its barriers, tensor memory
  // and resident weights on the side while this pass runs (it waits for this grid's completion before reading)
  ptx::grid_dep_launch();
  constexpr int CG = C / 8;
  const long long total = (long long)p.n_boards * 256 * CG;
  const long long stride = (long long)gridDim.x * kEwThreads;          // a multiple of CG: the channel group of a thread is fixed
  long long idx = (long long)blockIdx.x * kEwThreads + threadIdx.x;
  const int cg = (int)(idx % CG);
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cg * 8 + i;
    sc[i] = p.gamma[c] * p.stats[C + c];
    sh[i] = p.beta[c] - p.stats[c] * sc[i];
  }
  for (; idx < total; idx += stride) {
    const long long r = idx / CG;
    const size_t off = ((size_t)AZG_NET_FRONT + (size_t)r) * C + (size_t)cg * 8;
    float y[8];
    unsigned bits = 0;
    if (is_pad_row((int)(r & 255))) {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = 0.f;
    } else {
      float zf[8];
      unpack8(ptx::ldg128(p.z + off), zf);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf(zf[i], sc[i], sh[i]);
      if (p.residual) {
        float rf[8];
        unpack8(ptx::ldg128(p.residual + off), rf);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] += rf[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], 0.f);
    }
    const uint4 packed = pack8(y);
    ptx::stg128(p.out + off, packed);
    if (p.mask) {                         // the mask is taken from the ROUNDED output: exactly [a > 0]
      float yr[8];
      unpack8(packed, yr);
#pragma unroll
      for (int i = 0; i < 8; ++i) bits |= (yr[i] > 0.f ? 1u : 0u) << i;
      p.mask[idx] = (uint8_t)bits;
    }
  }
}

// backward: sums of dy and dy * x_hat per channel, dgamma / dbeta
template <int C>
__global__ void __launch_bounds__(kRedThreads)
bn_bwd_reduce_kernel(BnBwdArgs p) {
  channel_reduce<C, true>(p.z, p.g, p.mask, p.stats, p.n_boards, p.partial);
  if (!last_block_arrives(p.counter)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = kRedThreads / 32, PER = 2 * C / NW;
  for (int j0 = 0; j0 < PER; j0 += 4) {                               // four outputs per round: few registers
    double sums[4];
    warp_sum_partials_n<4>(p.partial, warp + j0 * NW, NW, (int)gridDim.x, sums);
    if (lane != 0) continue;
    for (int j = 0; j < 4; ++j) {
      const int t = warp + (j0 + j) * NW, k = t / C, c = t % C;
      p.sums[k * C + c] = (float)sums[j];
      if (k == 0) p.dbeta[c] = (float)sums[j]; else p.dgamma[c] = (float)sums[j];
    }
  }
}

// dz = gamma * rstd * (dy - mean(dy) - x_hat * mean(dy * x_hat)), pad rows zero; gskip = dy
template <int C>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_kernel(BnBwdArgs p) {
  ptx::grid_dep_launch();                 // the input-gradient convolution follows: see bn_apply_kernel
  constexpr int CG = C / 8;
  const long long total = (long long)p.n_boards * 256 * CG;
  const long long stride = (long long)gridDim.x * kEwThreads;
  long long idx = (long long)blockIdx.x * kEwThreads + threadIdx.x;
  const int cg = (int)(idx % CG);
  const float inv_n = 1.0f / ((float)p.n_boards * 225.0f);
  float mean[8], rstd[8], k0[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cg * 8 + i;
    mean[i] = p.stats[c];
    rstd[i] = p.stats[C + c];
    k0[i] = p.gamma[c] * rstd[i];
    m1[i] = p.sums[c] * inv_n;
    m2[i] = p.sums[C + c] * inv_n;
  }
  for (; idx < total; idx += stride) {
    const long long r = idx / CG;
    const size_t off = ((size_t)AZG_NET_FRONT + (size_t)r) * C + (size_t)cg * 8;
    float dz[8], dy[8];
    if (is_pad_row((int)(r & 255))) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = dy[i] = 0.f;
    } else {
      float zf[8], gf[8];
      const unsigned m = p.mask[idx];
      unpack8(ptx::ldg128(p.z + off), zf);
      unpack8(ptx::ldg128(p.g + off), gf);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dy[i] = ((m >> i) & 1u) ? gf[i] : 0.f;
        const float xh = (zf[i] - mean[i]) * rstd[i];
        dz[i] = k0[i] * (dy[i] - m1[i] - xh * m2[i]);
      }
    }
    ptx::stg128(p.dz + off, pack8(dz));
    if (p.gskip) ptx::stg128(p.gskip + off, pack8(dy));
  }
}

// ------------------------------------------------------------------------------------------------
// stem: conv 3x3, 3 planes -> C (network.py:94)
// ------------------------------------------------------------------------------------------------
// Block = one board at a time; thread = (pixel slice ps, channel c): its 27 weights sit in registers, the zero-padded
// planes (17 x 17) in shared memory are read as broadcasts, a pixel row leaves as C consecutive bf16.
constexpr int kStemSlices = 2;

__device__ __forceinline__ void stem_load_planes(float (*xs)[17][17], const float* __restrict__ planes, int b) {
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 17 * 17; i += blockDim.x) (&xs[0][0][0])[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 675; i += blockDim.x) {
    const int pl = i / 225, pix = i % 225;
    xs[pl][pix / 15 + 1][pix % 15 + 1] = planes[(size_t)b * 675 + i];
  }
  __syncthreads();
}

template <int C>
__global__ void __launch_bounds__(C * kStemSlices)
stem_train_fwd_kernel(StemTrainArgs p) {
  __shared__ float xs[3][17][17];
  const int c = threadIdx.x % C, ps = threadIdx.x / C;
  float w[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) w[i] = p.w[c * 27 + i];
  for (int b = blockIdx.x; b < p.n_boards; b += gridDim.x) {
    stem_load_planes(xs, p.planes, b);
    for (int r = ps; r < 15; r += kStemSlices) {          // a board row per iteration; a 3-column window slides along it
      float c0[9], c1[9], c2[9];                            // [plane * 3 + kernel row]: columns cc-1, cc, cc+1 (padded coordinates)
#pragma unroll
      for (int q = 0; q < 9; ++q) { c0[q] = xs[q / 3][r + q % 3][0]; c1[q] = xs[q / 3][r + q % 3][1]; }
      __nv_bfloat16* zr = p.z + ((size_t)AZG_NET_FRONT + (size_t)b * 256 + (size_t)((r + 1) * 16)) * C + c;
#pragma unroll 3
      for (int cc = 0; cc < 15; ++cc) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          c2[q] = xs[q / 3][r + q % 3][cc + 2];
          const int wi = (q / 3) * 9 + (q % 3) * 3;       // weight index of (plane, kernel row, kernel column 0)
          acc = fmaf(c0[q], w[wi], acc); acc = fmaf(c1[q], w[wi + 1], acc); acc = fmaf(c2[q], w[wi + 2], acc);
        }
        zr[(size_t)cc * C] = __float2bfloat16_rn(acc);
#pragma unroll
        for (int q = 0; q < 9; ++q) { c0[q] = c1[q]; c1[q] = c2[q]; }
      }
    }
  }
}

// dW[plane*9+tap][c] = sum over boards, pixels of dz[pixel][c] * x[plane][pixel + tap].  One board per block,
// thread = (pixel slice, channel); the slices are combined in shared memory and the block's 27 x C sums go to
// partial[board][27 * C].  stem_reduce_kernel then adds the boards: thread = output, 64 boards per block in a fixed
// order, one floating-point atomic per output and block of 64 boards (summation order across those few is not fixed -
// like the tensor-core weight-gradient kernel).
template <int C>
__global__ void __launch_bounds__(256)
stem_train_wgrad_kernel(StemTrainArgs p) {
  constexpr int kStemWSlices = 256 / C;
  __shared__ float xs[3][17][17];
  __shared__ float red[kStemWSlices][27][C];
  const int c = threadIdx.x % C, ps = threadIdx.x / C;
  const int b = blockIdx.x;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  stem_load_planes(xs, p.planes, b);
  const __nv_bfloat16* dzb = p.dz + ((size_t)AZG_NET_FRONT + (size_t)b * 256) * C + c;
  for (int r = ps; r < 15; r += kStemWSlices) {            // a board row per iteration: its 15 gradients are fetched at once
    float d[15];
#pragma unroll
    for (int cc = 0; cc < 15; ++cc) d[cc] = __bfloat162float(dzb[(size_t)((r + 1) * 16 + cc) * C]);
    float c0[9], c1[9], c2[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { c0[q] = xs[q / 3][r + q % 3][0]; c1[q] = xs[q / 3][r + q % 3][1]; }
#pragma unroll
    for (int cc = 0; cc < 15; ++cc) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        c2[q] = xs[q / 3][r + q % 3][cc + 2];
        const int wi = (q / 3) * 9 + (q % 3) * 3;
        acc[wi] = fmaf(d[cc], c0[q], acc[wi]); acc[wi + 1] = fmaf(d[cc], c1[q], acc[wi + 1]); acc[wi + 2] = fmaf(d[cc], c2[q], acc[wi + 2]);
      }
#pragma unroll
      for (int q = 0; q < 9; ++q) { c0[q] = c1[q]; c1[q] = c2[q]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) red[ps][i][c] = acc[i];
  __syncthreads();
  for (int t = threadIdx.x; t < 27 * C; t += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kStemWSlices; ++k) v += (&red[k][0][0])[t];
    p.partial[(size_t)b * 27 * C + t] = v;
  }
}

__global__ void __launch_bounds__(128)
stem_reduce_kernel(const float* __restrict__ partial, int n_boards, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b0 = blockIdx.y * 64, b1 = min(n_boards, b0 + 64);
  float acc = 0.f;
#pragma unroll 8
  for (int b = b0; b < b1; ++b) acc += partial[(size_t)b * n + i];
  atomicAdd(out + i, acc);
}

// ------------------------------------------------------------------------------------------------
// heads, forward
// ------------------------------------------------------------------------------------------------
// 1x1 convolutions C -> 2 (policy) + 1 (value): warp per pixel row, lanes over channels.
template <int C>
__global__ void __launch_bounds__(256)
head_conv_fwd_kernel(HeadTrainArgs p) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  constexpr int PER = C / 32;              // channels per lane
  float w[3][PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    w[0][i] = p.w1p[lane * PER + i];
    w[1][i] = p.w1p[C + lane * PER + i];
    w[2][i] = p.w1v[lane * PER + i];
  }
  const int total = p.n_boards * 225;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += warps) {
    const int b = item / 225, pix = item % 225;
    const size_t row = (size_t)AZG_NET_FRONT + (size_t)b * 256 + (size_t)((pix / 15 + 1) * 16 + pix % 15);
    const __nv_bfloat16* src = p.act + row * C + lane * PER;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float x = __bfloat162float(src[i]);
      d0 = fmaf(x, w[0][i], d0); d1 = fmaf(x, w[1][i], d1); d2 = fmaf(x, w[2][i], d2);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, s);
      d1 += __shfl_xor_sync(0xffffffffu, d1, s);
      d2 += __shfl_xor_sync(0xffffffffu, d2, s);
    }
    if (lane == 0) {
      float* o = p.zh + (size_t)b * 675 + pix;
      o[0] = d0; o[225] = d1; o[450] = d2;
    }
  }
}

// deterministic block sum of two doubles (any block size up to 1024)
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[32], sb[32];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); b += __shfl_xor_sync(0xffffffffu, b, s); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  a = lane < nw ? sa[lane] : 0.0;
  b = lane < nw ? sb[lane] : 0.0;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); b += __shfl_xor_sync(0xffffffffu, b, s); }
}

// BatchNorm2d(2) / BatchNorm2d(1) of the heads in training mode: batch statistics of the three 1x1-conv channels
// (0, 1 policy; 2 value).  Blocks sum slices of the batch, the last block finishes (warp h = channel h).  The
// normalisation + ReLU itself is applied by head_fc_fwd_kernel while it loads its boards.
__global__ void __launch_bounds__(512)
head_stats_kernel(HeadTrainArgs p) {
  const int n = p.n_boards * 225;
  double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* zr = p.zh + (size_t)(i / 225) * 675 + i % 225;
#pragma unroll
    for (int h = 0; h < 3; ++h) { const float z = zr[h * 225]; acc[h][0] += z; acc[h][1] += (double)z * z; }
  }
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    block_sum2(acc[h][0], acc[h][1]);
    if (threadIdx.x == 0) {
      p.partial[(size_t)(h * 2) * AZG_TRAIN_PARTIALS + blockIdx.x] = (float)acc[h][0];
      p.partial[(size_t)(h * 2 + 1) * AZG_TRAIN_PARTIALS + blockIdx.x] = (float)acc[h][1];
    }
  }
  if (!last_block_arrives(p.counter)) return;
  const int h = threadIdx.x >> 5;
  if (h >= 3) return;
  const double s = warp_sum_partials(p.partial, h * 2, (int)gridDim.x), q = warp_sum_partials(p.partial, h * 2 + 1, (int)gridDim.x);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const int which = h < 2 ? 0 : 1, ch = h < 2 ? h : 0;
  p.hstats[h * 2] = (float)mean;
  p.hstats[h * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
  const double unbiased = n > 1 ? var * n / (n - 1.0) : var;
  p.bn_rmean[which][ch] = (1.f - p.momentum) * p.bn_rmean[which][ch] + p.momentum * (float)mean;
  p.bn_rvar[which][ch] = (1.f - p.momentum) * p.bn_rvar[which][ch] + p.momentum * (float)unbiased;
}

// dense layers + loss, NB boards per block: logits = hidden_p Wp^T + bp; h1 = relu(hidden_v Wv1^T + bv1);
// value = tanh(h1 w2 + b2); KL row sum, squared error, dlogits, dvpre.
constexpr int kFcBoards = 2;
__global__ void __launch_bounds__(256)
head_fc_fwd_kernel(HeadTrainArgs p) {
  __shared__ float hs[kFcBoards][675];
  __shared__ float h1s[kFcBoards][64];
  __shared__ float red[kFcBoards][8];
  __shared__ float red2[kFcBoards][8];
  const int b0 = blockIdx.x * kFcBoards;
  const int nb = min(kFcBoards, p.n_boards - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kFcBoards * 675; i += blockDim.x) {      // hidden = relu(bn(zh)), kept for the backward pass
    const int k = i / 675, f = i % 675, h = f / 225;
    float v = 0.f;
    if (k < nb) {
      const int which = h < 2 ? 0 : 1, ch = h < 2 ? h : 0;
      const float sc = p.bn_gamma[which][ch] * p.hstats[h * 2 + 1], sh = p.bn_beta[which][ch] - p.hstats[h * 2] * sc;
      v = fmaxf(fmaf(p.zh[(size_t)(b0 + k) * 675 + f], sc, sh), 0.f);
      p.hidden[(size_t)(b0 + k) * 675 + f] = v;
    }
    hs[k][f] = v;
  }
  __syncthreads();
  float logit[kFcBoards];
  const int j = tid;
  if (j < 225) {
#pragma unroll
    for (int k = 0; k < kFcBoards; ++k) logit[k] = p.bp[j];
    for (int i0 = 0; i0 < 450; i0 += 30) {      // 30 weight loads in flight per thread: the loop is a chain of L2 round trips otherwise
      float w[30];
#pragma unroll
      for (int u = 0; u < 30; ++u) w[u] = p.wp_t[(size_t)(i0 + u) * 225 + j];
#pragma unroll
      for (int u = 0; u < 30; ++u)
#pragma unroll
        for (int k = 0; k < kFcBoards; ++k) logit[k] = fmaf(hs[k][i0 + u], w[u], logit[k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < kFcBoards; ++k) logit[k] = -INFINITY;
  }
  if (tid < 64) {
    float acc[kFcBoards];
#pragma unroll
    for (int k = 0; k < kFcBoards; ++k) acc[k] = p.bv1[tid];
    for (int i0 = 0; i0 < 225; i0 += 25) {
      float w[25];
#pragma unroll
      for (int u = 0; u < 25; ++u) w[u] = p.wv1_t[(size_t)(i0 + u) * 64 + tid];
#pragma unroll
      for (int u = 0; u < 25; ++u)
#pragma unroll
        for (int k = 0; k < kFcBoards; ++k) acc[k] = fmaf(hs[k][450 + i0 + u], w[u], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kFcBoards; ++k) {
      h1s[k][tid] = fmaxf(acc[k], 0.f);
      if (k < nb) p.h1[(size_t)(b0 + k) * 64 + tid] = h1s[k][tid];
    }
  }
  // row maxima
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    float m = logit[k];
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane == 0) red[k][warp] = m;
  }
  __syncthreads();
  float mx[kFcBoards], ex[kFcBoards];
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    float m = red[k][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[k][w]);
    mx[k] = m;
    ex[k] = j < 225 ? expf(logit[k] - m) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    float s = ex[k];
#pragma unroll
    for (int t = 16; t >= 1; t >>= 1) s += __shfl_xor_sync(0xffffffffu, s, t);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  const float inv_b = 1.0f / (float)p.n_boards;
  float klp[kFcBoards], pis[kFcBoards], prob[kFcBoards], tgt[kFcBoards];
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    float se = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) se += red[k][w];
    klp[k] = pis[k] = prob[k] = tgt[k] = 0.f;
    if (j < 225 && k < nb) {
      const float logp = logit[k] - mx[k] - logf(se);
      const float t = p.pi[(size_t)(b0 + k) * 225 + j];
      tgt[k] = t;
      prob[k] = ex[k] / se;
      klp[k] = t > 0.f ? t * (logf(t) - logp) : 0.f;
      pis[k] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    float a = klp[k], c = pis[k];
#pragma unroll
    for (int t = 16; t >= 1; t >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, t); c += __shfl_xor_sync(0xffffffffu, c, t); }
    if (lane == 0) { red[k][warp] = a; red2[k][warp] = c; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kFcBoards; ++k) {
    if (k >= nb) continue;
    float kl = 0.f, sp = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { kl += red[k][w]; sp += red2[k][w]; }
    if (j < 225) p.dlogits[(size_t)(b0 + k) * 225 + j] = (prob[k] * sp - tgt[k]) * inv_b;
    if (warp == k) {                        // value head of board k: warp k
      float v = h1s[k][lane] * p.w2[lane] + h1s[k][lane + 32] * p.w2[lane + 32];
#pragma unroll
      for (int t = 16; t >= 1; t >>= 1) v += __shfl_xor_sync(0xffffffffu, v, t);
      if (lane == 0) {
        const float val = tanhf(v + p.b2[0]);
        const float err = val - p.zt[b0 + k];
        p.value[b0 + k] = val;
        p.dvpre[b0 + k] = 2.f * err * inv_b * (1.f - val * val);
        p.loss_parts[(size_t)(b0 + k) * 2] = kl;
        p.loss_parts[(size_t)(b0 + k) * 2 + 1] = err * err;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// heads, backward
// ------------------------------------------------------------------------------------------------
// gradient at the head BatchNorm outputs (ReLU mask applied): one board per block
__global__ void __launch_bounds__(512)
head_fc_bwd_data_kernel(HeadTrainArgs p) {
  __shared__ float dl[225];
  __shared__ float dh1[64];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 225) dl[tid] = p.dlogits[(size_t)b * 225 + tid];
  if (tid >= 256 && tid < 320) {
    const int t = tid - 256;
    dh1[t] = p.h1[(size_t)b * 64 + t] > 0.f ? p.dvpre[b] * p.w2[t] : 0.f;
  }
  __syncthreads();
  if (tid < 450) {
    float acc = 0.f;
    for (int j0 = 0; j0 < 225; j0 += 25) {        // 25 weight loads in flight (see head_fc_fwd_kernel)
      float w[25];
#pragma unroll
      for (int u = 0; u < 25; ++u) w[u] = p.wp[(size_t)(j0 + u) * 450 + tid];
#pragma unroll
      for (int u = 0; u < 25; ++u) acc = fmaf(dl[j0 + u], w[u], acc);
    }
    const size_t o = (size_t)b * 675 + tid;
    p.dhid[o] = p.hidden[o] > 0.f ? acc : 0.f;
  }
  if (tid < 225) {
    float acc = 0.f;
    for (int t0 = 0; t0 < 64; t0 += 32) {
      float w[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) w[u] = p.wv1[(size_t)(t0 + u) * 225 + tid];
#pragma unroll
      for (int u = 0; u < 32; ++u) acc = fmaf(dh1[t0 + u], w[u], acc);
    }
    const size_t o = (size_t)b * 675 + 450 + tid;
    p.dhid[o] = p.hidden[o] > 0.f ? acc : 0.f;
  }
}

// weight gradients of the dense layers: thread = output element, blockIdx.y = slice of the batch (a thread that walked
// the whole batch alone was bound by 512+ dependent L2 round trips); the slices meet in floating-point atomics on the
// zeroed gradient vector.
constexpr int kFcSplit = 16;
__global__ void __launch_bounds__(256)
head_fc_wgrad_kernel(HeadTrainArgs p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (p.n_boards + kFcSplit - 1) / kFcSplit;
  const int b0 = blockIdx.y * per, b1 = min(p.n_boards, b0 + per);
  if (b0 >= b1) return;
  constexpr int N_WP = 225 * 450, N_BP = 225, N_WV = 64 * 225, N_BV = 64, N_W2 = 64;
  if (i < N_WP) {
    const int j = i / 450, k = i % 450;
    float acc = 0.f;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) acc = fmaf(p.dlogits[(size_t)b * 225 + j], p.hidden[(size_t)b * 675 + k], acc);
    atomicAdd(p.d_wp + i, acc);
    return;
  }
  int r = i - N_WP;
  if (r < N_BP) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc += p.dlogits[(size_t)b * 225 + r];
    atomicAdd(p.d_bp + r, acc);
    return;
  }
  r -= N_BP;
  if (r < N_WV) {
    const int t = r / 225, k = r % 225;
    float acc = 0.f;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) {
      const float d = p.h1[(size_t)b * 64 + t] > 0.f ? p.dvpre[b] * p.w2[t] : 0.f;
      acc = fmaf(d, p.hidden[(size_t)b * 675 + 450 + k], acc);
    }
    atomicAdd(p.d_wv1 + r, acc);
    return;
  }
  r -= N_WV;
  if (r < N_BV) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc += p.h1[(size_t)b * 64 + r] > 0.f ? p.dvpre[b] * p.w2[r] : 0.f;
    atomicAdd(p.d_bv1 + r, acc);
    return;
  }
  r -= N_BV;
  if (r < N_W2) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc = fmaf(p.dvpre[b], p.h1[(size_t)b * 64 + r], acc);
    atomicAdd(p.d_w2 + r, acc);
    return;
  }
  r -= N_W2;
  if (r == 0) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc += p.dvpre[b];
    atomicAdd(p.d_b2, acc);
  }
}

// head BatchNorm backward, reduction part: sums of dy and dy * x_hat per head channel -> hsums, dgamma, dbeta
__global__ void __launch_bounds__(512)
head_bn_bwd_reduce_kernel(HeadTrainArgs p) {
  const int n = p.n_boards * 225;
  double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  float mean[3], rstd[3];
#pragma unroll
  for (int h = 0; h < 3; ++h) { mean[h] = p.hstats[h * 2]; rstd[h] = p.hstats[h * 2 + 1]; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)(i / 225) * 675 + i % 225;
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      const float dy = p.dhid[o + h * 225];
      acc[h][0] += dy;
      acc[h][1] += (double)dy * ((p.zh[o + h * 225] - mean[h]) * rstd[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    block_sum2(acc[h][0], acc[h][1]);
    if (threadIdx.x == 0) {
      p.partial[(size_t)(h * 2) * AZG_TRAIN_PARTIALS + blockIdx.x] = (float)acc[h][0];
      p.partial[(size_t)(h * 2 + 1) * AZG_TRAIN_PARTIALS + blockIdx.x] = (float)acc[h][1];
    }
  }
  if (!last_block_arrives(p.counter)) return;
  const int h = threadIdx.x >> 5;
  if (h >= 3) return;
  const double s1 = warp_sum_partials(p.partial, h * 2, (int)gridDim.x), s2 = warp_sum_partials(p.partial, h * 2 + 1, (int)gridDim.x);
  if ((threadIdx.x & 31) != 0) return;
  const int which = h < 2 ? 0 : 1, ch = h < 2 ? h : 0;
  p.hsums[h * 2] = (float)s1;
  p.hsums[h * 2 + 1] = (float)s2;
  p.d_bn_gamma[which][ch] = (float)s2;
  p.d_bn_beta[which][ch] = (float)s1;
}

// 1x1 convolutions backward: g[row][c] = sum_h dzh[h][row] * w1[h][c] (pad rows zero) and
// dw1[h][c] = sum_rows dzh[h][row] * a[row][c] (block partials, last block adds them in order)
template <int C>
__global__ void __launch_bounds__(kEwThreads)
head_conv_bwd_kernel(HeadTrainArgs p) {
  constexpr int CG = C / 8, RL = kEwThreads / CG;
  __shared__ float red[RL][3][C];
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  float w[3][8], acc[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    w[0][i] = p.w1p[cg * 8 + i]; w[1][i] = p.w1p[C + cg * 8 + i]; w[2][i] = p.w1v[cg * 8 + i];
    acc[0][i] = acc[1][i] = acc[2][i] = 0.f;
  }
  float hk0[3], hm1[3], hm2[3], hmean[3], hrstd[3];
  {
    const float inv_n = 1.0f / ((float)p.n_boards * 225.0f);
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      const int which = h < 2 ? 0 : 1, ch = h < 2 ? h : 0;
      hmean[h] = p.hstats[h * 2]; hrstd[h] = p.hstats[h * 2 + 1];
      hk0[h] = p.bn_gamma[which][ch] * hrstd[h];
      hm1[h] = p.hsums[h * 2] * inv_n; hm2[h] = p.hsums[h * 2 + 1] * inv_n;
    }
  }
  const long long n_rows = (long long)p.n_boards * 256;
  for (long long r = (long long)blockIdx.x * RL + rl; r < n_rows; r += (long long)gridDim.x * RL) {
    const int qi = (int)(r & 255);
    const size_t off = ((size_t)AZG_NET_FRONT + (size_t)r) * C + (size_t)cg * 8;
    float g[8];
    if (is_pad_row(qi)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = 0.f;
    } else {
      const int pix = ((qi >> 4) - 1) * 15 + (qi & 15);
      const size_t ho = (size_t)(r >> 8) * 675 + pix;
      float dh[3];
#pragma unroll
      for (int h = 0; h < 3; ++h)         // head BatchNorm backward, applied on the fly
        dh[h] = hk0[h] * (p.dhid[ho + h * 225] - hm1[h] - (p.zh[ho + h * 225] - hmean[h]) * hrstd[h] * hm2[h]);
      const float d0 = dh[0], d1 = dh[1], d2 = dh[2];
      float af[8];
      unpack8(ptx::ldg128(p.act + off), af);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g[i] = d0 * w[0][i] + d1 * w[1][i] + d2 * w[2][i];
        acc[0][i] = fmaf(d0, af[i], acc[0][i]);
        acc[1][i] = fmaf(d1, af[i], acc[1][i]);
        acc[2][i] = fmaf(d2, af[i], acc[2][i]);
      }
    }
    ptx::stg128(p.g + off, pack8(g));
  }
#pragma unroll
  for (int h = 0; h < 3; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][h][cg * 8 + i] = acc[h][i];
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * C; t += kEwThreads) {
    const int h = t / C, c = t % C;
    float a = 0.f;
#pragma unroll 4
    for (int jx = 0; jx < RL; ++jx) a += red[jx][h][c];
    p.partial[(size_t)t * AZG_TRAIN_PARTIALS + blockIdx.x] = a;
  }
  if (!last_block_arrives(p.counter)) return;
  constexpr int NW = kEwThreads / 32, BATCH = 4;
  for (int t0 = threadIdx.x >> 5; t0 < 3 * C; t0 += NW * BATCH) {          // 3C / (8 warps * 4) = 12 (C = 128) or 6 rounds
    double sums[BATCH];
    warp_sum_partials_n<BATCH>(p.partial, t0, NW, (int)gridDim.x, sums);
    if ((threadIdx.x & 31) != 0) continue;
    for (int j = 0; j < BATCH; ++j) {
      const int t = t0 + j * NW, h = t / C, c = t % C;
      if (h < 2) p.d_w1p[h * C + c] = (float)sums[j]; else p.d_w1v[c] = (float)sums[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// optimiser
// ------------------------------------------------------------------------------------------------
// ||g / world||_2 over the flat gradient vector -> clip coefficient min(1, clip / (norm + 1e-6)) / world
__global__ void __launch_bounds__(256)
grad_norm_kernel(AdamArgs p) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const double g = (double)p.grads[i] * (double)p.inv_world;
    s += g * g;
  }
  double dummy = 0.0;
  block_sum2(s, dummy);
  if (threadIdx.x == 0) p.norm_partial[blockIdx.x] = (float)s;
  if (!last_block_arrives(p.counter)) return;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) tot += (double)p.norm_partial[b];
    const float norm = (float)sqrt(tot);
    float coef = p.clip / (norm + 1e-6f);
    if (coef > 1.f) coef = 1.f;
    p.scal[0] = coef * p.inv_world;
    p.scal[1] = norm;
    *p.step += 1;
  }
}

// One block = 1024 consecutive elements of one segment.  update: Adam on the fp32 master copy; always: refresh the
// bf16 tap-major convolution weights (forward and transposed / flipped for the input gradient) and the transposed
// dense matrices from the (new) master values.
__global__ void __launch_bounds__(256)
adam_pack_kernel(AdamArgs p, int update) {
  const int2 bs = p.block_seg[blockIdx.x];
  const AdamSeg seg = p.segs[bs.x];
  const int C = p.C;
  float coef = 0.f, step_size = 0.f, inv_sqrt_bc2 = 0.f;
  if (update) {
    coef = p.scal[0];
    const double t = (double)*p.step;
    const double bc1 = 1.0 - pow((double)p.b1, t), bc2 = 1.0 - pow((double)p.b2, t);
    step_size = (float)((double)p.lr / bc1);
    inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = bs.y + u * 256 + threadIdx.x;         // element within the segment (parameter order)
    if (e >= seg.count) continue;
    int ge = e;                                          // element within the segment in GRADIENT order
    int co = 0, ci = 0, tap = 0;
    if (seg.kind == SEG_CONV3) {
      co = e / (C * 9); ci = (e / 9) % C; tap = e % 9;
      ge = (tap * C + co) * C + ci;                     // wgrad kernel layout [tap][co][ci]
    } else if (seg.kind == SEG_STEM) {
      ge = (e % 27) * C + e / 27;                       // stem wgrad layout [27][C]
    }
    const long long pi = seg.off + e;
    float w = p.params[pi];
    if (update) {
      const float g = fmaf(p.grads[seg.off + ge], coef, p.wd * w);
      const float m = p.b1 * p.m[pi] + (1.f - p.b1) * g;
      const float v = p.b2 * p.v[pi] + (1.f - p.b2) * g * g;
      p.m[pi] = m; p.v[pi] = v;
      w -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + p.eps);
      p.params[pi] = w;
    }
    if (seg.kind == SEG_CONV3) {
      const __nv_bfloat16 wb16 = __float2bfloat16_rn(w);
      p.wf[((size_t)(seg.layer * 9 + tap) * C + co) * C + ci] = wb16;
      p.wb[((size_t)(seg.layer * 9 + (8 - tap)) * C + ci) * C + co] = wb16;
    } else if (seg.kind == SEG_WP) {
      p.wp_t[(size_t)(e % 450) * 225 + e / 450] = w;
    } else if (seg.kind == SEG_WV1) {
      p.wv1_t[(size_t)(e % 225) * 64 + e / 225] = w;
    }
  }
}

template <typename F>
int dispatch_c(int C, F&& f) {
  if (C == 64) return f(std::integral_constant<int, 64>{});
  if (C == 128) return f(std::integral_constant<int, 128>{});
  if (C == 256) return f(std::integral_constant<int, 256>{});
  return azg_fail(AZG_E_ARG, "training kernels: channels must be 64, 128 or 256");
}

int ew_grid(int n_boards, int C, int n_sm) {
  const long long total = (long long)n_boards * 256 * (C / 8);
  long long blocks = (total + kEwThreads - 1) / kEwThreads;
  const long long cap = (long long)n_sm * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

int azg_train_red_blocks() {
  static int v = 0;
  if (!v) {
    // the backward reduction needs 126 registers x 512 threads = one block per SM: more blocks than SMs would run as
    // a second wave on a few SMs and double the kernel's time (measured: 160 blocks 38 us, SMs idle half of it)
    const char* e = getenv("AZG_TRAIN_RED_BLOCKS");
    int dev = 0, n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    v = e ? atoi(e) : n_sm;
    if (v < 1) v = 1;
    if (v > AZG_TRAIN_PARTIALS) v = AZG_TRAIN_PARTIALS;
  }
  return v;
}

int azg_bn_stats_launch(int C, const BnStatsArgs& a, cudaStream_t s) {
  return dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    int grid = a.n_boards * 256 / (kRedThreads / (CC / 8));
    if (grid > azg_train_red_blocks()) grid = azg_train_red_blocks();
    if (grid < 1) grid = 1;
    bn_stats_kernel<CC><<<grid, kRedThreads, 0, s>>>(a);
    return azg_check_launch("bn_stats_kernel");
  });
}

int azg_bn_finalize_launch(int C, const BnStatsArgs& a, int n_slots, cudaStream_t s) {
  bn_finalize_kernel<<<(C * 32 + 255) / 256, 256, 0, s>>>(a, C, n_slots);
  return azg_check_launch("bn_finalize_kernel");
}

int azg_bn_apply_launch(int C, const BnApplyArgs& a, int n_sm, cudaStream_t s) {
  return dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    bn_apply_kernel<CC><<<ew_grid(a.n_boards, CC, n_sm), kEwThreads, 0, s>>>(a);
    return azg_check_launch("bn_apply_kernel");
  });
}

int azg_bn_bwd_finalize_launch(int C, const BnBwdArgs& a, const float* conv_partial, int n_slots, cudaStream_t s) {
  bn_bwd_finalize_kernel<<<(C * 32 + 255) / 256, 256, 0, s>>>(a, conv_partial, C, n_slots);
  return azg_check_launch("bn_bwd_finalize_kernel");
}

int azg_bn_bwd_reduce_launch(int C, const BnBwdArgs& a, cudaStream_t s) {
  return dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    int grid = a.n_boards * 256 / (kRedThreads / (CC / 8));
    if (grid > azg_train_red_blocks()) grid = azg_train_red_blocks();
    if (grid < 1) grid = 1;
    bn_bwd_reduce_kernel<CC><<<grid, kRedThreads, 0, s>>>(a);
    return azg_check_launch("bn_bwd_reduce_kernel");
  });
}

int azg_bn_bwd_apply_launch(int C, const BnBwdArgs& a, int n_sm, cudaStream_t s) {
  return dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    bn_bwd_apply_kernel<CC><<<ew_grid(a.n_boards, CC, n_sm), kEwThreads, 0, s>>>(a);
    return azg_check_launch("bn_bwd_apply_kernel");
  });
}

int azg_stem_train_fwd_launch(int C, const StemTrainArgs& a, cudaStream_t s) {
  return dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    stem_train_fwd_kernel<CC><<<a.n_boards < 1184 ? a.n_boards : 1184, CC * kStemSlices, 0, s>>>(a);
    return azg_check_launch("stem_train_fwd_kernel");
  });
}

int azg_stem_train_wgrad_launch(int C, const StemTrainArgs& a, cudaStream_t s) {       // a.dw must be zero (it is: the step clears the gradient vector)
  int rc = dispatch_c(C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    stem_train_wgrad_kernel<CC><<<a.n_boards, 256, 0, s>>>(a);
    return azg_check_launch("stem_train_wgrad_kernel");
  });
  if (rc) return rc;
  const int n = 27 * C;
  stem_reduce_kernel<<<dim3((n + 127) / 128, (a.n_boards + 63) / 64), 128, 0, s>>>(a.partial, a.n_boards, n, a.dw);
  return azg_check_launch("stem_reduce_kernel");
}

int azg_head_train_fwd_launch(const HeadTrainArgs& a, int n_sm, cudaStream_t s) {
  int rc = dispatch_c(a.C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    int grid = (a.n_boards * 225 + 7) / 8;
    if (grid > n_sm * 8) grid = n_sm * 8;
    head_conv_fwd_kernel<CC><<<grid, 256, 0, s>>>(a);
    return azg_check_launch("head_conv_fwd_kernel");
  });
  if (rc) return rc;
  int hgrid = (a.n_boards * 225 + 2047) / 2048;
  if (hgrid > azg_train_red_blocks()) hgrid = azg_train_red_blocks();
  head_stats_kernel<<<hgrid, 512, 0, s>>>(a);
  if ((rc = azg_check_launch("head_stats_kernel"))) return rc;
  head_fc_fwd_kernel<<<(a.n_boards + kFcBoards - 1) / kFcBoards, 256, 0, s>>>(a);
  return azg_check_launch("head_fc_fwd_kernel");
}

int azg_head_train_bwd_launch(const HeadTrainArgs& a, int n_sm, cudaStream_t s) {
  head_fc_bwd_data_kernel<<<a.n_boards, 512, 0, s>>>(a);
  int rc = azg_check_launch("head_fc_bwd_data_kernel");
  if (rc) return rc;
  constexpr int n_out = 225 * 450 + 225 + 64 * 225 + 64 + 64 + 1;
  head_fc_wgrad_kernel<<<dim3((n_out + 255) / 256, kFcSplit), 256, 0, s>>>(a);
  if ((rc = azg_check_launch("head_fc_wgrad_kernel"))) return rc;
  int hgrid = (a.n_boards * 225 + 2047) / 2048;
  if (hgrid > azg_train_red_blocks()) hgrid = azg_train_red_blocks();
  head_bn_bwd_reduce_kernel<<<hgrid, 512, 0, s>>>(a);
  if ((rc = azg_check_launch("head_bn_bwd_reduce_kernel"))) return rc;
  return dispatch_c(a.C, [&](auto c) {
    constexpr int CC = decltype(c)::value;
    int grid = a.n_boards * 256 / (kEwThreads / (CC / 8));
    const int cap = 2 * azg_train_red_blocks() < AZG_TRAIN_PARTIALS ? 2 * azg_train_red_blocks() : AZG_TRAIN_PARTIALS;      // 62 registers, 256 threads: two blocks per SM
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    head_conv_bwd_kernel<CC><<<grid, kEwThreads, 0, s>>>(a);
    return azg_check_launch("head_conv_bwd_kernel");
  });
}

int azg_grad_norm_launch(const AdamArgs& a, cudaStream_t s) {
  grad_norm_kernel<<<AZG_TRAIN_PARTIALS, 256, 0, s>>>(a);
  return azg_check_launch("grad_norm_kernel");
}

int azg_adam_launch(const AdamArgs& a, bool update, cudaStream_t s) {
  adam_pack_kernel<<<a.n_blocks, 256, 0, s>>>(a, update ? 1 : 0);
  return azg_check_launch("adam_pack_kernel");
}

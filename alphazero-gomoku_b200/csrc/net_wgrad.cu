// Weight gradient of a 3x3 convolution as a tcgen05 GEMM over the pixel dimension (sm_100a).
//
// Backward of the trunk convolutions of AlphaZeroNet (network.py:17-26, 94-99) for
// PyTorchModel.train_batch (network.py:199-235; the reference lets torch autograd do this):
//
//     dW[tap][co][ci] = sum over pixels p of  dz[p][co] * a[p + off(tap)][ci]
//
// Both tensors stay in the activation layout of the forward kernel (net_conv.cu): bf16,
// channels-last, 256 padded rows per board (row = y*16 + x, zero pad row / column), so a tap is
// again a constant ROW offset off = dr*16 + dc and every halo read hits a zero row.  The
// contraction index K is the ROW (pixel) index, i.e. both operands are "MN-major" for the tensor
// core (the M / N index - the channel - is the contiguous one): the instruction descriptor sets
// a_major = b_major = 1 and the shared-memory descriptors describe 64-channel x 8-row SWIZZLE_128B
// atoms (leading byte offset = distance between 64-channel slices, stride byte offset = 1024 B
// between 8-row groups), which is exactly what TMA writes for a {64 channels, R rows} box.
//
// Work split: one CTA = one kernel row dr (three taps dc = -1, 0, +1 share one 66-row tile of `a`)
// x one slice of the boards; three accumulators D[tap] = [128 x C] fp32 live in TMEM for the whole
// kernel and are added to the global gradient at the end with vectorised reductions
// (red.global.add.v4.f32).  cta_group::1, UMMA 128 x C x 16; warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 2-5 = epilogue.  With 64 channels M is still 128: the upper 64 accumulator rows read
// whatever follows the dz tile and are never stored.
//
// Algorithmic FLOPs per launch = 2 * 9 * C * C * 225 * boards; algorithmic bytes = the two tensors
// once (2 * boards * 225 * C * 2 B); each CTA row re-reads them, i.e. 3x from L2.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include "net.h"
#include "ptx.cuh"

namespace {

template <int C>
struct WCfg {
  static constexpr int KC = C / 64;                                  // 64-channel slices
  static constexpr int A_SLICE = 64 * 128;                           // dz tile of one slice: 64 rows x 128 B
  static constexpr int B_ROWS = 66;                                  // 64 rows + one halo row on each side (dc = -1 .. +1)
  static constexpr int B_SLICE = ((B_ROWS * 128) + 1023) & ~1023;    // 9216
  static constexpr int A_BYTES = KC * A_SLICE, B_BYTES = KC * B_SLICE;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = C == 128 ? 6 : 8;                     // 204 KB / 136 KB in flight: the tiles come from L2 or HBM with ~1-2 us of latency
  static constexpr int TMEM_COLS = 3 * C <= 256 ? 256 : 512;
  static constexpr int TX = KC * (64 * 128 + B_ROWS * 128);          // bytes TMA delivers per stage
  static constexpr int SMEM = 1024 + STAGES * STAGE + 256;
};

constexpr int kWThreads = 192;

// MN-major operand tile, SWIZZLE_128B: 64-element (128 B) groups along M/N `lbo` bytes apart, 8-row groups
// along K 1024 B apart (cute::UMMA make_umma_desc<Major::MN>, LayoutType::B128).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

enum { WERR_EMPTY = 11, WERR_FULL = 12, WERR_TFULL = 13 };

template <int C>
__global__ void __launch_bounds__(kWThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_a, WgradArgs p) {
  using K = WCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::STAGES * K::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + K::STAGES;
  uint64_t* tfull = bars + 2 * K::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tg = (int)blockIdx.x % 3;                 // kernel row: dr = tg - 1
  const int s = (int)blockIdx.x / 3, S = (int)gridDim.x / 3;
  const int n_boards = p.n_boards;
  int my_boards = 0;
  for (int b = s; b < n_boards; b += S) ++my_boards;
  const int n_chunks = my_boards * 4;
  // Tensors wider than the tile (256 channels, C = 128 tiles): blockIdx.y picks the (output-channel, input-channel) quadrant
  const int ldw = p.ct > 0 ? p.ct : C, nq = ldw / C;
  const int co0 = ((int)blockIdx.y / nq) * C, ci0 = ((int)blockIdx.y % nq) * C;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_dz); ptx::prefetch_tmap(&tm_a); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < K::STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
      ptx::mbar_init(tfull, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<1>(tmem_slot, K::TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; bool ok = true;
      for (int b = s; b < n_boards && ok; b += S)
        for (int q = 0; q < 4 && ok; ++q) {
          if (!ptx::mbar_wait(&empty[stage], phase ^ 1u)) { atomicExch(p.error, WERR_EMPTY); ok = false; break; }
          const int r0 = AZG_NET_FRONT + b * 256 + q * 64;
          uint8_t* sA = smem + stage * K::STAGE;
          uint8_t* sB = sA + K::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full[stage], (uint32_t)K::TX);
          for (int kc = 0; kc < K::KC; ++kc) {
            ptx::tma_load_2d(sA + kc * K::A_SLICE, &tm_dz, &full[stage], co0 + kc * 64, r0);
            ptx::tma_load_2d(sB + kc * K::B_SLICE, &tm_a, &full[stage], ci0 + kc * 64, r0 + 16 * (tg - 1) - 1);
          }
          if (++stage == K::STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // whole warp runs the loop (uniform descriptors), one elected lane issues
    constexpr uint32_t idesc = ptx::idesc_bf16(128, C) | (1u << 15) | (1u << 16);      // A and B MN-major
    const uint32_t lbo_a = p.desc_variant == 1 ? 1024u : (uint32_t)K::A_SLICE;
    const uint32_t sbo_a = p.desc_variant == 1 ? (uint32_t)K::A_SLICE : 1024u;
    const uint32_t lbo_b = p.desc_variant == 1 ? 1024u : (uint32_t)K::B_SLICE;
    const uint32_t sbo_b = p.desc_variant == 1 ? (uint32_t)K::B_SLICE : 1024u;
    int stage = 0; uint32_t phase = 0; bool ok = true;
    for (int it = 0; it < n_chunks && ok; ++it) {
      if (!ptx::mbar_wait(&full[stage], phase)) { if (lane == 0) atomicExch(p.error, WERR_FULL); ok = false; break; }
      ptx::tc_fence_after();
      const uint32_t sA = ptx::smem_u32(smem + stage * K::STAGE);
      const uint32_t sB = sA + (uint32_t)K::A_BYTES;
      if (ptx::elect_one()) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = smem_desc_mn_sw128(sA + (uint32_t)(j * 16 * 128), lbo_a, sbo_a);
            const uint64_t bd = smem_desc_mn_sw128(sB + (uint32_t)((t + 16 * j) * 128), lbo_b, sbo_b);
            ptx::umma_bf16<1>(tmem_base + (uint32_t)(t * C), ad, bd, idesc, (it | j) != 0 ? 1u : 0u);
          }
        ptx::umma_commit(&empty[stage]);
        if (it == n_chunks - 1) ptx::umma_commit(tfull);
      }
      __syncwarp();
      if (++stage == K::STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (n_chunks > 0) {
    const int quad = warp & 3;
    const int co = quad * 32 + lane;
    if (!ptx::mbar_wait(tfull, 0)) { atomicExch(p.error, WERR_TFULL); }
    else {
      ptx::tc_fence_after();
      if (quad * 32 < C) {          // with 64 channels the upper two quadrants hold nothing
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
          float* dst = p.dw + ((size_t)(tg * 3 + t) * ldw + co0 + co) * ldw + ci0;
#pragma unroll 1
          for (int c = 0; c < C / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * C + 32 * c), v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              red_add_v4(dst + 32 * c + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, K::TMEM_COLS);
}

template <int C>
int launch_wgrad(const CUtensorMap& tm_dz, const CUtensorMap& tm_a, const WgradArgs& a, int n_sm, cudaStream_t stream) {
  using K = WCfg<C>;
  cudaError_t e = cudaFuncSetAttribute(wgrad3x3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  const int nq = a.ct > 0 ? a.ct / C : 1;           // 256-channel tensors: four quadrants of 128 x 128 channels (grid.y)
  int S = n_sm / (3 * nq * nq);
  if (S > a.n_boards) S = a.n_boards;
  if (S < 1) S = 1;
  wgrad3x3_kernel<C><<<dim3(3 * S, nq * nq), kWThreads, K::SMEM, stream>>>(tm_dz, tm_a, a);
  return azg_check_launch("wgrad3x3_kernel");
}


// ------------------------------------------------------------------------------------------------
// Cluster variant: the three CTAs of a cluster are the three kernel rows (dr = rank - 1) of the SAME board slice.
// They need the same dz tile and overlapping rows of `a` (the union of their 66-row tiles is 98 rows), so each CTA
// fetches one third of both and multicasts it to all three: 13.7 KB from L2 per CTA and chunk instead of 33 KB.
// Per stage and CTA: full barrier armed for the whole tile (bytes arrive from all three senders); a stage is refilled
// only after the MMAs of all three CTAs released it (empty barrier counts three multicast commits).
// ------------------------------------------------------------------------------------------------
template <int C>
struct WClusterCfg {
  static constexpr int KC = C / 64;
  static constexpr int A_SLICE = 64 * 128;
  static constexpr int B_ROWS = 98;                                  // rows r0 - 17 .. r0 + 80
  static constexpr int B_SLICE = ((B_ROWS * 128) + 1023) & ~1023;    // 13312
  static constexpr int A_BYTES = KC * A_SLICE, B_BYTES = KC * B_SLICE;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int TMEM_COLS = 3 * C <= 256 ? 256 : 512;
  static constexpr int TX = KC * (64 + B_ROWS) * 128;
  static constexpr int SMEM = 1024 + STAGES * STAGE + 256;
};

template <int C>
__global__ void __cluster_dims__(3, 1, 1) __launch_bounds__(kWThreads, 1)
wgrad3x3_cluster_kernel(const __grid_constant__ CUtensorMap tm_dz24, const __grid_constant__ CUtensorMap tm_dz16,
                        const __grid_constant__ CUtensorMap tm_a32, const __grid_constant__ CUtensorMap tm_a34, WgradArgs p) {
  using K = WClusterCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::STAGES * K::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + K::STAGES;
  uint64_t* tfull = bars + 2 * K::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();       // kernel row: dr = rank - 1
  const int s = (int)ptx::cluster_id_x(), S = (int)ptx::ncluster_x();
  const int n_boards = p.n_boards;
  int my_boards = 0;
  for (int b = s; b < n_boards; b += S) ++my_boards;
  const int n_chunks = my_boards * 4;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_dz24); ptx::prefetch_tmap(&tm_dz16); ptx::prefetch_tmap(&tm_a32); ptx::prefetch_tmap(&tm_a34); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < K::STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 3); }
      ptx::mbar_init(tfull, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<1>(tmem_slot, K::TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                 // every CTA's barriers exist before anyone multicasts into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int a_off = rank * 24, b_off = rank * 32;                      // this CTA's third of the dz tile (24 / 24 / 16 rows) and of the a tile (32 / 32 / 34)
      const CUtensorMap* tdz = rank == 2 ? &tm_dz16 : &tm_dz24;
      const CUtensorMap* ta = rank == 2 ? &tm_a34 : &tm_a32;
      int stage = 0; uint32_t phase = 0; bool ok = true;
      for (int b = s; b < n_boards && ok; b += S)
        for (int q = 0; q < 4 && ok; ++q) {
          if (!ptx::mbar_wait(&empty[stage], phase ^ 1u)) { atomicExch(p.error, WERR_EMPTY); ok = false; break; }
          const int r0 = AZG_NET_FRONT + b * 256 + q * 64;
          uint8_t* sA = smem + stage * K::STAGE;
          uint8_t* sB = sA + K::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full[stage], (uint32_t)K::TX);
          for (int kc = 0; kc < K::KC; ++kc) {
            ptx::tma_load_2d_multicast(sA + kc * K::A_SLICE + a_off * 128, tdz, &full[stage], kc * 64, r0 + a_off, (uint16_t)7);
            ptx::tma_load_2d_multicast(sB + kc * K::B_SLICE + b_off * 128, ta, &full[stage], kc * 64, r0 - 17 + b_off, (uint16_t)7);
          }
          if (++stage == K::STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::idesc_bf16(128, C) | (1u << 15) | (1u << 16);
    int stage = 0; uint32_t phase = 0; bool ok = true;
    for (int it = 0; it < n_chunks && ok; ++it) {
      if (!ptx::mbar_wait(&full[stage], phase)) { if (lane == 0) atomicExch(p.error, WERR_FULL); ok = false; break; }
      ptx::tc_fence_after();
      const uint32_t sA = ptx::smem_u32(smem + stage * K::STAGE);
      const uint32_t sB = sA + (uint32_t)K::A_BYTES;
      if (ptx::elect_one()) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = smem_desc_mn_sw128(sA + (uint32_t)(j * 16 * 128), (uint32_t)K::A_SLICE, 1024u);
            // union tile row of pixel i for tap (dr, dc): i + 17 + 16 dr + dc = i + 16 rank + t   (dr = rank - 1, dc = t - 1)
            const uint64_t bd = smem_desc_mn_sw128(sB + (uint32_t)((16 * rank + t + 16 * j) * 128), (uint32_t)K::B_SLICE, 1024u);
            ptx::umma_bf16<1>(tmem_base + (uint32_t)(t * C), ad, bd, idesc, (it | j) != 0 ? 1u : 0u);
          }
        ptx::umma_commit_multicast(&empty[stage], (uint16_t)7);     // releases the stage in all three CTAs
        if (it == n_chunks - 1) ptx::umma_commit(tfull);
      }
      __syncwarp();
      if (++stage == K::STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (n_chunks > 0) {
    const int quad = warp & 3;
    const int co = quad * 32 + lane;
    if (!ptx::mbar_wait(tfull, 0)) { atomicExch(p.error, WERR_TFULL); }
    else {
      ptx::tc_fence_after();
      if (quad * 32 < C) {
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
          float* dst = p.dw + ((size_t)(rank * 3 + t) * C + co) * C;
#pragma unroll 1
          for (int c = 0; c < C / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * C + 32 * c), v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              red_add_v4(dst + 32 * c + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                 // nobody leaves while a peer may still multicast into it or signal its barriers
  if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, K::TMEM_COLS);
}

template <int C>
int launch_wgrad_cluster(const CUtensorMap* tm, const WgradArgs& a, int n_sm, cudaStream_t stream) {
  using K = WClusterCfg<C>;
  cudaError_t e = cudaFuncSetAttribute(wgrad3x3_cluster_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  // Clusters of three do not tile every GPC: ask how many can be resident at once and launch exactly one wave
  // (n_sm / 3 clusters would leave the remainder for a second wave that doubles the kernel's duration).
  static int resident = 0;
  if (resident == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(3 * (n_sm / 3)); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = K::SMEM;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 3; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, wgrad3x3_cluster_kernel<C>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = n_sm / 3; }
    resident = n;
    if (getenv("AZG_WGRAD_VERBOSE")) fprintf(stderr, "wgrad3x3_cluster_kernel<%d>: %d clusters of 3 resident on %d SMs\n", C, n, n_sm);
  }
  int S = resident < n_sm / 3 ? resident : n_sm / 3;
  if (S > a.n_boards) S = a.n_boards;
  if (S < 1) S = 1;
  wgrad3x3_cluster_kernel<C><<<3 * S, kWThreads, K::SMEM, stream>>>(tm[0], tm[1], tm[2], tm[3], a);
  return azg_check_launch("wgrad3x3_cluster_kernel");
}

}  // namespace

int azg_wgrad3x3_a_rows() { return WCfg<128>::B_ROWS; }

int azg_wgrad3x3_launch(int C, const CUtensorMap& tm_dz, const CUtensorMap& tm_a, const WgradArgs& a, int n_sm, cudaStream_t stream) {
  if (C == 64) return launch_wgrad<64>(tm_dz, tm_a, a, n_sm, stream);
  if (C == 128) return launch_wgrad<128>(tm_dz, tm_a, a, n_sm, stream);
  if (C == 256) { WgradArgs b = a; b.ct = 256; return launch_wgrad<128>(tm_dz, tm_a, b, n_sm, stream); }
  return azg_fail(AZG_E_ARG, "wgrad3x3: channels must be 64, 128 or 256");
}

// Cluster variant: tm = {dz box 24 rows, dz box 16 rows, a box 32 rows, a box 34 rows} (all 64 channels wide, SWIZZLE_128B).
int azg_wgrad3x3_cluster_launch(int C, const CUtensorMap* tm, const WgradArgs& a, int n_sm, cudaStream_t stream) {
  if (C == 64) return launch_wgrad_cluster<64>(tm, a, n_sm, stream);
  if (C == 128) return launch_wgrad_cluster<128>(tm, a, n_sm, stream);
  return azg_fail(AZG_E_ARG, "wgrad3x3: channels must be 64 or 128");
}

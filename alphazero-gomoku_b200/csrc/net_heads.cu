// Policy / value heads, dense part: policy_fc + softmax and value_fc1 + ReLU + value_fc2 + tanh
// (network.py:104-115, 180) as two tcgen05 GEMMs per 128-board tile.
//
//   logits[128 x 240] = hidden[128 x 480] * Wp^T      (225 outputs, 450 policy features, zero padded)
//   v1    [128 x  64] = hidden[128 x 448..704] * Wv^T  (64 outputs, 225 value features)
//
// Operands are float32 in HBM and are multiplied as TF32 (kind::tf32, fp32 accumulation in TMEM): the
// head features come out of a bf16 trunk, so the 10-bit mantissa of the products is below the error the
// trunk already carries (tests/test_net_gpu.py states the tolerance).  One persistent CTA per SM:
// warp 0 = TMA producer (K slices of 32 floats = one 128-byte swizzled row), warp 1 = MMA issuer,
// warps 2-5 = epilogue (thread = board: bias, softmax over its 225 columns straight from TMEM in three
// passes, value MLP tail), with a shared-memory transpose so that probabilities leave as 128-byte rows.
// A row of the result depends on that row of `hidden` only, so outputs do not depend on the batch shape.
#include "common.cuh"
#include "net.h"
#include "ptx.cuh"

namespace {

constexpr int kTileM = 128;                       // boards per tile (UMMA M, cta_group::1)
constexpr int kNP = 240;                          // policy outputs padded to a multiple of 16
constexpr int kNV = 64;
constexpr int kPolSlices = AZG_HEAD_WP_K / 32;    // 15
constexpr int kValSlices = AZG_HEAD_WV_K / 32;    // 8
constexpr int kValCol0 = 448;                     // first hidden column of the value GEMM (450 rounded down to a slice)
constexpr int kStages = 4;
constexpr int kABytes = kTileM * 128;             // 16 KB
constexpr int kBBytes = kNP * 128;                // 30 KB (value slices use the first 8 KB)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;                    // policy accumulator [0,240), value accumulator [256,320)
constexpr int kValTmemCol = 256;
constexpr int kEpiPitch = 33;                     // floats per row of the transpose tile
constexpr int kSmem = 1024 + kStages * kStageBytes + 4 * 32 * kEpiPitch * 4 + 1024;
enum { ERR_HEAD_EMPTY = 21, ERR_HEAD_FULL = 22, ERR_HEAD_TFULL = 23, ERR_HEAD_TEMPTY = 24 };

// kind::tf32 instruction descriptor: tf32 x tf32 -> f32, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct HeadGemmArgs {
  const int* n_boards; int max_boards;
  const float* pol_b;             // [240] (zero padded)
  const float* v1_b;              // [64]
  const float* v2_w;              // [64]
  const float* v2_b;              // [1]
  float* probs; float* values; float* logits;
  int* error;
};

__global__ void __launch_bounds__(192, 1)
heads_gemm_kernel(const __grid_constant__ CUtensorMap tm_hid, const __grid_constant__ CUtensorMap tm_wp,
                  const __grid_constant__ CUtensorMap tm_wv, HeadGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  float* epi = reinterpret_cast<float*>(smem + kStages * kStageBytes);            // [4 warps][32][33]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + 4 * 32 * kEpiPitch * 4);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  __shared__ float s_pb[kNP], s_vb[kNV], s_vw[kNV];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n = *p.n_boards;
  if (n > p.max_boards) n = p.max_boards;
  const int n_tiles = (n + kTileM - 1) / kTileM;

  for (int i = threadIdx.x; i < kNP; i += blockDim.x) s_pb[i] = p.pol_b[i];
  for (int i = threadIdx.x; i < kNV; i += blockDim.x) { s_vb[i] = p.v1_b[i]; s_vw[i] = p.v2_w[i]; }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hid);
    ptx::prefetch_tmap(&tm_wp);
    ptx::prefetch_tmap(&tm_wv);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      ptx::mbar_init(tfull, 1);
      ptx::mbar_init(tempty, 4);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<1>(tmem_slot, kTmemCols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x) {
        for (int s = 0; s < kPolSlices + kValSlices && ok; ++s) {
          if (!ptx::mbar_wait(&empty[stage], phase ^ 1u)) { atomicExch(p.error, ERR_HEAD_EMPTY); ok = false; break; }
          uint8_t* sa = stages + stage * kStageBytes;
          if (s < kPolSlices) {
            ptx::mbar_arrive_expect_tx(&full[stage], kABytes + kBBytes);
            ptx::tma_load_2d(sa, &tm_hid, &full[stage], s * 32, t * kTileM);
            ptx::tma_load_2d(sa + kABytes, &tm_wp, &full[stage], s * 32, 0);
          } else {
            const int v = s - kPolSlices;
            ptx::mbar_arrive_expect_tx(&full[stage], kABytes + kNV * 128);
            ptx::tma_load_2d(sa, &tm_hid, &full[stage], kValCol0 + v * 32, t * kTileM);
            ptx::tma_load_2d(sa + kABytes, &tm_wv, &full[stage], v * 32, 0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (warp-uniform loop, one elected lane issues) ==============================
    constexpr uint32_t idesc_p = idesc_tf32(kTileM, kNP), idesc_v = idesc_tf32(kTileM, kNV);
    const uint64_t desc0 = ptx::smem_desc_sw128(ptx::smem_u32(stages));
    int stage = 0, it = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++it) {
      if (!ptx::mbar_wait(tempty, ((uint32_t)it & 1u) ^ 1u)) { if (lane == 0) atomicExch(p.error, ERR_HEAD_TEMPTY); ok = false; break; }
      ptx::tc_fence_after();
      for (int s = 0; s < kPolSlices + kValSlices && ok; ++s) {
        if (!ptx::mbar_wait(&full[stage], phase)) { if (lane == 0) atomicExch(p.error, ERR_HEAD_FULL); ok = false; break; }
        ptx::tc_fence_after();
        const uint64_t ad = desc0 + (uint64_t)((stage * kStageBytes) >> 4);
        const uint64_t bd = ad + (uint64_t)(kABytes >> 4);
        if (ptx::elect_one()) {
          if (s < kPolSlices) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_base, ad + (uint64_t)((k * 32) >> 4), bd + (uint64_t)((k * 32) >> 4), idesc_p, (s | k) != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_base + kValTmemCol, ad + (uint64_t)((k * 32) >> 4), bd + (uint64_t)((k * 32) >> 4), idesc_v,
                        (s != kPolSlices || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (ok && ptx::elect_one()) ptx::umma_commit(tfull);
      __syncwarp();
    }
  } else {
    // ============================== epilogue: thread = board ==============================
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    float* tile = epi + (warp - 2) * 32 * kEpiPitch;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float v2b = p.v2_b[0];
    int it = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++it) {
      if (!ptx::mbar_wait(tfull, (uint32_t)it & 1u)) { if (lane == 0) atomicExch(p.error, ERR_HEAD_TFULL); ok = false; break; }
      ptx::tc_fence_after();
      const int b0 = t * kTileM + q * 32;                   // first board of this warp
      uint32_t v[32];
      // pass 1: row maximum of the 225 logits
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        ptx::tmem_ld_32x32(lane_addr + (uint32_t)(c * 32), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < 225) mx = fmaxf(mx, __uint_as_float(v[j]) + s_pb[c * 32 + j]);
      }
      // pass 2: sum of exponentials
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        ptx::tmem_ld_32x32(lane_addr + (uint32_t)(c * 32), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < 225) sum += expf(__uint_as_float(v[j]) + s_pb[c * 32 + j] - mx);
      }
      const float inv = 1.0f / sum;
      // pass 3: probabilities (and raw logits on request), transposed through shared memory
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        ptx::tmem_ld_32x32(lane_addr + (uint32_t)(c * 32), v);
        ptx::tmem_ld_wait();
        const int ncol = c < 7 ? 32 : 1;                     // columns 224..255: only 224 is real
        if (p.logits) {
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[lane * kEpiPitch + j] = __uint_as_float(v[j]) + s_pb[(c * 32 + j) < kNP ? c * 32 + j : 0];
          __syncwarp();
          for (int r = 0; r < 32; ++r)
            if (b0 + r < n && lane < ncol) p.logits[(size_t)(b0 + r) * 225 + c * 32 + lane] = tile[r * kEpiPitch + lane];
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          tile[lane * kEpiPitch + j] = expf(__uint_as_float(v[j]) + s_pb[(c * 32 + j) < kNP ? c * 32 + j : 0] - mx) * inv;
        __syncwarp();
        for (int r = 0; r < 32; ++r)
          if (b0 + r < n && lane < ncol) p.probs[(size_t)(b0 + r) * 225 + c * 32 + lane] = tile[r * kEpiPitch + lane];
        __syncwarp();
      }
      // value head tail: relu(v1 + b1) . w2 + b2 -> tanh
      float acc = v2b;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        ptx::tmem_ld_32x32(lane_addr + (uint32_t)(kValTmemCol + c * 32), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(fmaxf(__uint_as_float(v[j]) + s_vb[c * 32 + j], 0.f), s_vw[c * 32 + j], acc);
      }
      if (p.values && b0 + lane < n) p.values[b0 + lane] = tanhf(acc);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

}  // namespace

int azg_heads_gemm_launch(const CUtensorMap& tm_hid, const CUtensorMap& tm_wp, const CUtensorMap& tm_wv, const HeadArgs& a,
                          int* error, int n_sm, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(heads_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);   // per device
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  HeadGemmArgs p;
  p.n_boards = a.n_boards; p.max_boards = a.max_boards; p.pol_b = a.pol_b; p.v1_b = a.v1_b; p.v2_w = a.v2_w; p.v2_b = a.v2_b;
  p.probs = a.probs; p.values = a.values; p.logits = a.logits; p.error = error;
  int tiles = (a.max_boards + kTileM - 1) / kTileM;
  int grid = tiles < n_sm ? tiles : n_sm;
  if (grid < 1) grid = 1;
  heads_gemm_kernel<<<grid, 192, kSmem, stream>>>(tm_hid, tm_wp, tm_wv, p);
  return azg_check_launch("heads_gemm_kernel");
}

// 3x3 convolution + folded BatchNorm (+ residual) + ReLU as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces the trunk of AlphaZeroNet.forward (network.py:94-99, ResidualBlock network.py:17-26)
// - 99.75 % of the network's FLOPs - for a batch of leaf positions.
//
// Data layout (DESIGN.md "leaf evaluation"): activations are bf16 channels-last in a PADDED
// flat pixel space: every board owns 256 consecutive rows of C channels, row qi = y*16 + x with
// y = 0 a zero row above the board, x = 15 a zero column; pixel (r, c) lives at (r+1)*16 + c.
// A 3x3 tap (dr, dc) is then the constant row offset dr*16 + dc, halo reads hit zero rows, and
// one board is exactly one M = 256 tile of a CTA pair (two M = 128 halves).
//
// Kernel shape: persistent, one 2-CTA cluster per SM pair (cta_group::2, UMMA 256 x C x 16).
//   * the layer's weights (9 taps x C x C bf16) stay RESIDENT in shared memory, split by output
//     channel across the pair (C = 128: 144 KB per CTA) - no weight traffic per tile; C = 256 streams
//     one 64-channel slice of all nine taps at a time through a 9-slot ring instead;
//   * activations arrive by TMA as ONE copy of 162 rows (tile + halo) per 64-channel slice; all nine
//     taps are row offsets of that tile in the UMMA shared-memory descriptor, so each activation byte
//     is fetched once from L2 (MODE 1/3/4/5; MODE 0, three column-shifted copies, is the recorded
//     first version);
//   * accumulators live in TMEM as a ring of NACC buffers (4 x 128 columns at C = 128, 8 x 64 at
//     C = 64, 2 x 256 at C = 256): the epilogue of board i overlaps the MMAs of boards i+1 ...;
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM allocator,
//     warps 2-9 epilogue (TMEM -> registers -> shift (+residual) -> ReLU (optional) -> bf16 -> HBM).
// The same kernel is the forward AND the input-gradient convolution of the training step (train_engine.cu):
// relu = 0, shift = 0, weights either raw (z = conv(a)) or transposed with flipped taps (da = conv^T(dz)).
#include <cuda_bf16.h>
#include <type_traits>
#include "net.h"
#include "ptx.cuh"

namespace {

// Activation staging modes (template parameter MODE):
//   0: three column-shifted copies (dc = -1, 0, +1) of 160 rows; row taps are 2 KB-aligned offsets.
//   1: ONE copy of 162 rows per 64-channel slice; every tap is a row offset of the same tile
//      (start address not aligned to the 1024-byte swizzle period, descriptor base_offset = 0).
//      (Setting base_offset = (start >> 7) & 7 instead was tried and gives WRONG results on B200: the
//      swizzle is a function of absolute shared-memory address bits.)
//   3: as 1 but with the direct (row-per-thread) epilogue instead of the staged one.
//   4: as 1 but the staged tile leaves by coalesced 16-byte global stores of the warp instead of a TMA store.
template <int MODE>
struct Stage {
  static constexpr int ROWS = MODE == 0 ? 160 : 162;
  static constexpr int BYTES = ROWS * 128;                         // TMA transaction size
  static constexpr int PITCH = (BYTES + 1023) & ~1023;             // 1024-aligned stage pitch
};
// warp 0 TMA, warp 1 MMA, then EG groups of 8 epilogue warps.  With 64 channels the MMAs of a board take 1.15 k cycles
// but one group of 8 warps needs ~2.3 k cycles for its epilogue (round 1: 1 178 cycles per board spent waiting for a free
// accumulator), so two groups work on alternate boards there.
template <int C> constexpr int epi_groups() { return C == 64 ? 2 : 1; }
template <int C> constexpr int conv_threads() { return 64 + 256 * epi_groups<C>(); }

template <int C, int MODE>
struct Cfg {
  static constexpr int KC = C / 64;                       // 64-channel K slices
  static constexpr int BBLK = (C / 2) * 128;              // one (tap, slice) weight block: C/2 rows x 128 B
  // C <= 128: all 9*KC weight blocks of the layer stay resident.  C = 256 (1.18 MB of weights per
  // layer) streams them through a ring of 9 blocks (one 64-channel slice of all taps) instead.
  static constexpr bool STREAM = C > 128;
  static constexpr int NBLK = STREAM ? 9 : 9 * KC;        // weight blocks held in shared memory
  static constexpr int BBYTES = NBLK * BBLK;
  // Accumulator ring in TMEM.  Two buffers only hide an epilogue that is SHORTER than the MMAs of a board;
  // on residual layers its latency is longer (measured 5.9 k vs 4.25 k cycles), so the whole 512 columns
  // are used: four buffers at 128 channels (MMA of board i+4 waits for the epilogue of board i).
  static constexpr int NACC = C == 256 ? 2 : (C == 64 ? 8 : 4);
  static constexpr int TMEM_COLS = NACC * C;              // 256 or 512 columns (a power of two)
  static constexpr int STAGES = MODE == 0 ? (C == 128 ? 4 : 6) : 3;   // activation tiles in flight
  static constexpr int EPI = (MODE == 1 || MODE == 4 || MODE == 5) ? 8 * epi_groups<C>() * 2048 : 0;                // per-warp epilogue staging tiles (32 rows x 64 B)
  static constexpr int SMEM = 1024 + BBYTES + STAGES * Stage<MODE>::PITCH + EPI + 512;
};

// Folded BatchNorm shift of the layer, passed by value so the epilogue reads it from the constant
// bank (no shared-memory or LSU traffic).  The BatchNorm scale is folded into the bf16 weights.
template <int C>
struct ConvShift {
  float v[C];
};

// 1x1 head convolutions fused into the epilogue of the last trunk layer (network.py:102-104,
// 109-111): rows 0-1 policy_conv, row 2 value_conv, with their folded BatchNorm.
template <int C>
struct HeadConst {
  float w[3][C];
  float scale[3];
  float shift[3];
};

enum { ERR_BFULL = 1, ERR_EMPTY = 2, ERR_FULL = 3, ERR_TEMPTY = 4, ERR_TFULL = 5 };

// Sum over the 32 lanes (rows) of each of 32 per-lane values (channels): afterwards lane L holds the total of v[L]
// in v[0].  31 shuffles: in every round a lane hands half of its values to its partner and adds the partner's.
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

template <int C, bool HEADS, int MODE, int STATS = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(conv_threads<C>(), 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_out, ConvArgs p,
                    const __grid_constant__ ConvShift<C> shift, const __grid_constant__ HeadConst<HEADS ? C : 1> head) {
  using K = Cfg<C, MODE>;
  constexpr int kStages = K::STAGES;
  constexpr int kStageBytes = Stage<MODE>::BYTES, kStagePitch = Stage<MODE>::PITCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + K::BBYTES;
  uint8_t* sEpi = sA + kStages * kStagePitch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + K::EPI);
  uint64_t* full = bars;                // [kStages]  leader: both CTAs' copies landed
  uint64_t* empty = bars + kStages;     // [kStages]  each CTA: MMAs reading the stage retired
  uint64_t* tfull = bars + 2 * kStages; // [NACC]     each CTA: accumulator complete
  uint64_t* tempty = tfull + K::NACC;   // [NACC]     leader: both epilogues drained the accumulator
  uint64_t* bfull = tempty + K::NACC;         // leader: both weight halves resident (resident mode)
  uint64_t* wfull = bfull + 1;          // [9] leader: weight ring slot filled by both CTAs (streaming mode)
  uint64_t* wempty = wfull + 9;         // [9] each CTA: MMAs reading the slot retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wempty + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cid = (int)ptx::cluster_id_x(), ncl = (int)ptx::ncluster_x();
  // Programmatic dependent launch: the next layer may start its prologue (barriers, TMEM, its own weight
  // loads) on idle SMs while this one computes; everything that touches activations waits below.
  ptx::grid_dep_launch();
  int n_boards = *p.n_boards;
  if (n_boards > p.max_boards) n_boards = p.max_boards;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_act);
    ptx::prefetch_tmap(&tm_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      for (int a = 0; a < K::NACC; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 16); }
      ptx::mbar_init(bfull, 1);
      for (int i = 0; i < 9; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<2>(tmem_slot, K::TMEM_COLS);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (lane == 0) {
      if constexpr (!K::STREAM) {
        if (rank == 0) ptx::mbar_arrive_expect_tx(bfull, 2u * K::BBYTES);
        for (int blk = 0; blk < 9 * K::KC; ++blk) {
          const int tap = blk / K::KC, kc = blk % K::KC;
          ptx::tma_load_2d_pair(sB + blk * K::BBLK, &tm_w, bfull, kc * 64, (p.layer * 9 + tap) * C + (int)rank * (C / 2));
        }
      }
      ptx::grid_dep_wait();                    // activations of the previous layer are complete from here on
      int stage = 0, ws = 0;
      uint32_t phase = 0, wphase = 0;
      bool ok = true;
      long long t_wait = 0;
      const long long t_begin = clock64();
      for (int b = cid; b < n_boards && ok; b += ncl) {
        const int row0 = AZG_NET_FRONT + (b * 2 + (int)rank) * 128 - (MODE == 0 ? 16 : 17);
        for (int kc = 0; kc < K::KC && ok; ++kc)
          for (int dci = 0; dci < (MODE == 0 ? 3 : 1); ++dci) {
            const long long t0 = clock64();
            if (!ptx::mbar_wait(&empty[stage], phase ^ 1u)) { atomicExch(p.error, ERR_EMPTY); ok = false; break; }
            t_wait += clock64() - t0;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2u * kStageBytes);
            ptx::tma_load_2d_pair(sA + stage * kStagePitch, &tm_act, &full[stage], kc * 64, row0 + (MODE == 0 ? dci - 1 : 0));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            if constexpr (K::STREAM) {          // the nine tap blocks of this slice follow their activations
              for (int tap = 0; tap < 9 && ok; ++tap) {
                if (!ptx::mbar_wait(&wempty[ws], wphase ^ 1u)) { atomicExch(p.error, ERR_EMPTY); ok = false; break; }
                if (rank == 0) ptx::mbar_arrive_expect_tx(&wfull[ws], 2u * K::BBLK);
                ptx::tma_load_2d_pair(sB + ws * K::BBLK, &tm_w, &wfull[ws], kc * 64, (p.layer * 9 + tap) * C + (int)rank * (C / 2));
                if (++ws == 9) { ws = 0; wphase ^= 1u; }
              }
            }
          }
      }
      if (p.prof && rank == 0) {
        atomicAdd(p.prof + 3, (unsigned long long)t_wait);
        atomicAdd(p.prof + 4, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA) ==============================
    // The whole warp runs this loop so that every descriptor / stage / phase value is provably
    // warp-uniform (kept in uniform registers, no per-MMA R2UR round trips); one elected lane
    // issues the tcgen05 instructions and the commits.
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16(256, C);
      const uint64_t a_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(sA));
      const uint64_t b_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(sB));
      bool ok = K::STREAM ? true : ptx::mbar_wait(bfull, 0);
      if (!ok && lane == 0) atomicExch(p.error, ERR_BFULL);
      int stage = 0, it = 0, ws = 0;
      uint32_t phase = 0, wphase = 0;
      long long t_full = 0, t_tempty = 0;
      const long long t_begin = clock64();
      for (int b = cid; b < n_boards && ok; b += ncl, ++it) {
        const int acc = it % K::NACC;
        long long t0 = clock64();
        if (!ptx::mbar_wait(&tempty[acc], ((uint32_t)(it / K::NACC) & 1u) ^ 1u)) { if (lane == 0) atomicExch(p.error, ERR_TEMPTY); ok = false; break; }
        t_tempty += clock64() - t0;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * C);
        for (int kc = 0; kc < K::KC && ok; ++kc)
          for (int dcs = 0; dcs < (MODE == 0 ? 3 : 1); ++dcs) {
            t0 = clock64();
            if (!ptx::mbar_wait(&full[stage], phase)) { if (lane == 0) atomicExch(p.error, ERR_FULL); ok = false; break; }
            t_full += clock64() - t0;
            ptx::tc_fence_after();
            const uint64_t a_stage = a_desc0 + (uint64_t)((stage * kStagePitch) >> 4);
            const uint64_t b_slice = b_desc0 + (uint64_t)((kc * K::BBLK) >> 4);
            if constexpr (!K::STREAM) {
              if (ptx::elect_one()) {
#pragma unroll
                for (int dri = 0; dri < 3; ++dri) {
#pragma unroll
                  for (int dcu = 0; dcu < (MODE == 0 ? 1 : 3); ++dcu) {
                    const int dci = MODE == 0 ? dcs : dcu;
                    // MODE 0: the stage IS the dc copy, the row tap is a 16-row (2 KB) offset.
                    // MODE 1/3: copy row 17 + 16*(dr) + dc holds the tap's first row (dr, dc in -1..1).
                    const int arow = MODE == 0 ? dri * 16 : 17 + 16 * (dri - 1) + (dcu - 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const uint64_t ad = a_stage + (uint64_t)((arow * 128 + k * 32) >> 4);
                      const uint64_t bd = b_slice + (uint64_t)((((dri * 3 + dci) * K::KC) * K::BBLK + k * 32) >> 4);
                      ptx::umma_bf16<2>(tmem_d, ad, bd, idesc, (kc | dcs | dri | dcu | k) != 0 ? 1u : 0u);
                    }
                  }
                }
                ptx::umma_commit_pair(&empty[stage], 3);        // frees the stage in both CTAs
              }
            } else {
              // streaming weights (single activation copy only): one ring slot per tap
#pragma unroll 1
              for (int tap = 0; tap < 9 && ok; ++tap) {
                if (!ptx::mbar_wait(&wfull[ws], wphase)) { if (lane == 0) atomicExch(p.error, ERR_BFULL); ok = false; break; }
                ptx::tc_fence_after();
                const int arow = 17 + 16 * (tap / 3 - 1) + (tap % 3 - 1);
                const uint64_t a_tap = a_stage + (uint64_t)((arow * 128) >> 4);
                const uint64_t b_tap = b_desc0 + (uint64_t)((ws * K::BBLK) >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    ptx::umma_bf16<2>(tmem_d, a_tap + (uint64_t)((k * 32) >> 4), b_tap + (uint64_t)((k * 32) >> 4), idesc,
                                      (kc | tap | k) != 0 ? 1u : 0u);
                  ptx::umma_commit_pair(&wempty[ws], 3);
                  if (tap == 8) ptx::umma_commit_pair(&empty[stage], 3);
                }
                __syncwarp();
                if (++ws == 9) { ws = 0; wphase ^= 1u; }
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        if (ok && ptx::elect_one()) ptx::umma_commit_pair(&tfull[acc], 3);       // accumulator ready in both CTAs
        __syncwarp();
      }
      if (p.prof && lane == 0) {
        atomicAdd(p.prof + 0, (unsigned long long)t_full);
        atomicAdd(p.prof + 1, (unsigned long long)t_tempty);
        atomicAdd(p.prof + 2, (unsigned long long)(clock64() - t_begin));
        atomicAdd(p.prof + 7, (unsigned long long)it);
      }
    }
  } else {
    // ============================== epilogue (warps 2..9, both CTAs) ==============================
    // Two warps per TMEM lane quadrant; each takes half of the channels of its 32 rows.  In the
    // fused-heads variant (last layer) one warp per quadrant takes all channels of its rows so
    // that the three head dot products stay inside a thread; the other warp only signals.
    ptx::grid_dep_wait();                      // residual reads and output writes follow the previous layer
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 3;                 // epilogue group: takes the boards with it % EG == grp
    const int half = ((warp - 2) & 7) >> 2;
    constexpr int EG = epi_groups<C>();
    constexpr int NCH = HEADS ? C : C / 2;           // channels handled by a working warp
    const int ch0 = HEADS ? 0 : half * (C / 2);
    const bool working = !HEADS || half == 0;
    const int row = quad * 32 + lane;
    const int qi = (int)rank * 128 + row;
    const bool pad = (qi < 16) || ((qi & 15) == 15);
    int it = 0;
    bool ok = true;
    long long t_tfull = 0, t_wr = 0, t_rt = 0, t_ld = 0, t_cs = 0, t_st = 0;
    // STATS 1 (training forward): per-channel sum and sum of squares of the bf16-rounded outputs, accumulated by this
    // warp over all its boards (lane = channel within a 32-channel chunk) and left in stat_partial at the end.
    // STATS 2 (input gradient): the output is g = dL/da of the layer below; with that layer's ReLU bit mask and its
    // pre-BatchNorm z the same two slots collect sum(dy) and sum(dy * z), dy = g * [a > 0] - the reductions of its
    // BatchNorm backward pass, which then needs no pass of its own over g and z.
    float st_sum[STATS ? (HEADS ? C : C / 2) / 32 : 1], st_sq[STATS ? (HEADS ? C : C / 2) / 32 : 1];
    if constexpr (STATS != 0) {
#pragma unroll
      for (int c = 0; c < (HEADS ? C : C / 2) / 32; ++c) st_sum[c] = st_sq[c] = 0.f;
    }
    const bool detail = p.prof != nullptr && p.prof_detail != 0;      // per-phase clocks only on request (AZG_CONV_PHASES=1)
    const float lo = p.relu ? 0.f : -INFINITY;       // relu == 0: linear output (training: pre-BatchNorm z, input gradients)
    const long long t_begin = clock64();
    // MODE >= 1: global traffic of the epilogue is coalesced through a per-warp 2 KB staging tile
    // (32 rows x 32 channels, 64-byte rows, 16-byte units XOR-swizzled with (row >> 1) & 3, which is
    // both bank-conflict free and the layout of a SWIZZLE_64B TMA box): residual rows are read with
    // coalesced 16-byte loads (8 rows x 64 B per instruction) and transposed through the tile, results
    // leave by one TMA store per tile.  MODE 0 keeps the direct row-per-thread accesses.
    // MODE 5: results staged + TMA store as in MODE 1, but the residual comes straight from global memory in the
    // accumulator's row-per-thread layout (256-bit loads, one full sector per lane): no shared-memory round trip
    // for it - on residual layers the shared memory is the busiest unit.
    constexpr bool STAGED = MODE == 1 || MODE == 4 || MODE == 5;
    constexpr bool STAGED_RES = MODE == 1 || MODE == 4;
    constexpr bool TMA_OUT = MODE == 1 || MODE == 5;
    constexpr int NCHUNK = NCH / 32;
    const uint32_t tile = ptx::smem_u32(sEpi) + (uint32_t)(warp - 2) * 2048u;
    const uint32_t own = tile + (uint32_t)lane * 64u;                      // this thread's row in the tile
    const uint32_t own_sw = (uint32_t)((lane >> 1) & 3);
    const int crow = lane >> 2, cunit = lane & 3;                          // coalesced mapping: row 8i + crow, unit cunit
    for (int b = cid; b < n_boards && ok; b += ncl, ++it) {
      if (EG > 1 && (it % EG) != grp) continue;
      const int acc = it % K::NACC;
      const size_t grow0 = (size_t)AZG_NET_FRONT + (size_t)b * 256 + (size_t)rank * 128 + (size_t)quad * 32;   // first row of this warp
      const size_t grow = grow0 + (size_t)lane;
      __nv_bfloat16* orow = p.out ? p.out + grow * C : nullptr;
      const bool has_res = p.residual != nullptr && working;
      // the residual is fetched while the MMAs of this board are still running
      uint4 rc[STAGED_RES ? NCHUNK * 4 : 1];
      uint32_t res[STAGED_RES ? 16 : NCH / 2];
      if (has_res) {
        if constexpr (STAGED_RES) {
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              rc[c * 4 + i] = ptx::ldg128(p.residual + (grow0 + (size_t)(8 * i + crow)) * C + ch0 + 32 * c + 8 * cunit);
        } else {
          const __nv_bfloat16* rrow = p.residual + grow * C;
#pragma unroll
          for (int j = 0; j < NCH / 16; ++j) ptx::ldg256(rrow + ch0 + 16 * j, &res[8 * j]);
        }
      }
      uint32_t bz[STATS == 2 ? NCH / 2 : 1], bm[STATS == 2 ? NCHUNK : 1];
      if constexpr (STATS == 2) {
        if (working) {
          const __nv_bfloat16* zrow = p.bwd_z + grow * C + ch0;
#pragma unroll
          for (int j = 0; j < NCH / 16; ++j) ptx::ldg256(zrow + 16 * j, &bz[8 * j]);
          const uint8_t* mrow = p.bwd_mask + ((size_t)b * 256 + (size_t)rank * 128 + (size_t)row) * (C / 8) + ch0 / 8;
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) bm[c] = *reinterpret_cast<const uint32_t*>(mrow + 4 * c);
        }
      }
      const long long t0 = clock64();
      if (!ptx::mbar_wait(&tfull[acc], (uint32_t)(it / K::NACC) & 1u)) { atomicExch(p.error, ERR_TFULL); ok = false; break; }
      t_tfull += clock64() - t0;
      ptx::tc_fence_after();
      if (working) {
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int cc = 32 * c;
          const int ch = ch0 + cc;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * C + ch), v);
          if constexpr (STAGED) {
            long long tp = detail ? clock64() : 0;
            if (TMA_OUT && STAGED_RES && p.out) {   // the previous TMA store must have finished reading the tile
              if (lane == 0) ptx::bulk_wait_read0();
              __syncwarp();
            }
            if (detail) { t_wr += clock64() - tp; tp = clock64(); }
            if (STAGED_RES && has_res) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = 8 * i + crow;
                ptx::sts128(tile + (uint32_t)r * 64u + (uint32_t)((cunit ^ ((r >> 1) & 3)) * 16), rc[STAGED_RES ? c * 4 + i : 0]);
              }
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 q = ptx::lds128(own + (uint32_t)(((uint32_t)u ^ own_sw) * 16u));
                res[4 * u] = q.x; res[4 * u + 1] = q.y; res[4 * u + 2] = q.z; res[4 * u + 3] = q.w;
              }
              __syncwarp();
            }
            if (detail) t_rt += clock64() - tp;
          }
          long long tq = detail ? clock64() : 0;
          ptx::tmem_ld_wait();
          if (detail) { t_ld += clock64() - tq; tq = clock64(); }
          uint32_t outv[16];
          // shift (constant bank, compile-time index) -> + residual -> ReLU -> bf16 pair; instantiated
          // per (channel half, residual) so that the inner loop carries no selects
          auto math = [&](auto SHOFF, auto RES) {
#pragma unroll
            for (int h = 0; h < 16; ++h) {
              float y0 = __uint_as_float(v[2 * h]) + shift.v[decltype(SHOFF)::value + cc + 2 * h];
              float y1 = __uint_as_float(v[2 * h + 1]) + shift.v[decltype(SHOFF)::value + cc + 2 * h + 1];
              if constexpr (decltype(RES)::value) {
                const uint32_t rw = STAGED_RES ? res[h] : res[(cc / 2 + h) % (STAGED_RES ? 16 : NCH / 2)];
                y0 += __uint_as_float(rw << 16);
                y1 += __uint_as_float(rw & 0xffff0000u);
              }
              const __nv_bfloat162 pk = __floats2bfloat162_rn(fmaxf(y0, lo), fmaxf(y1, lo));
              outv[h] = pad ? 0u : *reinterpret_cast<const uint32_t*>(&pk);
            }
          };
          using I0 = std::integral_constant<int, 0>;
          using IH = std::integral_constant<int, HEADS ? 0 : C / 2>;
          if (!HEADS && half) { if (has_res) math(IH{}, std::true_type{}); else math(IH{}, std::false_type{}); }
          else { if (has_res) math(I0{}, std::true_type{}); else math(I0{}, std::false_type{}); }
          if constexpr (STATS != 0) {
            float zs[32], zq[32];
#pragma unroll
            for (int h = 0; h < 16; ++h) {
              zs[2 * h] = __uint_as_float(outv[h] << 16); zs[2 * h + 1] = __uint_as_float(outv[h] & 0xffff0000u);
              if constexpr (STATS == 2) {
                const uint32_t zw = bz[(STATS == 2 ? 16 * c : 0) + h];
                if (!((bm[STATS == 2 ? c : 0] >> (2 * h)) & 1u)) zs[2 * h] = 0.f;
                if (!((bm[STATS == 2 ? c : 0] >> (2 * h + 1)) & 1u)) zs[2 * h + 1] = 0.f;
                zq[2 * h] = zs[2 * h] * __uint_as_float(zw << 16); zq[2 * h + 1] = zs[2 * h + 1] * __uint_as_float(zw & 0xffff0000u);
              } else {
                zq[2 * h] = zs[2 * h] * zs[2 * h]; zq[2 * h + 1] = zs[2 * h + 1] * zs[2 * h + 1];
              }
            }
            warp_transpose_sum(zs, lane);
            warp_transpose_sum(zq, lane);
            st_sum[c] += zs[0];
            st_sq[c] += zq[0];
          }
          if constexpr (HEADS) {
            // the heads see the bf16-rounded activations, exactly like the unfused path
#pragma unroll
            for (int h = 0; h < 16; ++h) {
              const float z0 = __uint_as_float(outv[h] << 16), z1 = __uint_as_float(outv[h] & 0xffff0000u);
              d0 = fmaf(z0, head.w[0][cc + 2 * h], d0); d0 = fmaf(z1, head.w[0][cc + 2 * h + 1], d0);
              d1 = fmaf(z0, head.w[1][cc + 2 * h], d1); d1 = fmaf(z1, head.w[1][cc + 2 * h + 1], d1);
              d2 = fmaf(z0, head.w[2][cc + 2 * h], d2); d2 = fmaf(z1, head.w[2][cc + 2 * h + 1], d2);
            }
          }
          if (p.out) {
            if constexpr (STAGED) {
              if constexpr (TMA_OUT && !STAGED_RES) {   // MODE 5: the tile is first touched here, after the arithmetic,
                if (lane == 0) ptx::bulk_wait_read0();  // so the previous store's read-out has had that long to finish
                __syncwarp();
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                ptx::sts128(own + (uint32_t)(((uint32_t)u ^ own_sw) * 16u),
                            make_uint4(outv[4 * u], outv[4 * u + 1], outv[4 * u + 2], outv[4 * u + 3]));
              if (detail) { t_cs += clock64() - tq; tq = clock64(); }
              if constexpr (TMA_OUT) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  ptx::tma_store_2d(&tm_out, sEpi + (warp - 2) * 2048, ch, (int)grow0);
                  ptx::bulk_commit();
                }
              } else {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {      // 8 rows x 64 B per instruction
                  const int r = 8 * i + crow;
                  const uint4 q = ptx::lds128(tile + (uint32_t)r * 64u + (uint32_t)((cunit ^ ((r >> 1) & 3)) * 16));
                  ptx::stg128(p.out + (grow0 + (size_t)r) * C + ch + 8 * cunit, q);
                }
                __syncwarp();
              }
              if (detail) t_st += clock64() - tq;
            } else {
              ptx::stg256(orow + ch, &outv[0]);
              ptx::stg256(orow + ch + 16, &outv[8]);
            }
          }
        }
        if constexpr (HEADS) {
          if (!pad) {
            const int pix = ((qi >> 4) - 1) * 15 + (qi & 15);
            float* hid = p.hidden + (size_t)b * AZG_HEAD_PITCH + pix;        // consecutive lanes = consecutive pixels
            hid[0] = fmaxf(fmaf(d0, head.scale[0], head.shift[0]), 0.f);
            hid[225] = fmaxf(fmaf(d1, head.scale[1], head.shift[1]), 0.f);
            hid[450] = fmaxf(fmaf(d2, head.scale[2], head.shift[2]), 0.f);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_leader(&tempty[acc]);
    }
    if constexpr (STATS != 0) {
      // slot = one epilogue warp's share of the rows: (cluster, CTA, quadrant[, group]); channel = ch0 + 32 c + lane
      const int slot = ((cid * 2 + (int)rank) * 4 + quad) * EG + grp;
      if (working) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          p.stat_partial[(size_t)(ch0 + 32 * c + lane) * AZG_CONV_STAT_SLOTS + slot] = st_sum[c];
          p.stat_partial[(size_t)(C + ch0 + 32 * c + lane) * AZG_CONV_STAT_SLOTS + slot] = st_sq[c];
        }
      }
    }
    if (TMA_OUT && lane == 0) ptx::bulk_wait0();            // all stores of this warp have landed before the CTA exits
    if (p.prof && rank == 0 && warp == 2 && lane == 0) {   // one representative epilogue warp
      atomicAdd(p.prof + 5, (unsigned long long)t_tfull);
      atomicAdd(p.prof + 6, (unsigned long long)(clock64() - t_begin));
      atomicAdd(p.prof + 8, (unsigned long long)t_wr);
      atomicAdd(p.prof + 9, (unsigned long long)t_rt);
      atomicAdd(p.prof + 10, (unsigned long long)t_ld);
      atomicAdd(p.prof + 11, (unsigned long long)t_cs);
      atomicAdd(p.prof + 12, (unsigned long long)t_st);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc<2>(tmem_base, K::TMEM_COLS);
}

template <int C, bool HEADS, int MODE, int STATS = 0>
int launch_conv(const CUtensorMap& tm_act, const CUtensorMap& tm_w, const CUtensorMap& tm_out, const ConvArgs& args, int n_sm,
                cudaStream_t stream) {
  using K = Cfg<C, MODE>;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_pair_kernel<C, HEADS, MODE, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  int grid = n_sm & ~1;
  const int want = 2 * args.max_boards;
  if (grid > want) grid = want < 2 ? 2 : want;
  ConvShift<C> shift;
  for (int i = 0; i < C; ++i) shift.v[i] = args.shift_host[i];
  HeadConst<HEADS ? C : 1> head;
  if constexpr (HEADS) {
    for (int r = 0; r < 3; ++r) {
      for (int i = 0; i < C; ++i) head.w[r][i] = args.head_host[r * C + i];
      head.scale[r] = args.head_host[3 * C + r];
      head.shift[r] = args.head_host[3 * C + 3 + r];
    }
  } else {
    head.w[0][0] = head.w[1][0] = head.w[2][0] = 0.f;
    for (int r = 0; r < 3; ++r) head.scale[r] = head.shift[r] = 0.f;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(conv_threads<C>());
  cfg.dynamicSmemBytes = K::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see grid_dep_launch / grid_dep_wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<C, HEADS, MODE, STATS>, tm_act, tm_w, tm_out, args, shift, head);
  if (e != cudaSuccess) return azg_fail(AZG_E_CUDA, cudaGetErrorString(e));
  return azg_check_launch("conv3x3_pair_kernel");
}

template <int C, int MODE>
int launch_conv_heads(const CUtensorMap& tm_act, const CUtensorMap& tm_w, const CUtensorMap& tm_out, const ConvArgs& args, int n_sm,
                      cudaStream_t stream) {
  return args.head_host ? launch_conv<C, true, MODE>(tm_act, tm_w, tm_out, args, n_sm, stream)
                        : launch_conv<C, false, MODE>(tm_act, tm_w, tm_out, args, n_sm, stream);
}

}  // namespace

int azg_conv3x3_rows(int mode) { return mode == 0 ? Stage<0>::ROWS : Stage<1>::ROWS; }

int azg_conv3x3_stat_slots(int max_boards, int n_sm, int C) {
  int grid = n_sm & ~1;
  const int want = 2 * max_boards;
  if (grid > want) grid = want < 2 ? 2 : want;
  return grid * 4 * (C == 64 ? epi_groups<64>() : 1);
}

int azg_conv3x3_launch(int C, int mode, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const CUtensorMap& tm_out,
                       const ConvArgs& args, int n_sm, cudaStream_t stream) {
  if (args.stat_partial) {          // training: per-channel statistics in the epilogue (forward: of z; input gradient: BatchNorm backward sums)
    if (args.head_host) return azg_fail(AZG_E_ARG, "conv3x3: statistics and fused heads are separate variants");
    if (args.bwd_z) {
      if (!args.bwd_mask) return azg_fail(AZG_E_ARG, "conv3x3: bwd_z needs bwd_mask");
      if (C == 128 && mode == 1) return launch_conv<128, false, 1, 2>(tm_act, tm_w, tm_out, args, n_sm, stream);
      if (C == 64 && mode == 3) return launch_conv<64, false, 3, 2>(tm_act, tm_w, tm_out, args, n_sm, stream);
      if (C == 256 && mode == 1) return launch_conv<256, false, 1, 2>(tm_act, tm_w, tm_out, args, n_sm, stream);
      return azg_fail(AZG_E_ARG, "conv3x3: the statistics epilogue is built for (128 / 256 channels, mode 1) and (64 channels, mode 3)");
    }
    if (C == 128 && mode == 1) return launch_conv<128, false, 1, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (C == 64 && mode == 3) return launch_conv<64, false, 3, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (C == 256 && mode == 1) return launch_conv<256, false, 1, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
    return azg_fail(AZG_E_ARG, "conv3x3: the statistics epilogue is built for (128 / 256 channels, mode 1) and (64 channels, mode 3)");
  }
  if (C == 128) {
    if (mode == 0) return launch_conv_heads<128, 0>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 1) return launch_conv_heads<128, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 3) return launch_conv_heads<128, 3>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 4) return launch_conv_heads<128, 4>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 5) return launch_conv_heads<128, 5>(tm_act, tm_w, tm_out, args, n_sm, stream);
    return azg_fail(AZG_E_ARG, "conv3x3: unknown staging mode");
  }
  if (C == 64) {
    if (mode == 0) return launch_conv_heads<64, 0>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 1) return launch_conv_heads<64, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 3) return launch_conv_heads<64, 3>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 4) return launch_conv_heads<64, 4>(tm_act, tm_w, tm_out, args, n_sm, stream);
    return azg_fail(AZG_E_ARG, "conv3x3: unknown staging mode");
  }
  if (C == 256) {
    if (args.head_host) return azg_fail(AZG_E_ARG, "conv3x3: the fused-heads epilogue is built for 64 and 128 channels");
    if (mode == 3) return launch_conv<256, false, 3>(tm_act, tm_w, tm_out, args, n_sm, stream);
    if (mode == 4) return launch_conv<256, false, 4>(tm_act, tm_w, tm_out, args, n_sm, stream);
    return launch_conv<256, false, 1>(tm_act, tm_w, tm_out, args, n_sm, stream);
  }
  return azg_fail(AZG_E_ARG, "conv3x3: channels must be 64, 128 or 256");
}

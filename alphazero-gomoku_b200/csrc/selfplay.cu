// Self-play driver kernels: on-device Philox for root Dirichlet noise and move sampling,
// example capture, outcome labelling and 8-fold symmetry expansion.
//
// Reference: mcts/new_mcts_alpha.py:170-174 (noise), train.py:252-266 (temperature sampling),
// train.py:360-412 (play_game_and_collect), mcts/new_mcts_alpha.py:42-56 (symmetries).
// numpy's MT19937 stream cannot be reproduced by a counter-based generator; these kernels
// are validated statistically (tests/test_selfplay_gpu.py), the deterministic parts
// (argmax at T = 0, z labels, symmetry order) exactly.
#include "common.cuh"
#include "engine.h"
#include "host.h"
#include "rules.cuh"

// ------------------------------------------------------------------------------------------------
// Philox4x32-10
// ------------------------------------------------------------------------------------------------
struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  __device__ Philox(unsigned long long seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3 = 0) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = c0; ctr[1] = c1; ctr[2] = c2; ctr[3] = c3;      // next() increments ctr[3]: keep the low 8 bits of c3 zero
  }
  __device__ uint4 next() {                       // one block of 4 x 32 random bits; bumps ctr[3]
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3]++;
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  __device__ double uniform53(uint32_t a, uint32_t b) {   // (0, 1)
    const unsigned long long x = (((unsigned long long)a << 32) | b) >> 11;
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
  }
};

// log of a Gamma(alpha, 1) variate, alpha > 0 (Marsaglia-Tsang on alpha+1, boosted by U^(1/alpha));
// returned in log space because alpha ~ 0.03 underflows float64 otherwise.
__device__ double log_gamma_variate(Philox& rng, double alpha) {
  const double d = alpha + 1.0 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int attempt = 0; attempt < 64; ++attempt) {
    const uint4 r0 = rng.next(), r1 = rng.next();
    const double u1 = rng.uniform53(r0.x, r0.y), u2 = rng.uniform53(r0.z, r0.w);
    const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);          // standard normal
    const double t = 1.0 + c * x;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    const double u = rng.uniform53(r1.x, r1.y), ub = rng.uniform53(r1.z, r1.w);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return log(d * v) + log(ub) / alpha;
  }
  return log(d);
}

// noise[g][0..224] ~ Dirichlet(alpha, ..., alpha) over all 225 actions (new_mcts_alpha.py:172).
// The stream of a draw is keyed by (global game id, action, caller's draw index, games finished by
// this slot, ply of the current game): nothing has to change on the host from ply to ply, which is
// what lets a whole ply be captured once in a CUDA graph.
__global__ void __launch_bounds__(128)
noise_kernel(azg_dev e, azg_selfplay_buf sp, unsigned long long draw, double* __restrict__ noise) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  const int l = lane_id();
  const uint32_t epoch = ((uint32_t)(sp.n_done ? sp.n_done[g] : 0) * 1024u + (uint32_t)(sp.n_plies ? sp.n_plies[g] : 0)) << 8;
  double lg[8], mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int a = l + 32 * j;
    lg[j] = -INFINITY;
    if (a < AZG_A) {
      Philox rng(e.seed ^ 0xD1B54A32D192ED03ULL, (uint32_t)(g + e.game_base), (uint32_t)a, (uint32_t)draw, epoch);
      lg[j] = log_gamma_variate(rng, e.alpha);
      mx = fmax(mx, lg[j]);
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) mx = fmax(mx, __shfl_xor_sync(AZG_FULL, mx, s));
  double sum = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { lg[j] = (l + 32 * j < AZG_A) ? exp(lg[j] - mx) : 0.0; sum += lg[j]; }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(AZG_FULL, sum, s);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (l + 32 * j < AZG_A) noise[(size_t)g * AZG_A + l + 32 * j] = lg[j] / sum;
}

// ------------------------------------------------------------------------------------------------
// move choice + example capture (train.py:370-388)
// ------------------------------------------------------------------------------------------------
// temperature = max(0, 1 - move_number/temp_threshold) (train.py:647-648); T == 0 -> first argmax,
// otherwise sample from softmax(log(pi + 1e-15)/T) (train.py:252-266); an illegal pick falls back
// to the argmax (train.py:380-382).  The root position and pi are stored as the ply's example.
__global__ void __launch_bounds__(128)
choose_kernel(azg_dev e, azg_selfplay_buf sp, const float* __restrict__ pi, float temp_threshold, unsigned long long draw,
              int32_t* __restrict__ actions) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  const int l = lane_id();
  if (sp.active && sp.active[g] == 0) {          // retired slot: no move, no example, no ply count
    if (l == 0) actions[g] = -1;
    return;
  }
  const azg_ctl* ctl = e.ctl + g;
  const WPos root = wpos_load(&ctl->root);
  const int ply = sp.n_plies[g];
  float p[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int a = 8 * l + j; p[j] = a < AZG_A ? pi[(size_t)g * AZG_A + a] : -1.f; }
  // first argmax
  float best = -1.f; int best_i = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (8 * l + j < AZG_A && p[j] > best) { best = p[j]; best_i = 8 * l + j; }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const float ob = __shfl_xor_sync(AZG_FULL, best, s);
    const int oi = __shfl_xor_sync(AZG_FULL, best_i, s);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  int action = best_i;
  const float temp = fmaxf(0.f, 1.0f - (float)ply / temp_threshold);
  if (temp > 0.f) {
    float w[8], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) { w[j] = 8 * l + j < AZG_A ? logf(p[j] + 1e-15f) / temp : -INFINITY; mx = fmaxf(mx, w[j]); }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(AZG_FULL, mx, s));
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { w[j] = 8 * l + j < AZG_A ? expf(w[j] - mx) : 0.f; mine += w[j]; }
    float incl = mine;                                   // inclusive scan over lanes
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const float o = __shfl_up_sync(AZG_FULL, incl, s); if (l >= s) incl += o; }
    const float total = __shfl_sync(AZG_FULL, incl, 31);
    Philox rng(e.seed ^ 0x8CB92BA72F3D8DD7ULL, (uint32_t)(g + e.game_base), (uint32_t)ply, (uint32_t)draw, (uint32_t)sp.n_done[g] << 8);
    const uint4 r = rng.next();
    const float u = (float)(rng.uniform53(r.x, r.y) * (double)total);
    float run = incl - mine;
    int pick = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      run += w[j];
      if (pick == 0x7fffffff && 8 * l + j < AZG_A && w[j] > 0.f && run > u) pick = 8 * l + j;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) pick = min(pick, __shfl_xor_sync(AZG_FULL, pick, s));
    if (pick != 0x7fffffff) action = pick;
    if (wpos_at(root, action) != 0) action = best_i;     // illegal pick -> argmax (train.py:380-382)
  }
  // example capture: stones, side to move, pi
  if (ply < sp.max_plies) {
    const size_t slot = (size_t)g * sp.max_plies + ply;
    if ((l & 3) == 0) { sp.ex_key[slot * 16 + (l >> 2)] = root.w0; sp.ex_key[slot * 16 + 8 + (l >> 2)] = root.w1; }
    if (l == 0) sp.ex_player[slot] = root.player;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (8 * l + j < AZG_A) sp.ex_pi[slot * AZG_A + 8 * l + j] = p[j];
  }
  if (l == 0) { actions[g] = action; sp.n_plies[g] = ply + 1; }
}

// ------------------------------------------------------------------------------------------------
// game end: z labels + symmetry expansion (train.py:392-410, new_mcts_alpha.py:42-56)
// ------------------------------------------------------------------------------------------------
// status: azg_rules status bits after the move.  A game ends when it is over or has reached
// max_moves plies.  Every stored ply yields `n_sym` rows (8 or 1) of 901 floats in `out`:
// planes[3][225], pi[225], z.  Rows are reserved with one atomicAdd per game; rows beyond
// `capacity` are dropped and counted in *dropped.
__device__ __forceinline__ int sym_source(int k, int i, int j) {
  // source cell of output (i, j) for symmetry k: k>>1 quarter turns (np.rot90), odd k mirrored left-right
  if (k & 1) j = AZG_N - 1 - j;
  switch (k >> 1) {
    case 0: return i * AZG_N + j;
    case 1: return j * AZG_N + (AZG_N - 1 - i);
    case 2: return (AZG_N - 1 - i) * AZG_N + (AZG_N - 1 - j);
    default: return (AZG_N - 1 - j) * AZG_N + i;
  }
}

// One expanded row (901 floats: planes[3][225], pi[225], z) of symmetry s from a stored ply.
__device__ __forceinline__ void write_example_row(float* __restrict__ o, const uint32_t* __restrict__ k, int player,
                                                  const float* __restrict__ ppi, float z, int s) {
  for (int a = threadIdx.x; a < AZG_A; a += blockDim.x) {
    const int src = sym_source(s, a / AZG_N, a % AZG_N);
    const uint32_t b1 = (k[src >> 5] >> (src & 31)) & 1u, b2 = (k[8 + (src >> 5)] >> (src & 31)) & 1u;
    o[a] = (float)(player == 1 ? b1 : b2);
    o[AZG_A + a] = (float)(player == 1 ? b2 : b1);
    o[2 * AZG_A + a] = 1.0f;
    o[3 * AZG_A + a] = ppi[src];
  }
  if (threadIdx.x == 0) o[900] = z;
}

// packed != 0: out is uint32[capacity][AZG_PACKED_WORDS], one row per PLY (stones 16 words, side to move, z bits,
// pi[225], pad) - 976 bytes instead of 8 x 3604: what travels between GPUs; azg_examples_expand makes the rows.
__global__ void __launch_bounds__(256)
finish_games_kernel(azg_dev e, azg_selfplay_buf sp, const int32_t* __restrict__ status, int max_moves, int n_sym,
                    float* __restrict__ out, long long capacity, unsigned long long* __restrict__ cursor,
                    int32_t* __restrict__ done_mask, int32_t* __restrict__ winners, int packed) {
  const int g = blockIdx.x;
  if (g >= e.G) return;
  __shared__ long long s_base;
  if (sp.active && sp.active[g] == 0) {
    if (threadIdx.x == 0) { done_mask[g] = 0; if (winners) winners[g] = -1; }
    return;
  }
  const int st = status[g];
  const int plies = sp.n_plies[g];
  const bool over = (st & 4) != 0 || plies >= max_moves;
  if (threadIdx.x == 0) {
    done_mask[g] = over ? 1 : 0;
    if (winners) winners[g] = over ? (st & 3) : -1;
  }
  if (!over) return;
  const int won = st & 3;
  const int stored = plies < sp.max_plies ? plies : sp.max_plies;
  const long long rows = packed ? (long long)stored : (long long)stored * n_sym;
  if (threadIdx.x == 0) s_base = out ? (long long)atomicAdd(cursor, (unsigned long long)rows) : 0;
  __syncthreads();
  const long long base = s_base;
  if (out) {
    for (int i = 0; i < stored; ++i) {
      const size_t slot = (size_t)g * sp.max_plies + i;
      const int player = sp.ex_player[slot];
      const float z = won == 0 ? 0.f : (won == player ? 1.f : -1.f);
      const uint32_t* k = sp.ex_key + slot * 16;
      const float* ppi = sp.ex_pi + slot * AZG_A;
      if (packed) {
        const long long row = base + i;
        if (row >= capacity) continue;
        uint32_t* o = reinterpret_cast<uint32_t*>(out) + row * AZG_PACKED_WORDS;
        for (int a = threadIdx.x; a < AZG_PACKED_WORDS; a += blockDim.x)
          o[a] = a < 16 ? k[a] : a == 16 ? (uint32_t)player : a == 17 ? __float_as_uint(z) : a < 18 + AZG_A ? __float_as_uint(ppi[a - 18]) : 0u;
        continue;
      }
      for (int s = 0; s < n_sym; ++s) {
        const long long row = base + (long long)i * n_sym + s;
        if (row >= capacity) continue;
        write_example_row(out + row * 901, k, player, ppi, z, s);
      }
    }
  }
  if (threadIdx.x == 0) { sp.n_plies[g] = 0; sp.n_done[g] += 1; }
}

// packed plies -> example rows (train.py:392-410 on the receiving side of the exchange): block = one ply
__global__ void __launch_bounds__(256)
expand_examples_kernel(const uint32_t* __restrict__ packed, long long n, int n_sym, float* __restrict__ out) {
  const long long i = blockIdx.x;
  if (i >= n) return;
  __shared__ uint32_t row[AZG_PACKED_WORDS];
  for (int a = threadIdx.x; a < AZG_PACKED_WORDS; a += blockDim.x) row[a] = packed[i * AZG_PACKED_WORDS + a];
  __syncthreads();
  const int player = (int)row[16];
  const float z = __uint_as_float(row[17]);
  const float* ppi = reinterpret_cast<const float*>(row + 18);
  for (int s = 0; s < n_sym; ++s) write_example_row(out + (i * n_sym + s) * 901, row, player, ppi, z, s);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline dim3 warp_grid(int n) { return dim3((unsigned)((n + 3) / 4)); }

extern "C" int azg_selfplay_enable(azg_engine* e, int max_plies) {
  if (!e || max_plies < 1) return azg_fail(AZG_E_ARG, "azg_selfplay_enable: bad argument");
  if (e->sp.ex_key) return AZG_OK;
  AZG_CUDA(cudaSetDevice(e->cfg.device));
  const size_t slots = (size_t)e->dev.G * max_plies;
  if (cudaMalloc((void**)&e->sp.ex_key, slots * 16 * 4) != cudaSuccess || cudaMalloc((void**)&e->sp.ex_player, slots * 4) != cudaSuccess ||
      cudaMalloc((void**)&e->sp.ex_pi, slots * AZG_A * 4) != cudaSuccess || cudaMalloc((void**)&e->sp.n_plies, (size_t)e->dev.G * 4) != cudaSuccess ||
      cudaMalloc((void**)&e->sp.n_done, (size_t)e->dev.G * 4) != cudaSuccess) {
    cudaGetLastError();
    return azg_fail(AZG_E_NOMEM, "azg_selfplay_enable: cudaMalloc failed");
  }
  e->bytes += (int64_t)(slots * (16 * 4 + 4 + AZG_A * 4) + (size_t)e->dev.G * 4);
  e->sp.max_plies = max_plies;
  AZG_CUDA(cudaMemset(e->sp.n_plies, 0, (size_t)e->dev.G * 4));
  AZG_CUDA(cudaMemset(e->sp.n_done, 0, (size_t)e->dev.G * 4));
  return AZG_OK;
}

extern "C" int azg_selfplay_set_active(azg_engine* e, const int32_t* active) {
  if (!e) return azg_fail(AZG_E_ARG, "null argument");
  e->sp.active = active;
  return AZG_OK;
}

extern "C" int azg_selfplay_noise(azg_engine* e, uint64_t draw, double* noise) {
  if (!e || !noise) return azg_fail(AZG_E_ARG, "null argument");
  AZG_USE_DEVICE(e->cfg.device);
  noise_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, e->sp, draw, noise);
  return azg_check_launch("noise_kernel");
}

extern "C" int azg_selfplay_choose(azg_engine* e, const float* pi, float temp_threshold, uint64_t draw, int32_t* actions) {
  if (!e || !pi || !actions || !e->sp.ex_key) return azg_fail(AZG_E_ARG, "azg_selfplay_choose: bad argument (call azg_selfplay_enable first)");
  AZG_USE_DEVICE(e->cfg.device);
  choose_kernel<<<warp_grid(e->dev.G), 128, 0, e->stream>>>(e->dev, e->sp, pi, temp_threshold, draw, actions);
  return azg_check_launch("choose_kernel");
}

extern "C" int azg_selfplay_finish(azg_engine* e, const int32_t* status, int max_moves, int use_symmetries, float* out,
                                   int64_t capacity, uint64_t* cursor, int32_t* done_mask, int32_t* winners) {
  if (!e || !status || !done_mask || !e->sp.ex_key || (out && !cursor)) return azg_fail(AZG_E_ARG, "azg_selfplay_finish: bad argument");
  AZG_USE_DEVICE(e->cfg.device);
  finish_games_kernel<<<e->dev.G, 256, 0, e->stream>>>(e->dev, e->sp, status, max_moves, use_symmetries ? 8 : 1, out,
                                                        (long long)capacity, (unsigned long long*)cursor, done_mask, winners, 0);
  return azg_check_launch("finish_games_kernel");
}

extern "C" int azg_selfplay_finish_packed(azg_engine* e, const int32_t* status, int max_moves, uint32_t* out, int64_t capacity,
                                          uint64_t* cursor, int32_t* done_mask, int32_t* winners) {
  if (!e || !status || !done_mask || !e->sp.ex_key || (out && !cursor)) return azg_fail(AZG_E_ARG, "azg_selfplay_finish_packed: bad argument");
  AZG_USE_DEVICE(e->cfg.device);
  finish_games_kernel<<<e->dev.G, 256, 0, e->stream>>>(e->dev, e->sp, status, max_moves, 1, reinterpret_cast<float*>(out),
                                                        (long long)capacity, (unsigned long long*)cursor, done_mask, winners, 1);
  return azg_check_launch("finish_games_kernel");
}

extern "C" int azg_examples_expand(const uint32_t* packed, int64_t n, int use_symmetries, float* out, void* stream) {
  if (n < 0 || (n > 0 && (!packed || !out))) return azg_fail(AZG_E_ARG, "azg_examples_expand: bad argument");
  if (n == 0) return AZG_OK;
  if (n > 0x7fffffffLL) return azg_fail(AZG_E_ARG, "azg_examples_expand: too many rows for one call");
  expand_examples_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(packed, (long long)n, use_symmetries ? 8 : 1, out);
  return azg_check_launch("expand_examples_kernel");
}

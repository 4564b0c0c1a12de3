// Warp-resident bitboard rules for Gomoku / Pente (sm_100a).
//
// One warp owns one position.  The 2 x 256-bit stone sets are spread over the
// warp: lane L keeps word (L >> 2) of each colour, so the byte of the board that
// holds children 8L..8L+7 of a search node is already in lane L's registers and
// move application / legality need no shared memory.  Cell probes for the
// five-in-a-row and custodial-capture tests are one shuffle each.
//
// Semantics restated from the reference:
//   do_move            games/gomoku.py:60-78, games/pente.py:57-79
//   capture            games/pente.py:114-152 (me, foe, foe, me on the 8 rays)
//   check_winner       games/gomoku.py:155-193, games/pente.py:199-233
//   is_game_over       games/gomoku.py:195-197
//   get_valid_moves    games/gomoku.py:109-121
#pragma once
#include "common.cuh"

struct WPos {
  uint32_t w0, w1;                       // this lane's word of player-1 / player-2 stones
  int player, last, cap0, cap1, plies;   // warp-uniform
};

__device__ __forceinline__ WPos wpos_load(const azg_pos* p) {
  WPos q;
  const int w = lane_id() >> 2;
  q.w0 = __ldcg(&p->stones[0][w]);
  q.w1 = __ldcg(&p->stones[1][w]);
  q.player = __ldcg(&p->player);
  q.last = __ldcg(&p->last);
  q.cap0 = __ldcg(&p->caps[0]);
  q.cap1 = __ldcg(&p->caps[1]);
  q.plies = __ldcg(&p->plies);
  return q;
}

__device__ __forceinline__ void wpos_store(azg_pos* p, const WPos& q) {
  const int l = lane_id();
  if ((l & 3) == 0) {
    p->stones[0][l >> 2] = q.w0;
    p->stones[1][l >> 2] = q.w1;
  }
  if (l == 0) {
    p->player = q.player; p->last = q.last; p->caps[0] = q.cap0; p->caps[1] = q.cap1; p->plies = q.plies;
  }
}

// Empty cells of this lane's word.
__device__ __forceinline__ uint32_t wpos_empty_word(const WPos& q) {
  return ~(q.w0 | q.w1) & board_word_mask(lane_id() >> 2);
}

// Legal-move bits of children 8L .. 8L+7 (bit j == child 8L+j).
__device__ __forceinline__ uint32_t wpos_legal_byte(const WPos& q) {
  return (wpos_empty_word(q) >> ((lane_id() & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ int wpos_count_empty(const WPos& q) {
  int c = (lane_id() & 3) == 0 ? __popc(wpos_empty_word(q)) : 0;
  return __reduce_add_sync(AZG_FULL, c);
}

// Is there a stone of colour `col` (0/1, warp-uniform) on `cell` (per lane, valid)?
__device__ __forceinline__ bool wpos_probe(const WPos& q, int col, int cell, bool valid) {
  const uint32_t mine = col ? q.w1 : q.w0;
  const int c = valid ? cell : 0;
  const uint32_t word = __shfl_sync(AZG_FULL, mine, (c >> 5) << 2);
  return valid && ((word >> (c & 31)) & 1u);
}

// Colour (0 none, 1, 2) of a warp-uniform cell.
__device__ __forceinline__ int wpos_at(const WPos& q, int cell) {
  const uint32_t a = __shfl_sync(AZG_FULL, q.w0, (cell >> 5) << 2);
  const uint32_t b = __shfl_sync(AZG_FULL, q.w1, (cell >> 5) << 2);
  const uint32_t bit = 1u << (cell & 31);
  return (a & bit) ? 1 : ((b & bit) ? 2 : 0);
}

// Winner judged through the last stone only.
__device__ __forceinline__ int wpos_winner(const WPos& q, int rule) {
  if (q.last < 0) return 0;
  const int who = wpos_at(q, q.last);
  if (who == 0) return 0;
  if (rule == 1 && (who == 1 ? q.cap0 : q.cap1) >= 5) return who;
  const int r = q.last / AZG_N, c = q.last - r * AZG_N;
  const int l = lane_id();
  const int d = l >> 3, t = (l & 3) + 1, sgn = (l & 4) ? -1 : 1;
  const int dr = (d == 1) ? 0 : 1;                 // (1,0) (0,1) (1,1) (1,-1)
  const int dc = (d == 0) ? 0 : ((d == 3) ? -1 : 1);
  const int rr = r + sgn * t * dr, cc = c + sgn * t * dc;
  const bool inb = rr >= 0 && rr < AZG_N && cc >= 0 && cc < AZG_N;
  const bool hit = wpos_probe(q, who - 1, rr * AZG_N + cc, inb);
  const uint32_t m = __ballot_sync(AZG_FULL, hit);
  bool won = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t b = (m >> (8 * k)) & 0xffu;
    const int fwd = __ffs(~(b & 0xfu)) - 1;        // contiguous stones in the + direction (max 4)
    const int bwd = __ffs(~(b >> 4)) - 1;
    won = won || (1 + fwd + bwd >= 5);
  }
  return won ? who : 0;
}

__device__ __forceinline__ bool wpos_any_empty(const WPos& q) {
  return __ballot_sync(AZG_FULL, wpos_empty_word(q) != 0u) != 0u;
}

// Place a stone for the side to move on an EMPTY on-board cell (warp-uniform).
__device__ __forceinline__ void wpos_play(WPos& q, int rule, int a) {
  const int me = q.player;
  const int l = lane_id();
  const uint32_t bit = 1u << (a & 31);
  if ((l >> 2) == (a >> 5)) { if (me == 1) q.w0 |= bit; else q.w1 |= bit; }
  q.last = a;
  q.plies += 1;
  if (rule == 1) {
    const int r = a / AZG_N, c = a - r * AZG_N;
    const int d = l & 7;                            // ray index, order as in pente.py:124-129
    const int dr = (d == 2 || d == 3) ? 0 : ((d == 1 || d == 5 || d == 7) ? -1 : 1);
    const int dc = (d < 2) ? 0 : ((d == 3 || d == 5 || d == 6) ? -1 : 1);
    const int r3 = r + 3 * dr, c3 = c + 3 * dc;
    const bool inb = r3 >= 0 && r3 < AZG_N && c3 >= 0 && c3 < AZG_N;
    const int a1 = (r + dr) * AZG_N + c + dc, a2 = (r + 2 * dr) * AZG_N + c + 2 * dc, a3 = r3 * AZG_N + c3;
    const bool f1 = wpos_probe(q, 2 - me, a1, inb);  // foe colour index = (3-me)-1
    const bool f2 = wpos_probe(q, 2 - me, a2, inb);
    const bool m3 = wpos_probe(q, me - 1, a3, inb);
    uint32_t take = __ballot_sync(AZG_FULL, f1 && f2 && m3) & 0xffu;
    const int pairs = __popc(take);
    while (take) {
      const int k = __ffs(take) - 1;
      take &= take - 1;
      const int x1 = __shfl_sync(AZG_FULL, a1, k), x2 = __shfl_sync(AZG_FULL, a2, k);
      uint32_t clr = 0;
      if ((l >> 2) == (x1 >> 5)) clr |= 1u << (x1 & 31);
      if ((l >> 2) == (x2 >> 5)) clr |= 1u << (x2 & 31);
      if (me == 1) q.w1 &= ~clr; else q.w0 &= ~clr;
    }
    if (me == 1) q.cap0 += pairs; else q.cap1 += pairs;
  }
  q.player = 3 - me;
}

// Position hash over stones + side to move: the low word selects the probe window, the high word is
// the slot tag (the full key is always compared, new_mcts_alpha.py:190-197).  Each owner lane mixes its
// word pair into two 32-bit streams (murmur3 finaliser), the streams are XOR-reduced with redux.sync
// and finalised once more - 32-bit multiplies only.
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ unsigned long long wpos_hash(const WPos& q) {
  const int l = lane_id();
  uint32_t ha = 0, hb = 0;
  if ((l & 3) == 0) {
    const uint32_t k = 0x9E3779B9u * (uint32_t)((l >> 2) + 1);
    const uint32_t x = q.w0 * 0xcc9e2d51u + k;
    const uint32_t y = q.w1 * 0x1b873593u + (k ^ 0x7f4a7c15u);
    ha = fmix32(x ^ __funnelshift_l(y, y, 15));
    hb = fmix32((y + __funnelshift_l(x, x, 7)) ^ 0x52dce729u);
  }
  ha = __reduce_xor_sync(AZG_FULL, ha);
  hb = __reduce_xor_sync(AZG_FULL, hb);
  const uint32_t lo = fmix32(ha ^ (hb >> 3) ^ (uint32_t)q.player);
  const uint32_t hi = fmix32(hb + lo * 0x9E3779B9u + 0x85ebca6bu);
  return ((unsigned long long)hi << 32) | lo;
}

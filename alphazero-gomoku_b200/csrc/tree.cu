// PUCT search kernels: one warp per game, structure-of-arrays node slabs in HBM.
//
// Restates the reference search (mcts/new_mcts_alpha.py:77-185) as the per-game
// state machine FILL -> EVAL -> COMMIT described in DESIGN.md:
//   azg_fill_kernel    runs simulations until the game's leaf queue holds
//                      queue_len nodes (the parked simulation is resumed later,
//                      new_mcts_alpha.py:121-135) or the simulation budget is spent;
//   azg_commit_kernel  stores masked priors, zeroes N and W of every queued node
//                      (new_mcts_alpha.py:163-185) and re-arms the game;
//   azg_finish_kernel  root visit counts -> pi (new_mcts_alpha.py:88-97);
//   azg_advance_kernel plays the chosen move on the root and reclaims nodes that
//                      can never be looked up again.
// Arithmetic of the selection rule is float32 in the reference's exact operation
// order (new_mcts_alpha.py:136-140); float64 at a Dirichlet-noised root.
#include "common.cuh"
#include "rules.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------------
// transposition table: per-game open addressing in windows of 32 slots (one coalesced probe)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_equal(const azg_dev& e, int g, int node, const WPos& p) {
  const int l = lane_id();
  const size_t off = azg_node_off(e, g, node);
  const uint32_t x0 = __shfl_sync(AZG_FULL, p.w0, (l & 7) << 2);
  const uint32_t x1 = __shfl_sync(AZG_FULL, p.w1, (l & 7) << 2);
  bool ok = true;
  if (l < 16) ok = __ldcg(&e.key[off * 16 + l]) == (l < 8 ? x0 : x1);
  else if (l == 16) ok = ((__ldcg(&e.meta[off]) >> 1) & 3u) == (uint32_t)p.player;
  return __all_sync(AZG_FULL, ok);
}

__device__ __forceinline__ int table_find(const azg_dev& e, int g, const WPos& p, unsigned long long h, int* ins) {
  const unsigned long long* tab = e.slots + (size_t)g * (size_t)e.hcap;
  const uint32_t tag = (uint32_t)(h >> 32);
  const int nwin = e.hcap >> 5;
  int win = (int)((uint32_t)h & (uint32_t)(nwin - 1));
  const int l = lane_id();
  for (int t = 0; t < nwin; ++t) {
    const unsigned long long s = __ldcg(&tab[(win << 5) + l]);
    uint32_t mm = __ballot_sync(AZG_FULL, s != 0ULL && (uint32_t)(s >> 32) == tag);
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      const int node = (int)__shfl_sync(AZG_FULL, (uint32_t)s, src) - 1;
      if (key_equal(e, g, node, p)) return node;
    }
    const uint32_t em = __ballot_sync(AZG_FULL, s == 0ULL);
    if (em) { *ins = (win << 5) + __ffs(em) - 1; return -1; }
    win = (win + 1) & (nwin - 1);
  }
  *ins = -1;
  return -1;
}

__device__ __forceinline__ void table_put(const azg_dev& e, int g, int slot, unsigned long long h, int node) {
  if (lane_id() == 0)
    e.slots[(size_t)g * (size_t)e.hcap + slot] = ((h >> 32) << 32) | (unsigned long long)(uint32_t)(node + 1);
}

__device__ __forceinline__ void node_write_key(const azg_dev& e, int g, int node, const WPos& p) {
  const int l = lane_id();
  const size_t off = azg_node_off(e, g, node);
  const uint32_t x0 = __shfl_sync(AZG_FULL, p.w0, (l & 7) << 2);
  const uint32_t x1 = __shfl_sync(AZG_FULL, p.w1, (l & 7) << 2);
  if (l < 16) e.key[off * 16 + l] = (l < 8 ? x0 : x1);
  if (l == 16) e.meta[off] = AZG_META_ALIVE | ((uint32_t)p.player << 1);
  if (e.child) {                                   // a new node knows nothing about its children yet
    const size_t base = off * AZG_ROW;
    int4* c = reinterpret_cast<int4*>(e.child + base + 8 * l);
    if (l < 28) { c[0] = make_int4(0, 0, 0, 0); c[1] = make_int4(0, 0, 0, 0); }
    else if (l == 28) e.child[base + 224] = 0;
  }
}

// children arrays of a node: lane L < 28 owns elements 8L..8L+7, lane 28 owns element 224.
__device__ __forceinline__ void node_fill(const azg_dev& e, int g, int node, uint32_t legal, float pv) {
  const int l = lane_id();
  const size_t base = azg_node_off(e, g, node) * AZG_ROW;
  if (l < 28) {
    float4 a, b;
    a.x = (legal & 1u) ? pv : 0.f;   a.y = (legal & 2u) ? pv : 0.f;   a.z = (legal & 4u) ? pv : 0.f;   a.w = (legal & 8u) ? pv : 0.f;
    b.x = (legal & 16u) ? pv : 0.f;  b.y = (legal & 32u) ? pv : 0.f;  b.z = (legal & 64u) ? pv : 0.f;  b.w = (legal & 128u) ? pv : 0.f;
    float4* P = reinterpret_cast<float4*>(e.P + base + 8 * l);
    P[0] = a; P[1] = b;
    int4 z = make_int4(0, 0, 0, 0);
    int4* Nn = reinterpret_cast<int4*>(e.Nv + base + 8 * l);
    int4* Ww = reinterpret_cast<int4*>(e.W + base + 8 * l);
    Nn[0] = z; Nn[1] = z; Ww[0] = z; Ww[1] = z;
  } else if (l == 28) {
    e.P[base + 224] = (legal & 1u) ? pv : 0.f;
    e.Nv[base + 224] = 0;
    e.W[base + 224] = 0;
  }
}

// Loads of a game's slab inside FILL.  One warp owns a game and is the only reader and writer of its slab while the
// kernel runs.  L1 = false: every load goes to L2 (ld.cg) because the back-up updates N and W with L2 atomics that an
// L1 line would not see.  L1 = true (Gomoku): the back-up is a plain load-add-store by the owning warp - a path never
// contains the same (node, action) twice there, which Pente's captures could in principle produce - so ordinary
// L1-cached loads stay coherent (same SM, write-through) and the root and the upper levels of the tree, read by every
// simulation of a launch, come from L1 instead of paying an L2 round trip per level.
template <bool L1, typename T>
__device__ __forceinline__ T ldx(const T* p) { return L1 ? __ldca(p) : __ldcg(p); }

// Everything the selection step needs from one node, fetched with ONE round trip: the child
// arrays (lane L: children 8L..8L+7), the flags, and the stored key for the transposition check.
struct NodeData {
  float p[8];
  int n[8], w[8];
  int c[8];          // child codes (only loaded when the engine keeps them)
  uint32_t meta;
  uint32_t keyw;     // lanes 0..15: key word l of the node
};
enum { AZG_CHILD_UNKNOWN = 0, AZG_CHILD_WON = 1, AZG_CHILD_DRAW = 2, AZG_CHILD_NODE0 = 4 };

template <bool L1 = false, bool CC = false>
__device__ __forceinline__ void node_load(const azg_dev& e, int g, int node, NodeData& nd) {
  const int l = lane_id();
  const size_t off = azg_node_off(e, g, node);
  const size_t base = off * AZG_ROW;
  nd.meta = ldx<L1>(&e.meta[off]);
  nd.keyw = l < 16 ? ldx<L1>(&e.key[off * 16 + l]) : 0u;
  if (l < 28) {
    const float4* P = reinterpret_cast<const float4*>(e.P + base + 8 * l);
    const int4* Nn = reinterpret_cast<const int4*>(e.Nv + base + 8 * l);
    const int4* Ww = reinterpret_cast<const int4*>(e.W + base + 8 * l);
    const float4 pa = ldx<L1>(P), pb = ldx<L1>(P + 1);
    const int4 na = ldx<L1>(Nn), nb = ldx<L1>(Nn + 1);
    const int4 wa = ldx<L1>(Ww), wb = ldx<L1>(Ww + 1);
    nd.p[0] = pa.x; nd.p[1] = pa.y; nd.p[2] = pa.z; nd.p[3] = pa.w; nd.p[4] = pb.x; nd.p[5] = pb.y; nd.p[6] = pb.z; nd.p[7] = pb.w;
    nd.n[0] = na.x; nd.n[1] = na.y; nd.n[2] = na.z; nd.n[3] = na.w; nd.n[4] = nb.x; nd.n[5] = nb.y; nd.n[6] = nb.z; nd.n[7] = nb.w;
    nd.w[0] = wa.x; nd.w[1] = wa.y; nd.w[2] = wa.z; nd.w[3] = wa.w; nd.w[4] = wb.x; nd.w[5] = wb.y; nd.w[6] = wb.z; nd.w[7] = wb.w;
    if (CC) {
      const int4* Cc = reinterpret_cast<const int4*>(e.child + base + 8 * l);
      const int4 ca = ldx<L1>(Cc), cb = ldx<L1>(Cc + 1);
      nd.c[0] = ca.x; nd.c[1] = ca.y; nd.c[2] = ca.z; nd.c[3] = ca.w; nd.c[4] = cb.x; nd.c[5] = cb.y; nd.c[6] = cb.z; nd.c[7] = cb.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { nd.p[j] = 0.f; nd.n[j] = 0; nd.w[j] = 0; nd.c[j] = 0; }
    if (l == 28) {
      nd.p[0] = ldx<L1>(e.P + base + 224); nd.n[0] = ldx<L1>(e.Nv + base + 224); nd.w[0] = ldx<L1>(e.W + base + 224);
      if (CC) nd.c[0] = ldx<L1>(e.child + base + 224);
    }
  }
}

__device__ __forceinline__ bool node_key_matches(const NodeData& nd, const WPos& p) {
  const int l = lane_id();
  const uint32_t x0 = __shfl_sync(AZG_FULL, p.w0, (l & 7) << 2);
  const uint32_t x1 = __shfl_sync(AZG_FULL, p.w1, (l & 7) << 2);
  bool ok = ((nd.meta >> 1) & 3u) == (uint32_t)p.player;
  if (l < 16) ok = ok && nd.keyw == (l < 8 ? x0 : x1);
  return __all_sync(AZG_FULL, ok);
}

// Transposition lookup that also fetches the node: the candidate's arrays are requested together
// with its key, so a hit costs two dependent memory round trips (probe window, node) instead of
// three.  The full key is always compared (the tag only selects candidates).
// This lane's slot of the first probe window of hash h: issued by the caller BEFORE the win test of the position, so
// that the test's ~300 cycles of shuffles and ballots run under the load's latency (table_find_load takes it as `first`).
template <bool L1 = false>
__device__ __forceinline__ unsigned long long table_first_window(const azg_dev& e, int g, unsigned long long h) {
  const unsigned long long* tab = e.slots + (size_t)g * (size_t)e.hcap;
  const int nwin = e.hcap >> 5;
  const int win = (int)((uint32_t)h & (uint32_t)(nwin - 1));
  return ldx<L1>(&tab[(win << 5) + lane_id()]);
}

template <bool L1 = false, bool CC = false, bool PRE = false>
__device__ __forceinline__ int table_find_load(const azg_dev& e, int g, const WPos& p, unsigned long long h, int* ins,
                                               NodeData& nd, unsigned long long first = 0ULL, unsigned long long* snap = nullptr) {
  const unsigned long long* tab = e.slots + (size_t)g * (size_t)e.hcap;
  const uint32_t tag = (uint32_t)(h >> 32);
  const int nwin = e.hcap >> 5;
  int win = (int)((uint32_t)h & (uint32_t)(nwin - 1));
  const int l = lane_id();
  for (int t = 0; t < nwin; ++t) {
    const unsigned long long s = (PRE && t == 0) ? first : ldx<L1>(&tab[(win << 5) + l]);
    uint32_t mm = __ballot_sync(AZG_FULL, s != 0ULL && (uint32_t)(s >> 32) == tag);
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      const int node = (int)__shfl_sync(AZG_FULL, (uint32_t)s, src) - 1;
      node_load<L1, CC>(e, g, node, nd);
      if (node_key_matches(nd, p)) return node;
    }
    const uint32_t em = __ballot_sync(AZG_FULL, s == 0ULL);
    if (em) { *ins = (win << 5) + __ffs(em) - 1; if (snap) *snap = s; return -1; }      // snap: this lane's slot of the window of *ins
    win = (win + 1) & (nwin - 1);
  }
  *ins = -1;
  return -1;
}

// First-index argmax of the reference's PUCT score over legal children.
__device__ __forceinline__ int puct_select(const azg_dev& e, int g, const NodeData& nd, uint32_t legal) {
  const int l = lane_id();
  const int slot64 = (int)((nd.meta >> 4) & 15u) - 1;
  const float* p = nd.p;
  const int* n = nd.n;
  const int* w = nd.w;
  int tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) tot += n[j];
  tot = __reduce_add_sync(AZG_FULL, tot);

  int best_i = 0x7fffffff;
  if (slot64 < 0) {
    // float32: ucb = W/(1+N) + ((cpuct*P)*sqrt_sum)/(1+N)     (new_mcts_alpha.py:136-137)
    // Unvisited children divide by exactly 1.0f and a zero numerator gives +0, so both cases skip
    // div.rn (whose zero-numerator slow path was 19 % of this kernel's instructions); results are
    // bit-identical to the straight formula.
    const float sq = __fsqrt_rn((float)tot);
    float best = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (legal & (1u << j)) {
        const float x = __fmul_rn(__fmul_rn(e.cpuct, p[j]), sq);
        float q = (float)w[j], u = x;
        if (n[j] != 0) {
          const float n1 = __fadd_rn(1.0f, (float)n[j]);
          if (w[j] != 0) q = __fdiv_rn(q, n1);
          if (x != 0.0f) u = __fdiv_rn(x, n1);
        }
        const float s = __fadd_rn(q, u);
        if (s > best) { best = s; best_i = 8 * l + j; }
      }
    }
    // warp argmax, first index on ties: lanes own ascending index ranges, so the lowest lane holding the
    // maximum wins.  Scores are never NaN and never -0 (q + u with u >= +0), so the usual order-preserving
    // map float -> uint32 is exact.
    const uint32_t bits = __float_as_uint(best);
    const uint32_t key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
    const uint32_t top = __reduce_max_sync(AZG_FULL, key);
    const int src = __ffs(__ballot_sync(AZG_FULL, key == top)) - 1;
    best_i = __shfl_sync(AZG_FULL, best_i, src);
  } else {
    // float64 at a noised root: P is float64 there, so numpy promotes the exploration term
    // (SURVEY 0.6); the W/(1+N) term is still rounded to float32 first.
    const double* P64 = e.P64 + ((size_t)g * AZG_P64_SLOTS + slot64) * AZG_ROW;
    const double sq = sqrt((double)tot);
    double best = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = 8 * l + j;
      if ((legal & (1u << j)) && idx < AZG_A) {
        const double x = __dmul_rn(__dmul_rn(e.cpuct64, __ldcg(P64 + idx)), sq);
        float q = (float)w[j];
        double u = x;
        if (n[j] != 0) {
          const float n1 = __fadd_rn(1.0f, (float)n[j]);
          if (w[j] != 0) q = __fdiv_rn(q, n1);
          if (x != 0.0) u = __ddiv_rn(x, (double)n1);
        }
        const double s = __dadd_rn((double)q, u);
        if (s > best) { best = s; best_i = idx; }
      }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const double ob = __shfl_xor_sync(AZG_FULL, best, s);
      const int oi = __shfl_xor_sync(AZG_FULL, best_i, s);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
  }
  return best_i;
}

// ------------------------------------------------------------------------------------------------
// FILL
// ------------------------------------------------------------------------------------------------
template <bool L1, bool CC>
__device__ __forceinline__ void fill_body(const azg_dev& e) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  const int l = lane_id();
  azg_ctl* ctl = e.ctl + g;
  int state = __ldcg(&ctl->state);
  if (state != AZG_ST_RUN) return;

  int sims_left = __ldcg(&ctl->sims_left);
  int root_node = __ldcg(&ctl->root_node);
  int n_pending = __ldcg(&ctl->n_pending);
  int n_nodes = __ldcg(&ctl->n_nodes);
  int n_free = __ldcg(&ctl->n_free);
  int n_live = __ldcg(&ctl->n_live);
  int err = 0;
  bool resume = __ldcg(&ctl->susp) != 0;
  unsigned long long visits = 0, sims = 0;
  uint32_t* path = e.path + (size_t)g * AZG_MAX_DEPTH;
  int32_t* freelist = e.freelist + (size_t)g * e.cap;
  const WPos root = wpos_load(&ctl->root);

  while (true) {
    WPos pos;
    NodeData nd;
    int depth, node = -1, v = 0;
    int par_node = -1, par_a = 0;              // where this position was reached from (child codes)
    bool select_here = false;
    if (resume) {
      pos = wpos_load(&ctl->scratch);
      depth = __ldcg(&ctl->depth);
      node = __ldcg(&ctl->resume_node);
      node_load<L1, CC>(e, g, node, nd);
      select_here = true;                       // fall through to selection at the evaluated leaf
      resume = false;
    } else {
      if (sims_left <= 0) break;
      pos = root;
      depth = 0;
    }
    bool parked = false;
    for (;;) {
      if (!select_here) {
        ++visits;
        // hash and first probe window first: the load is in flight while the win test runs (a terminal position wastes it)
        const bool at_root = depth == 0 && root_node >= 0;
        const unsigned long long h = wpos_hash(pos);
        const unsigned long long first = at_root ? 0ULL : table_first_window<L1>(e, g, h);
        const int won = wpos_winner(pos, e.rule);            // new_mcts_alpha.py:106-112
        if (won != 0) { v = -1; if (CC && par_node >= 0 && l == 0) e.child[azg_node_off(e, g, par_node) * AZG_ROW + par_a] = AZG_CHILD_WON; break; }
        if (!wpos_any_empty(pos)) { v = 0; if (CC && par_node >= 0 && l == 0) e.child[azg_node_off(e, g, par_node) * AZG_ROW + par_a] = AZG_CHILD_DRAW; break; }
        int ins = -1;
        if (at_root) { node = root_node; node_load<L1, CC>(e, g, node, nd); }   // the root key is fixed for the run
        else node = table_find_load<L1, CC, true>(e, g, pos, h, &ins, nd, first);
        if (CC && node >= 0 && par_node >= 0 && l == 0) e.child[azg_node_off(e, g, par_node) * AZG_ROW + par_a] = node + AZG_CHILD_NODE0;
        if (node < 0) {                                        // new_mcts_alpha.py:114-132
          if (ins < 0) { err |= AZG_ERR_HASH; break; }
          if (n_free > 0) node = __ldcg(&freelist[--n_free]);
          else if (n_nodes < e.cap) node = n_nodes++;
          else { err |= AZG_ERR_NODES; break; }
          ++n_live;
          node_write_key(e, g, node, pos);
          table_put(e, g, ins, h, node);
          if (CC && par_node >= 0 && l == 0) e.child[azg_node_off(e, g, par_node) * AZG_ROW + par_a] = node + AZG_CHILD_NODE0;
          if (depth == 0) root_node = node;
          if (l == 0) ctl->pending[n_pending] = node;
          ++n_pending;
          if (n_pending >= e.queue_len) {                     // park: the queue is evaluated first
            wpos_store(&ctl->scratch, pos);
            if (l == 0) { ctl->depth = depth; ctl->resume_node = node; ctl->susp = 1; }
            parked = true;
            break;
          }
          const int cnt = wpos_count_empty(pos);
          node_fill(e, g, node, wpos_legal_byte(pos), __fdiv_rn(1.0f, (float)cnt));
          v = 0;
          break;
        }
      }
      select_here = false;
      const int a = puct_select(e, g, nd, wpos_legal_byte(pos));
      if (depth >= AZG_MAX_DEPTH) { err |= AZG_ERR_DEPTH; break; }
      if (l == 0) __stcg(&path[depth], ((uint32_t)node << 8) | (uint32_t)a);
      ++depth;
      wpos_play(pos, e.rule, a);
      if (CC) {
        // Child codes (Gomoku): what this (node, action) led to the last time.  A terminal child needs no win test, a known
        // child no hash and no table probe - its arrays are fetched at once and the stored key is compared as always; a
        // stale code (slot reused, never the case for a live parent in Gomoku) just falls back to the full path.
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) if ((a & 7) == j) mine = nd.c[j];
        const int code = __shfl_sync(AZG_FULL, mine, a >> 3);
        par_node = node; par_a = a;
        if (code == AZG_CHILD_WON) { ++visits; v = -1; break; }
        if (code == AZG_CHILD_DRAW) { ++visits; v = 0; break; }
        if (code >= AZG_CHILD_NODE0) {
          const int child = code - AZG_CHILD_NODE0;
          node_load<L1, CC>(e, g, child, nd);
          if ((nd.meta & AZG_META_ALIVE) && node_key_matches(nd, pos)) { ++visits; node = child; select_here = true; }
        }
      }
    }
    if (parked || err) break;
    // back-up: the leaf value alternates sign up the path (new_mcts_alpha.py:146-151)
    __threadfence_block();
    __syncwarp();
    for (int d = l; d < depth; d += 32) {
      const uint32_t pe = __ldcg(&path[d]);
      const size_t idx = azg_node_off(e, g, (int)(pe >> 8)) * AZG_ROW + (pe & 255u);
      if (L1) {           // the owning warp is the only writer, and a Gomoku path holds every (node, action) once
        e.Nv[idx] = __ldca(&e.Nv[idx]) + 1;
        if (v != 0) e.W[idx] = __ldca(&e.W[idx]) + (((depth - d) & 1) ? -v : v);
      } else {
        atomicAdd(&e.Nv[idx], 1);
        if (v != 0) atomicAdd(&e.W[idx], ((depth - d) & 1) ? -v : v);
      }
    }
    __threadfence_block();
    __syncwarp();
    --sims_left;
    ++sims;
  }

  if (err) state = AZG_ST_ERROR;
  else if (n_pending >= e.queue_len) state = AZG_ST_NEED_EVAL;
  else if (sims_left <= 0) state = (n_pending > 0) ? AZG_ST_NEED_FINAL : AZG_ST_DONE;
  if (l == 0) {
    ctl->state = state;
    ctl->err = __ldcg(&ctl->err) | err;
    ctl->sims_left = sims_left;
    ctl->root_node = root_node;
    ctl->n_pending = n_pending;
    ctl->n_nodes = n_nodes;
    ctl->n_free = n_free;
    ctl->n_live = n_live;
    ctl->visits += visits;
    ctl->sims += sims;
  }
}

extern "C" __global__ void __launch_bounds__(128) azg_fill_kernel(azg_dev e) {
  if (e.rule == AZG_RULE_GOMOKU && e.fill_l1) {
    if (e.child) fill_body<true, true>(e); else fill_body<true, false>(e);
  } else fill_body<false, false>(e);
}
// ------------------------------------------------------------------------------------------------
// FILL, fast mode (NOT the reference's algorithm: an explicit non-parity option, azg_config.fast_warps > 0)
// ------------------------------------------------------------------------------------------------
// The reference walks its simulations one after the other, which is what the kernel above reproduces exactly and what
// bounds the latency of a single game (one warp, ~3 us per tree level).  Here `fast_warps` warps of one block walk ONE
// game's tree at the same time, kept apart by a virtual loss: on the way down every selected (node, action) gets
// N += vl and W -= vl with L2 atomics, the back-up takes it out again and adds the real visit.  Lookups and selection
// are lock-free; creating a node (slab allocation, key, placeholder priors, table entry, leaf queue) happens under a
// per-game spin lock and the node is published in the table last.  Kept from the reference: the deferred-evaluation
// queue (placeholder nodes with a uniform-legal prior, evaluated `queue_len` at a time, N and W reset on evaluation),
// terminal values only, masked priors, root noise.  Dropped: the strict order of simulations and the quirk that the
// simulation which fills the queue continues below its leaf.  Visit counts therefore differ from the reference's; the
// tests check invariants, tactics and agreement of the chosen move instead (tests/test_fast_mode_gpu.py).
#define AZG_FAST_MAX_WARPS 16
#define AZG_FAST_MAX_DEPTH 160

extern "C" __global__ void __launch_bounds__(32 * AZG_FAST_MAX_WARPS) azg_fill_fast_kernel(azg_dev e) {
  __shared__ int s_state, s_sims_left, s_n_pending, s_lock, s_n_nodes, s_n_free, s_n_live, s_err, s_stop, s_root_node;
  __shared__ unsigned int s_visits, s_sims;
  __shared__ uint32_t s_path[AZG_FAST_MAX_WARPS][AZG_FAST_MAX_DEPTH];
  __shared__ int s_spare[AZG_FAST_MAX_WARPS];
  const int g = blockIdx.x;
  const int l = lane_id(), wib = threadIdx.x >> 5;
  azg_ctl* ctl = e.ctl + g;
  if (threadIdx.x == 0) {
    s_state = __ldcg(&ctl->state);
    s_sims_left = __ldcg(&ctl->sims_left); s_n_pending = __ldcg(&ctl->n_pending); s_n_nodes = __ldcg(&ctl->n_nodes);
    s_n_free = __ldcg(&ctl->n_free); s_n_live = __ldcg(&ctl->n_live); s_root_node = __ldcg(&ctl->root_node);
    s_lock = 0; s_err = 0; s_stop = 0; s_visits = 0; s_sims = 0;
  }
  __syncthreads();
  if (s_state != AZG_ST_RUN) return;
  const int vl = e.virtual_loss > 0 ? e.virtual_loss : 1;
  int32_t* freelist = e.freelist + (size_t)g * e.cap;
  const WPos root = wpos_load(&ctl->root);
  uint32_t* path = s_path[wib];
  int spare = -1;                         // a slab node this warp built but did not publish (somebody else made the position first)

  while (true) {
    int go = 0;
    if (l == 0) {
      if (!*(volatile int*)&s_stop && !*(volatile int*)&s_err) {
        if (atomicSub(&s_sims_left, 1) > 0) go = 1; else atomicAdd(&s_sims_left, 1);
      }
    }
    go = __shfl_sync(AZG_FULL, go, 0);
    if (!go) break;
    WPos pos = root;
    NodeData nd;
    int depth = 0, v = 0, visits = 0;
    bool abort = false;
    for (;;) {
      ++visits;
      const int won = wpos_winner(pos, e.rule);
      if (won != 0) { v = -1; break; }
      if (!wpos_any_empty(pos)) { v = 0; break; }
      const unsigned long long h = wpos_hash(pos);
      int ins = -1, node;
      unsigned long long snap = 0ULL;
      const int rn = *(volatile int*)&s_root_node;
      if (depth == 0 && rn >= 0) { node = rn; node_load(e, g, node, nd); }
      else node = table_find_load(e, g, pos, h, &ins, nd, 0ULL, &snap);
      if (node < 0) {
        // ---- create the leaf.  The node is taken from the slab and BUILT (key, placeholder priors, zeroed statistics)
        // outside the lock - nobody can find it yet; the lock only covers "is the position still missing?" (one reload of
        // the probe window: slots are only ever filled, and every warp looking for this position ends at this window),
        // the table entry and the queue entry.  A node that turns out not to be needed stays with the warp as its spare.
        if (ins < 0) { if (l == 0) atomicOr(&s_err, AZG_ERR_HASH); abort = true; break; }
        int nn = spare;
        if (nn < 0) {
          int got = -1, fail = 0;
          if (l == 0) {
            const int old = atomicSub(&s_n_free, 1);
            if (old > 0) got = __ldcg(&freelist[old - 1]);
            else {
              atomicAdd(&s_n_free, 1);
              const int n = atomicAdd(&s_n_nodes, 1);
              if (n < e.cap) got = n; else { atomicSub(&s_n_nodes, 1); fail = AZG_ERR_NODES; }
            }
          }
          nn = __shfl_sync(AZG_FULL, got, 0); fail = __shfl_sync(AZG_FULL, fail, 0);
          if (fail) { if (l == 0) atomicOr(&s_err, fail); abort = true; break; }
        }
        spare = nn;
        node_write_key(e, g, nn, pos);
        node_fill(e, g, nn, wpos_legal_byte(pos), __fdiv_rn(1.0f, (float)wpos_count_empty(pos)));
        __threadfence_block();
        __syncwarp();
        int got = 0;
        if (l == 0) {
          for (int spin = 0; spin < (1 << 24); ++spin)
            if (atomicCAS(&s_lock, 0, 1) == 0) { got = 1; break; }
        }
        got = __shfl_sync(AZG_FULL, got, 0);
        if (!got) { if (l == 0) atomicOr(&s_err, AZG_ERR_HASH); abort = true; break; }
        __threadfence_block();
        const unsigned long long now = __ldcg(&e.slots[(size_t)g * (size_t)e.hcap + (size_t)((ins & ~31) + l)]);
        node = -1;
        if (__any_sync(AZG_FULL, now != snap)) node = table_find_load(e, g, pos, h, &ins, nd);   // the window changed: look again
        bool done = false;
        if (node < 0) {
          int np = 0, fail = 0;
          if (l == 0) {
            np = *(volatile int*)&s_n_pending;
            if (ins < 0) fail = AZG_ERR_HASH;
            else if (np >= e.queue_len) fail = -1;                // the queue filled up: this simulation is given back
          }
          np = __shfl_sync(AZG_FULL, np, 0); fail = __shfl_sync(AZG_FULL, fail, 0);
          if (fail) { if (fail > 0 && l == 0) atomicOr(&s_err, fail); abort = true; done = true; }
          else {
            table_put(e, g, ins, h, nn);                          // published last: lock-free readers find a complete node
            if (l == 0) {
              ctl->pending[np] = nn;
              *(volatile int*)&s_n_pending = np + 1;
              atomicAdd(&s_n_live, 1);
              if (depth == 0) *(volatile int*)&s_root_node = nn;
              if (np + 1 >= e.queue_len) *(volatile int*)&s_stop = 1;
            }
            spare = -1;
            v = 0;
            done = true;
          }
        }
        __threadfence_block();
        __syncwarp();
        if (l == 0) atomicExch(&s_lock, 0);
        __syncwarp();
        if (done) break;
      }
      const int a = puct_select(e, g, nd, wpos_legal_byte(pos));
      if (depth >= AZG_FAST_MAX_DEPTH) { if (l == 0) atomicOr(&s_err, AZG_ERR_DEPTH); abort = true; break; }
      if (l == 0) {
        path[depth] = ((uint32_t)node << 8) | (uint32_t)a;
        const size_t idx = azg_node_off(e, g, node) * AZG_ROW + a;
        atomicAdd(&e.Nv[idx], vl);                              // virtual loss: this edge looks visited and lost to the others
        atomicAdd(&e.W[idx], -vl);
      }
      ++depth;
      wpos_play(pos, e.rule, a);
    }
    __syncwarp();
    // back-up: take the virtual loss out, add the real visit (value alternates sign up the path)
    for (int d = l; d < depth; d += 32) {
      const uint32_t pe = path[d];
      const size_t idx = azg_node_off(e, g, (int)(pe >> 8)) * AZG_ROW + (pe & 255u);
      if (abort) { atomicAdd(&e.Nv[idx], -vl); atomicAdd(&e.W[idx], vl); }
      else {
        atomicAdd(&e.Nv[idx], 1 - vl);
        atomicAdd(&e.W[idx], vl + (v != 0 ? (((depth - d) & 1) ? -v : v) : 0));
      }
    }
    __syncwarp();
    if (l == 0) {
      atomicAdd(&s_visits, (unsigned)visits);
      if (abort) atomicAdd(&s_sims_left, 1); else atomicAdd(&s_sims, 1u);
    }
    if (abort) break;                                           // queue full or error: this warp is done for this launch
  }
  if (l == 0) s_spare[wib] = spare;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w)      // unused nodes go back to the free stack, marked dead for the next GC
      if (s_spare[w] >= 0) { e.meta[azg_node_off(e, g, s_spare[w])] = 0u; freelist[s_n_free++] = s_spare[w]; }
    int state = AZG_ST_RUN;
    const int left = s_sims_left < 0 ? 0 : s_sims_left;
    if (s_err) state = AZG_ST_ERROR;
    else if (s_n_pending >= e.queue_len) state = AZG_ST_NEED_EVAL;
    else if (left <= 0) state = s_n_pending > 0 ? AZG_ST_NEED_FINAL : AZG_ST_DONE;
    ctl->state = state;
    ctl->err = __ldcg(&ctl->err) | s_err;
    ctl->sims_left = left;
    ctl->root_node = s_root_node;
    ctl->n_pending = s_n_pending;
    ctl->n_nodes = s_n_nodes;
    ctl->n_free = s_n_free;
    ctl->n_live = s_n_live;
    ctl->susp = 0;
    ctl->visits += s_visits;
    ctl->sims += s_sims;
  }
}

// ------------------------------------------------------------------------------------------------
// leaf batch assembly: exclusive scan of the per-game queue lengths (single block)
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(1024) azg_scan_kernel(azg_dev e) {
  __shared__ int part[1024], s_n[1024], s_off[1024], s_rn[1024];
  __shared__ int carry, s_active, s_errors, s_roots;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  if (t == 0) { carry = 0; s_active = 0; s_errors = 0; s_roots = 0; }
  __syncthreads();
  for (int base = 0; base < e.G; base += 1024) {
    const int g = base + t;
    int n = 0, st = AZG_ST_DONE, rn = -1;
    if (g < e.G) {
      st = e.ctl[g].state;
      if (st == AZG_ST_NEED_EVAL || st == AZG_ST_NEED_FINAL) n = e.ctl[g].n_pending;
      rn = e.ctl[g].root_node;
    }
    const int act = __popc(__ballot_sync(AZG_FULL, st == AZG_ST_RUN || st == AZG_ST_NEED_EVAL));   // needs another fill after the commit
    const int bad = __popc(__ballot_sync(AZG_FULL, st == AZG_ST_ERROR));
    if (l == 0 && (act | bad)) { atomicAdd(&s_active, act); atomicAdd(&s_errors, bad); }
    part[t] = n;
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {
      const int x = (t >= s) ? part[t - s] : 0;
      __syncthreads();
      part[t] += x;
      __syncthreads();
    }
    const int off = carry + part[t] - n;
    if (g < e.G) e.ctl[g].leaf_off = off;
    s_n[t] = n; s_off[t] = off; s_rn[t] = rn;
    __syncthreads();
    // the queues are copied by warps: lane i moves entry i of one game (coalesced on both sides)
    int roots = 0;
    for (int k = 0; k < 32; ++k) {
      const int j = w * 32 + k, gg = base + j;
      const int nn = s_n[j];
      for (int i = l; i < nn; i += 32) {
        const int node = e.ctl[gg].pending[i];
        e.leaf_game[s_off[j] + i] = gg;
        e.leaf_node[s_off[j] + i] = node;
        roots += node == s_rn[j];
      }
    }
    roots = __reduce_add_sync(AZG_FULL, roots);
    if (l == 0 && roots) atomicAdd(&s_roots, roots);
    __syncthreads();
    if (t == 1023) carry += part[t];
    __syncthreads();
  }
  if (t == 0) { e.counters[0] = carry; e.counters[1] = s_active; e.counters[2] = s_errors; e.counters[3] = s_roots; }
}

// Encoded planes of every queued leaf, float32 NCHW exactly as get_encoded_state()
// (games/gomoku.py:130-150): side-to-move stones, opponent stones, ones.
extern "C" __global__ void azg_leaf_planes_kernel(azg_dev e, float* out) {
  const int n = e.counters[0];
  for (int leaf = blockIdx.x; leaf < n; leaf += gridDim.x) {
    const int g = e.leaf_game[leaf], node = e.leaf_node[leaf];
    const size_t off = azg_node_off(e, g, node);
    const int player = (e.meta[off] >> 1) & 3;
    const uint32_t* k = e.key + off * 16;
    float* o = out + (size_t)leaf * 3 * AZG_A;
    for (int a = threadIdx.x; a < AZG_A; a += blockDim.x) {
      const uint32_t b1 = (k[a >> 5] >> (a & 31)) & 1u, b2 = (k[8 + (a >> 5)] >> (a & 31)) & 1u;
      o[a] = (float)(player == 1 ? b1 : b2);
      o[AZG_A + a] = (float)(player == 1 ? b2 : b1);
      o[2 * AZG_A + a] = 1.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// COMMIT
// ------------------------------------------------------------------------------------------------
// numpy's pairwise float sum for n = 225 (n > 128: halves of 112 and 113; each half keeps
// eight running partials; the odd last element is added at the end).  Bit-exact with
// np.sum on a contiguous vector, which is what new_mcts_alpha.py:167 and :174 call.
template <typename T>
__device__ __forceinline__ T numpy_sum225(const T* sm) {
  const int l = lane_id();
  T r = (T)0;
  if (l < 16) {
    const T* a = sm + (l >> 3) * 112 + (l & 7);
    r = a[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) r = r + a[8 * i];
  }
  // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) inside each group of 8 lanes
  r = r + __shfl_down_sync(AZG_FULL, r, 1);
  r = r + __shfl_down_sync(AZG_FULL, r, 2);
  r = r + __shfl_down_sync(AZG_FULL, r, 4);
  T lo = __shfl_sync(AZG_FULL, r, 0);
  T hi = __shfl_sync(AZG_FULL, r, 8);
  hi = hi + sm[224];
  return lo + hi;
}

// One block per game, one warp per queued leaf (the leaves of a queue are distinct nodes, so they commit
// independently; at most one of them is the root, which alone may take a float64 slot).
extern "C" __global__ void __launch_bounds__(1024)
azg_commit_kernel(azg_dev e, const float* __restrict__ probs, const double* __restrict__ noise) {
  __shared__ float sm_f[32][AZG_ROW];
  __shared__ double sm_d[AZG_ROW];
  __shared__ int s_err;
  const int g = blockIdx.x;
  const int l = lane_id();
  const int wib = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  azg_ctl* ctl = e.ctl + g;
  const int state = ctl->state;
  if (state != AZG_ST_NEED_EVAL && state != AZG_ST_NEED_FINAL) return;      // block-uniform
  const int n_pending = ctl->n_pending;
  const int leaf_off = ctl->leaf_off;
  const int root_node = ctl->root_node;
  const bool noisy_run = e.noise_on && ctl->ply < e.noise_plies && noise != nullptr;
  const int p64_used = ctl->p64_used;
  if (threadIdx.x == 0) s_err = 0;
  __syncthreads();                                                          // every read of ctl precedes the final update
  float* smf = sm_f[wib];
  double* smd = sm_d;

  for (int i = wib; i < n_pending; i += n_warps) {
    const int node = ctl->pending[i];
    const size_t off = azg_node_off(e, g, node);
    const size_t base = off * AZG_ROW;
    const uint32_t occ = e.key[off * 16 + (l >> 2)] | e.key[off * 16 + 8 + (l >> 2)];
    const uint32_t empty = ~occ & board_word_mask(l >> 2);
    const uint32_t legal = (empty >> ((l & 3) * 8)) & 0xffu;
    const float* row = probs + (size_t)(leaf_off + i) * AZG_A;
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = 8 * l + j;
      p[j] = 0.f;
      if (idx < AZG_A) {
        p[j] = __fmul_rn(row[idx], (legal & (1u << j)) ? 1.0f : 0.0f);     // p * valid (new_mcts_alpha.py:166)
        smf[idx] = p[j];
      }
    }
    __syncwarp();
    const float total = numpy_sum225<float>(smf);
    __syncwarp();
    if (total < 1e-8f) {                                                    // uniform-legal fallback (:167-168)
      int c = (l & 3) == 0 ? __popc(empty) : 0;
      c = __reduce_add_sync(AZG_FULL, c);
      const float u = __fdiv_rn(1.0f, (float)c);
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = (legal & (1u << j)) ? u : 0.f;
    }
    uint32_t meta = AZG_META_ALIVE | (e.meta[off] & 6u);
    if (noisy_run && node == root_node) {                                   // root Dirichlet mix (:171-174)
      int slot = -1;
      for (int s = 0; s < AZG_P64_SLOTS; ++s) if (!(p64_used & (1 << s))) { slot = s; break; }
      if (slot < 0) { if (l == 0) atomicOr(&s_err, AZG_ERR_P64); }
      else {
        if (l == 0) ctl->p64_used = p64_used | (1 << slot);                 // only this warp touches the field
        meta |= (uint32_t)(slot + 1) << 4;
        const double* nz = noise + (size_t)g * AZG_A;
        const float keep = (float)(1.0 - e.eps);                            // (1-eps)*p stays float32 (NEP 50)
        double q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = 8 * l + j;
          q[j] = 0.0;
          if (idx < AZG_A) {
            q[j] = __dadd_rn((double)__fmul_rn(keep, p[j]), __dmul_rn(e.eps, nz[idx]));
            smd[idx] = q[j];
          }
        }
        __syncwarp();
        const double tot = numpy_sum225<double>(smd);
        __syncwarp();
        double* P64 = e.P64 + ((size_t)g * AZG_P64_SLOTS + slot) * AZG_ROW;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = 8 * l + j;
          if (idx < AZG_A) { P64[idx] = __ddiv_rn(q[j], tot); p[j] = (float)P64[idx]; }
        }
      }
    }
    if (l < 28) {
      float4* P = reinterpret_cast<float4*>(e.P + base + 8 * l);
      P[0] = make_float4(p[0], p[1], p[2], p[3]);
      P[1] = make_float4(p[4], p[5], p[6], p[7]);
      const int4 z = make_int4(0, 0, 0, 0);
      int4* Nn = reinterpret_cast<int4*>(e.Nv + base + 8 * l);
      int4* Ww = reinterpret_cast<int4*>(e.W + base + 8 * l);
      Nn[0] = z; Nn[1] = z; Ww[0] = z; Ww[1] = z;
    } else if (l == 28) {
      e.P[base + 224] = p[0]; e.Nv[base + 224] = 0; e.W[base + 224] = 0;
    }
    if (l == 0) e.meta[off] = meta;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int err = s_err;
    ctl->evals += (unsigned long long)n_pending;
    ctl->n_pending = 0;
    if (err) { ctl->err |= err; ctl->state = AZG_ST_ERROR; }
    else ctl->state = (state == AZG_ST_NEED_FINAL) ? AZG_ST_DONE : AZG_ST_RUN;
  }
}

// ------------------------------------------------------------------------------------------------
// run control
// ------------------------------------------------------------------------------------------------
// Start a run on every game (new_mcts_alpha.py:77-83): fix the root key, look it up.
extern "C" __global__ void __launch_bounds__(128)
azg_begin_kernel(azg_dev e, const int32_t* __restrict__ plies, int n_sims, const int32_t* __restrict__ mask) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  azg_ctl* ctl = e.ctl + g;
  if (mask && !mask[g]) {            // this game sits the run out (evaluation arena: the other model is to move)
    if (lane_id() == 0) { ctl->state = AZG_ST_DONE; ctl->err = 0; ctl->sims_left = 0; ctl->root_node = -1; ctl->susp = 0; ctl->n_pending = 0; }
    return;
  }
  const WPos root = wpos_load(&ctl->root);
  int state = AZG_ST_RUN, err = 0, root_node = -1;
  if (wpos_winner(root, e.rule) != 0 || !wpos_any_empty(root)) {
    // the reference would raise KeyError at new_mcts_alpha.py:89; reported as an error status
    state = AZG_ST_ERROR; err = AZG_ERR_ROOT_TERMINAL;
  } else {
    int ins;
    root_node = table_find(e, g, root, wpos_hash(root), &ins);
  }
  if (lane_id() == 0) {
    ctl->state = state; ctl->err = err; ctl->sims_left = n_sims;
    ctl->ply = plies ? plies[g] : root.plies;
    ctl->root_node = root_node; ctl->susp = 0; ctl->depth = 0; ctl->resume_node = -1; ctl->n_pending = 0;
  }
}

// pi = N[root]/sum(N[root]) in float32, uniform over legal moves when the sum is zero
// (new_mcts_alpha.py:88-97).  Also exports the raw visit counts.
extern "C" __global__ void __launch_bounds__(128)
azg_finish_kernel(azg_dev e, float* __restrict__ pi, int32_t* __restrict__ visits) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  const int l = lane_id();
  const azg_ctl* ctl = e.ctl + g;
  const int node = ctl->root_node;
  if (node < 0 || ctl->state == AZG_ST_ERROR) {
    for (int a = l; a < AZG_A; a += 32) { if (pi) pi[(size_t)g * AZG_A + a] = 0.f; if (visits) visits[(size_t)g * AZG_A + a] = 0; }
    return;
  }
  const size_t base = azg_node_off(e, g, node) * AZG_ROW;
  const WPos root = wpos_load(&ctl->root);
  int n[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int idx = 8 * l + j; n[j] = idx < AZG_A ? e.Nv[base + idx] : 0; tot += n[j]; }
  tot = __reduce_add_sync(AZG_FULL, tot);
  const uint32_t legal = wpos_legal_byte(root);
  const int cnt = wpos_count_empty(root);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int idx = 8 * l + j;
    if (idx < AZG_A) {
      float v;
      if (tot > 0) v = __fdiv_rn((float)n[j], (float)tot);
      else v = (legal & (1u << j)) ? __fdiv_rn(1.0f, (float)cnt) : 0.f;
      if (pi) pi[(size_t)g * AZG_A + idx] = v;
      if (visits) visits[(size_t)g * AZG_A + idx] = n[j];
    }
  }
}

// Play `actions[g]` on the root (skipped when < 0) and reclaim dead nodes.
//   Gomoku: stones are never removed, so a stored position can be reached again only
//           if it contains every stone of the new root.
//   Pente : a root stone missing from a stored position must have been captured on the
//           way, and no path the search follows lets a player exceed four captured
//           pairs before the position is terminal (games/pente.py:209), so a stored
//           position missing more than 2*(4-caps) root stones of a colour is dead.
// Everything that survives is re-inserted into a cleared table; freed slots go to the
// free stack.  gc == 0 keeps every node.
extern "C" __global__ void __launch_bounds__(128)
azg_advance_kernel(azg_dev e, const int32_t* __restrict__ actions, int gc, int reserve, int32_t* __restrict__ status) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  const int l = lane_id();
  azg_ctl* ctl = e.ctl + g;
  WPos root = wpos_load(&ctl->root);
  const int a = actions ? actions[g] : -1;
  int ok = 1;
  if (a >= 0) {
    const bool empty = a < AZG_A && wpos_at(root, a < AZG_A ? a : 0) == 0;
    if (empty) { wpos_play(root, e.rule, a); wpos_store(&ctl->root, root); }
    else ok = 0;
  }
  const int won = wpos_winner(root, e.rule);
  const bool over = won != 0 || !wpos_any_empty(root);
  if (status && l == 0) status[g] = (ok ? 0 : 8) | (over ? 4 : 0) | won;
  if (!gc || !ok) return;

  const int n_nodes = ctl->n_nodes;
  int n_free = 0, n_live = 0, p64_used = 0;
  int32_t* freelist = e.freelist + (size_t)g * e.cap;
  unsigned long long* tab = e.slots + (size_t)g * (size_t)e.hcap;
  for (int i = l; i < e.hcap; i += 32) __stcg(&tab[i], 0ULL);
  __threadfence_block();
  __syncwarp();
  const int lim1 = e.rule == 1 ? 2 * max(0, 4 - root.cap1) : 0;   // player-1 stones that player 2 may still capture
  const int lim2 = e.rule == 1 ? 2 * max(0, 4 - root.cap0) : 0;
  for (int node = 0; node < n_nodes; ++node) {
    const size_t off = azg_node_off(e, g, node);
    const uint32_t meta = e.meta[off];
    // this lane's word of the stored key, same distribution as WPos
    WPos k;
    k.w0 = e.key[off * 16 + (l >> 2)];
    k.w1 = e.key[off * 16 + 8 + (l >> 2)];
    k.player = (meta >> 1) & 3; k.last = -1; k.cap0 = k.cap1 = 0; k.plies = 0;
    int d1 = (l & 3) == 0 ? __popc(root.w0 & ~k.w0) : 0;
    int d2 = (l & 3) == 0 ? __popc(root.w1 & ~k.w1) : 0;
    d1 = __reduce_add_sync(AZG_FULL, d1);
    d2 = __reduce_add_sync(AZG_FULL, d2);
    const bool alive = (meta & AZG_META_ALIVE) && d1 <= lim1 && d2 <= lim2;
    if (alive) {
      const unsigned long long h = wpos_hash(k);
      int ins = -1;
      // keys are unique, so only the insertion point is needed
      const int nwin = e.hcap >> 5;
      int win = (int)((uint32_t)h & (uint32_t)(nwin - 1));
      for (int t = 0; t < nwin; ++t) {
        const unsigned long long s = __ldcg(&tab[(win << 5) + l]);
        const uint32_t em = __ballot_sync(AZG_FULL, s == 0ULL);
        if (em) { ins = (win << 5) + __ffs(em) - 1; break; }
        win = (win + 1) & (nwin - 1);
      }
      if (ins >= 0 && l == 0) __stcg(&tab[ins], ((h >> 32) << 32) | (unsigned long long)(uint32_t)(node + 1));
      __threadfence_block();
      __syncwarp();
      ++n_live;
      const int s64 = (int)((meta >> 4) & 15u) - 1;
      if (s64 >= 0) p64_used |= 1 << s64;
    } else {
      if (l == 0) { e.meta[off] = 0u; freelist[n_free] = node; }
      ++n_free;
    }
  }
  // Not enough room for the next run (it can add up to `reserve` nodes): drop the whole tree
  // rather than fail mid-run.  This loses tree reuse for this game (a deviation from the
  // reference, which never frees) and is counted in azg_search_stats.
  if (reserve > 0 && e.cap - n_live < reserve) {
    for (int i = l; i < e.hcap; i += 32) __stcg(&tab[i], 0ULL);
    for (int i = l; i < n_nodes; i += 32) e.meta[azg_node_off(e, g, i)] = 0u;
    if (l == 0) { ctl->n_nodes = 0; ctl->n_free = 0; ctl->n_live = 0; ctl->p64_used = 0; ctl->resets += 1; }
    return;
  }
  if (l == 0) { ctl->n_free = n_free; ctl->n_live = n_live; ctl->p64_used = p64_used; }
}

// Forget every node of the selected games (MCTS.clear_tree, new_mcts_alpha.py:58-72)
// and optionally load new root positions.
extern "C" __global__ void __launch_bounds__(128)
azg_reset_kernel(azg_dev e, const int32_t* __restrict__ mask, const azg_pos* __restrict__ roots, int clear_tree) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= e.G) return;
  if (mask && !mask[g]) return;
  const int l = lane_id();
  azg_ctl* ctl = e.ctl + g;
  if (roots) {
    const WPos p = wpos_load(roots + g);
    wpos_store(&ctl->root, p);
  }
  if (clear_tree) {
    unsigned long long* tab = e.slots + (size_t)g * (size_t)e.hcap;
    for (int i = l; i < e.hcap; i += 32) tab[i] = 0ULL;
    if (l == 0) {
      ctl->n_nodes = 0; ctl->n_free = 0; ctl->n_live = 0; ctl->p64_used = 0; ctl->n_pending = 0; ctl->susp = 0;
      ctl->root_node = -1; ctl->state = AZG_ST_IDLE; ctl->err = 0;
    }
  }
}

// Aggregate counters over all games (single block).
extern "C" __global__ void __launch_bounds__(256) azg_stats_kernel(azg_dev e, unsigned long long* out) {
  __shared__ unsigned long long acc[8];
  if (threadIdx.x < 8) acc[threadIdx.x] = 0ULL;
  __syncthreads();
  for (int g = threadIdx.x; g < e.G; g += blockDim.x) {
    const azg_ctl* c = e.ctl + g;
    atomicAdd(&acc[0], c->sims);
    atomicAdd(&acc[1], c->visits);
    atomicAdd(&acc[2], c->evals);
    atomicAdd(&acc[3], (unsigned long long)c->n_live);
    atomicMax(&acc[4], (unsigned long long)c->n_nodes);
    if (c->state == AZG_ST_ERROR || c->err) atomicAdd(&acc[5], 1ULL);
    atomicOr(&acc[6], (unsigned long long)c->err);
    atomicAdd(&acc[7], (unsigned long long)c->resets);
  }
  __syncthreads();
  if (threadIdx.x < 8) out[threadIdx.x] = acc[threadIdx.x];
}

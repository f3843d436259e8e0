// search_layer / search_knn for one warp per query (reference: src/hnsw/core.rs:607-675, 865-892).
//
// Formulation.  The reference keeps two heaps, `c` (candidates, unbounded) and `w` (results, <= ef).  Every
// member of `w` is also pushed to `c`, an element leaves `w` only by being the worst when a better one arrives,
// and the loop stops when the nearest element of `c` is farther than the worst of `w` (core.rs:635).  An element
// evicted from `w` is never expanded afterwards (everything left in `w` is nearer), so the state is exactly:
// the ef best nodes seen so far, each with an "expanded" flag; expand the nearest unexpanded one until none is
// left.  A candidate's neighbours can be evaluated as a batch because the final content of `w` after the batch
// is the top-ef of (w ∪ batch) whatever the order (core.rs:657-664) — as long as no two sims tie.
//
// State per warp: a sorted candidate list distributed over registers (entry e lives in lane e % 32, register
// e / 32), an exact visited hash set in shared memory (or global memory for the large-ef / retry pass), and the
// query chunks in registers (or shared memory for the generic / scalar distance modes).
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "distance.cuh"

namespace hnsw {

struct Counters {
  uint32_t n_dist, n_adj, n_hops;
};

// ---------------------------------------------------------------- sorted candidate list in registers

template <int EFR>
struct CandList {
  float sim[EFR];
  uint32_t id[EFR];  // kEmpty = unused; bit 31 = expanded
  int len;           // warp-uniform
  float worst;       // warp-uniform; sim of entry ef-1 once len == ef, else -inf

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < EFR; ++r) sim[r] = -CUDART_INF_F, id[r] = kEmpty;
    len = 0;
    worst = -CUDART_INF_F;
    wpos = 0;
  }

  __device__ __forceinline__ bool admits(float s, int ef) const { return len < ef || s > worst; }  // core.rs:657

  // Insert (s, nid) keeping descending order (new entry goes after equal sims).  Caller checked admits().
  __device__ __forceinline__ void insert(float s, uint32_t nid, int ef, int lane) {
    int p = 0;
#pragma unroll
    for (int r = 0; r < EFR; ++r) p += __popc(__ballot_sync(kFull, id[r] != kEmpty && sim[r] >= s));
#pragma unroll
    for (int r = EFR - 1; r >= 0; --r) {
      if (r * 32 + 31 < p) continue;  // warp-uniform: nothing at or after p in this register
      float us = __shfl_up_sync(kFull, sim[r], 1);
      uint32_t ui = __shfl_up_sync(kFull, id[r], 1);
      if (r > 0) {
        float cs = __shfl_sync(kFull, sim[r - 1], 31);
        uint32_t ci = __shfl_sync(kFull, id[r - 1], 31);
        if (lane == 0) us = cs, ui = ci;
      }
      int e = r * 32 + lane;
      if (e > p) sim[r] = us, id[r] = ui;
      else if (e == p) sim[r] = s, id[r] = nid;
    }
    if (len < ef) ++len;
    // drop what fell past ef (only possible when ef is not the full register capacity)
    if (ef < EFR * 32) {
#pragma unroll
      for (int r = 0; r < EFR; ++r)
        if (r * 32 + lane >= ef) sim[r] = -CUDART_INF_F, id[r] = kEmpty;
    }
    if (len == ef) {
      float v = 0.f;
      int wr = (ef - 1) >> 5;
#pragma unroll
      for (int r = 0; r < EFR; ++r)
        if (r == wr) v = sim[r];
      worst = __shfl_sync(kFull, v, (ef - 1) & 31);
    }
  }

  // index of the nearest unexpanded entry, or -1
  __device__ __forceinline__ int first_unexpanded() const {
    int pos = -1;
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      uint32_t b = __ballot_sync(kFull, id[r] != kEmpty && !(id[r] & kExpanded));
      if (pos < 0 && b) pos = r * 32 + __ffs(b) - 1;
    }
    return pos;
  }

  // ---- WIDE mode (EFR >= 4, staged search kernels) ------------------------------------------------------------------
  // Keeping the list sorted costs ~12 instructions per list register per admitted candidate (profiles/r1e_curve.md: the
  // ef = 200 / 400 classes ran at 0.68 / 0.43 of the roofline because of it).  Nothing in search_level needs the order
  // while the search runs — only the ef-th best value (core.rs:651,657), the nearest unexpanded entry (core.rs:631) and,
  // at the end, the ranking.  So during a search the entries sit in arrival order; an admitted candidate overwrites the
  // worst entry (one warp arg-min over EFR registers per lane), the next candidate is a warp arg-max over the unexpanded
  // entries, and ONE bitonic sort at the end of the level restores the sorted layout every consumer of the list expects.
  // MEASURED (r2 call H, 1M x 128): slower, not faster — ef 96 / 200 / 400 ran at 3.5 / 2.0 / 1.0 TB/s algorithmic against
  // 6.0 / 4.5 / 2.8 TB/s for the sorted list.  The sorted insert is cheaper than it looks (registers before the insert
  // position are skipped, its ballots and shifts are independent), while the arg-min / arg-max butterflies are dependent
  // shuffle chains, and under the register caps of Search2Bounds the sort's swaps spill (2.3 KB at EFR = 16).  The mode is
  // kept compiled out as the record of the experiment (profiles/r2_experiments.md).
  static constexpr bool kWide = false;
  int wpos;          // wide mode: slot of the worst entry once len == ef (warp-uniform)

  __device__ __forceinline__ void insert_wide(float s, uint32_t nid, int ef, int lane) {
    const int slot = len < ef ? len : wpos;
#pragma unroll
    for (int r = 0; r < EFR; ++r)
      if (r == (slot >> 5) && lane == (slot & 31)) sim[r] = s, id[r] = nid;
    if (len < ef) ++len;
    if (len == ef) {                                           // the ef-th best value and where it sits
      float m = CUDART_INF_F;
      int mp = 0;
#pragma unroll
      for (int r = 0; r < EFR; ++r) {
        const int e = r * 32 + lane;
        if (e < ef && sim[r] < m) m = sim[r], mp = e;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float om = __shfl_xor_sync(kFull, m, off);
        const int op = __shfl_xor_sync(kFull, mp, off);
        if (om < m || (om == m && op > mp)) m = om, mp = op;
      }
      worst = m, wpos = mp;
    }
  }

  // nearest unexpanded entry -> (nid, s), marked expanded; false if there is none (core.rs:631-638)
  __device__ __forceinline__ bool pop_wide(uint32_t& nid, float& s, int lane) {
    float b = -CUDART_INF_F;
    int bp = -1;
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      const bool open = id[r] != kEmpty && !(id[r] & kExpanded);
      if (open && (bp < 0 || sim[r] > b)) b = sim[r], bp = r * 32 + lane;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(kFull, b, off);
      const int op = __shfl_xor_sync(kFull, bp, off);
      if (op >= 0 && (bp < 0 || ob > b || (ob == b && op < bp))) b = ob, bp = op;
    }
    if (bp < 0) return false;
    uint32_t vi = kEmpty;
#pragma unroll
    for (int r = 0; r < EFR; ++r)
      if (r == (bp >> 5)) {
        vi = id[r];
        if (lane == (bp & 31)) id[r] |= kExpanded;
      }
    nid = __shfl_sync(kFull, vi, bp & 31) & ~kExpanded;
    s = b;
    return true;
  }

  // bitonic sort of the EFR * 32 slots into the sorted layout (entry of rank e in lane e % 32, register e / 32, nearest
  // first; unused slots carry -inf and end up last)
  __device__ __forceinline__ void sort_wide(int lane) {
    constexpr int N = EFR * 32;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j >= 32; j >>= 1) {                 // partner in another register of the same lane (static indices)
#pragma unroll
        for (int r = 0; r < EFR; ++r) {
          const int pr = r ^ (j >> 5);
          if (pr > r) {
            const bool desc = ((r * 32 + lane) & k) == 0;      // this block sorts descending
            const bool swap = desc ? (sim[r] < sim[pr]) : (sim[r] > sim[pr]);
            if (swap) {
              const float ts = sim[r];
              sim[r] = sim[pr], sim[pr] = ts;
              const uint32_t ti = id[r];
              id[r] = id[pr], id[pr] = ti;
            }
          }
        }
      }
      // partner in another lane, same register: a run-time loop keeps the code small (the compile time of the fully
      // unrolled network was 30 minutes for the library)
#pragma unroll 1
      for (int j = (k >> 1) < 16 ? (k >> 1) : 16; j > 0; j >>= 1) {
        const bool lower = (lane & j) == 0;                    // the member of the pair with the smaller index
#pragma unroll
        for (int r = 0; r < EFR; ++r) {
          const float os = __shfl_xor_sync(kFull, sim[r], j);
          const uint32_t oi = __shfl_xor_sync(kFull, id[r], j);
          const bool desc = ((r * 32 + lane) & k) == 0;
          // descending block: the smaller index keeps the larger sim
          const bool take = (desc == lower) ? (os > sim[r]) : (os < sim[r]);
          if (take) sim[r] = os, id[r] = oi;
        }
      }
    }
  }

  // ---- mode-independent interface used by the staged kernels
  __device__ __forceinline__ void push(float s, uint32_t nid, int ef, int lane) {
    if constexpr (kWide) insert_wide(s, nid, ef, lane);
    else insert(s, nid, ef, lane);
  }
  __device__ __forceinline__ bool pop(uint32_t& nid, float& s, int lane) {
    if constexpr (kWide) {
      return pop_wide(nid, s, lane);
    } else {
      const int pos = first_unexpanded();
      if (pos < 0) return false;
      get(pos, lane, true, nid, s);
      return true;
    }
  }
  __device__ __forceinline__ void finish(int lane) {
    if constexpr (kWide) sort_wide(lane);
  }
  // id of the nearest unexpanded entry without marking it (the lookahead's prediction, search_la.cuh)
  __device__ __forceinline__ bool peek(uint32_t& nid, int lane) {
    if constexpr (kWide) {
      float b = -CUDART_INF_F;
      int bp = -1;
#pragma unroll
      for (int r = 0; r < EFR; ++r) {
        const bool open = id[r] != kEmpty && !(id[r] & kExpanded);
        if (open && (bp < 0 || sim[r] > b)) b = sim[r], bp = r * 32 + lane;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(kFull, b, off);
        const int op = __shfl_xor_sync(kFull, bp, off);
        if (op >= 0 && (bp < 0 || ob > b || (ob == b && op < bp))) b = ob, bp = op;
      }
      if (bp < 0) return false;
      uint32_t vi = kEmpty;
#pragma unroll
      for (int r = 0; r < EFR; ++r)
        if (r == (bp >> 5)) vi = id[r];
      nid = __shfl_sync(kFull, vi, bp & 31) & ~kExpanded;
      return true;
    } else {
      const int pos = first_unexpanded();
      if (pos < 0) return false;
      float s;
      get(pos, lane, false, nid, s);
      return true;
    }
  }

  __device__ __forceinline__ void bind(uint32_t*, uint32_t) {}  // registers: nothing to bind (see CandList<0>)
  // f(e, id, sim) for the entries e < n this lane owns (e % 32 == lane)
  template <class F>
  __device__ __forceinline__ void for_each_prefix(int n, int lane, F f) const {
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      const int e = r * 32 + lane;
      if (e < n) f(e, id[r] & ~kExpanded, sim[r]);
    }
  }
  __device__ __forceinline__ bool contains(uint32_t x, int) const {
    bool hit = false;
#pragma unroll
    for (int r = 0; r < EFR; ++r) hit |= (id[r] & ~kExpanded) == x;   // an empty slot masks to 0x7FFFFFFF, not a node id
    return __any_sync(kFull, hit);
  }

  // id of entry `pos` (warp-uniform), on every lane
  __device__ __forceinline__ uint32_t entry_at(int pos, int) const {
    uint32_t x = kEmpty;
#pragma unroll
    for (int r = 0; r < EFR; ++r)
      if (r == (pos >> 5)) x = id[r];
    return __shfl_sync(kFull, x, pos & 31) & ~kExpanded;
  }

  // read entry `pos` (warp-uniform) and optionally mark it expanded
  __device__ __forceinline__ void get(int pos, int lane, bool mark, uint32_t& nid, float& s) {
    uint32_t vi = kEmpty;
    float vs = 0.f;
    int rr = pos >> 5, l = pos & 31;
#pragma unroll
    for (int r = 0; r < EFR; ++r)
      if (r == rr) {
        vi = id[r], vs = sim[r];
        if (mark && lane == l) id[r] |= kExpanded;
      }
    nid = __shfl_sync(kFull, vi, l) & ~kExpanded;
    s = __shfl_sync(kFull, vs, l);
  }
};

// ---------------------------------------------------------------- candidate list in memory (ef beyond the register classes)

// CandList<0>: the same sorted list, kept in the warp's slice of a global-memory scratch buffer instead of registers.
// The reference accepts any EFCON (lib.rs:53, core.rs:322-346); the register classes end at 1024 entries, and this class
// takes over beyond that (register-staged kernels only: search_knn_kernel, search_level_kernel, the one-warp insert and
// delete).  Every operation is O(ef / 32) warp steps — a compatibility path, not a fast one.
template <>
struct CandList<0> {
  float* sim;        // [cap] descending
  uint32_t* id;      // [cap], bit 31 = expanded
  int len;           // warp-uniform
  float worst;
  int ucur;          // every entry before this position is expanded
  static constexpr bool kWide = false;

  __device__ __forceinline__ void bind(uint32_t* mem, uint32_t cap) {
    sim = reinterpret_cast<float*>(mem);
    id = mem + cap;
  }
  __device__ __forceinline__ void init() {
    len = 0, ucur = 0;
    worst = -CUDART_INF_F;
  }
  __device__ __forceinline__ bool admits(float s, int ef) const { return len < ef || s > worst; }

  __device__ __forceinline__ void insert(float s, uint32_t nid, int ef, int lane) {
    int p = 0;                                                 // entries with sim >= s (new entry goes after equal sims)
    for (int base = 0; base < len; base += 32) {
      const uint32_t b = __ballot_sync(kFull, base + lane < len && sim[base + lane] >= s);
      p += __popc(b);
      if (b != kFull) break;
    }
    const int newlen = len < ef ? len + 1 : ef;
    for (int hi = newlen - 1; hi > p; hi -= 32) {              // shift [p, newlen - 1) one slot to the right, from the end
      const int e = hi - lane;
      const bool act = e > p;
      float vs = 0.f;
      uint32_t vi = kEmpty;
      if (act) vs = sim[e - 1], vi = id[e - 1];
      __syncwarp();
      if (act) sim[e] = vs, id[e] = vi;
      __syncwarp();
    }
    if (lane == 0 && p < newlen) sim[p] = s, id[p] = nid;
    __syncwarp();
    len = newlen;
    if (len == ef) worst = sim[ef - 1];
    if (p < ucur) ucur = p;
  }
  __device__ __forceinline__ int first_unexpanded() {
    for (int base = ucur & ~31; base < len; base += 32) {
      const int e = base + lane_id();
      const uint32_t b = __ballot_sync(kFull, e < len && e >= ucur && !(id[e] & kExpanded));
      if (b) {
        ucur = base + __ffs(b) - 1;
        return ucur;
      }
    }
    ucur = len;
    return -1;
  }
  __device__ __forceinline__ void get(int pos, int lane, bool mark, uint32_t& nid, float& s) {
    nid = id[pos] & ~kExpanded;
    s = sim[pos];
    __syncwarp();
    if (mark && lane == 0) id[pos] |= kExpanded;
    __syncwarp();
  }
  // f(e, id, sim) for the entries e < n this lane owns (e % 32 == lane)
  template <class F>
  __device__ __forceinline__ void for_each_prefix(int n, int lane, F f) const {
    for (int e = lane; e < n; e += 32) f(e, id[e] & ~kExpanded, sim[e]);
  }
  __device__ __forceinline__ bool contains(uint32_t x, int lane) const {
    bool hit = false;
    for (int e = lane; e < len; e += 32) hit |= (id[e] & ~kExpanded) == x;
    return __any_sync(kFull, hit);
  }
  __device__ __forceinline__ uint32_t entry_at(int pos, int) const { return id[pos] & ~kExpanded; }
};

// ---------------------------------------------------------------- exact visited set (open addressing)

struct Visited {
  uint32_t* tab;    // slots (shared or global memory), kEmpty = free
  uint32_t mask;    // slots - 1
  uint32_t count;   // warp-uniform number of stored ids
  uint32_t limit;   // abort threshold (load factor 3/4)
  int shift;        // 32 - log2(slots)
};

__device__ __forceinline__ void visited_clear(Visited& v, int lane) {
  uint4 e = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  uint4* t = reinterpret_cast<uint4*>(v.tab);
  for (uint32_t i = lane; i < (v.mask + 1) / 4; i += 32) t[i] = e;
  v.count = 0;
  __syncwarp();
}

// Warp-collective test-and-insert: lanes with `active` offer distinct ids; returns true where the id was new.
// Lanes race for free slots with plain stores and re-read after a warp barrier; the loser moves on.
__device__ __forceinline__ bool visited_insert(Visited& v, uint32_t nid, bool active) {
  bool pending = active, is_new = false;
  uint32_t slot = (nid * 2654435761u) >> v.shift;
  while (__any_sync(kFull, pending)) {
    uint32_t cur = pending ? v.tab[slot] : 0u;
    bool try_write = pending && cur == kEmpty;
    if (pending && cur == nid) pending = false;  // seen before
    if (try_write) v.tab[slot] = nid;
    __syncwarp();
    if (try_write) {
      if (v.tab[slot] == nid) is_new = true, pending = false;
      else slot = (slot + 1) & v.mask;
    } else if (pending) {
      slot = (slot + 1) & v.mask;
    }
    __syncwarp();
  }
  v.count += __popc(__ballot_sync(kFull, is_new));
  return is_new;
}

// ---------------------------------------------------------------- distance providers

// dim = 32*C known at compile time: query chunks in registers, rows fetched with one V-wide load per group.
template <int C_>
struct DistReg {
  static constexpr int C = C_;
  static constexpr int U = (C <= 4) ? 4 : ((C <= 8) ? 2 : 1);  // rows in flight per lane
  static constexpr bool kNeedsSmemQuery = false;
  static constexpr bool kStaged = true;  // search2.cuh has a TMA-staged kernel for this dimension
  float q[C];
  __device__ __forceinline__ void load_query(const float* __restrict__ qn, float*, uint32_t, int lane) {
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = qn[32 * c + lane];
  }
  // query = the stored vector of `node` (insert path: the new node / the node being re-selected)
  __device__ __forceinline__ void load_query_slab(const Graph& g, uint32_t node, float*, int lane) {
    RowRegs<C> r;
    load_row_regs<C>(g.vecs + (size_t)node * (32 * C), lane, r);
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = r.x[c];
  }
  __device__ __forceinline__ float one(const Graph& g, uint32_t nid, int lane) const {
    RowRegs<C> r;
    load_row_regs<C>(g.vecs + (size_t)nid * (32 * C), lane, r);
    return warp_hsum_avx_order(lane_partial<C>(q, r));
  }
  // sims of the neighbours flagged in `mask` (lane j holds neighbour id `nb`); lane j gets its own
  __device__ __forceinline__ float batch(const Graph& g, uint32_t nb, uint32_t mask, int lane) const {
    float mine = -CUDART_INF_F;
    while (mask) {
      int j[U];
      RowRegs<C> rr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        j[u] = mask ? (__ffs(mask) - 1) : -1;
        mask &= mask - 1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j[u] >= 0) {
          uint32_t nid = __shfl_sync(kFull, nb, j[u]);
          load_row_regs<C>(g.vecs + (size_t)nid * (32 * C), lane, rr[u]);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j[u] >= 0) {
          float s = warp_hsum_avx_order(lane_partial<C>(q, rr[u]));
          if (lane == j[u]) mine = s;
        }
    }
    return mine;
  }
};

// dim % 32 == 0, any size: permuted query in shared memory.
struct DistGeneric {
  static constexpr bool kNeedsSmemQuery = true;
  static constexpr bool kStaged = false;
  const float* qs;
  uint32_t C;
  int V;
  __device__ __forceinline__ void load_query(const float* __restrict__ qn, float* smem_q, uint32_t dim, int lane) {
    C = dim / 32;
    V = dist_vec_width(dim);
    __syncwarp();
    for (uint32_t i = lane; i < dim; i += 32) smem_q[permuted_pos(i, V)] = qn[i];
    qs = smem_q;
    __syncwarp();
  }
  __device__ __forceinline__ void load_query_slab(const Graph& g, uint32_t node, float* smem_q, int lane) {
    C = g.dim / 32;
    V = dist_vec_width(g.dim);
    const float* row = g.vecs + (size_t)node * g.dim;  // slab rows are already lane-permuted
    __syncwarp();
    for (uint32_t i = lane; i < g.dim; i += 32) smem_q[i] = row[i];
    qs = smem_q;
    __syncwarp();
  }
  __device__ __forceinline__ float one(const Graph& g, uint32_t nid, int lane) const {
    return warp_hsum_avx_order(lane_partial_generic(qs, g.vecs + (size_t)nid * (32 * C), C, V, lane));
  }
  __device__ __forceinline__ float batch(const Graph& g, uint32_t nb, uint32_t mask, int lane) const {
    float mine = -CUDART_INF_F;
    while (mask) {
      int j0 = __ffs(mask) - 1;
      mask &= mask - 1;
      int j1 = mask ? (__ffs(mask) - 1) : -1;
      mask &= mask - 1;
      uint32_t n0 = __shfl_sync(kFull, nb, j0);
      float p0 = lane_partial_generic(qs, g.vecs + (size_t)n0 * (32 * C), C, V, lane);
      float p1 = 0.f;
      if (j1 >= 0) {
        uint32_t n1 = __shfl_sync(kFull, nb, j1);
        p1 = lane_partial_generic(qs, g.vecs + (size_t)n1 * (32 * C), C, V, lane);
      }
      float s0 = warp_hsum_avx_order(p0);
      if (lane == j0) mine = s0;
      if (j1 >= 0) {
        float s1 = warp_hsum_avx_order(p1);
        if (lane == j1) mine = s1;
      }
    }
    return mine;
  }
};

// dim % 32 != 0: reference scalar fold, one lane per row, natural-order query in shared memory.
struct DistScalar {
  static constexpr bool kNeedsSmemQuery = true;
  static constexpr bool kStaged = false;
  const float* qs;
  uint32_t dim;
  __device__ __forceinline__ void load_query(const float* __restrict__ qn, float* smem_q, uint32_t d, int lane) {
    dim = d;
    __syncwarp();
    for (uint32_t i = lane; i < d; i += 32) smem_q[i] = qn[i];
    qs = smem_q;
    __syncwarp();
  }
  __device__ __forceinline__ void load_query_slab(const Graph& g, uint32_t node, float* smem_q, int lane) {
    dim = g.dim;
    const float* row = g.vecs + (size_t)node * g.dim;
    __syncwarp();
    for (uint32_t i = lane; i < g.dim; i += 32) smem_q[i] = row[i];
    qs = smem_q;
    __syncwarp();
  }
  __device__ __forceinline__ float one(const Graph& g, uint32_t nid, int) const {
    return scalar_sim(qs, g.vecs + (size_t)nid * dim, dim);  // every lane computes the same value
  }
  __device__ __forceinline__ float batch(const Graph& g, uint32_t nb, uint32_t mask, int lane) const {
    if ((mask >> lane) & 1u) return scalar_sim(qs, g.vecs + (size_t)nb * dim, dim);
    return -CUDART_INF_F;
  }
};

// ---------------------------------------------------------------- search_layer (core.rs:607-675)

// Evaluate one chunk of <= 32 neighbour ids (lane j holds nb or kEmpty) against the list.
// Returns false if the visited table passed its load limit (caller aborts and retries with a larger table).
template <int EFR, class Dist>
__device__ __forceinline__ bool expand_chunk(const Graph& g, const Dist& dist, uint32_t nb, int ef, CandList<EFR>& L,
                                             Visited& vis, Counters& cnt, int lane) {
  bool valid = nb != kEmpty;
  uint32_t vmask = __ballot_sync(kFull, valid);
  if (!vmask) return true;
  cnt.n_adj += __popc(vmask);                                  // core.rs:646
  bool is_new = visited_insert(vis, nb, valid);                // core.rs:648-649
  uint32_t newmask = __ballot_sync(kFull, is_new);
  if (vis.count > vis.limit) return false;
  if (!newmask) return true;
  cnt.n_dist += __popc(newmask);
  float mine = dist.batch(g, nb, newmask, lane);               // core.rs:652-656
  uint32_t cand = __ballot_sync(kFull, is_new && L.admits(mine, ef));
  while (cand) {                                               // list order, threshold re-read (core.rs:651,657)
    int j = __ffs(cand) - 1;
    cand &= cand - 1;
    float s = __shfl_sync(kFull, mine, j);
    uint32_t nid = __shfl_sync(kFull, nb, j);
    if (L.admits(s, ef)) L.insert(s, nid, ef, lane);           // core.rs:658-664
  }
  return true;
}

// Greedy best-first search of one level from entry point `ep`.  On return L holds the result set `w`
// nearest-first.  `expanded_out` (optional, global memory) receives the ids whose adjacency rows were read.
template <int EFR, class Dist>
__device__ __forceinline__ bool search_layer(const Graph& g, const Dist& dist, uint32_t ep, int ef, uint32_t level,
                                             CandList<EFR>& L, Visited& vis, Counters& cnt, int lane,
                                             uint32_t* expanded_out = nullptr, uint32_t expanded_cap = 0,
                                             uint32_t* n_expanded = nullptr) {
  visited_clear(vis, lane);
  visited_insert(vis, ep, lane == 0);                          // core.rs:617
  float s_ep = dist.one(g, ep, lane);                          // core.rs:621
  cnt.n_dist += 1;
  L.init();
  L.insert(s_ep, ep, ef, lane);                                // core.rs:627-628
  uint32_t nexp = 0;
  for (;;) {
    int pos = L.first_unexpanded();                            // core.rs:631-638
    if (pos < 0) break;
    uint32_t cid;
    float cs;
    L.get(pos, lane, true, cid, cs);
    cnt.n_hops += 1;
    if (expanded_out) {
      if (nexp < expanded_cap && lane == 0) expanded_out[nexp] = cid;
      ++nexp;
    }
    uint32_t* ovf;
    const uint32_t* row = row_ptr(g, cid, level, &ovf);        // core.rs:642-645
    if (!row) continue;
    uint32_t link = *ovf;
    bool more = true;
    for (uint32_t w = 0; w < g.W / 32 && more; ++w) {
      uint32_t nb = row[w * 32 + lane];
      more = __shfl_sync(kFull, nb, 31) != kEmpty;             // rows are compact: an empty tail ends the list
      if (!expand_chunk<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, lane)) return false;
    }
    while (more && link != kEmpty) {                           // overflow rows (degree is unbounded)
      uint32_t nb = g.pool[(size_t)link * 32 + lane];
      link = __shfl_sync(kFull, nb, 31);
      if (lane == 31) nb = kEmpty;
      if (!expand_chunk<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, lane)) return false;
    }
  }
  if (n_expanded) *n_expanded = nexp;
  return true;
}

// ---------------------------------------------------------------- search_knn kernel (core.rs:477-486, 865-892)

struct SearchArgs {
  const float* queries;   // [nq][dim] natural order
  uint32_t nq;
  uint32_t k;
  uint32_t ef;
  uint32_t* ids;          // [nq][k]
  float* sims;            // [nq][k]
  uint32_t* counts;       // [nq]
  uint32_t* stats;        // [nq][4] or null
  uint32_t* work_counter; // dynamic query scheduler
  uint32_t* retry_list;   // queries whose visited table overflowed in the shared-memory pass
  uint32_t* retry_count;
  uint32_t vis_slots;     // per-warp visited slots (power of two)
  uint32_t* vis_global;   // [warps][vis_slots] when the table lives in global memory
  int retry_pass;         // 1: take query indices from retry_list
  uint32_t* list_mem;     // CandList<0> only: [warps of the grid][2 * list_cap] words
  uint32_t list_cap;
};

template <int EFR, class Dist, bool VIS_SMEM>
__global__ void __launch_bounds__(256) search_knn_kernel(Graph g, SearchArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;

  Visited vis;
  vis.mask = a.vis_slots - 1;
  vis.shift = 32 - (31 - __clz(a.vis_slots));
  vis.limit = a.vis_slots - a.vis_slots / 4;
  float* smem_q = nullptr;
  if (VIS_SMEM) {
    vis.tab = smem + (size_t)warp * a.vis_slots;
    if (Dist::kNeedsSmemQuery) smem_q = reinterpret_cast<float*>(smem + (size_t)warps * a.vis_slots) + (size_t)warp * g.dim;
  } else {
    vis.tab = a.vis_global + ((size_t)blockIdx.x * warps + warp) * a.vis_slots;
    if (Dist::kNeedsSmemQuery) smem_q = reinterpret_cast<float*>(smem) + (size_t)warp * g.dim;
  }

  const uint32_t total = a.retry_pass ? *a.retry_count : a.nq;
  Dist dist;
  CandList<EFR> L;
  L.bind(a.list_mem + ((size_t)blockIdx.x * warps + warp) * 2 * a.list_cap, a.list_cap);
  constexpr uint32_t kRegCap = EFR * 32;                       // 0 for the memory-backed class

  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.work_counter, 1u);
    wi = __shfl_sync(kFull, wi, 0);
    if (wi >= total) break;
    const uint32_t qi = a.retry_pass ? a.retry_list[wi] : wi;

    Counters cnt = {0, 0, 0};
    dist.load_query(a.queries + (size_t)qi * g.dim, smem_q, g.dim, lane);
    const int32_t entry = g.meta[kMetaEntry];
    bool ok = true;
    uint32_t n_out = 0;
    if (entry >= 0) {                                          // core.rs:481-483 (empty index -> no results)
      uint32_t ep = (uint32_t)entry;
      // core.rs:869-876: ef = 1 on the upper levels, taking the nearest as the next entry point; `ef` on level 0
      for (int lc = g.meta[kMetaMaxLayer]; lc >= 0 && ok; --lc) {
        ok = search_layer<EFR, Dist>(g, dist, ep, lc > 0 ? 1 : (int)a.ef, (uint32_t)lc, L, vis, cnt, lane);
        float s;
        if (ok && lc > 0) L.get(0, lane, false, ep, s);
      }
      if (ok) n_out = min((uint32_t)L.len, a.k);               // core.rs:879
    }
    if (!ok && !a.retry_pass) {                                // visited table too small: queue for the retry pass
      if (lane == 0) a.retry_list[atomicAdd(a.retry_count, 1u)] = qi;
      continue;
    }
    // results nearest-first (core.rs:878-891); unused slots are padded
    if (!ok) n_out = 0;
    L.for_each_prefix((int)n_out, lane, [&](int e, uint32_t nid, float sv) {
      a.ids[(size_t)qi * a.k + e] = nid;
      a.sims[(size_t)qi * a.k + e] = sv;
    });
    for (uint32_t e = n_out + lane; e < a.k; e += 32) {        // unused slots are padded
      a.ids[(size_t)qi * a.k + e] = kEmpty;
      a.sims[(size_t)qi * a.k + e] = -CUDART_INF_F;
    }
    (void)kRegCap;
    if (lane == 0) {
      a.counts[qi] = n_out;
      if (a.stats) {
        a.stats[(size_t)qi * 4 + 0] = cnt.n_dist;
        a.stats[(size_t)qi * 4 + 1] = cnt.n_adj;
        a.stats[(size_t)qi * 4 + 2] = cnt.n_hops;
        a.stats[(size_t)qi * 4 + 3] = (a.retry_pass ? 1u : 0u) | (ok ? 0u : 2u);
      }
      if (!ok) atomicOr((unsigned int*)(g.meta + kMetaError), (unsigned int)kErrVisitedOverflow);
    }
  }
}

// ---------------------------------------------------------------- search_level on its own (parity tests)

struct LevelArgs {
  const float* query;
  uint32_t entry, ef, level;
  uint32_t* ids;    // [ef]
  float* sims;      // [ef]
  uint32_t* n_out;  // [1]; 0xFFFFFFFF on visited overflow
  uint32_t vis_slots;
  uint32_t* vis_global;
  uint32_t* list_mem;  // CandList<0> only: [2 * list_cap] words
  uint32_t list_cap;
};

template <int EFR, class Dist>
__global__ void __launch_bounds__(32) search_level_kernel(Graph g, LevelArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  Visited vis;
  vis.tab = a.vis_global;
  vis.mask = a.vis_slots - 1;
  vis.shift = 32 - (31 - __clz(a.vis_slots));
  vis.limit = a.vis_slots - a.vis_slots / 4;
  Dist dist;
  dist.load_query(a.query, reinterpret_cast<float*>(smem), g.dim, lane);
  CandList<EFR> L;
  L.bind(a.list_mem, a.list_cap);
  Counters cnt = {0, 0, 0};
  bool ok = search_layer<EFR, Dist>(g, dist, a.entry, (int)a.ef, a.level, L, vis, cnt, lane);
  if (!ok) {
    if (lane == 0) *a.n_out = 0xFFFFFFFFu;
    return;
  }
  L.for_each_prefix(L.len, lane, [&](int e, uint32_t nid, float sv) {
    a.ids[e] = nid;
    a.sims[e] = sv;
  });
  if (lane == 0) *a.n_out = (uint32_t)L.len;
}

}  // namespace hnsw

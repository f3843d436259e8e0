// Insert path on the device (reference: src/hnsw/core.rs:383-412 add_node, :489-599 insert, :677-757
// select_neighbors, :759-774 connect_neighbors, :776-822 update_node_connections).
//
// What the reference does per insert, and what survives here
//   * per level lc <= l:  W = search_level(q, ep, ef_construction, lc)                      (core.rs:524)
//   * R = select_neighbors(q, W, m): top-m by sim over W ∪ N(W) \ {q}                       (core.rs:531)
//     When the search stops every member of W has been expanded, so every node of N(W) was evaluated by the
//     search itself and either sits in W or lost against W's worst.  Hence top-m(W ∪ N(W)) == the first
//     min(m, |W|) entries of W: the reference's 2-hop sweep at this call site is redundant work and is skipped.
//     The one exception is ef_construction < m with W full (|W| = ef_construction < m): nodes turned away only because
//     W was full can still be selected, so there the sweep IS computed (EXACT kernels; FAST builds run EXACT then).
//   * connect: q.list = R nearest-first; q appended to the tail of every r in R              (core.rs:532, 759-774)
//   * shrink: for e in R nearest-first, if |N(e)| > cap (m_max_0 on level 0, m_max above):   (core.rs:540-574)
//       E' = top-cap by sim(e, .) over N(e) ∪ N(N(e)) \ {e}   -- a real 2-hop distance sweep (core.rs:568)
//       e.list = (old order, minus N(e) \ E') ++ (E' \ N(e) nearest-first); mirrored appends (no cap check on the
//       other side, core.rs:793-795) and mirrored removals (core.rs:808-816).
//
// Two builders share these device functions:
//   EXACT  one warp runs the NODE.ADD stream strictly in order (sequentially consistent, graph identical to the
//          reference's for tie-free data, including adjacency-list order).
//   FAST   a batch of new nodes is searched against the frozen graph by one warp each (K1), linked under row locks
//          (K2), over-full rows are re-selected against the linked snapshot (K3) and the edge deltas applied under
//          row locks (K4).  Same per-insert algorithm; nodes of one batch do not see each other.
#pragma once
#include "search.cuh"

namespace hnsw {

// control words of a build launch sequence (device memory, zeroed per batch)
enum BuildCtl : int {
  kCtlWorkSearch = 0,
  kCtlWorkRetry = 1,
  kCtlRetryCount = 2,
  kCtlWorkLink = 3,
  kCtlWlCount = 4,      // re-prune worklist length
  kCtlWorkReprune = 5,
  kCtlWorkApply = 6,
  kCtlWlDropped = 7,    // rows that were over cap but did not fit the worklist (left over-full; benign)
  kCtlDistEvals = 8,
  kCtlReprunes = 9,
  kCtlSkipped = 10,     // re-prunes skipped (list too long / visited overflow)
  kCtlProgress = 11,    // EXACT: inserts completed
  kCtlTouched = 12,     // EXACT: entries written to the touched buffer
  kCtlRefused = 13,     // FAST: edges dropped on both sides because a hub row already held `lcap` ids
  kCtlWords = 32,
};

struct BuildShared {
  // common to all build kernels
  uint32_t* ctl;
  uint32_t m, cap0, capU, efc;
  uint32_t lcap;         // ids a list buffer holds (shared memory words per list)
  uint32_t epoch;        // batch stamp for worklist de-duplication
  uint32_t* stamp0;      // [n]   last epoch the level-0 row was queued
  uint32_t* stampU;      // [nU]
};

// ---------------------------------------------------------------- rows: locks, whole-list load / store

__device__ __forceinline__ uint32_t row_key(const Graph& g, uint32_t node, uint32_t level) {
  return level == 0 ? node : (0x80000000u | (g.upper_base[node] + level - 1));
}

__device__ __forceinline__ uint32_t* lock_of(const Graph& g, uint32_t key) {
  return g.locks + ((key * 2654435761u) >> g.lock_shift);
}

// One lane spins; the lock table is hashed (two rows may share a lock: false contention only, locks never nest).
__device__ __forceinline__ void row_lock(uint32_t* l, int lane) {
  if (lane == 0) {
    while (atomicCAS(l, 0u, 1u) != 0u) __nanosleep(100);
    __threadfence();
  }
  __syncwarp();
}

__device__ __forceinline__ void row_unlock(uint32_t* l, int lane) {
  __threadfence();  // every lane publishes its row writes before the release
  __syncwarp();
  if (lane == 0) atomicExch(l, 0u);
  __syncwarp();
}

// Whole adjacency list (fixed row + overflow chain) -> buf (shared memory, `lcap` words).  Reads bypass L1
// (rows are edited by other SMs under locks).  Returns the length, or kEmpty if it does not fit.
__device__ __forceinline__ uint32_t list_load(const Graph& g, const uint32_t* row, const uint32_t* ovf, uint32_t* buf,
                                              uint32_t lcap, int lane) {
  uint32_t len = 0;
  bool more = true;
  for (uint32_t w = 0; w < g.W / 32 && more; ++w) {
    uint32_t nb = __ldcg(row + w * 32 + lane);
    uint32_t cnt = __popc(__ballot_sync(kFull, nb != kEmpty));
    if (len + cnt > lcap) return kEmpty;
    if (nb != kEmpty) buf[len + lane] = nb;  // rows are compact: valid ids form a prefix
    len += cnt;
    more = cnt == 32;
  }
  uint32_t link = more ? __ldcg(ovf) : kEmpty;
  while (link != kEmpty) {
    uint32_t nb = __ldcg(g.pool + (size_t)link * 32 + lane);
    uint32_t next = __shfl_sync(kFull, nb, 31);
    bool valid = lane < kPoolIds && nb != kEmpty;
    uint32_t cnt = __popc(__ballot_sync(kFull, valid));
    if (len + cnt > lcap) return kEmpty;
    if (valid) buf[len + lane] = nb;
    len += cnt;
    link = (cnt == (uint32_t)kPoolIds) ? next : kEmpty;
  }
  __syncwarp();
  return len;
}

// buf[0..len) -> fixed row + overflow chain (extended from the pool when needed; chain rows are kept when the
// list shrinks).  Returns false when the pool is exhausted.
__device__ __forceinline__ bool list_store(const Graph& g, uint32_t* row, uint32_t* ovf, const uint32_t* buf,
                                           uint32_t len, int lane) {
  __syncwarp();
  for (uint32_t w = 0; w < g.W / 32; ++w) {
    uint32_t i = w * 32 + lane;
    row[i] = i < len ? buf[i] : kEmpty;
  }
  uint32_t done = g.W;
  uint32_t link = __ldcg(ovf);
  uint32_t* prev = ovf;
  while (done < len || link != kEmpty) {
    uint32_t next;
    uint32_t i = done + lane;
    if (link == kEmpty) {
      uint32_t nr = 0;
      if (lane == 0) nr = atomicAdd(reinterpret_cast<unsigned int*>(g.meta + kMetaPoolUsed), 1u);
      nr = __shfl_sync(kFull, nr, 0);
      if (nr >= g.pool_cap) {
        if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrPoolExhausted);
        return false;
      }
      link = nr;
      next = kEmpty;
      g.pool[(size_t)link * 32 + lane] = (lane < kPoolIds && i < len) ? buf[i] : kEmpty;  // word 31: no next row
      __syncwarp();
      if (lane == 0) *prev = link;
    } else {
      next = __ldcg(g.pool + (size_t)link * 32 + 31);
      if (lane < kPoolIds) g.pool[(size_t)link * 32 + lane] = i < len ? buf[i] : kEmpty;
    }
    done += kPoolIds;
    prev = g.pool + (size_t)link * 32 + 31;
    link = next;
  }
  __syncwarp();
  return true;
}

// position of `x` in buf[0..len) or -1 (warp-uniform)
__device__ __forceinline__ int list_find(const uint32_t* buf, uint32_t len, uint32_t x, int lane) {
  for (uint32_t i = 0; i < len; i += 32) {
    uint32_t b = __ballot_sync(kFull, i + lane < len && buf[i + lane] == x);
    if (b) return (int)i + __ffs(b) - 1;
  }
  return -1;
}

// order-preserving removal of position p (Vec::remove, core.rs:151)
__device__ __forceinline__ void list_erase(uint32_t* buf, uint32_t len, int p, int lane) {
  for (uint32_t i = (uint32_t)p; i + 1 < len; i += 32) {
    uint32_t j = i + lane;
    uint32_t v = (j + 1 < len) ? buf[j + 1] : kEmpty;
    __syncwarp();
    if (j + 1 < len) buf[j] = v;
    __syncwarp();
  }
}

// add_neighbor (core.rs:137-143) on the row of (node, level): append `x` unless present.
// Returns the new length (kEmpty on failure).  Caller holds the row lock (FAST) or is the only writer (EXACT).
// `full` (optional) is set when the list already holds `lcap` ids: the batched builder then drops the edge on both
// sides (a hub row is bounded by the edit buffer); without it a full list raises the sticky kErrListTooLong flag.
__device__ __forceinline__ uint32_t row_append_unique(const Graph& g, uint32_t node, uint32_t level, uint32_t x,
                                                      uint32_t* buf, uint32_t lcap, int lane, bool* full = nullptr) {
  uint32_t* ovf;
  uint32_t* row = row_ptr(g, node, level, &ovf);
  if (full) *full = false;
  if (!row) return kEmpty;
  uint32_t len = list_load(g, row, ovf, buf, lcap, lane);
  if (len != kEmpty && list_find(buf, len, x, lane) >= 0) return len;
  if (len == kEmpty || len + 1 > lcap) {
    if (full) *full = true;
    else if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
    return kEmpty;
  }
  if (lane == 0) buf[len] = x;
  __syncwarp();
  if (!list_store(g, row, ovf, buf, len + 1, lane)) return kEmpty;
  return len + 1;
}

// rm_neighbor (core.rs:145-152): order-preserving removal of `x` if present.
__device__ __forceinline__ void row_remove(const Graph& g, uint32_t node, uint32_t level, uint32_t x, uint32_t* buf,
                                           uint32_t lcap, int lane) {
  uint32_t* ovf;
  uint32_t* row = row_ptr(g, node, level, &ovf);
  if (!row) return;
  uint32_t len = list_load(g, row, ovf, buf, lcap, lane);
  if (len == kEmpty) return;
  int p = list_find(buf, len, x, lane);
  if (p < 0) return;
  list_erase(buf, len, p, lane);
  list_store(g, row, ovf, buf, len - 1, lane);
}

// ---------------------------------------------------------------- re-selection of an over-full row

// is `x` (warp-uniform) one of the entries of L?
template <int EFR>
__device__ __forceinline__ bool cand_contains(const CandList<EFR>& L, uint32_t x) {
  return L.contains(x, lane_id());
}

// expand_chunk with a LOSSY visited table (direct-mapped, most recent id per slot): a hit proves "already
// evaluated", a conflict forgets the older id.  Used by the batched builder's re-selection, whose 2-hop sweeps
// (|N(e)| * degree ids) would need a very large exact table for big m: a forgotten id is evaluated again and the
// membership test below keeps it from entering the list twice, so the selected set is the same top-cap set.
// `skip_a` / `skip_b` (kEmpty = none) are never candidates (the node itself and the ignored node, core.rs:704-708).
template <int EFR, class Dist>
__device__ __forceinline__ void expand_chunk_lossy(const Graph& g, const Dist& dist, uint32_t nb, int ef, CandList<EFR>& L,
                                                   Visited& vis, Counters& cnt, uint32_t skip_a, uint32_t skip_b, int lane) {
  bool valid = nb != kEmpty && nb != skip_a && nb != skip_b;
  if (!__any_sync(kFull, valid)) return;
  bool is_new = false;
  if (valid) {
    uint32_t slot = (nb * 2654435761u) >> vis.shift;
    is_new = vis.tab[slot] != nb;
    if (is_new) vis.tab[slot] = nb;
  }
  __syncwarp();
  uint32_t newmask = __ballot_sync(kFull, is_new);
  if (!newmask) return;
  cnt.n_dist += __popc(newmask);
  float mine = dist.batch(g, nb, newmask, lane);
  uint32_t cand = __ballot_sync(kFull, is_new && L.admits(mine, ef));
  while (cand) {
    int j = __ffs(cand) - 1;
    cand &= cand - 1;
    float s = __shfl_sync(kFull, mine, j);
    uint32_t nid = __shfl_sync(kFull, nb, j);
    if (L.admits(s, ef) && !cand_contains<EFR>(L, nid)) L.insert(s, nid, ef, lane);
  }
}

// walk the adjacency list of (node, level) chunk by chunk through expand_chunk
template <int EFR, class Dist, bool LOSSY = false>
__device__ __forceinline__ bool expand_row(const Graph& g, const Dist& dist, uint32_t node, uint32_t level, int ef,
                                           CandList<EFR>& L, Visited& vis, Counters& cnt, int lane,
                                           uint32_t skip_a = kEmpty, uint32_t skip_b = kEmpty) {
  uint32_t* ovf;
  const uint32_t* row = row_ptr(g, node, level, &ovf);
  if (!row) return true;
  uint32_t link = *ovf;
  bool more = true;
  for (uint32_t w = 0; w < g.W / 32 && more; ++w) {
    uint32_t nb = row[w * 32 + lane];
    more = __shfl_sync(kFull, nb, 31) != kEmpty;
    if constexpr (LOSSY) expand_chunk_lossy<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, skip_a, skip_b, lane);
    else if (!expand_chunk<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, lane)) return false;
  }
  while (more && link != kEmpty) {
    uint32_t nb = g.pool[(size_t)link * 32 + lane];
    link = __shfl_sync(kFull, nb, 31);
    if (lane == 31) nb = kEmpty;
    if constexpr (LOSSY) expand_chunk_lossy<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, skip_a, skip_b, lane);
    else if (!expand_chunk<EFR, Dist>(g, dist, nb, ef, L, vis, cnt, lane)) return false;
  }
  return true;
}

// select_neighbors(e, N(e), cap, lc) as insert calls it (core.rs:544-568): on return L holds the top-`cap` of
// N(e) ∪ N(N(e)) \ {e} by sim(e, .), nearest-first.  `old` = N_lc(e) (shared memory), `dist` holds e's vector.
// `ignored` (kEmpty = none) is the node being deleted when delete_node_from_neighbors makes the call
// (core.rs:853): never a candidate (core.rs:704-708, 728-731), but its row IS swept when it is a member of `old`.
// LOSSY = false: exact visited set, returns false if the table overflowed.  LOSSY = true: direct-mapped table
// (see expand_chunk_lossy), never fails.
template <int EFR, class Dist, bool LOSSY = false>
__device__ __forceinline__ bool reprune_select(const Graph& g, const Dist& dist, uint32_t e, uint32_t level, int cap,
                                               const uint32_t* old, uint32_t n_old, CandList<EFR>& L, Visited& vis,
                                               Counters& cnt, int lane, uint32_t ignored = kEmpty) {
  visited_clear(vis, lane);
  if constexpr (!LOSSY) {
    visited_insert(vis, e, lane == 0);  // e itself is never a candidate (core.rs:704, 728)
    if (ignored != kEmpty) visited_insert(vis, ignored, lane == 0);
  }
  L.init();
  for (uint32_t i = 0; i < n_old; i += 32) {  // econn: sims of the current neighbours (core.rs:549-557)
    uint32_t nb = (i + lane < n_old) ? old[i + lane] : kEmpty;
    if constexpr (LOSSY) expand_chunk_lossy<EFR, Dist>(g, dist, nb, cap, L, vis, cnt, e, ignored, lane);
    else if (!expand_chunk<EFR, Dist>(g, dist, nb, cap, L, vis, cnt, lane)) return false;
  }
  for (uint32_t j = 0; j < n_old; ++j)        // extend_candidates (core.rs:698-721)
    if (!expand_row<EFR, Dist, LOSSY>(g, dist, old[j], level, cap, L, vis, cnt, lane, e, ignored)) return false;
  return true;
}

// Split the outcome of a re-selection against the old list:
//   keep_add[0..n_keep)            old entries that stay, old order
//   keep_add[n_keep..n_keep+n_add) entries of L that are new, nearest-first          (core.rs:790-796)
//   rem[0..n_rem)                  old entries that go                                 (core.rs:799-816)
// All three buffers are shared memory; keep_add needs n_old + cap words.
template <int EFR>
__device__ __forceinline__ void reprune_delta(const CandList<EFR>& L, const uint32_t* old, uint32_t n_old,
                                              uint32_t* keep_add, uint32_t* rem, uint32_t& n_keep, uint32_t& n_add,
                                              uint32_t& n_rem, int lane) {
  n_keep = n_rem = n_add = 0;
  for (uint32_t j = 0; j < n_old; ++j) {
    uint32_t x = old[j];
    bool kept = cand_contains<EFR>(L, x);
    if (lane == 0) {
      if (kept) keep_add[n_keep] = x;
      else rem[n_rem] = x;
    }
    if (kept) ++n_keep;
    else ++n_rem;
  }
  __syncwarp();
  for (int p = 0; p < L.len; ++p) {
    const uint32_t x = L.entry_at(p, lane);
    if (list_find(old, n_old, x, lane) < 0) {
      if (lane == 0) keep_add[n_keep + n_add] = x;
      ++n_add;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------- EXACT: the NODE.ADD stream on one warp

struct ExactArgs {
  uint32_t first, count;   // ids first .. first+count-1 are inserted in order
  uint32_t m, cap0, capU, efc;
  uint32_t lcap;
  uint32_t vis_slots;      // global-memory visited table (one warp)
  uint32_t* vis;
  uint32_t* ctl;
  uint32_t* touched;       // optional: ids reported through update_fn (core.rs:580-584), duplicates allowed
  uint32_t touched_cap;
  uint32_t* list_mem;      // CandList<0> (ef_construction or 2m beyond the register classes): [2 * list_cap] words
  uint32_t list_cap;
};

template <int EFR, class Dist>
__global__ void __launch_bounds__(32) insert_exact_kernel(Graph g, ExactArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  // shared memory: sel[m] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap] | query[dim] (if the metric needs it)
  uint32_t* sel = smem;
  uint32_t* old = sel + ((a.m + 31) & ~31u);
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;
  float* smem_q = reinterpret_cast<float*>(edit + a.lcap);

  Visited vis;
  vis.tab = a.vis;
  vis.mask = a.vis_slots - 1;
  vis.shift = 32 - (31 - __clz(a.vis_slots));
  vis.limit = a.vis_slots - a.vis_slots / 4;

  Dist dist;
  CandList<EFR> L;
  L.bind(a.list_mem, a.list_cap);
  Counters cnt = {0, 0, 0};
  uint32_t n_touched = 0, n_reprunes = 0;
  auto touch = [&](uint32_t id) {
    if (a.touched) {
      if (lane == 0 && n_touched < a.touched_cap) a.touched[n_touched] = id;
      ++n_touched;
    }
  };

  for (uint32_t it = 0; it < a.count; ++it) {
    const uint32_t q = a.first + it;
    const int l = g.level[q];
    const int l_max = g.meta[kMetaMaxLayer];                    // core.rs:496
    uint32_t ep = (uint32_t)g.meta[kMetaEntry];                 // core.rs:508
    bool ok = true;
    for (int lc = l_max; lc >= 0 && ok; --lc) {
      const bool link = lc <= l;
      const uint32_t cap = lc == 0 ? a.cap0 : a.capU;           // core.rs:560
      dist.load_query_slab(g, q, smem_q, lane);
      ok = search_layer<EFR, Dist>(g, dist, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, vis, cnt, lane);  // :513, :524
      if (!ok) break;
      float s;
      L.get(0, lane, false, ep, s);                             // :514 / :576 nearest of w
      if (!link) continue;
      // select_neighbors(q, w, m) == first min(m, |w|) entries of w (see the header of this file)    core.rs:531
      // ... unless ef_construction < m and w came back full: then N(w) can hold nodes the search turned away only
      // because w was full, and the reference's sweep (core.rs:698-721) picks them up until m are selected.
      if (a.efc < a.m && (uint32_t)L.len == a.efc) {
        const uint32_t n_w = (uint32_t)L.len;
        L.for_each_prefix((int)n_w, lane, [&](int e, uint32_t nid, float) { old[e] = nid; });
        __syncwarp();
        ok = reprune_select<EFR, Dist>(g, dist, q, (uint32_t)lc, (int)a.m, old, n_w, L, vis, cnt, lane);  // dist holds q
        if (!ok) break;
      }
      const uint32_t n_sel = min((uint32_t)L.len, a.m);
      L.for_each_prefix((int)n_sel, lane, [&](int e, uint32_t nid, float) { sel[e] = nid; });
      __syncwarp();
      // connect_neighbors (core.rs:759-774): q's list is R nearest-first; q goes to the tail of each r
      {
        uint32_t* ovf;
        uint32_t* row = row_ptr(g, q, (uint32_t)lc, &ovf);
        list_store(g, row, ovf, sel, n_sel, lane);
      }
      for (uint32_t i = 0; i < n_sel; ++i) {
        row_append_unique(g, sel[i], (uint32_t)lc, q, edit, a.lcap, lane);
        touch(sel[i]);                                          // core.rs:535-537
      }
      // shrink connections (core.rs:540-574), nearest-first
      for (uint32_t i = 0; i < n_sel && ok; ++i) {
        const uint32_t e = sel[i];
        uint32_t* eovf;
        uint32_t* erow = row_ptr(g, e, (uint32_t)lc, &eovf);
        uint32_t n_old = list_load(g, erow, eovf, old, a.lcap, lane);
        if (n_old == kEmpty) {
          if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
          continue;
        }
        if (n_old <= cap) continue;                             // core.rs:561
        dist.load_query_slab(g, e, smem_q, lane);
        ok = reprune_select<EFR, Dist>(g, dist, e, (uint32_t)lc, (int)cap, old, n_old, L, vis, cnt, lane);  // :568
        if (!ok) break;
        ++n_reprunes;
        uint32_t n_keep, n_add, n_rem;
        reprune_delta<EFR>(L, old, n_old, keep_add, rem, n_keep, n_add, n_rem, lane);
        // update_node_connections (core.rs:776-822)
        list_store(g, erow, eovf, keep_add, n_keep + n_add, lane);
        touch(e);
        for (uint32_t t = 0; t < n_add; ++t) {                  // :793-796 (no cap check on the other side)
          row_append_unique(g, keep_add[n_keep + t], (uint32_t)lc, e, edit, a.lcap, lane);
          touch(keep_add[n_keep + t]);
        }
        for (uint32_t t = 0; t < n_keep; ++t) touch(keep_add[t]);  // :796 every member of the new set is reported
        for (uint32_t t = 0; t < n_rem; ++t) {                  // :805-816
          row_remove(g, rem[t], (uint32_t)lc, e, edit, a.lcap, lane);
          touch(rem[t]);
        }
      }
    }
    if (!ok) {                                                  // visited table too small: host enlarges it and resumes
      if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrVisitedOverflow);
      break;
    }
    if (l > l_max && lane == 0) {                               // core.rs:587-593
      g.meta[kMetaMaxLayer] = l;
      g.meta[kMetaEntry] = (int32_t)q;
    }
    __syncwarp();
    if (lane == 0) a.ctl[kCtlProgress] = it + 1;
  }
  if (lane == 0) {
    a.ctl[kCtlDistEvals] += cnt.n_dist;
    a.ctl[kCtlReprunes] += n_reprunes;
    a.ctl[kCtlTouched] = n_touched;
  }
}

// ---------------------------------------------------------------- delete_node (core.rs:414-475, 824-863)

// One warp removes node a.first from the graph exactly as the reference does: for every level of the victim and every
// neighbour n of the victim IN LIST ORDER (core.rs:829), n's whole list is re-selected with the victim ignored —
// top-cap of N(n) ∪ N(N(n)) \ {n, victim}, whether or not n was over its cap (core.rs:846-853) — and applied through
// update_node_connections (core.rs:856), whose removal of the victim from n's list is not mirrored (core.rs:810-813):
// the victim's own lists stay intact while they are being walked and are cleared at the end.
template <int EFR, class Dist>
__global__ void __launch_bounds__(32) delete_exact_kernel(Graph g, ExactArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  // shared memory: vlist[lcap] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap] | query[dim] (if needed)
  uint32_t* vlist = smem;
  uint32_t* old = vlist + a.lcap;
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;
  float* smem_q = reinterpret_cast<float*>(edit + a.lcap);

  Visited vis;
  vis.tab = a.vis;
  vis.mask = a.vis_slots - 1;
  vis.shift = 32 - (31 - __clz(a.vis_slots));
  vis.limit = a.vis_slots - a.vis_slots / 4;

  Dist dist;
  CandList<EFR> L;
  L.bind(a.list_mem, a.list_cap);
  Counters cnt = {0, 0, 0};
  uint32_t n_touched = 0, n_reprunes = 0;
  auto touch = [&](uint32_t id) {
    if (a.touched) {
      if (lane == 0 && n_touched < a.touched_cap) a.touched[n_touched] = id;
      ++n_touched;
    }
  };
  auto flag = [&](int err) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)err);
  };

  const uint32_t victim = a.first;
  const int top = g.level[victim];
  bool ok = true;
  for (int lc = 0; lc <= top && ok; ++lc) {                       // core.rs:434-440
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;               // core.rs:846
    uint32_t* vovf;
    uint32_t* vrow = row_ptr(g, victim, (uint32_t)lc, &vovf);
    if (!vrow) continue;
    const uint32_t n_v = list_load(g, vrow, vovf, vlist, a.lcap, lane);
    if (n_v == kEmpty) {
      flag(kErrListTooLong);
      ok = false;
      break;
    }
    for (uint32_t i = 0; i < n_v && ok; ++i) {                    // core.rs:829 list order
      const uint32_t n = vlist[i];
      uint32_t* novf;
      uint32_t* nrow = row_ptr(g, n, (uint32_t)lc, &novf);
      if (!nrow) continue;
      const uint32_t n_old = list_load(g, nrow, novf, old, a.lcap, lane);   // nconn (core.rs:834-844)
      if (n_old == kEmpty) {
        flag(kErrListTooLong);
        ok = false;
        break;
      }
      dist.load_query_slab(g, n, smem_q, lane);
      ok = reprune_select<EFR, Dist>(g, dist, n, (uint32_t)lc, (int)cap, old, n_old, L, vis, cnt, lane, victim);  // :853
      if (!ok) {
        flag(kErrVisitedOverflow);
        break;
      }
      ++n_reprunes;
      uint32_t n_keep, n_add, n_rem;
      reprune_delta<EFR>(L, old, n_old, keep_add, rem, n_keep, n_add, n_rem, lane);
      // update_node_connections(n, new, nconn, lc, Some(victim))  core.rs:776-822
      list_store(g, nrow, novf, keep_add, n_keep + n_add, lane);
      touch(n);                                                   // core.rs:855, 787
      for (uint32_t t = 0; t < n_add; ++t) {                      // core.rs:793-796
        row_append_unique(g, keep_add[n_keep + t], (uint32_t)lc, n, edit, a.lcap, lane);
        touch(keep_add[n_keep + t]);
      }
      for (uint32_t t = 0; t < n_keep; ++t) touch(keep_add[t]);
      for (uint32_t t = 0; t < n_rem; ++t) {                      // core.rs:805-816
        if (rem[t] == victim) continue;                           // core.rs:810-813: not mirrored, not reported
        row_remove(g, rem[t], (uint32_t)lc, n, edit, a.lcap, lane);
        touch(rem[t]);
      }
    }
  }
  if (ok) {
    for (int lc = 0; lc <= top; ++lc) {                           // the node is dropped (core.rs:419)
      uint32_t* vovf;
      uint32_t* vrow = row_ptr(g, victim, (uint32_t)lc, &vovf);
      if (vrow) list_store(g, vrow, vovf, vlist, 0, lane);
    }
    __syncwarp();
    if (lane == 0) {
      g.level[victim] = -1;
      a.ctl[kCtlProgress] = 1;
    }
  }
  if (lane == 0) {
    a.ctl[kCtlDistEvals] += cnt.n_dist;
    a.ctl[kCtlReprunes] += n_reprunes;
    a.ctl[kCtlTouched] = n_touched;
  }
}

// ---------------------------------------------------------------- FAST: batch kernels

struct FastArgs {
  uint32_t first;          // first new node id of the batch
  uint32_t n_new;          // nodes in the batch
  uint32_t n_tasks;        // (node, level) link tasks
  const uint32_t* task_base;   // [n_new]   first task of the node (task index = task_base + level)
  const uint32_t* task_node;   // [n_tasks]
  const uint32_t* task_level;  // [n_tasks]
  uint32_t* sel_ids;       // [n_tasks][m]
  uint32_t* sel_cnt;       // [n_tasks]
  uint32_t m, cap0, capU, efc, lcap;
  uint32_t* ctl;
  // K1 visited tables / retry
  uint32_t vis_slots;
  uint32_t* vis_global;
  uint32_t* retry_list;
  int retry_pass;
  // worklist of over-full rows
  uint32_t epoch;
  uint32_t* stamp0;
  uint32_t* stampU;
  uint32_t wl_cap;
  uint32_t* wl_node;
  uint32_t* wl_level;
  uint32_t* wl_old;        // [wl_cap][lcap]  snapshot of the row at re-selection time
  uint32_t* wl_new;        // [wl_cap][W]     re-selected list nearest-first
  uint32_t* wl_len;        // [wl_cap][2]     {n_old, n_new}; n_new = kEmpty when the re-selection was skipped
};

// K1: searches of the batch against the frozen graph (core.rs:511-531 per node)
template <int EFR, class Dist, bool VIS_SMEM>
__global__ void __launch_bounds__(256) build_search_kernel(Graph g, FastArgs a) {
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
  const uint32_t total = a.retry_pass ? a.ctl[kCtlRetryCount] : a.n_new;
  uint32_t* work = a.ctl + (a.retry_pass ? kCtlWorkRetry : kCtlWorkSearch);
  Dist dist;
  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  const int l_max = g.meta[kMetaMaxLayer];
  const uint32_t entry = (uint32_t)g.meta[kMetaEntry];
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(work, 1u);
    wi = __shfl_sync(kFull, wi, 0);
    if (wi >= total) break;
    const uint32_t b = a.retry_pass ? a.retry_list[wi] : wi;
    const uint32_t q = a.first + b;
    const int l = g.level[q];
    const uint32_t tb = a.task_base[b];
    dist.load_query_slab(g, q, smem_q, lane);
    uint32_t ep = entry;
    bool ok = true;
    for (int lc = l_max; lc >= 0 && ok; --lc) {
      const bool link = lc <= l;
      ok = search_layer<EFR, Dist>(g, dist, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, vis, cnt, lane);
      if (!ok) break;
      float s;
      L.get(0, lane, false, ep, s);
      if (!link) continue;
      const uint32_t n_sel = min((uint32_t)L.len, a.m);
      uint32_t* out = a.sel_ids + (size_t)(tb + lc) * a.m;
      L.for_each_prefix((int)n_sel, lane, [&](int e, uint32_t nid, float) { out[e] = nid; });
      if (lane == 0) a.sel_cnt[tb + lc] = n_sel;
    }
    if (!ok) {
      if (!a.retry_pass) {
        if (lane == 0) a.retry_list[atomicAdd(a.ctl + kCtlRetryCount, 1u)] = b;
      } else if (lane == 0) {
        atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrVisitedOverflow);
        for (int lc = min(l, l_max); lc >= 0; --lc) a.sel_cnt[tb + lc] = 0;
      }
    }
  }
  if (lane == 0) atomicAdd(a.ctl + kCtlDistEvals, cnt.n_dist);
}

// K2: connect_neighbors for every (node, level) task; rows that end up over their cap are queued once
#ifdef HNSW_PLAIN_BUILD_KERNELS  // no distance arithmetic: defined once, in build_host.cu
__global__ void __launch_bounds__(256) build_link_kernel(Graph g, FastArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  uint32_t* edit = smem + (size_t)warp * a.lcap;
  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ctl + kCtlWorkLink, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (t >= a.n_tasks) break;
    const uint32_t q = a.task_node[t], lc = a.task_level[t];
    const uint32_t n_sel = a.sel_cnt[t];
    const uint32_t* ids = a.sel_ids + (size_t)t * a.m;
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;
    {  // q's own row: nobody else can reach q during this batch
      uint32_t* ovf;
      uint32_t* row = row_ptr(g, q, lc, &ovf);
      for (uint32_t i = lane; i < n_sel; i += 32) edit[i] = ids[i];
      __syncwarp();
      list_store(g, row, ovf, edit, n_sel, lane);
    }
    bool refused = false;  // some ids[i] is a hub whose list is full: that edge is dropped on both sides
    for (uint32_t i = 0; i < n_sel; ++i) {
      const uint32_t r = ids[i];
      const uint32_t key = row_key(g, r, lc);
      uint32_t* lk = lock_of(g, key);
      row_lock(lk, lane);
      bool full;
      uint32_t len = row_append_unique(g, r, lc, q, edit, a.lcap, lane, &full);
      if (full) {
        refused = true;
        if (lane == 0) a.sel_ids[(size_t)t * a.m + i] = kEmpty, atomicAdd(a.ctl + kCtlRefused, 1u);
      }
      if (len != kEmpty && len > cap && lane == 0) {
        uint32_t* st = (key & 0x80000000u) ? a.stampU + (key & 0x7FFFFFFFu) : a.stamp0 + key;
        if (*reinterpret_cast<volatile uint32_t*>(st) != a.epoch) {   // protected by the row lock
          *reinterpret_cast<volatile uint32_t*>(st) = a.epoch;
          uint32_t w = atomicAdd(a.ctl + kCtlWlCount, 1u);
          if (w < a.wl_cap) a.wl_node[w] = r, a.wl_level[w] = lc;
          else atomicAdd(a.ctl + kCtlWlDropped, 1u);
        }
      }
      row_unlock(lk, lane);
    }
    if (refused) {  // rewrite q's own row without the refused hubs
      uint32_t* ovf;
      uint32_t* row = row_ptr(g, q, lc, &ovf);
      uint32_t kept = 0;
      __syncwarp();
      if (lane == 0) {
        for (uint32_t i = 0; i < n_sel; ++i) {
          const uint32_t v = a.sel_ids[(size_t)t * a.m + i];
          if (v != kEmpty) edit[kept++] = v;
        }
      }
      kept = __shfl_sync(kFull, kept, 0);
      __syncwarp();
      list_store(g, row, ovf, edit, kept, lane);
    }
  }
}

#endif  // HNSW_PLAIN_BUILD_KERNELS

// K3: re-select every queued row against the linked snapshot (read-only on the graph)
template <int EFR, class Dist>
__global__ void __launch_bounds__(256) build_reprune_kernel(Graph g, FastArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  // shared memory per warp: visited[vis_slots] | old[lcap] | query[dim] (if needed)
  uint32_t* base = smem + (size_t)warp * (a.vis_slots + a.lcap);
  Visited vis;
  vis.tab = base;
  vis.mask = a.vis_slots - 1;
  vis.shift = 32 - (31 - __clz(a.vis_slots));
  vis.limit = a.vis_slots - a.vis_slots / 4;
  uint32_t* old = base + a.vis_slots;
  float* smem_q = Dist::kNeedsSmemQuery
                      ? reinterpret_cast<float*>(smem + (size_t)warps * (a.vis_slots + a.lcap)) + (size_t)warp * g.dim
                      : nullptr;
  const uint32_t total = min(a.ctl[kCtlWlCount], a.wl_cap);
  Dist dist;
  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  uint32_t n_done = 0, n_skip = 0;
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(a.ctl + kCtlWorkReprune, 1u);
    w = __shfl_sync(kFull, w, 0);
    if (w >= total) break;
    const uint32_t e = a.wl_node[w], lc = a.wl_level[w];
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;
    uint32_t* ovf;
    uint32_t* row = row_ptr(g, e, lc, &ovf);
    uint32_t n_old = list_load(g, row, ovf, old, a.lcap, lane);
    bool ok = n_old != kEmpty;
    if (ok) {
      dist.load_query_slab(g, e, smem_q, lane);
      ok = reprune_select<EFR, Dist, true>(g, dist, e, lc, (int)cap, old, n_old, L, vis, cnt, lane);
    }
    if (!ok) {
      if (lane == 0) a.wl_len[2 * w] = 0, a.wl_len[2 * w + 1] = kEmpty;
      ++n_skip;
      continue;
    }
    ++n_done;
    for (uint32_t i = lane; i < n_old; i += 32) a.wl_old[(size_t)w * a.lcap + i] = old[i];
    L.for_each_prefix(L.len, lane, [&](int p, uint32_t nid, float) { a.wl_new[(size_t)w * g.W + p] = nid; });
    if (lane == 0) a.wl_len[2 * w] = n_old, a.wl_len[2 * w + 1] = (uint32_t)L.len;
  }
  if (lane == 0) {
    atomicAdd(a.ctl + kCtlDistEvals, cnt.n_dist);
    atomicAdd(a.ctl + kCtlReprunes, n_done);
    atomicAdd(a.ctl + kCtlSkipped, n_skip);
  }
}

// K4: update_node_connections for every re-selected row; every edit of a row happens under that row's lock
#ifdef HNSW_PLAIN_BUILD_KERNELS
__global__ void __launch_bounds__(256) build_apply_kernel(Graph g, FastArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  // shared memory per warp: old[lcap] | newl[W] | edit[lcap]
  uint32_t* old = smem + (size_t)warp * (2 * a.lcap + g.W);
  uint32_t* newl = old + a.lcap;
  uint32_t* edit = newl + g.W;
  const uint32_t total = min(a.ctl[kCtlWlCount], a.wl_cap);
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(a.ctl + kCtlWorkApply, 1u);
    w = __shfl_sync(kFull, w, 0);
    if (w >= total) break;
    const uint32_t n_old = a.wl_len[2 * w], n_new = a.wl_len[2 * w + 1];
    if (n_new == kEmpty) continue;
    const uint32_t e = a.wl_node[w], lc = a.wl_level[w];
    __syncwarp();
    for (uint32_t i = lane; i < n_old; i += 32) old[i] = a.wl_old[(size_t)w * a.lcap + i];
    for (uint32_t i = lane; i < n_new; i += 32) newl[i] = a.wl_new[(size_t)w * g.W + i];
    __syncwarp();
    uint32_t* elock = lock_of(g, row_key(g, e, lc));
    // e's own row: drop what the re-selection dropped, append what it added (rows of other nodes may be editing
    // e's row concurrently with mirrored operations, so this is an edit of the current row, not an overwrite)
    row_lock(elock, lane);
    {
      uint32_t* ovf;
      uint32_t* row = row_ptr(g, e, lc, &ovf);
      uint32_t len = list_load(g, row, ovf, edit, a.lcap, lane);
      if (len != kEmpty) {
        for (uint32_t j = 0; j < n_old; ++j) {
          if (list_find(newl, n_new, old[j], lane) >= 0) continue;
          int p = list_find(edit, len, old[j], lane);
          if (p >= 0) list_erase(edit, len, p, lane), --len;
        }
        for (uint32_t j = 0; j < n_new; ++j) {
          if (list_find(old, n_old, newl[j], lane) >= 0) continue;
          if (list_find(edit, len, newl[j], lane) >= 0 || len + 1 > a.lcap) continue;
          __syncwarp();
          if (lane == 0) edit[len] = newl[j];
          ++len;
          __syncwarp();
        }
        list_store(g, row, ovf, edit, len, lane);
      }
    }
    row_unlock(elock, lane);
    // mirrored removals (core.rs:808-816) and appends (core.rs:793-795)
    for (uint32_t j = 0; j < n_old; ++j) {
      const uint32_t x = old[j];
      if (list_find(newl, n_new, x, lane) >= 0) continue;
      uint32_t* lk = lock_of(g, row_key(g, x, lc));
      row_lock(lk, lane);
      row_remove(g, x, lc, e, edit, a.lcap, lane);
      row_unlock(lk, lane);
    }
    for (uint32_t j = 0; j < n_new; ++j) {
      const uint32_t x = newl[j];
      if (list_find(old, n_old, x, lane) >= 0) continue;
      uint32_t* lk = lock_of(g, row_key(g, x, lc));
      row_lock(lk, lane);
      bool full;
      row_append_unique(g, x, lc, e, edit, a.lcap, lane, &full);
      row_unlock(lk, lane);
      if (full) {  // x is a hub whose list is full: drop the edge on e's side too (keeps the graph symmetric)
        if (lane == 0) atomicAdd(a.ctl + kCtlRefused, 1u);
        row_lock(elock, lane);
        row_remove(g, e, lc, x, edit, a.lcap, lane);
        row_unlock(elock, lane);
      }
    }
  }
}

#endif  // HNSW_PLAIN_BUILD_KERNELS

}  // namespace hnsw

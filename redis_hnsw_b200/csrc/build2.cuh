// K1 of the batched builder, second generation: the per-insert searches of a batch (core.rs:511-531) on the
// TMA-staged machinery of search2.cuh — bulk-async row staging with an mbarrier per warp, transposed bit-exact
// reduction, lossy exact-tag visited table — instead of the register-staged kernel with an exact visited set.
//
// Why: with ef_construction = 200 the exact set needs 8192 slots x 4 B = 32 KB per warp, which leaves 4 warps per SM
// (measured: 12 % of the HBM roofline for K1, 81 % of the build's kernel time; profiles/r1d_build_launches.md).  The
// lossy table (4096 16-bit tags = 8 KB) plus an S-row stage fits 16+ warps per SM.  The result of a search is the same
// top-ef set either way (search2.cuh explains why forgetting a visited id is harmless), so `sel` is unchanged.
#pragma once
#include "build.cuh"
#include "search2.cuh"
#include "search_la.cuh"

namespace hnsw {

// rows per stage for the builder's searches (one choice per dimension keeps the number of instantiations down)
template <int C>
struct BuildStage {
  static constexpr int S = (C == 1) ? 32 : ((C <= 4) ? 8 : 4);
};

template <int EFR, int C, class T>
__global__ void __launch_bounds__(128, Search2Bounds<EFR, C>::kMinBlocks) build_search2_kernel(Graph g, FastArgs a) {
  constexpr int S = BuildStage<C>::S;
  constexpr int V = RowRegs<C>::V;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  unsigned char* base = smem2 + (size_t)warp * warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T));
  Warp2<C, S, T> w;
  warp2_setup<C, S, T>(w, base, a.vis_slots, lane);

  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  const int l_max = g.meta[kMetaMaxLayer];
  const uint32_t entry = (uint32_t)g.meta[kMetaEntry];
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.ctl + kCtlWorkSearch, 1u);
    wi = __shfl_sync(kFull, wi, 0);
    if (wi >= a.n_new) break;
    const uint32_t q = a.first + wi;
    const int l = g.level[q];
    const uint32_t tb = a.task_base[wi];
    {  // the query is the node's own (lane-permuted) slab row: chunk c of lane t sits at ((c / V) * 32 + t) * V + c % V
      const float* row = g.vecs + (size_t)q * (32 * C);
#pragma unroll
      for (int c = 0; c < C; ++c) w.q[c] = row[((c / V) * 32 + lane) * V + (c % V)];
    }
    uint32_t ep = entry;
    for (int lc = l_max; lc >= 0; --lc) {
      const bool link = lc <= l;
      search_layer2<EFR, C, S, T, RowCopy<C>::kDefault>(g, w, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane);  // core.rs:513, :524
      float s;
      L.get(0, lane, false, ep, s);                                                              // core.rs:514, :576
      if (!link) continue;
      const uint32_t n_sel = min((uint32_t)L.len, a.m);                                          // core.rs:531 (build.cuh header)
      uint32_t* out = a.sel_ids + (size_t)(tb + lc) * a.m;
#pragma unroll
      for (int r = 0; r < EFR; ++r) {
        uint32_t e = r * 32 + lane;
        if (e < n_sel) out[e] = L.id[r] & ~kExpanded;
      }
      if (lane == 0) a.sel_cnt[tb + lc] = n_sel;
    }
  }
  if (lane == 0) atomicAdd(a.ctl + kCtlDistEvals, cnt.n_dist);
}

}  // namespace hnsw

// ================================================================ EXACT insert / delete on the staged machinery
//
// insert_exact_kernel / delete_exact_kernel (build.cuh) are one warp walking a chain of dependent hops; with the
// register-staged distance provider a hop fetches its <= 32 new rows four at a time (five memory round trips per hop).
// Here a hop stages all of them with one round of bulk-async copies (S = 32 rows for dim <= 128), and the re-selection
// sweeps (core.rs:568, :853) run on the same stage.  The visited set is the lossy exact-tag table; candidate lists are
// protected by the explicit membership test of eval_and_admit, so the selected sets are the reference's.
namespace hnsw {

template <int C>
struct ExactStage {
  static constexpr int S = (C <= 4) ? 32 : 8;
};

template <int C, int S, class T>
__device__ __forceinline__ void load_q_from_slab(Warp2<C, S, T>& w, const Graph& g, uint32_t node, int lane) {
  constexpr int V = RowRegs<C>::V;
  const float* row = g.vecs + (size_t)node * (32 * C);
#pragma unroll
  for (int c = 0; c < C; ++c) w.q[c] = row[((c / V) * 32 + lane) * V + (c % V)];
}

// address of the first 128-byte line of the adjacency row of (node, level) for an L2 prefetch, or null
__device__ __forceinline__ const void* row_line(const Graph& g, uint32_t node, uint32_t level) {
  if (level == 0) return g.adj0 + (size_t)node * g.W;
  const uint32_t base = g.upper_base[node];
  if (base == kEmpty || (int32_t)level > g.level[node]) return nullptr;
  return g.adjU + (size_t)(base + level - 1) * g.W;
}

// select_neighbors(e, N(e), cap, lc, ignored) as insert (core.rs:568) and delete (core.rs:853) call it; see
// reprune_select in build.cuh.  w.q holds e's vector.  The sweep is a SET computation (top-cap by sim of everything
// reachable in two hops), so unseen ids are collected across adjacency rows in `pend` (shared memory, >= 64 words) and
// evaluated 32 at a time: one staging round per 32 candidates instead of one per row walked.
template <int EFR, int C, int S, class T>
__device__ __forceinline__ void reprune_select2(const Graph& g, Warp2<C, S, T>& w, uint32_t e, uint32_t level, int cap,
                                                const uint32_t* old, uint32_t n_old, CandList<EFR>& L, Counters& cnt,
                                                int lane, uint32_t ignored, uint32_t* pend) {
  w.seen.clear(lane);
  L.init();
  for (uint32_t i = 0; i < n_old; i += 32)                       // the rows the sweep is about to walk
    if (i + lane < n_old) {
      const void* p = row_line(g, old[i + lane], level);
      if (p) prefetch_l2(p);
    }
  uint32_t np = 0;                                                // warp-uniform number of pending ids
  auto flush = [&](uint32_t n) {                                  // evaluate pend[0..n), n <= 32
    __syncwarp();
    const uint32_t nb = lane < (int)n ? pend[lane] : kEmpty;
    const uint32_t rest = lane + 32 < (int)np ? pend[lane + 32] : kEmpty;
    __syncwarp();
    cnt.n_dist += n;
    eval_and_admit<EFR, C, S, T, RowCopy<C>::kDefault>(g, w, nb, n >= 32 ? kFull : ((1u << n) - 1u), cap, L, nullptr, lane);
    if (lane + 32 < (int)np) pend[lane] = rest;
    np -= n;
    __syncwarp();
  };
  auto feed = [&](uint32_t nb) {
    const bool valid = nb != kEmpty && nb != e && nb != ignored;  // never candidates (core.rs:704-708, 728-731)
    const bool is_new = valid && w.seen.test_and_set(nb);
    const uint32_t mask = __ballot_sync(kFull, is_new);
    if (!mask) return;
    if (is_new) pend[np + __popc(mask & ((1u << lane) - 1u))] = nb;
    np += __popc(mask);
    if (np >= 32) flush(32);
  };
  for (uint32_t i = 0; i < n_old; i += 32) feed((i + lane < n_old) ? old[i + lane] : kEmpty);   // core.rs:549-557
  for (uint32_t j = 0; j < n_old; ++j) {                         // extend_candidates (core.rs:698-721)
    uint32_t* ovf;
    const uint32_t* row = row_ptr(g, old[j], level, &ovf);
    if (!row) continue;
    uint32_t link = *ovf;
    bool more = true;
    for (uint32_t c = 0; c < g.W / 32 && more; ++c) {
      const uint32_t nb = row[c * 32 + lane];
      more = __shfl_sync(kFull, nb, 31) != kEmpty;
      feed(nb);
    }
    while (more && link != kEmpty) {
      uint32_t nb = g.pool[(size_t)link * 32 + lane];
      link = __shfl_sync(kFull, nb, 31);
      if (lane == 31) nb = kEmpty;
      feed(nb);
    }
  }
  if (np) flush(np);
  L.finish(lane);                                                // wide lists (EFR >= 4): back to the sorted layout
}

// update_node_connections (core.rs:776-822) for node `e` whose re-selected list is in L; shared by insert and delete.
template <int EFR, class Touch>
__device__ __forceinline__ void apply_reselection(const Graph& g, const CandList<EFR>& L, uint32_t e, uint32_t lc, uint32_t* erow,
                                                  uint32_t* eovf, const uint32_t* old, uint32_t n_old, uint32_t* keep_add,
                                                  uint32_t* rem, uint32_t* edit, uint32_t lcap, uint32_t victim, Touch touch,
                                                  int lane) {
  uint32_t n_keep, n_add, n_rem;
  reprune_delta<EFR>(L, old, n_old, keep_add, rem, n_keep, n_add, n_rem, lane);
  for (uint32_t t = lane; t < n_add + n_rem; t += 32) {           // the rows the mirrored edits below will touch
    const void* p = row_line(g, t < n_add ? keep_add[n_keep + t] : rem[t - n_add], lc);
    if (p) prefetch_l2(p);
  }
  list_store(g, erow, eovf, keep_add, n_keep + n_add, lane);
  touch(e);
  for (uint32_t t = 0; t < n_add; ++t) {                          // :793-796 (no cap check on the other side)
    row_append_unique(g, keep_add[n_keep + t], lc, e, edit, lcap, lane);
    touch(keep_add[n_keep + t]);
  }
  for (uint32_t t = 0; t < n_keep; ++t) touch(keep_add[t]);
  for (uint32_t t = 0; t < n_rem; ++t) {                          // :805-816
    if (rem[t] == victim) continue;                               // delete: not mirrored, not reported (core.rs:810-813)
    row_remove(g, rem[t], lc, e, edit, lcap, lane);
    touch(rem[t]);
  }
}

// core.rs:383-412 + 489-599, one warp, the NODE.ADD stream in order (see insert_exact_kernel)
// SMALL: the re-selection list (<= m_max_0 entries) lives in min(EFR, 2) registers per lane instead of EFR, which cuts
// the per-candidate list maintenance of the 2-hop sweeps by EFR/2 (the host picks it when m_max_0 <= 64).
template <int EFR, int C, bool SMALL>
__global__ void __launch_bounds__(32) insert_exact2_kernel(Graph g, ExactArgs a) {
  constexpr int ER = SMALL ? (EFR < 2 ? EFR : 2) : EFR;
  constexpr int S = ExactStage<C>::S;
  using T = uint32_t;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  Warp2<C, S, T> w;
  unsigned char* after = warp2_setup<C, S, T>(w, smem2, a.vis_slots, lane);
  constexpr bool kLookahead = kLookaheadInBuilders && S == 32 && RowCopy<C>::kOk;          // search_la.cuh
  LaBuf<C> lb;
  if constexpr (kLookahead) after = la_setup<C, S, T>(lb, w, after, lane);
  uint32_t* lists = reinterpret_cast<uint32_t*>(after);
  // sel[m] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap]
  uint32_t* sel = lists;
  uint32_t* old = sel + ((a.m + 31) & ~31u);
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;

  CandList<EFR> L;
  CandList<ER> R;
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
    const int l_max = g.meta[kMetaMaxLayer];                      // core.rs:496
    uint32_t ep = (uint32_t)g.meta[kMetaEntry];                   // core.rs:508
    for (int lc = l_max; lc >= 0; --lc) {
      const bool link = lc <= l;
      const uint32_t cap = lc == 0 ? a.cap0 : a.capU;             // core.rs:560
      load_q_from_slab<C, S, T>(w, g, q, lane);
      if constexpr (kLookahead) search_layer2_la<EFR, C, S, T>(g, w, lb, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane);   // :513, :524
      else search_layer2<EFR, C, S, T, RowCopy<C>::kDefault>(g, w, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane);
      float s;
      L.get(0, lane, false, ep, s);                               // :514 / :576
      if (!link) continue;
      uint32_t n_sel = min((uint32_t)L.len, a.m);                 // core.rs:531 (build.cuh header)
      if (a.efc < a.m && (uint32_t)L.len == a.efc) {
        // ef_construction < m and w came back full: N(w) can hold nodes the search turned away only because w was
        // full, and the reference's sweep (core.rs:698-721) picks them up until m are selected.  Same set computation
        // as a re-selection: top-m by sim(q, .) over w U N(w) \ {q}; w.q still holds q's vector.
        const uint32_t n_w = (uint32_t)L.len;
#pragma unroll
        for (int r = 0; r < EFR; ++r) {
          uint32_t e = r * 32 + lane;
          if (e < n_w) old[e] = L.id[r] & ~kExpanded;
        }
        __syncwarp();
        reprune_select2<ER, C, S, T>(g, w, q, (uint32_t)lc, (int)a.m, old, n_w, R, cnt, lane, kEmpty, keep_add);
        n_sel = min((uint32_t)R.len, a.m);
#pragma unroll
        for (int r = 0; r < ER; ++r) {
          uint32_t e = r * 32 + lane;
          if (e < n_sel) sel[e] = R.id[r] & ~kExpanded;
        }
      } else {
#pragma unroll
        for (int r = 0; r < EFR; ++r) {
          uint32_t e = r * 32 + lane;
          if (e < n_sel) sel[e] = L.id[r] & ~kExpanded;
        }
      }
      __syncwarp();
      {                                                           // connect_neighbors (core.rs:759-774)
        uint32_t* ovf;
        uint32_t* row = row_ptr(g, q, (uint32_t)lc, &ovf);
        list_store(g, row, ovf, sel, n_sel, lane);
      }
      for (uint32_t i = 0; i < n_sel; ++i) {
        row_append_unique(g, sel[i], (uint32_t)lc, q, edit, a.lcap, lane);
        touch(sel[i]);                                            // core.rs:535-537
      }
      for (uint32_t i = 0; i < n_sel; ++i) {                      // shrink connections (core.rs:540-574), nearest-first
        const uint32_t e = sel[i];
        uint32_t* eovf;
        uint32_t* erow = row_ptr(g, e, (uint32_t)lc, &eovf);
        const uint32_t n_old = list_load(g, erow, eovf, old, a.lcap, lane);
        if (n_old == kEmpty) {
          if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
          continue;
        }
        if (n_old <= cap) continue;                               // core.rs:561
        load_q_from_slab<C, S, T>(w, g, e, lane);
        reprune_select2<ER, C, S, T>(g, w, e, (uint32_t)lc, (int)cap, old, n_old, R, cnt, lane, kEmpty, keep_add);   // :568
        ++n_reprunes;
        apply_reselection<ER>(g, R, e, (uint32_t)lc, erow, eovf, old, n_old, keep_add, rem, edit, a.lcap, kEmpty, touch, lane);
      }
    }
    if (l > l_max && lane == 0) {                                 // core.rs:587-593
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

// core.rs:414-475 + 824-863 (see delete_exact_kernel)
template <int EFR, int C>
__global__ void __launch_bounds__(32) delete_exact2_kernel(Graph g, ExactArgs a) {
  constexpr int S = ExactStage<C>::S;
  using T = uint32_t;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  Warp2<C, S, T> w;
  uint32_t* lists = reinterpret_cast<uint32_t*>(warp2_setup<C, S, T>(w, smem2, a.vis_slots, lane));
  // vlist[lcap] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap]
  uint32_t* vlist = lists;
  uint32_t* old = vlist + a.lcap;
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;

  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  uint32_t n_touched = 0, n_reprunes = 0;
  auto touch = [&](uint32_t id) {
    if (a.touched) {
      if (lane == 0 && n_touched < a.touched_cap) a.touched[n_touched] = id;
      ++n_touched;
    }
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
        ok = false;
        break;
      }
      load_q_from_slab<C, S, T>(w, g, n, lane);
      reprune_select2<EFR, C, S, T>(g, w, n, (uint32_t)lc, (int)cap, old, n_old, L, cnt, lane, victim, keep_add);   // :853
      ++n_reprunes;
      apply_reselection<EFR>(g, L, n, (uint32_t)lc, nrow, novf, old, n_old, keep_add, rem, edit, a.lcap, victim, touch, lane);  // :856
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
  } else if (lane == 0) {
    atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
  }
  if (lane == 0) {
    a.ctl[kCtlDistEvals] += cnt.n_dist;
    a.ctl[kCtlReprunes] += n_reprunes;
    a.ctl[kCtlTouched] = n_touched;
  }
}

}  // namespace hnsw

// ================================================================ K3 of the batched builder on the staged machinery
//
// build_reprune_kernel (build.cuh) with reprune_select2: unseen ids are collected across the rows of a 2-hop sweep and
// evaluated a full stage at a time.  Same worklist protocol, same outputs (wl_old / wl_new / wl_len).
namespace hnsw {

template <int EFR, int C, class T>
__global__ void __launch_bounds__(128, (C > 8 ? 4 : 6)) build_reprune2_kernel(Graph g, FastArgs a) {
  constexpr int S = BuildStage<C>::S;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  // per warp: Warp2 region | old[lcap] | pend[64]
  const size_t per_warp = warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T)) + ((size_t)a.lcap + 64) * 4;
  unsigned char* base = smem2 + (size_t)warp * per_warp;
  Warp2<C, S, T> w;
  uint32_t* old = reinterpret_cast<uint32_t*>(warp2_setup<C, S, T>(w, base, a.vis_slots, lane));
  uint32_t* pend = old + a.lcap;
  const uint32_t total = min(a.ctl[kCtlWlCount], a.wl_cap);
  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  uint32_t n_done = 0, n_skip = 0;
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.ctl + kCtlWorkReprune, 1u);
    wi = __shfl_sync(kFull, wi, 0);
    if (wi >= total) break;
    const uint32_t e = a.wl_node[wi], lc = a.wl_level[wi];
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;
    uint32_t* ovf;
    uint32_t* row = row_ptr(g, e, lc, &ovf);
    const uint32_t n_old = list_load(g, row, ovf, old, a.lcap, lane);
    if (n_old == kEmpty) {
      if (lane == 0) a.wl_len[2 * wi] = 0, a.wl_len[2 * wi + 1] = kEmpty;
      ++n_skip;
      continue;
    }
    load_q_from_slab<C, S, T>(w, g, e, lane);
    reprune_select2<EFR, C, S, T>(g, w, e, lc, (int)cap, old, n_old, L, cnt, lane, kEmpty, pend);
    ++n_done;
    for (uint32_t i = lane; i < n_old; i += 32) a.wl_old[(size_t)wi * a.lcap + i] = old[i];
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      int p = r * 32 + lane;
      if (p < L.len) a.wl_new[(size_t)wi * g.W + p] = L.id[r] & ~kExpanded;
    }
    if (lane == 0) a.wl_len[2 * wi] = n_old, a.wl_len[2 * wi + 1] = (uint32_t)L.len;
    __syncwarp();
  }
  if (lane == 0) {
    atomicAdd(a.ctl + kCtlDistEvals, cnt.n_dist);
    atomicAdd(a.ctl + kCtlReprunes, n_done);
    atomicAdd(a.ctl + kCtlSkipped, n_skip);
  }
}

}  // namespace hnsw

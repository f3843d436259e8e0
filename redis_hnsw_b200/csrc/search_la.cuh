// search_layer with one-hop LOOKAHEAD (reference: src/hnsw/core.rs:607-675) — the latency flavour of search_layer2.
//
// A single search is a chain of dependent hops, and a hop is two dependent memory round trips: the adjacency row of the
// candidate, then the vectors of its unseen neighbours (profiles/r1e_search.md §3: one HNSW.SEARCH is latency-bound, not
// bandwidth-bound).  The candidate of the NEXT hop is almost always known before the current hop's rows have arrived — it is
// the nearest unexpanded entry of the list, unless one of the few nodes admitted by this hop beats it.  So while the
// current hop's row copies are in flight the warp reads the predicted candidate's adjacency row, filters it against the
// visited table WITHOUT marking anything, and issues the row copies of its unseen neighbours into a second stage
// (cp.async groups: wait_group 1 waits for the current hop only).  If the prediction holds, the next hop starts with its
// vectors already in shared memory; if not, the prefetched rows are dropped (nothing was marked, nothing is undone).
//
// Results are bit-identical to search_layer2: the only difference a hit can make is that ids marked by the remaining
// chunks of the previous hop's row (rows longer than 32 ids) are evaluated once more — and an evaluated node is either in
// the list (membership test before insertion) or lost against a threshold that has only risen since (core.rs:657).
//
// Used where latency is the product: one query per call / slices of a sharded batch (search_knn2_la_kernel), the one-warp
// EXACT insert and delete, and K1 of the SPEC builder.  Requires a 32-row stage and cp.async row copies (32-d / 128-d).
#pragma once
#include "search2.cuh"

namespace hnsw {

// Measured (r2 call G, 1M x 128, ef 64): one HNSW.SEARCH 211 us without / 224 us with the lookahead, a 1250-query slice
// 450 / 476 us, SPEC K1 unchanged — a hop is bound by its ~1100 dependent instructions (list maintenance, shuffles,
// ballots), not by the two memory round trips the lookahead overlaps (profiles/r2_latency.md).  The kernel stays behind
// the `lookahead` option (default off) as the record of that experiment; the builders do not use it.
constexpr bool kLookaheadInBuilders = false;

template <int C>
struct LaBuf {
  const float* stage[2];   // [32][32 * C] floats each
  uint32_t stage_s[2];     // shared-space addresses
  uint32_t* ids[2];        // [32] compacted ids of the rows in flight
};

// extra shared memory per warp for the second stage (the first one is the Warp2 region's)
__host__ __device__ inline size_t la_smem_bytes(uint32_t dim) { return (size_t)32 * dim * 4 + 128; }

template <int C, int S, class T>
__device__ __forceinline__ unsigned char* la_setup(LaBuf<C>& lb, const Warp2<C, S, T>& w, unsigned char* extra, int lane) {
  static_assert(S == 32, "the lookahead takes a whole adjacency chunk per round");
  lb.stage[0] = reinterpret_cast<const float*>(w.stage);
  lb.stage_s[0] = w.stage_s;
  lb.ids[0] = w.ids;
  lb.stage[1] = reinterpret_cast<const float*>(extra);
  lb.stage_s[1] = smem_u32(extra);
  lb.ids[1] = reinterpret_cast<uint32_t*>(extra + (size_t)32 * C * 128);
  uint4* z = reinterpret_cast<uint4*>(extra);
  for (uint32_t i = lane; i < 32u * C * 8u; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
  lb.ids[1][lane] = 0u;
  w.ids[lane] = 0u;
  __syncwarp();
  return extra + la_smem_bytes(32 * C);
}

// compact the ids flagged in `newmask` (lane j holds nb) into ids[], one cp.async row copy each, ONE commit group
template <int C>
__device__ __forceinline__ void la_issue(const Graph& g, uint32_t stage_s, uint32_t* ids, uint32_t nb, uint32_t newmask, int lane) {
  constexpr int V = RowRegs<C>::V;
  constexpr uint32_t RB = 128u * C;
  const int n_new = __popc(newmask);
  const int rank = __popc(newmask & ((1u << lane) - 1u));
  __syncwarp();                                              // earlier reads of ids[] and of this stage are complete
  if ((newmask >> lane) & 1u) ids[rank] = nb;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const uint32_t rid = ids[r];                             // stale beyond n_new (a valid id; the copy is predicated off)
    const float* src = g.vecs + (size_t)rid * (32 * C) + lane * V;
#pragma unroll
    for (int q = 0; q < C / V; ++q)
      cp_async_if<4 * V>(stage_s + (uint32_t)r * RB + (uint32_t)(q * 128 * V) + (uint32_t)(lane * 4 * V), src + q * 32 * V, r < n_new);
  }
  cp_async_commit();
}

// partials, transposed reduction and in-order admission of the `nr` rows of one stage (the second half of eval_and_admit)
template <int EFR, int C>
__device__ __forceinline__ void la_finish(const Graph& g, const float (&q)[C], const float* st, const uint32_t* ids, int nr, int ef,
                                          CandList<EFR>& L, const uint32_t* adj_prefetch, int lane) {
  constexpr int S = 32;
  constexpr int G = partial_group(C, S);
  float acc[S];
#pragma unroll
  for (int g0 = 0; g0 < S; g0 += G) {
    if (g0 == 0 || g0 < nr) {
#pragma unroll
      for (int r = g0; r < g0 + G; ++r) acc[r] = staged_partial<C>(q, st + (size_t)r * (32 * C), lane);
    } else {
#pragma unroll
      for (int r = g0; r < g0 + G; ++r) acc[r] = 0.0f;
    }
  }
  const float s = reduce_rows<S>(acc, lane);
  const uint32_t id = (lane < nr) ? ids[lane] : kEmpty;
  uint32_t cand = __ballot_sync(kFull, lane < nr && L.admits(s, ef));
  while (cand) {
    const int j = __ffs(cand) - 1;
    cand &= cand - 1;
    const float sj = __shfl_sync(kFull, s, j);
    const uint32_t idj = __shfl_sync(kFull, id, j);
    if (L.admits(sj, ef) && !list_has<EFR>(L, idj)) {          // core.rs:657 (threshold re-read per neighbour)
      L.push(sj, idj, ef, lane);                               // core.rs:658-664
      if (adj_prefetch && lane < (int)(g.W / 32)) prefetch_l2(adj_prefetch + (size_t)idj * g.W + lane * 32);
    }
  }
}

// core.rs:607-675 with lookahead; same contract as search_layer2
template <int EFR, int C, int S, class T, class Hook = NoSearchHook>
__device__ __forceinline__ void search_layer2_la(const Graph& g, Warp2<C, S, T>& w, const LaBuf<C>& lb, uint32_t ep, int ef,
                                                 uint32_t level, CandList<EFR>& L, Counters& cnt, int lane,
                                                 const Hook& hook = Hook()) {
  static_assert(S == 32 && RowCopy<C>::kOk, "lookahead needs a 32-row stage and cp.async row copies");
  w.seen.clear(lane);
  L.init();
  const uint32_t* adj_prefetch = (level == 0 && ef > 1) ? g.adj0 : nullptr;
  {
    const uint32_t nb = lane == 0 ? ep : kEmpty;                 // core.rs:617-628
    if (lane == 0) w.seen.test_and_set(ep);
    cnt.n_dist += 1;
    eval_and_admit<EFR, C, S, T, 1>(g, w, nb, 1u, ef, L, adj_prefetch, lane);
  }
  bool pf_valid = false, pf_more = false;
  uint32_t pf_cid = kEmpty, pf_nb = kEmpty, pf_mask = 0, pf_vcnt = 0;
  int pf_buf = 0;
  for (;;) {
    uint32_t cid;
    float cs;
    if (!L.pop(cid, cs, lane)) break;                            // core.rs:631-638
    cnt.n_hops += 1;
    uint32_t* ovf;
    const uint32_t* row = row_ptr(g, cid, level, &ovf);          // core.rs:642-645
    if (!row) continue;
    hook.expand(cid, level, L.worst);   // (the ids of the row are not handed to the hook here: see spec.cuh)
    hook.done();
    uint32_t nb, newmask;
    bool more;
    int buf;
    if (pf_valid && pf_cid == cid) {                             // the prediction held: this hop's rows are already on their way
      nb = pf_nb, newmask = pf_mask, more = pf_more, buf = pf_buf;
      cnt.n_adj += pf_vcnt;
      if ((newmask >> lane) & 1u) w.seen.set(nb);                // core.rs:648-649, deferred from the lookahead
      __syncwarp();
    } else {
      if (pf_valid) cp_async_wait_all();                         // a dropped lookahead may still be landing in the other stage
      nb = row[lane];
      more = __shfl_sync(kFull, nb, 31) != kEmpty;               // rows are compact: an empty tail ends the list
      const bool valid = nb != kEmpty;
      cnt.n_adj += __popc(__ballot_sync(kFull, valid));          // core.rs:646
      const bool is_new = valid && w.seen.test_and_set(nb);      // core.rs:648-649
      newmask = __ballot_sync(kFull, is_new);
      buf = 0;
      if (newmask) la_issue<C>(g, lb.stage_s[0], lb.ids[0], nb, newmask, lane);
    }
    pf_valid = false;
    const int n_new = __popc(newmask);
    cnt.n_dist += n_new;                                         // core.rs:652-656
    if (!more) {                                                 // lookahead: the nearest unexpanded entry as the list stands now
      uint32_t pid = kEmpty;
      if (L.peek(pid, lane)) {
        uint32_t* povf;
        const uint32_t* prow = row_ptr(g, pid, level, &povf);
        if (prow) {
          pf_nb = prow[lane];
          pf_more = __shfl_sync(kFull, pf_nb, 31) != kEmpty;
          const bool pvalid = pf_nb != kEmpty;
          pf_vcnt = __popc(__ballot_sync(kFull, pvalid));
          pf_mask = __ballot_sync(kFull, pvalid && w.seen.test(pf_nb));   // tested, NOT marked
          pf_cid = pid;
          pf_buf = buf ^ 1;
          pf_valid = true;
          la_issue<C>(g, lb.stage_s[pf_buf], lb.ids[pf_buf], pf_nb, pf_mask, lane);   // commits a group even when it is empty
        }
      }
    }
    if (n_new) {
      if (pf_valid) cp_async_wait_group<1>();                    // everything but the lookahead group = this hop's rows
      else cp_async_wait_all();
      __syncwarp();
      la_finish<EFR, C>(g, w.q, lb.stage[buf], lb.ids[buf], n_new, ef, L, adj_prefetch, lane);
    }
    if (more) {                                                  // a full first chunk: the rest of the row the plain way (no lookahead
      for (uint32_t c = 1; c < g.W / 32 && more; ++c) {          // was issued, so stage 0 is free)
        const uint32_t nb2 = row[c * 32 + lane];
        more = __shfl_sync(kFull, nb2, 31) != kEmpty;
        expand_chunk2<EFR, C, S, T, 1>(g, w, nb2, ef, L, cnt, adj_prefetch, lane);
      }
      if (more) {                                                // overflow rows (degree is unbounded); rare
        uint32_t link = *ovf;
        while (link != kEmpty) {
          uint32_t nb2 = g.pool[(size_t)link * 32 + lane];
          link = __shfl_sync(kFull, nb2, 31);
          if (lane == 31) nb2 = kEmpty;
          expand_chunk2<EFR, C, S, T, 1>(g, w, nb2, ef, L, cnt, adj_prefetch, lane);
        }
      }
    }
  }
  if (pf_valid) cp_async_wait_all();                             // nothing may still be landing when the stages are reused
  L.finish(lane);
}

// core.rs:477-486, 865-892 — search_knn2_kernel's latency flavour: 32-row stage, cp.async rows, one-hop lookahead.
// Chosen by the host when every query of the call is resident at once (one HNSW.SEARCH, a slice of a sharded batch).
template <int EFR, int C, class T>
__global__ void __launch_bounds__(128, 2) search_knn2_la_kernel(Graph g, SearchArgs a) {
  constexpr int S = 32;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const size_t per_warp = warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T)) + la_smem_bytes(32 * C);
  unsigned char* base = smem2 + (size_t)warp * per_warp;
  Warp2<C, S, T> w;
  LaBuf<C> lb;
  la_setup<C, S, T>(lb, w, warp2_setup<C, S, T>(w, base, a.vis_slots, lane), lane);

  CandList<EFR> L;
  uint32_t evals = 0;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(a.work_counter, 1u);
    qi = __shfl_sync(kFull, qi, 0);
    if (qi >= a.nq) break;
    Counters cnt = {0, 0, 0};
    const float* qn = a.queries + (size_t)qi * (32 * C);
#pragma unroll
    for (int c = 0; c < C; ++c) w.q[c] = qn[32 * c + lane];
    const int32_t entry = g.meta[kMetaEntry];
    uint32_t n_out = 0;
    if (entry >= 0) {                                            // core.rs:481-483
      uint32_t ep = (uint32_t)entry;
      for (int lc = g.meta[kMetaMaxLayer]; lc >= 0; --lc) {      // core.rs:869-876
        search_layer2_la<EFR, C, S, T>(g, w, lb, ep, lc > 0 ? 1 : (int)a.ef, (uint32_t)lc, L, cnt, lane);
        float s;
        if (lc > 0) L.get(0, lane, false, ep, s);
      }
      n_out = min((uint32_t)L.len, a.k);                         // core.rs:879
    }
#pragma unroll
    for (int r = 0; r < EFR; ++r) {                              // core.rs:878-891 nearest-first
      uint32_t e = r * 32 + lane;
      if (e < a.k) {
        bool have = e < n_out;
        a.ids[(size_t)qi * a.k + e] = have ? (L.id[r] & ~kExpanded) : kEmpty;
        a.sims[(size_t)qi * a.k + e] = have ? L.sim[r] : -CUDART_INF_F;
      }
    }
    for (uint32_t e = EFR * 32 + lane; e < a.k; e += 32) {
      a.ids[(size_t)qi * a.k + e] = kEmpty;
      a.sims[(size_t)qi * a.k + e] = -CUDART_INF_F;
    }
    if (lane == 0) {
      a.counts[qi] = n_out;
      if (a.stats) {
        a.stats[(size_t)qi * 4 + 0] = cnt.n_dist;                // evaluations performed (>= the reference's count)
        a.stats[(size_t)qi * 4 + 1] = cnt.n_adj;
        a.stats[(size_t)qi * 4 + 2] = cnt.n_hops;
        a.stats[(size_t)qi * 4 + 3] = 4u;                        // bit2: lossy visited table
      }
    }
    evals += cnt.n_dist;
  }
  if (lane == 0 && a.retry_count) atomicAdd(a.retry_count + 1, evals);
}

}  // namespace hnsw

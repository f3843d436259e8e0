// search_knn, second generation: TMA-staged vector rows (reference: src/hnsw/core.rs:607-675, 865-892).
//
// Same formulation as search.cuh (one warp per query, sorted ef-list with "expanded" flags in registers), but the
// distance batch of a hop is restructured around the memory system of a B200 SM:
//   * the neighbour vectors of one adjacency chunk (<= 32 rows) are fetched with one bulk-async copy each
//     (cp.async.bulk, the non-tensor TMA path; SASS UBLKCP) into a per-warp shared-memory stage and complete on a
//     per-warp mbarrier, so a warp keeps up to S rows (S * 4 * dim bytes) in flight without holding them in
//     registers;
//   * every lane reads its 16-byte slices of each staged row (LDS.128, conflict-free thanks to the lane-permuted
//     slab layout), forms the per-lane partial of metrics.rs:55-69, and the S partials are reduced TOGETHER by a
//     transposed butterfly: the same five exchange steps (xor 8, 16, 4, 1, 2 — the reference's hsum tree,
//     metrics.rs:25-42,71-74) but with rows paired so each step halves the number of live registers.
//     31 shuffles for 32 rows instead of 160, and bit-identical sums (every add has the reference's operand pair);
//   * the visited set is a direct-mapped exact-tag table: a hit proves "already evaluated", a conflict simply
//     forgets the older id.  Forgetting is harmless: a node that was evaluated and is not in the list lost against
//     the list's worst entry and will lose again (core.rs:657), and a node that is still in the list is caught by an
//     explicit membership test before insertion.  Results are identical to the exact set; only the number of
//     distance evaluations can exceed the reference's (reported by the kernel);
//   * the adjacency row of every admitted candidate is prefetched into L2 (all members of the final list get
//     expanded), taking one DRAM round trip out of the dependent chain of a hop.
#pragma once
#include <math_constants.h>

#include "search.cuh"

namespace hnsw {

// ---------------------------------------------------------------- mbarrier / bulk copy (PTX)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy completing on an mbarrier (bytes: multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// BYTES-wide (4 / 8 / 16) asynchronous copy global -> shared (LDGSTS), skipped when `on` is false; completion with
// cp_async_wait_all
template <int BYTES>
__device__ __forceinline__ void cp_async_if(uint32_t dst, const void* src, bool on) {
  if constexpr (BYTES == 16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %2, 0;\n"
        "@p cp.async.cg.shared.global [%0], [%1], 16;\n"
        "}\n" ::"r"(dst),
        "l"(src), "r"((int)on)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %2, 0;\n"
        "@p cp.async.ca.shared.global [%0], [%1], %3;\n"
        "}\n" ::"r"(dst),
        "l"(src), "r"((int)on), "n"(BYTES)
        : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most N of the most recently committed cp.async groups are still in flight
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------- transposed hsum of S row partials

// On entry acc[r] is this lane's partial of row r (metrics.rs:55-69).  On return lane r (< S) holds the complete
// reference-ordered sum of row r, negated (metrics.rs:75).
template <int S>
__device__ __forceinline__ float reduce_rows(float (&acc)[S], int lane) {
  constexpr int kOff[5] = {8, 16, 4, 1, 2};  // metrics.rs:71-74 (e1+e2)+(e3+e4) | :37-39 lo+hi | :25-32
#pragma unroll
  for (int lev = 0; lev < 5; ++lev) {
    const int off = kOff[lev];
    const bool up = (lane & off) != 0;
    const int n = S >> lev;  // live registers before this step
    if (n >= 2) {
#pragma unroll
      for (int i = 0; i < n / 2; ++i) {
        float a = acc[2 * i], b = acc[2 * i + 1];
        float send = up ? a : b, keep = up ? b : a;
        acc[i] = __fadd_rn(keep, __shfl_xor_sync(kFull, send, off));
      }
    } else {
      acc[0] = __fadd_rn(acc[0], __shfl_xor_sync(kFull, acc[0], off));
    }
  }
  // lane t = (b4 b3 b2 b1 b0) now holds row 16*b1 + 8*b0 + 4*b2 + 2*b4 + b3 (missing bits for S < 32 are free);
  // route row r to lane r
  const int r = lane;
  const int src = (((r >> 1) & 1) << 4) | ((r & 1) << 3) | (((r >> 2) & 1) << 2) | (((r >> 4) & 1) << 1) | ((r >> 3) & 1);
  return -__shfl_sync(kFull, acc[0], src);
}

// ---------------------------------------------------------------- lossy exact-tag visited table

// Direct-mapped table of the most recently evaluated ids.  An entry identifies its id EXACTLY (never a false
// "seen"); a conflicting id simply replaces the older one (a false "new" costs one redundant evaluation).
//   T = uint32_t : the entry is the id itself.
//   T = uint16_t : f(id) = id * odd mod 2^(B+15) is a bijection on ids < 2^(B+15) (B = log2 slots); slot = the top B
//                  bits of f, entry = 0x8000 | low 15 bits of f, so (slot, entry) still determines the id.  Half the
//                  shared memory per slot; the host picks it only when every id is below 2^(B+15).
template <class T>
struct Recent {
  T* tab;
  uint32_t n16;   // table bytes / 16
  int bits;       // B = log2(slots)
  __device__ __forceinline__ void clear(int lane) {
    __syncwarp();  // every lane's test_and_set stores of the previous search are ordered before the wipe
    const uint32_t fill = sizeof(T) == 4 ? kEmpty : 0u;
    uint4 e = make_uint4(fill, fill, fill, fill);
    uint4* t = reinterpret_cast<uint4*>(tab);
    for (uint32_t i = lane; i < n16; i += 32) t[i] = e;
    __syncwarp();
  }
  // true if `nid` was NOT found (and is now remembered)
  __device__ __forceinline__ bool test_and_set(uint32_t nid) {
    if constexpr (sizeof(T) == 4) {
      uint32_t slot = (nid * 2654435761u) >> (32 - bits);
      if (tab[slot] == nid) return false;
      tab[slot] = nid;
    } else {
      uint32_t f = (nid * 2654435761u) & ((1u << (bits + 15)) - 1u);
      uint32_t slot = f >> 15;
      T tag = (T)(0x8000u | (f & 0x7FFFu));
      if (tab[slot] == tag) return false;
      tab[slot] = tag;
    }
    return true;
  }
  // the two halves of test_and_set, for the lookahead (search_la.cuh): true if `nid` is NOT in the table / remember it
  __device__ __forceinline__ bool test(uint32_t nid) const {
    if constexpr (sizeof(T) == 4) {
      return tab[(nid * 2654435761u) >> (32 - bits)] != nid;
    } else {
      uint32_t f = (nid * 2654435761u) & ((1u << (bits + 15)) - 1u);
      return tab[f >> 15] != (T)(0x8000u | (f & 0x7FFFu));
    }
  }
  __device__ __forceinline__ void set(uint32_t nid) {
    if constexpr (sizeof(T) == 4) {
      tab[(nid * 2654435761u) >> (32 - bits)] = nid;
    } else {
      uint32_t f = (nid * 2654435761u) & ((1u << (bits + 15)) - 1u);
      tab[f >> 15] = (T)(0x8000u | (f & 0x7FFFu));
    }
  }
};

// DRAFT (not run on hardware): two-way set-associative flavour in the same shared memory.  A set is one 32-bit word
// holding two 16-bit exact tags, most recent in the low half; a miss moves the newcomer to the front and drops the older
// of the two (a 2-entry LRU).  Same exactness argument as the 16-bit direct-mapped table: (set, tag) determines the id for
// ids < 2^(log2(sets) + 15).  Aim: fewer forgotten ids -> fewer of the 6 % re-evaluations.
struct Way2 {
  uint32_t w;
};
template <>
struct Recent<Way2> {
  Way2* tab;
  uint32_t n16;   // table bytes / 16
  int bits;       // log2(sets)
  __device__ __forceinline__ void clear(int lane) {
    __syncwarp();
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    uint4* t = reinterpret_cast<uint4*>(tab);
    for (uint32_t i = lane; i < n16; i += 32) t[i] = e;
    __syncwarp();
  }
  __device__ __forceinline__ bool test_and_set(uint32_t nid) {
    const uint32_t f = (nid * 2654435761u) & ((1u << (bits + 15)) - 1u);
    const uint32_t set = f >> 15;
    const uint32_t tag = 0x8000u | (f & 0x7FFFu);
    const uint32_t cur = tab[set].w;
    const uint32_t t0 = cur & 0xFFFFu, t1 = cur >> 16;
    if (t0 == tag) return false;
    tab[set].w = (t0 << 16) | tag;  // newcomer (or the hit in way 1) moves to the front
    return t1 != tag;
  }
};

template <int EFR>
__device__ __forceinline__ bool list_has(const CandList<EFR>& L, uint32_t x) {
  bool hit = false;   // node ids are < 2^31, so an empty slot (0xFFFFFFFF) never equals x once the flag bit is masked... it
#pragma unroll        // would equal 0x7FFFFFFF, which is not a node id either
  for (int r = 0; r < EFR; ++r) hit |= (L.id[r] & ~kExpanded) == x;
  return __any_sync(kFull, hit);
}

// ---------------------------------------------------------------- per-warp context

template <int C, int S, class T>
struct Warp2 {
  static constexpr uint32_t kRowBytes = 128u * C;
  float q[C];          // q[c] = query[32c + lane]
  const float4* stage; // [S][C*8] float4 per... (S rows of 4*dim bytes)
  uint32_t stage_s;    // shared-space address of the stage
  uint32_t* ids;       // [32] compacted neighbour ids of the round
  uint32_t bar;        // shared-space address of the mbarrier
  uint32_t parity;
  Recent<T> seen;
  float* sims;         // COPY == 2 (CTA-cooperative rounds): [32] sims of the round's rows
  uint32_t* ctl;       // COPY == 2: ctl[0] = rows of the round (kEmpty = the query is finished)
};

// per-lane partial of one staged row (lane-permuted layout: group g of V chunks at float-V index g*32 + lane)
template <int C>
__device__ __forceinline__ float staged_partial(const float (&q)[C], const float* row, int lane) {
  constexpr int V = RowRegs<C>::V;
  float acc = 0.0f;
#pragma unroll
  for (int g = 0; g < C / V; ++g) {
    if constexpr (V == 4) {
      float4 v = reinterpret_cast<const float4*>(row)[g * 32 + lane];
      float d;
      d = __fsub_rn(q[4 * g + 0], v.x), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q[4 * g + 1], v.y), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q[4 * g + 2], v.z), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q[4 * g + 3], v.w), acc = __fmaf_rn(d, d, acc);
    } else if constexpr (V == 2) {
      float2 v = reinterpret_cast<const float2*>(row)[g * 32 + lane];
      float d;
      d = __fsub_rn(q[2 * g + 0], v.x), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q[2 * g + 1], v.y), acc = __fmaf_rn(d, d, acc);
    } else {
      float d = __fsub_rn(q[g], row[g * 32 + lane]);
      acc = __fmaf_rn(d, d, acc);
    }
  }
  return acc;
}

// Evaluate the ids flagged new (lane j holds nb) and apply them to the list in lane order (= adjacency-list order,
// core.rs:646-667).  `adj_prefetch` = base of the level-0 adjacency rows (or null) for the L2 prefetch of admitted ids.
// ---------------------------------------------------------------- CTA-cooperative rounds (COPY == 2)  -- DRAFT
// One query per CTA of 4 warps.  Warp 0 owns the search (list, visited table, adjacency walk); for every round it
// publishes the compacted ids, and all four warps copy, multiply and reduce 8 rows each, so the ~450 dependent
// instructions a 32-row round costs one warp (copies 147, partials 200, reduction 96; profiles/r1e_search.md) shrink
// to ~120 on the critical path.  Two named barriers per round.  NOT YET RUN ON HARDWARE.
// every warp arrives converged (compute-sanitizer synccheck flags a warp that reaches bar.sync in pieces, e.g. right after
// an `if (lane == 0)` store: found by the r2 sanitizer run, harmless on sm_100 but not something to rely on)
__device__ __forceinline__ void cta_bar(int id) {
  __syncwarp();
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

template <int C, int S, class T>
__device__ __forceinline__ void cta_round(const Graph& g, Warp2<C, S, T>& w, int nr, int wi, int lane) {
  static_assert(S == 32, "4 warps x 8 rows");
  constexpr int RPW = 8;
  constexpr int V = RowRegs<C>::V;
  constexpr uint32_t RB = Warp2<C, S, T>::kRowBytes;
  const int r0 = wi * RPW;
  if (r0 >= nr) return;
#pragma unroll
  for (int j = 0; j < RPW; ++j) {
    const int r = r0 + j;
    const uint32_t rid = w.ids[r];                         // stale beyond nr (copy predicated off)
    const float* src = g.vecs + (size_t)rid * (32 * C) + lane * V;
#pragma unroll
    for (int q = 0; q < C / V; ++q)
      cp_async_if<4 * V>(w.stage_s + (uint32_t)r * RB + (uint32_t)(q * 128 * V) + (uint32_t)(lane * 4 * V), src + q * 32 * V, r < nr);
  }
  cp_async_wait_all();
  __syncwarp();
  float acc[RPW];
  const float* st = reinterpret_cast<const float*>(w.stage);
#pragma unroll
  for (int j = 0; j < RPW; ++j) acc[j] = staged_partial<C>(w.q, st + (size_t)(r0 + j) * (32 * C), lane);
  const float s = reduce_rows<RPW>(acc, lane);            // lane j < 8 holds row r0 + j
  if (lane < RPW) w.sims[r0 + lane] = s;
}

// rows per unconditional group of eval_and_admit: about 32 floats of row data per lane, a power of two, 1 <= G <= min(S, 8)
__host__ __device__ constexpr int partial_group(int C, int S) {
  int g = 1;
  while (g * 2 * C <= 32 && g * 2 <= S && g * 2 <= 8) g *= 2;
  return g;
}

// COPY selects how the rows of a round reach the stage:
//   0  one bulk-async copy per row (UBLKCP) completing on the warp's mbarrier.  Bulk copies are warp-uniform
//      instructions, so a warp issues its rows one after the other (~9 instructions per row: ELECT, 4 x R2UR, ...).
//   1  one cp.async per lane and lane-permuted group of the row (LDGSTS, 4 * V bytes per lane): a 128-d row is ONE
//      warp-wide instruction, the ids come from the compacted list in shared memory; completion = cp.async.wait_all.
//      Measured at 1M x 128, ef 64: 9.29 -> 9.81 M QPS (profiles/r1e_search.md).  The default where a row is at most two
//      instructions (RowCopy<C>); 768-d rows (6 per row) stay on bulk copies.
template <int C>
struct RowCopy {
  static constexpr int kGroups = C / RowRegs<C>::V;       // cp.async instructions per row
  static constexpr bool kOk = kGroups <= 2;
  static constexpr int kDefault = kOk ? 1 : 0;
};

template <int EFR, int C, int S, class T, int COPY = 0>
__device__ __forceinline__ void eval_and_admit(const Graph& g, Warp2<C, S, T>& w, uint32_t nb, uint32_t newmask, int ef,
                                               CandList<EFR>& L, const uint32_t* adj_prefetch, int lane) {
  constexpr uint32_t RB = Warp2<C, S, T>::kRowBytes;
  const int n_new = __popc(newmask);
  const int rank = __popc(newmask & ((1u << lane) - 1u));
  const bool mine_new = (newmask >> lane) & 1u;
  for (int base = 0; base < n_new; base += S) {
    const int nr = min(S, n_new - base);
    __syncwarp();  // the previous round's reads of the stage and of ids[] are complete
    float s;
    if constexpr (COPY == 2) {
      if (mine_new && rank >= base && rank < base + nr) w.ids[rank - base] = nb;
      if (lane == 0) w.ctl[0] = (uint32_t)nr;
      cta_bar(1);                                          // ids and nr are published: the other warps start
      cta_round<C, S, T>(g, w, nr, 0, lane);
      cta_bar(2);                                          // every warp has written its sims
      s = lane < nr ? w.sims[lane] : 0.0f;
    } else {
    if constexpr (COPY == 0) {
      if (lane == 0) mbar_expect_tx(w.bar, (uint32_t)nr * RB);
      __syncwarp();
      if (mine_new && rank >= base && rank < base + nr) {
        w.ids[rank - base] = nb;
        bulk_g2s(w.stage_s + (uint32_t)(rank - base) * RB, g.vecs + (size_t)nb * (32 * C), RB, w.bar);
      }
      mbar_wait(w.bar, w.parity);
      w.parity ^= 1u;
    } else {
      if (mine_new && rank >= base && rank < base + nr) w.ids[rank - base] = nb;
      __syncwarp();
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const uint32_t rid = w.ids[r];                     // broadcast read; stale beyond nr (copy predicated off)
        constexpr int V = RowRegs<C>::V;                   // a group = 32 lanes x V floats, lane t's slice at float t * V
        const float* src = g.vecs + (size_t)rid * (32 * C) + lane * V;
#pragma unroll
        for (int q = 0; q < C / V; ++q)
          cp_async_if<4 * V>(w.stage_s + (uint32_t)r * RB + (uint32_t)(q * 128 * V) + (uint32_t)(lane * 4 * V), src + q * 32 * V,
                             r < nr);
      }
      cp_async_wait_all();
    }
    __syncwarp();
    // Partials are formed for whole groups of G rows without looking at nr: the loads of a group issue back to back
    // and its arithmetic interleaves (a per-row test costs 5 instructions and serialises the shared-memory latencies).
    // Rows at or beyond nr hold older rows (zeros before the first use); their sums are never looked at.
    constexpr int G = partial_group(C, S);
    float acc[S];
    const float* st = reinterpret_cast<const float*>(w.stage);
#pragma unroll
    for (int g0 = 0; g0 < S; g0 += G) {
      if (g0 == 0 || g0 < nr) {
#pragma unroll
        for (int r = g0; r < g0 + G; ++r) acc[r] = staged_partial<C>(w.q, st + (size_t)r * (32 * C), lane);
      } else {
#pragma unroll
        for (int r = g0; r < g0 + G; ++r) acc[r] = 0.0f;
      }
    }
    s = reduce_rows<S>(acc, lane);
    }
    const uint32_t id = (lane < nr) ? w.ids[lane] : kEmpty;
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
}

template <int EFR, int C, int S, class T, int COPY = 0>
__device__ __forceinline__ void expand_chunk2(const Graph& g, Warp2<C, S, T>& w, uint32_t nb, int ef, CandList<EFR>& L,
                                              Counters& cnt, const uint32_t* adj_prefetch, int lane) {
  const bool valid = nb != kEmpty;
  const uint32_t vmask = __ballot_sync(kFull, valid);
  if (!vmask) return;
  cnt.n_adj += __popc(vmask);                                    // core.rs:646
  const bool is_new = valid && w.seen.test_and_set(nb);          // core.rs:648-649 (ids of one list are distinct)
  const uint32_t newmask = __ballot_sync(kFull, is_new);
  if (!newmask) return;
  cnt.n_dist += __popc(newmask);                                 // core.rs:652-656
  eval_and_admit<EFR, C, S, T, COPY>(g, w, nb, newmask, ef, L, adj_prefetch, lane);
}

// observer of the rows a search expands (the SPEC builder logs them as its read set, spec.cuh); the default does nothing
struct NoSearchHook {
  __device__ __forceinline__ void expand(uint32_t, uint32_t, float) const {}   // (node, level, admission threshold of the moment)
  __device__ __forceinline__ void ids(uint32_t) const {}                       // a chunk of the row's ids, one per lane (kEmpty = none)
  __device__ __forceinline__ void done() const {}
};

// core.rs:607-675
template <int EFR, int C, int S, class T, int COPY = 0, class Hook = NoSearchHook>
__device__ __forceinline__ void search_layer2(const Graph& g, Warp2<C, S, T>& w, uint32_t ep, int ef, uint32_t level,
                                              CandList<EFR>& L, Counters& cnt, int lane, const Hook& hook = Hook()) {
  w.seen.clear(lane);
  L.init();
  const uint32_t* adj_prefetch = (level == 0 && ef > 1) ? g.adj0 : nullptr;
  {
    const uint32_t nb = lane == 0 ? ep : kEmpty;                 // core.rs:617-628
    if (lane == 0) w.seen.test_and_set(ep);
    cnt.n_dist += 1;
    eval_and_admit<EFR, C, S, T, COPY>(g, w, nb, 1u, ef, L, adj_prefetch, lane);
  }
  for (;;) {
    uint32_t cid;
    float cs;
    if (!L.pop(cid, cs, lane)) break;                            // core.rs:631-638
    cnt.n_hops += 1;
    uint32_t* ovf;
    const uint32_t* row = row_ptr(g, cid, level, &ovf);          // core.rs:642-645
    if (!row) continue;
    hook.expand(cid, level, L.worst);
    bool more = true;
    for (uint32_t c = 0; c < g.W / 32 && more; ++c) {
      const uint32_t nb = row[c * 32 + lane];
      more = __shfl_sync(kFull, nb, 31) != kEmpty;               // rows are compact: an empty tail ends the list
      hook.ids(nb);
      expand_chunk2<EFR, C, S, T, COPY>(g, w, nb, ef, L, cnt, adj_prefetch, lane);
    }
    if (more) {                                                  // overflow rows (degree is unbounded); rare
      uint32_t link = *ovf;
      while (link != kEmpty) {
        uint32_t nb = g.pool[(size_t)link * 32 + lane];
        link = __shfl_sync(kFull, nb, 31);
        if (lane == 31) nb = kEmpty;
        hook.ids(nb);
        expand_chunk2<EFR, C, S, T, COPY>(g, w, nb, ef, L, cnt, adj_prefetch, lane);
      }
    }
    hook.done();
  }
  L.finish(lane);                                                // wide lists: back to the sorted layout (search.cuh)
}

// shared memory per warp (bytes), 128-byte aligned pieces: stage | visited | ids | mbarrier
__host__ __device__ inline size_t warp2_smem_bytes(uint32_t dim, int S, uint32_t vis_slots, uint32_t slot_bytes) {
  return (size_t)S * dim * 4 + (((size_t)vis_slots * slot_bytes + 127) & ~(size_t)127) + 128 + 128;
}

template <int C, int S, class T>
__device__ __forceinline__ unsigned char* warp2_setup(Warp2<C, S, T>& w, unsigned char* base, uint32_t vis_slots, int lane) {
  w.stage = reinterpret_cast<const float4*>(base);
  w.stage_s = smem_u32(base);
  const uint32_t tab_bytes = (vis_slots * (uint32_t)sizeof(T) + 127u) & ~127u;
  w.seen.tab = reinterpret_cast<T*>(base + (size_t)S * C * 128);
  w.seen.n16 = tab_bytes / 16;
  w.seen.bits = 31 - __clz(vis_slots);
  w.ids = reinterpret_cast<uint32_t*>(base + (size_t)S * C * 128 + tab_bytes);
  w.bar = smem_u32(w.ids + 32);
  w.parity = 0;
  if (lane == 0) mbar_init(w.bar, 1);
  // the stage starts as zeros: eval_and_admit forms partials for whole row groups, also beyond the rows of a round
  uint4* z = reinterpret_cast<uint4*>(base);
  for (uint32_t i = lane; i < (uint32_t)S * C * 8; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  return base + warp2_smem_bytes(32 * C, S, vis_slots, sizeof(T));
}

// core.rs:477-486, 865-892
// Register budget per list size: the kernel's throughput follows resident warps (profiles/tune_r1.md,
// profiles/r1d_curve.md), and without a cap ptxas takes 105 / 120 registers for 8 / 16 list registers per lane.
// __launch_bounds__(128, B) caps registers at 65536 / (128 * B): 64 for B = 8, 72 for 7, 80 for 6, 128 for 4.
// Rows of 768 floats (C = 24) keep 24 query registers per lane and are bandwidth/power-bound with few warps: a cap
// there only adds spills (measured: 256 k -> 232 k QPS at 1M x 768, ef = 200), so large C keeps 128 registers.
template <int EFR, int C = 4>
struct Search2Bounds {
  static constexpr int kMinBlocks = EFR >= 32 ? 3 : (C > 8 ? 4 : (EFR <= 4 ? 8 : (EFR == 8 ? 7 : 6)));
};

template <int EFR, int C, int S, class T, int COPY = 0>
__global__ void __launch_bounds__(128, Search2Bounds<EFR, C>::kMinBlocks) search_knn2_kernel(Graph g, SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  unsigned char* base = smem2 + (size_t)warp * warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T));
  Warp2<C, S, T> w;
  warp2_setup<C, S, T>(w, base, a.vis_slots, lane);

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
        search_layer2<EFR, C, S, T, COPY>(g, w, ep, lc > 0 ? 1 : (int)a.ef, (uint32_t)lc, L, cnt, lane);
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
        a.stats[(size_t)qi * 4 + 3] = 4u;                        // bit2: lossy visited table (n_dist may exceed the reference's)
      }
    }
    evals += cnt.n_dist;
  }
  if (lane == 0 && a.retry_count) atomicAdd(a.retry_count + 1, evals);  // ctl[3]: evaluations of the launch (low 32 bits)
}

// ---------------------------------------------------------------- one query per CTA (COPY == 2)  -- DRAFT, see cta_round
// grid = nq, block = 128.  Shared memory: one Warp2 region with a 32-row stage | sims[32] | ctl.
template <int EFR, int C, class T>
__global__ void __launch_bounds__(128, 1) search_knn2_cta_kernel(Graph g, SearchArgs a) {
  constexpr int S = 32;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int wi = threadIdx.x >> 5;
  Warp2<C, S, T> w;
  unsigned char* extra;
  if (wi == 0) {
    extra = warp2_setup<C, S, T>(w, smem2, a.vis_slots, lane);   // tags are cleared per search_layer2, stage zeroed here
  } else {                                                        // same pointers, no initialisation
    w.stage = reinterpret_cast<const float4*>(smem2);
    w.stage_s = smem_u32(smem2);
    const uint32_t tab_bytes = (a.vis_slots * (uint32_t)sizeof(T) + 127u) & ~127u;
    w.ids = reinterpret_cast<uint32_t*>(smem2 + (size_t)S * C * 128 + tab_bytes);
    extra = smem2 + warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T));
  }
  w.sims = reinterpret_cast<float*>(extra);
  w.ctl = reinterpret_cast<uint32_t*>(extra + 128);
  __syncthreads();
  const uint32_t qi = blockIdx.x;
  if (qi >= a.nq) return;                                         // whole CTA
  const float* qn = a.queries + (size_t)qi * (32 * C);
#pragma unroll
  for (int c = 0; c < C; ++c) w.q[c] = qn[32 * c + lane];
  if (wi != 0) {                                                  // workers: serve rounds until the owner says stop
    for (;;) {
      cta_bar(1);
      const uint32_t nr = w.ctl[0];
      if (nr == kEmpty) return;
      cta_round<C, S, T>(g, w, (int)nr, wi, lane);
      cta_bar(2);
    }
  }
  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  const int32_t entry = g.meta[kMetaEntry];
  uint32_t n_out = 0;
  if (entry >= 0) {                                               // core.rs:481-483
    uint32_t ep = (uint32_t)entry;
    for (int lc = g.meta[kMetaMaxLayer]; lc >= 0; --lc) {         // core.rs:869-876
      search_layer2<EFR, C, S, T, 2>(g, w, ep, lc > 0 ? 1 : (int)a.ef, (uint32_t)lc, L, cnt, lane);
      float s;
      if (lc > 0) L.get(0, lane, false, ep, s);
    }
    n_out = min((uint32_t)L.len, a.k);                            // core.rs:879
  }
  if (lane == 0) w.ctl[0] = kEmpty;                               // release the workers
  cta_bar(1);
#pragma unroll
  for (int r = 0; r < EFR; ++r) {                                 // core.rs:878-891 nearest-first
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
      a.stats[(size_t)qi * 4 + 0] = cnt.n_dist;
      a.stats[(size_t)qi * 4 + 1] = cnt.n_adj;
      a.stats[(size_t)qi * 4 + 2] = cnt.n_hops;
      a.stats[(size_t)qi * 4 + 3] = 4u;
    }
  }
}

}  // namespace hnsw

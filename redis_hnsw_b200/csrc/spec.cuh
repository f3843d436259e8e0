// SPEC builder: the NODE.ADD stream of the reference (core.rs:383-412, 489-599) executed speculatively in windows and
// committed strictly in stream order, so that the graph is the sequential one (list for list) while the long part of
// an insert — the ef_construction search, ~220 dependent hops — runs for many inserts at once.
//
//   K1  spec_exec_kernel     one warp per insert of the window [f, f + B).  The whole insert (searches, select, connect,
//                            re-selection of over-full rows) runs against the graph as it stands — every insert < f is
//                            committed, nothing else is — WITHOUT writing to it: new row contents go to a private write
//                            log (read back by the insert itself: read-your-own-writes), and every graph row the insert
//                            looked at goes to a read log.
//   K2  spec_commit_kernel   one CTA walks the window in order.  Insert q commits iff no row of its read log was written
//                            since its snapshot (row stamps: ver[row] = 1 + id of the last insert that wrote the row).  A
//                            valid insert saw exactly the rows the sequential execution would have seen, and the insert is a
//                            deterministic function of those rows, so its write log IS the sequential result; it is copied
//                            into the graph and the rows are stamped.  The walk stops at the first invalid insert, which is
//                            re-executed by the next K1 (now as the head of its window, where it cannot fail again).
//   Inserts that were executed but not reached keep their logs; the next K1 re-validates them against the stamps and only
//   re-executes the ones that lost a row.
//
// Exactness does not rest on any probability: an insert is committed only when its whole read set is untouched.  What
// speculation buys is measured (tools/sim_spec_build.cpp): the dependency chains between consecutive
// inserts are dense (every insert rewrites ~46 rows and reads ~350), so a round commits a prefix of O(sqrt N)-ish inserts.
#pragma once
#include "build2.cuh"
#include "search_la.cuh"

namespace hnsw {

enum SpecHdr : int {
  kSpecState = 0,     // 0 = needs execution, 1 = executed (logs valid for `snap`)
  kSpecSnap = 1,      // every insert with id < snap was committed when the logs were made
  kSpecNode = 2,
  kSpecReads = 3,
  kSpecEntries = 4,
  kSpecFlags = 5,     // 1 = read log overflowed (valid only as the head of a window), 2 = write log overflowed (unusable)
  kSpecDist = 6,
  kSpecReprunes = 7,
  kSpecT0 = 8,        // diagnostics: %globaltimer (low word, ns) when the warp entered K1,
  kSpecDur = 9,       //              ns it spent there,
  kSpecSm = 10,       //              SM it ran on | 0x80000000 when it executed (not just validated)
  kSpecHdrWords = 12,
};
constexpr uint32_t kSpecRdOverflow = 1, kSpecWrOverflow = 2;

enum SpecCtl : int {
  kSpecCommitted = 0,  // inserts committed by this K2
  kSpecReason = 1,     // why the walk stopped: 0 end of window, 1 not executed, 2 write log overflow, 3 invalid, 4 pool low
  kSpecExecuted = 2,   // inserts (re-)executed by K1 (accumulates)
  kSpecDistEvals = 3,  // distance evaluations of committed inserts (accumulates)
  kSpecReprunesDone = 4,
  kSpecDistWasted = 5, // distance evaluations of executions that were thrown away
  kSpecPoolUsed = 6,   // copies of the index scalars after the commit walk (one transfer per round)
  kSpecMaxLayer = 7,
  kSpecEntry = 8,
  kSpecError = 9,
  kSpecCtlWords = 16,
};

struct SpecArgs {
  uint32_t frontier;   // first insert of the window; every id below is committed
  uint32_t count;      // window size
  uint32_t ring;       // slots (power of two); slot = id & (ring - 1)
  uint32_t m, cap0, capU, efc, lcap, vis_slots;
  uint32_t rcap, wcap, wmaxe;
  uint32_t* hdr;       // [ring][kSpecHdrWords]
  uint32_t* rd;        // [ring][rcap]   row keys
  uint32_t* wkey;      // [ring][wmaxe]  row key of entry e (kEmpty = dead)
  uint32_t* woff;      // [ring][wmaxe]  word offset of entry e in wdata: {reserved, len, ids...}
  uint32_t* wdata;     // [ring][wcap]
  uint32_t* ver0;      // [n]   1 + id of the last insert that wrote the level-0 row
  uint32_t* verU;      // [nU]
  uint32_t* ctl;
};

__device__ __forceinline__ uint32_t spec_ver(const SpecArgs& a, uint32_t key) {
  return (key & 0x80000000u) ? __ldcg(a.verU + (key & 0x7FFFFFFFu)) : __ldcg(a.ver0 + key);
}

// ---------------------------------------------------------------- per-warp logs

struct SpecLog {
  uint32_t* rd;        // global
  uint32_t* wdata;     // global
  uint32_t* wkey_s;    // shared [wmaxe]
  uint32_t* woff_s;    // shared [wmaxe]
  uint32_t rcap, wcap, wmaxe;
  uint32_t n_reads, n_entries, used, flags;  // warp-uniform

  __device__ __forceinline__ void read(uint32_t key, int lane) {
    if (n_reads < rcap) {
      if (lane == 0) rd[n_reads] = key;
      ++n_reads;
    } else {
      flags |= kSpecRdOverflow;
    }
  }
  // index of the live entry of `key`, or -1
  __device__ __forceinline__ int find(uint32_t key, int lane) const {
    for (uint32_t i = 0; i < n_entries; i += 32) {
      const uint32_t b = __ballot_sync(kFull, i + lane < n_entries && wkey_s[i + lane] == key);
      if (b) return (int)i + __ffs(b) - 1;
    }
    return -1;
  }
};

// hook of search_layer2: every expanded row is a read (core.rs:642-646)
struct SpecSearchHook {
  const uint32_t* upper_base;
  SpecLog* lg;
  int lane;
  __device__ __forceinline__ void expand(uint32_t node, uint32_t level) const {
    lg->read(level == 0 ? node : (0x80000000u | (upper_base[node] + level - 1)), lane);   // row_key()
  }
};

// the insert's view of the adjacency list of (node, level): its own latest version, else the graph's (logged as a read)
__device__ __forceinline__ uint32_t view_load(const Graph& g, SpecLog& lg, uint32_t node, uint32_t level, uint32_t* buf,
                                              uint32_t lcap, int lane) {
  const uint32_t key = row_key(g, node, level);
  const int e = lg.find(key, lane);
  if (e >= 0) {
    const uint32_t* p = lg.wdata + lg.woff_s[e];
    const uint32_t len = p[1];
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) buf[i] = p[2 + i];
    __syncwarp();
    return len;
  }
  lg.read(key, lane);
  uint32_t* ovf;
  const uint32_t* row = row_ptr(g, node, level, &ovf);
  if (!row) return 0;
  return list_load(g, row, ovf, buf, lcap, lane);
}

// new content of (node, level) -> write log (in place when the row already has an entry that is large enough)
__device__ __forceinline__ void view_store(const Graph& g, SpecLog& lg, uint32_t node, uint32_t level, const uint32_t* buf,
                                           uint32_t len, int lane) {
  const uint32_t key = row_key(g, node, level);
  __syncwarp();
  int e = lg.find(key, lane);
  uint32_t off;
  if (e >= 0 && lg.wdata[lg.woff_s[e]] >= len) {
    off = lg.woff_s[e];
  } else {
    if (e >= 0 && lane == 0) lg.wkey_s[e] = kEmpty;            // superseded
    const uint32_t reserve = len + 8;
    if (lg.n_entries >= lg.wmaxe || lg.used + 2 + reserve > lg.wcap) {
      lg.flags |= kSpecWrOverflow;
      __syncwarp();
      return;
    }
    off = lg.used;
    if (lane == 0) {
      lg.wkey_s[lg.n_entries] = key;
      lg.woff_s[lg.n_entries] = off;
      lg.wdata[off] = reserve;
    }
    lg.n_entries += 1;
    lg.used += 2 + reserve;
  }
  if (lane == 0) lg.wdata[off + 1] = len;
  for (uint32_t i = lane; i < len; i += 32) lg.wdata[off + 2 + i] = buf[i];
  __syncwarp();
}

// reprune_select2 (build2.cuh) over the insert's view: rows of the sweep come through view_load into `tmp`
template <int EFR, int C, int S, class T>
__device__ __forceinline__ void reprune_select2v(const Graph& g, SpecLog& lg, Warp2<C, S, T>& w, uint32_t e, uint32_t level,
                                                 int cap, const uint32_t* old, uint32_t n_old, CandList<EFR>& L,
                                                 Counters& cnt, int lane, uint32_t* pend, uint32_t* tmp, uint32_t lcap) {
  w.seen.clear(lane);
  L.init();
  for (uint32_t i = 0; i < n_old; i += 32)
    if (i + lane < n_old) {
      const void* p = row_line(g, old[i + lane], level);
      if (p) prefetch_l2(p);
    }
  uint32_t np = 0;
  auto flush = [&](uint32_t n) {
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
    const bool valid = nb != kEmpty && nb != e;                   // core.rs:704-708, 728-731
    const bool is_new = valid && w.seen.test_and_set(nb);
    const uint32_t mask = __ballot_sync(kFull, is_new);
    if (!mask) return;
    if (is_new) pend[np + __popc(mask & ((1u << lane) - 1u))] = nb;
    np += __popc(mask);
    if (np >= 32) flush(32);
  };
  for (uint32_t i = 0; i < n_old; i += 32) feed((i + lane < n_old) ? old[i + lane] : kEmpty);   // core.rs:549-557
  for (uint32_t j = 0; j < n_old; ++j) {                         // extend_candidates (core.rs:698-721)
    const uint32_t n_row = view_load(g, lg, old[j], level, tmp, lcap, lane);
    if (n_row == kEmpty) {
      if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
      continue;
    }
    for (uint32_t i = 0; i < n_row; i += 32) feed((i + lane < n_row) ? tmp[i + lane] : kEmpty);
  }
  if (np) flush(np);
  L.finish(lane);                                                // wide lists (EFR >= 4): back to the sorted layout
}

// ---------------------------------------------------------------- K1

template <int EFR, int C, bool SMALL>
__global__ void __launch_bounds__(32) spec_exec_kernel(Graph g, SpecArgs a) {
  constexpr int ER = SMALL ? (EFR < 2 ? EFR : 2) : EFR;
  constexpr int S = ExactStage<C>::S;
  using T = uint32_t;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const uint32_t q = a.frontier + blockIdx.x;
  const uint32_t slot = q & (a.ring - 1);
  uint32_t* hdr = a.hdr + (size_t)slot * kSpecHdrWords;
  uint32_t* rd = a.rd + (size_t)slot * a.rcap;
  uint64_t t_in;
  uint32_t smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (lane == 0) hdr[kSpecT0] = (uint32_t)t_in, hdr[kSpecDur] = 0, hdr[kSpecSm] = smid;

  // executed earlier and still valid?  (rows stamped after the snapshot invalidate the logs)
  if (__ldcg(hdr + kSpecState) == 1u && __ldcg(hdr + kSpecNode) == q) {
    const uint32_t snap = __ldcg(hdr + kSpecSnap), n_reads = __ldcg(hdr + kSpecReads), flags = __ldcg(hdr + kSpecFlags);
    if (flags & kSpecWrOverflow) return;                          // unusable either way: the host runs it through EXACT
    // Warp-uniform control flow on purpose: a per-lane early exit from this loop left the warp split into groups that
    // ran the whole insert below one after the other (measured: 4.1 ms instead of 0.9 ms per execution, r2 call D).
    bool bad = (flags & kSpecRdOverflow) && snap != q;
    for (uint32_t i = 0; i < n_reads && !bad; i += 32) {
      const bool mine = i + lane < n_reads && spec_ver(a, __ldcg(rd + i + lane)) > snap;
      bad = __any_sync(kFull, mine);
    }
    if (!bad) return;
    if (lane == 0) atomicAdd(a.ctl + kSpecDistWasted, hdr[kSpecDist]);
    __syncwarp();
  }

  Warp2<C, S, T> w;
  unsigned char* after = warp2_setup<C, S, T>(w, smem2, a.vis_slots, lane);
  constexpr bool kLookahead = kLookaheadInBuilders && S == 32 && RowCopy<C>::kOk;          // search_la.cuh: second stage for the next hop's rows
  LaBuf<C> lb;
  if constexpr (kLookahead) after = la_setup<C, S, T>(lb, w, after, lane);
  uint32_t* lists = reinterpret_cast<uint32_t*>(after);
  // sel[m] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap] | tmp[lcap] | wkey[wmaxe] | woff[wmaxe]
  uint32_t* sel = lists;
  uint32_t* old = sel + ((a.m + 31) & ~31u);
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;
  uint32_t* tmp = edit + a.lcap;
  SpecLog lg;
  lg.rd = rd;
  lg.wdata = a.wdata + (size_t)slot * a.wcap;
  lg.wkey_s = tmp + a.lcap;
  lg.woff_s = lg.wkey_s + a.wmaxe;
  lg.rcap = a.rcap, lg.wcap = a.wcap, lg.wmaxe = a.wmaxe;
  lg.n_reads = lg.n_entries = lg.used = lg.flags = 0;
  SpecSearchHook hook{g.upper_base, &lg, lane};

  CandList<EFR> L;
  CandList<ER> R;
  Counters cnt = {0, 0, 0};
  uint32_t n_reprunes = 0;

  const int l = g.level[q];
  const int l_max = g.meta[kMetaMaxLayer];                        // core.rs:496
  uint32_t ep = (uint32_t)g.meta[kMetaEntry];                     // core.rs:508
  for (int lc = l_max; lc >= 0; --lc) {
    const bool link = lc <= l;
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;               // core.rs:560
    load_q_from_slab<C, S, T>(w, g, q, lane);
    if constexpr (kLookahead) search_layer2_la<EFR, C, S, T>(g, w, lb, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane, hook);   // :513, :524
    else search_layer2<EFR, C, S, T, RowCopy<C>::kDefault>(g, w, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane, hook);
    float s;
    L.get(0, lane, false, ep, s);                                 // :514 / :576
    if (!link) continue;
    const uint32_t n_sel = min((uint32_t)L.len, a.m);             // core.rs:531 (build.cuh header; the host sends ef_construction < m to EXACT)
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      uint32_t e = r * 32 + lane;
      if (e < n_sel) sel[e] = L.id[r] & ~kExpanded;
    }
    __syncwarp();
    view_store(g, lg, q, (uint32_t)lc, sel, n_sel, lane);         // connect_neighbors (core.rs:759-774)
    for (uint32_t i = 0; i < n_sel; ++i) {
      const uint32_t r = sel[i];
      uint32_t len = view_load(g, lg, r, (uint32_t)lc, edit, a.lcap, lane);
      if (len == kEmpty || len + 1 > a.lcap) {
        if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
        continue;
      }
      if (list_find(edit, len, q, lane) < 0) {
        if (lane == 0) edit[len] = q;
        ++len;
      }
      view_store(g, lg, r, (uint32_t)lc, edit, len, lane);
    }
    for (uint32_t i = 0; i < n_sel; ++i) {                        // shrink connections (core.rs:540-574), nearest-first
      const uint32_t e = sel[i];
      const uint32_t n_old = view_load(g, lg, e, (uint32_t)lc, old, a.lcap, lane);
      if (n_old == kEmpty) {
        if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
        continue;
      }
      if (n_old <= cap) continue;                                 // core.rs:561
      load_q_from_slab<C, S, T>(w, g, e, lane);
      reprune_select2v<ER, C, S, T>(g, lg, w, e, (uint32_t)lc, (int)cap, old, n_old, R, cnt, lane, keep_add, tmp, a.lcap);   // :568
      ++n_reprunes;
      uint32_t n_keep, n_add, n_rem;                              // update_node_connections (core.rs:776-822)
      reprune_delta<ER>(R, old, n_old, keep_add, rem, n_keep, n_add, n_rem, lane);
      view_store(g, lg, e, (uint32_t)lc, keep_add, n_keep + n_add, lane);
      for (uint32_t t = 0; t < n_add; ++t) {                      // :793-796 (no cap check on the other side)
        const uint32_t x = keep_add[n_keep + t];
        uint32_t len = view_load(g, lg, x, (uint32_t)lc, edit, a.lcap, lane);
        if (len == kEmpty || len + 1 > a.lcap) {
          if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
          continue;
        }
        if (list_find(edit, len, e, lane) < 0) {
          if (lane == 0) edit[len] = e;
          ++len;
          view_store(g, lg, x, (uint32_t)lc, edit, len, lane);
        }
      }
      for (uint32_t t = 0; t < n_rem; ++t) {                      // :805-816
        const uint32_t x = rem[t];
        const uint32_t len = view_load(g, lg, x, (uint32_t)lc, edit, a.lcap, lane);
        if (len == kEmpty) continue;
        const int p = list_find(edit, len, e, lane);
        if (p < 0) continue;
        list_erase(edit, len, p, lane);
        view_store(g, lg, x, (uint32_t)lc, edit, len - 1, lane);
      }
    }
  }
  __syncwarp();
  uint32_t* wkey = a.wkey + (size_t)slot * a.wmaxe;
  uint32_t* woff = a.woff + (size_t)slot * a.wmaxe;
  for (uint32_t i = lane; i < lg.n_entries; i += 32) wkey[i] = lg.wkey_s[i], woff[i] = lg.woff_s[i];
  if (lane == 0) {
    hdr[kSpecSnap] = a.frontier;
    hdr[kSpecNode] = q;
    hdr[kSpecReads] = lg.n_reads;
    hdr[kSpecEntries] = lg.n_entries;
    hdr[kSpecFlags] = lg.flags;
    hdr[kSpecDist] = cnt.n_dist;
    hdr[kSpecReprunes] = n_reprunes;
    uint64_t t_out;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_out));
    hdr[kSpecDur] = (uint32_t)(t_out - t_in);
    hdr[kSpecSm] = smid | 0x80000000u;
    __threadfence();
    hdr[kSpecState] = 1u;
    atomicAdd(a.ctl + kSpecExecuted, 1u);
  }
}

}  // namespace hnsw

// ---------------------------------------------------------------- K2
#ifdef HNSW_PLAIN_BUILD_KERNELS  // no distance arithmetic: defined once, in build_host.cu
namespace hnsw {

// One CTA commits the longest valid prefix of the window, in stream order.
__global__ void __launch_bounds__(256) spec_commit_kernel(Graph g, SpecArgs a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
  __shared__ uint32_t s_need;
  uint32_t committed = 0, reason = 0, dist = 0, reprunes = 0;
  for (uint32_t q = a.frontier; q < a.frontier + a.count; ++q) {
    const uint32_t slot = q & (a.ring - 1);
    uint32_t* hdr = a.hdr + (size_t)slot * kSpecHdrWords;
    const uint32_t state = __ldcg(hdr + kSpecState), node = __ldcg(hdr + kSpecNode), snap = __ldcg(hdr + kSpecSnap);
    const uint32_t n_reads = __ldcg(hdr + kSpecReads), n_entries = __ldcg(hdr + kSpecEntries), flags = __ldcg(hdr + kSpecFlags);
    if (state != 1u || node != q) {
      reason = 1;
      break;
    }
    if (flags & kSpecWrOverflow) {
      reason = 2;
      break;
    }
    // valid iff no row the insert looked at was written after its snapshot
    const uint32_t* rd = a.rd + (size_t)slot * a.rcap;
    int bad = ((flags & kSpecRdOverflow) && snap != q) ? 1 : 0;
    for (uint32_t i = tid; i < n_reads; i += blockDim.x) bad |= spec_ver(a, __ldcg(rd + i)) > snap ? 1 : 0;
    if (tid == 0) s_need = 0;
    bad = __syncthreads_or(bad);
    if (bad) {
      if (tid == 0) {
        hdr[kSpecState] = 0u;
        atomicAdd(a.ctl + kSpecDistWasted, __ldcg(hdr + kSpecDist));
      }
      reason = 3;
      break;
    }
    const uint32_t* wkey = a.wkey + (size_t)slot * a.wmaxe;
    const uint32_t* woff = a.woff + (size_t)slot * a.wmaxe;
    const uint32_t* wdata = a.wdata + (size_t)slot * a.wcap;
    // overflow rows the copy can allocate at most (chains already in place are not counted: an upper bound)
    uint32_t need = 0;
    for (uint32_t e = tid; e < n_entries; e += blockDim.x)
      if (__ldcg(wkey + e) != kEmpty) {
        const uint32_t len = __ldcg(wdata + __ldcg(woff + e) + 1);
        if (len > g.W) need += (len - g.W + kPoolIds - 1) / kPoolIds;
      }
    if (need) atomicAdd(&s_need, need);
    __syncthreads();
    if ((uint32_t)__ldcg(g.meta + kMetaPoolUsed) + s_need > g.pool_cap) {
      reason = 4;
      break;
    }
    for (uint32_t e = warp; e < n_entries; e += warps) {
      const uint32_t key = __ldcg(wkey + e);
      if (key == kEmpty) continue;
      const uint32_t* p = wdata + __ldcg(woff + e);
      const uint32_t len = __ldcg(p + 1);
      uint32_t *row, *ovf;
      if (key & 0x80000000u) {
        const uint32_t r = key & 0x7FFFFFFFu;
        row = g.adjU + (size_t)r * g.W, ovf = g.ovfU + r;
      } else {
        row = g.adj0 + (size_t)key * g.W, ovf = g.ovf0 + key;
      }
      list_store(g, row, ovf, p + 2, len, lane);
      if (lane == 0) {
        if (key & 0x80000000u) a.verU[key & 0x7FFFFFFFu] = q + 1;
        else a.ver0[key] = q + 1;
      }
    }
    if (tid == 0) {
      const int l = g.level[q];
      if (l > g.meta[kMetaMaxLayer]) {                            // core.rs:587-593 (the host ends the window at such a node)
        g.meta[kMetaMaxLayer] = l;
        g.meta[kMetaEntry] = (int32_t)q;
      }
      hdr[kSpecState] = 0u;                                       // the slot is free
    }
    dist += __ldcg(hdr + kSpecDist);
    reprunes += __ldcg(hdr + kSpecReprunes);
    ++committed;
    __threadfence();
    __syncthreads();
  }
  if (tid == 0) {
    a.ctl[kSpecCommitted] = committed;
    a.ctl[kSpecReason] = reason;
    a.ctl[kSpecDistEvals] += dist;
    a.ctl[kSpecReprunesDone] += reprunes;
    a.ctl[kSpecPoolUsed] = (uint32_t)g.meta[kMetaPoolUsed];
    a.ctl[kSpecMaxLayer] = (uint32_t)g.meta[kMetaMaxLayer];
    a.ctl[kSpecEntry] = (uint32_t)g.meta[kMetaEntry];
    a.ctl[kSpecError] = (uint32_t)g.meta[kMetaError];
  }
}

}  // namespace hnsw
#endif  // HNSW_PLAIN_BUILD_KERNELS

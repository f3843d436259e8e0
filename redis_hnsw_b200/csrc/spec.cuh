// SPEC builder: the NODE.ADD stream of the reference (core.rs:383-412, 489-599) executed speculatively in windows and
// committed strictly in stream order, so that the graph is the sequential one (list for list) while the long part of
// an insert — the ef_construction search, ~220 dependent hops — runs for many inserts at once.
//
//   K1  spec_exec_kernel     one warp per insert of the window [f, f + B).  The whole insert (searches, select, connect,
//                            re-selection of over-full rows) runs against the graph as it stands — every insert < f is
//                            committed, nothing else is — WITHOUT writing to it: new row contents go to a private write
//                            log (read back by the insert itself: read-your-own-writes), and every graph row the insert
//                            looked at goes to a read log.
//   K2  spec_commit_kernel   one CTA walks the window in order.  Insert q commits iff nothing it DEPENDED on was written
//                            since its snapshot (row stamps: ver[row] = 1 + id of the last insert that wrote the row).  A
//                            valid insert took every decision the sequential execution would have taken, and the insert is a
//                            deterministic function of those decisions, so its write log IS the sequential result; it is
//                            applied to the graph and the rows are stamped.  The walk stops at the first invalid insert,
//                            which is re-executed by the next K1 (now as the head of its window, where it cannot fail).
//   Inserts that were executed but not reached keep their logs; the next K1 re-validates them and only re-executes the
//   ones that lost a dependency.
//
// What "depended on" means (a.fine; a.fine = 0 is row-level validation: any write to any row the insert touched):
//   search read   (kind 1)  a row expanded by search_level with the admission threshold of that moment (core.rs:657) and
//                           the ids it held.  A later write to the row matters only if an id that came or went could have
//                           been admitted: sim(q, id) above the threshold (the threshold only rises during a search).
//   sweep read    (kind 2)  a row swept by a re-selection of e (core.rs:698-721), with the sim of the last selected id.
//   strict read   (kind 0)  the row a re-selection replaces: its whole content decides the outcome.
//   length bound  (kind 6)  a row that got the new node appended and still fit its cap (core.rs:561): the row may have
//                           grown meanwhile, as long as the check still says "fits".
//   Appends and removals of one id (core.rs:137-152) are logged as OPERATIONS besides the new row content; at commit a
//   row that was not written since the snapshot takes the content, a row that was takes the operations, applied to what
//   it holds now — which is what the sequential execution does.  (Whether the id is present cannot have changed: it is
//   only ever added or removed together with a write to the strict row of the same re-selection.)
//   tools/sim_spec_build.cpp (SIM_VERIFY) replays these rules on the CPU: every accepted execution equals the sequential one.
//
// Exactness does not rest on any probability: the sims of validation are compared with a margin far above f32 rounding,
// so an insert is committed only when every decision it took is provably the sequential one.  What speculation buys is
// measured (tools/sim_spec_build.cpp): a round commits a prefix of O(sqrt N)-ish inserts.
#pragma once
#include "build2.cuh"
#include "search_la.cuh"

namespace hnsw {

static_assert(!kLookaheadInBuilders, "search_layer2_la does not report row ids to the search hook: the SPEC read log needs them");

enum SpecHdr : int {
  kSpecState = 0,     // 0 = needs execution, 1 = executed, 2 = only the levels above 0 are done (checkpoint, see K1),
                      // 3 = suspended between two re-selections of level 0 (time budget, see K1)
  kSpecSnap = 1,      // every insert with id < snap was committed when the level-0 part of the logs was made
  kSpecNode = 2,
  kSpecReads = 3,
  kSpecEntries = 4,
  kSpecFlags = 5,     // 1 = read log overflowed (valid only as the head of a window), 2 = write log overflowed (unusable)
  kSpecDist = 6,
  kSpecReprunes = 7,
  kSpecT0 = 8,        // diagnostics: %globaltimer (low word, ns) when the warp entered K1,
  kSpecDur = 9,       //              ns it spent there,
  kSpecSm = 10,       //              SM it ran on | 0x80000000 when it executed (not just validated)
  kSpecOps = 11,      // operations logged (fine validation)
  kSpecSnapU = 12,    // the same snapshot id for the part above level 0 (rows with bit 31 set in their key)
  kSpecCpReads = 13,  // checkpoint = the logs and counters as they stood when level 0 began: read records,
  kSpecCpRused = 14,  //   id words behind them,
  kSpecCpEntries = 15,  // write-log entries,
  kSpecCpUsed = 16,   //   words behind them,
  kSpecCpOps = 17,    //   operations,
  kSpecCpEp = 18,     //   entry point of the level-0 search,
  kSpecCpDist = 19,
  kSpecCpReprunes = 20,
  kSpecTSearch = 21,  // diagnostics: ns of this execution inside search_level,
  kSpecTSelect = 22,  //              ns inside the re-selections (sweep + top-m)
  kSpecRused = 23,    // id words behind the read records  } only needed to continue a suspended execution
  kSpecUsed = 24,     // words of the write log            }
  kSpecSusI = 25,     // position in the selected list at which a suspended execution continues
  kSpecSusSel = 26,   // number of selected neighbours (ssel)
  kSpecHdrWords = 32,
};
constexpr uint32_t kSpecRdOverflow = 1, kSpecWrOverflow = 2, kSpecOpOverflow = 4;
// read kinds (word y of a read record = kind | len << 3)
constexpr uint32_t kRdStrict = 0, kRdSearch = 1, kRdSweep = 2, kRdLenBound = 6;

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
  kSpecOpRows = 10,    // rows committed as operations on newer content (accumulates)
  kSpecPrepared = 11,  // checkpoints made behind the window (accumulates)
  kSpecSuspended = 12, // executions stopped by the time budget (accumulates)
  kSpecCtlWords = 16,
};

struct SpecArgs {
  uint32_t frontier;   // first insert of the window; every id below is committed
  uint32_t count;      // window size (K1 may be launched with more blocks: nodes behind the window only prepare, see K1)
  uint32_t ring;       // slots (power of two); slot = id & (ring - 1)
  uint32_t m, cap0, capU, efc, lcap, vis_slots;
  uint32_t rcap, wcap, wmaxe;
  uint32_t rmax, ocap; // read records / operations per slot
  uint32_t fine;       // 1 = dependency-level validation (see header), 0 = row-level
  uint32_t budget_ns;  // an execution that has run this long stops before its next re-selection and continues in the next K1 (0 = never)
  uint32_t* ssel;      // [ring][m rounded up to 32]  selected neighbours of a suspended execution
  uint32_t* hdr;       // [ring][kSpecHdrWords]
  uint4* rdh;          // [ring][rmax]   read records {row key, kind | len << 3, query node | base length, threshold | growth}
  uint32_t* rdo;       // [ring][rmax]   offset of the record's ids in rd
  uint32_t* rd;        // [ring][rcap]   ids the rows held when they were read (kinds 1, 2)
  uint32_t* okey;      // [ring][ocap]   operations in program order: row key,
  uint32_t* oval;      // [ring][ocap]   id to append, or id | 0x80000000 to remove
  uint32_t* wkey;      // [ring][wmaxe]  row key of entry e (kEmpty = dead)
  uint32_t* woff;      // [ring][wmaxe]  word offset of entry e in wdata: {reserved, len, ids...}
  uint32_t* wbase;     // [ring][wmaxe]  length of the graph's row when entry e was made
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
  uint4* rdh;          // global
  uint32_t* rdo;       // global
  uint32_t* rd;        // global
  uint32_t* wdata;     // global
  uint32_t* okey;      // global
  uint32_t* oval;      // global
  uint32_t* wkey_s;    // shared [wmaxe]
  uint32_t* woff_s;    // shared [wmaxe]
  uint32_t* wbase_s;   // shared [wmaxe]  length of the row in the graph when the entry was made
  uint32_t rmax, rcap, wcap, wmaxe, ocap, fine, self;
  uint32_t n_reads, rused, n_entries, used, n_ops, flags;  // warp-uniform
  uint32_t cur_key, cur_kind, cur_aux, cur_off;             // record being written (begin .. end)
  float cur_thr;
  bool cur_open;

  // a record without ids (strict read, length bound)
  __device__ __forceinline__ void read_plain(uint32_t key, uint32_t kind, uint32_t aux, uint32_t w, int lane) {
    if (n_reads < rmax) {
      if (lane == 0) rdh[n_reads] = make_uint4(key, kind, aux, w), rdo[n_reads] = 0;
      ++n_reads;
    } else {
      flags |= kSpecRdOverflow;
    }
  }
  // a record with the ids of the row: begin, ids (chunks of <= 32, kEmpty = no id), end
  __device__ __forceinline__ void begin(uint32_t key, uint32_t kind, uint32_t aux, float thr) {
    cur_key = key, cur_kind = kind, cur_aux = aux, cur_thr = thr, cur_off = rused;
    cur_open = n_reads < rmax;
    if (!cur_open) flags |= kSpecRdOverflow;
  }
  __device__ __forceinline__ void ids(uint32_t nb, int lane) {
    if (!cur_open || !fine) return;
    const uint32_t mask = __ballot_sync(kFull, nb != kEmpty);
    const uint32_t n = __popc(mask);
    if (rused + n > rcap) {
      flags |= kSpecRdOverflow;
      cur_open = false;
      return;
    }
    if (nb != kEmpty) rd[rused + __popc(mask & ((1u << lane) - 1u))] = nb;
    rused += n;
  }
  __device__ __forceinline__ void end(int lane) {
    if (!cur_open) return;
    if (lane == 0) rdh[n_reads] = make_uint4(cur_key, cur_kind | ((rused - cur_off) << 3), cur_aux, __float_as_uint(cur_thr)), rdo[n_reads] = cur_off;
    ++n_reads;
    cur_open = false;
  }
  // the row of (node, level) as the graph holds it -> one record (K1 never writes the graph: the same at any time)
  __device__ __forceinline__ void read_row(const Graph& g, uint32_t node, uint32_t level, uint32_t key, uint32_t kind, uint32_t aux,
                                           float thr, int lane) {
    begin(key, kind, aux, thr);
    uint32_t* ovf;
    const uint32_t* row = row_ptr(g, node, level, &ovf);
    if (row && fine) {
      bool more = true;
      for (uint32_t c = 0; c < g.W / 32 && more; ++c) {
        const uint32_t nb = __ldcg(row + c * 32 + lane);
        more = __shfl_sync(kFull, nb, 31) != kEmpty;
        ids(nb, lane);
      }
      uint32_t link = more ? __ldcg(ovf) : kEmpty;
      while (link != kEmpty) {
        uint32_t nb = __ldcg(g.pool + (size_t)link * 32 + lane);
        link = __shfl_sync(kFull, nb, 31);
        if (lane == 31) nb = kEmpty;
        ids(nb, lane);
      }
    }
    end(lane);
  }
  // thresholds of the sweep records [from, n_reads) are known once the re-selection is over
  __device__ __forceinline__ void patch_thr(uint32_t from, float thr, int lane) {
    __syncwarp();
    for (uint32_t i = from + lane; i < n_reads; i += 32)
      if ((rdh[i].y & 7u) == kRdSweep) rdh[i].w = __float_as_uint(thr);
    __syncwarp();
  }
  __device__ __forceinline__ void op(uint32_t key, uint32_t val, int lane) {
    if (n_ops < ocap) {
      if (lane == 0) okey[n_ops] = key, oval[n_ops] = val;
      ++n_ops;
    } else {
      flags |= kSpecOpOverflow;
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

// hook of search_layer2: every expanded row is a read (core.rs:642-646), with the threshold of the moment and its ids
struct SpecSearchHook {
  const uint32_t* upper_base;
  SpecLog* lg;
  int lane;
  __device__ __forceinline__ void expand(uint32_t node, uint32_t level, float thr) const {
    const uint32_t key = level == 0 ? node : (0x80000000u | (upper_base[node] + level - 1));   // row_key()
    lg->begin(key, lg->fine ? kRdSearch : kRdStrict, lg->self, thr);
  }
  __device__ __forceinline__ void ids(uint32_t nb) const { lg->ids(nb, lane); }
  __device__ __forceinline__ void done() const { lg->end(lane); }
};

// the insert's view of the adjacency list of (node, level): its own latest version, else the graph's.  `ent` = the entry
// of the write log the list came from, or -1.  Row-level validation logs every row that comes from the graph; the
// dependency-level one leaves it to the caller to say what the row was read FOR.
__device__ __forceinline__ uint32_t view_load(const Graph& g, SpecLog& lg, uint32_t node, uint32_t level, uint32_t* buf,
                                              uint32_t lcap, int lane, int* ent = nullptr) {
  const uint32_t key = row_key(g, node, level);
  const int e = lg.find(key, lane);
  if (ent) *ent = e;
  if (e >= 0) {
    const uint32_t* p = lg.wdata + lg.woff_s[e];
    const uint32_t len = p[1];
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) buf[i] = p[2 + i];
    __syncwarp();
    return len;
  }
  if (!lg.fine) lg.read_plain(key, kRdStrict, 0, 0, lane);
  uint32_t* ovf;
  const uint32_t* row = row_ptr(g, node, level, &ovf);
  if (!row) return 0;
  return list_load(g, row, ovf, buf, lcap, lane);
}

// new content of (node, level) -> write log (in place when the row already has an entry that is large enough).
// `base_len` = what the graph's row held (only used when the entry is new).
__device__ __forceinline__ void view_store(const Graph& g, SpecLog& lg, uint32_t node, uint32_t level, const uint32_t* buf,
                                           uint32_t len, uint32_t base_len, int lane) {
  const uint32_t key = row_key(g, node, level);
  __syncwarp();
  int e = lg.find(key, lane);
  uint32_t off;
  if (e >= 0 && lg.wdata[lg.woff_s[e]] >= len) {
    off = lg.woff_s[e];
  } else {
    const uint32_t base = e >= 0 ? lg.wbase_s[e] : base_len;
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
      lg.wbase_s[lg.n_entries] = base;
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
  const uint32_t first_sweep = lg.n_reads;
  // extend_candidates (core.rs:698-721): the rows of the old neighbours, in order.  A row comes from the insert's own write
  // log when it has an entry there, else from the graph.  The lookups are done for 32 rows at once (one lane per row) and
  // the first 32 ids of 8 rows are loaded back to back, so that the sweep waits for memory once per 8 rows, not per row
  // (measured before: ~110 us per re-selection, most of it 33 dependent row reads).
  uint32_t* s_key = tmp + 256;                                    // per row of the group of 32: key (kEmpty = the node has no
  uint32_t* s_ent = tmp + 288;                                    // row at this level, core.rs:642), log entry or kEmpty,
  uint32_t* s_len = tmp + 320;                                    // length and offset of the entry's content
  uint32_t* s_off = tmp + 352;
  for (uint32_t j0 = 0; j0 < n_old; j0 += 32) {
    {
      const uint32_t id_l = j0 + lane < n_old ? old[j0 + lane] : kEmpty;
      uint32_t key_l = kEmpty, len_l = 0, off_l = 0, ent_l = kEmpty;
      if (id_l != kEmpty && !(level != 0 && (g.upper_base[id_l] == kEmpty || (int32_t)level > g.level[id_l]))) {
        key_l = row_key(g, id_l, level);
        for (uint32_t i = 0; i < lg.n_entries; ++i)
          if (lg.wkey_s[i] == key_l) ent_l = i;                   // (superseded entries hold kEmpty)
        if (ent_l != kEmpty) off_l = lg.woff_s[ent_l], len_l = lg.wdata[off_l + 1];
      }
      __syncwarp();
      s_key[lane] = key_l, s_ent[lane] = ent_l, s_len[lane] = len_l, s_off[lane] = off_l;
      __syncwarp();
    }
    const uint32_t nrows = min(32u, n_old - j0);
    for (uint32_t r0 = 0; r0 < nrows; r0 += 8) {
      {
        uint32_t nb8[8];                                          // 8 independent loads in flight, then parked in `tmp`
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t src = min(r0 + r, 31u);
          const uint32_t key = s_key[src];
          nb8[r] = kEmpty;
          if (r0 + r < nrows && key != kEmpty) {
            if (s_ent[src] != kEmpty) {
              if ((uint32_t)lane < s_len[src]) nb8[r] = lg.wdata[s_off[src] + 2 + lane];
            } else {
              const uint32_t* row = (key & 0x80000000u) ? g.adjU + (size_t)(key & 0x7FFFFFFFu) * g.W : g.adj0 + (size_t)key * g.W;
              nb8[r] = __ldcg(row + lane);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) tmp[r * 32 + lane] = nb8[r];
        __syncwarp();
      }
#pragma unroll 1
      for (uint32_t r = 0; r < 8 && r0 + r < nrows; ++r) {
        const uint32_t src = r0 + r;
        const uint32_t key = s_key[src];
        if (key == kEmpty) continue;
        const uint32_t id = old[j0 + src];
        // what the sweep depends on is the row in the GRAPH (the insert's own edits on top of it are the same in any
        // order); the new node's own rows have no earlier writer
        const bool log_it = lg.fine && id != lg.self;
        if (s_ent[src] != kEmpty) {
          const uint32_t len = s_len[src], off = s_off[src];
          if (log_it) lg.read_row(g, id, level, key, kRdSweep, e, 0.f, lane);
          feed(tmp[r * 32 + lane]);
          for (uint32_t c = 32; c < len; c += 32) feed(c + lane < len ? lg.wdata[off + 2 + c + lane] : kEmpty);
          continue;
        }
        if (!lg.fine) lg.read_plain(key, kRdStrict, 0, 0, lane);
        const uint32_t* row = (key & 0x80000000u) ? g.adjU + (size_t)(key & 0x7FFFFFFFu) * g.W : g.adj0 + (size_t)key * g.W;
        const uint32_t* ovf = (key & 0x80000000u) ? g.ovfU + (key & 0x7FFFFFFFu) : g.ovf0 + key;
        if (log_it) lg.begin(key, kRdSweep, e, 0.f);
        uint32_t nb = tmp[r * 32 + lane];
        bool more = __shfl_sync(kFull, nb, 31) != kEmpty;         // rows are compact: an empty tail ends the list
        if (log_it) lg.ids(nb, lane);
        feed(nb);
        for (uint32_t c = 1; c < g.W / 32 && more; ++c) {
          nb = __ldcg(row + c * 32 + lane);
          more = __shfl_sync(kFull, nb, 31) != kEmpty;
          if (log_it) lg.ids(nb, lane);
          feed(nb);
        }
        uint32_t link = more ? __ldcg(ovf) : kEmpty;
        while (link != kEmpty) {                                  // overflow rows: as list_load walks them
          nb = __ldcg(g.pool + (size_t)link * 32 + lane);
          const uint32_t next = __shfl_sync(kFull, nb, 31);
          if (lane >= kPoolIds) nb = kEmpty;
          const uint32_t n_valid = __popc(__ballot_sync(kFull, nb != kEmpty));
          if (log_it) lg.ids(nb, lane);
          feed(nb);
          link = n_valid == (uint32_t)kPoolIds ? next : kEmpty;
        }
        if (log_it) lg.end(lane);
      }
      __syncwarp();
    }
  }
  if (np) flush(np);
  L.finish(lane);                                                // wide lists (EFR >= 4): back to the sorted layout
  if (lg.fine) lg.patch_thr(first_sweep, L.worst, lane);         // -inf while fewer than `cap` candidates exist
}

__device__ __forceinline__ uint64_t spec_now() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- validation (K1 for kept logs, K2 before a commit)

// sim of two slab rows, any summation order: only ever compared with a margin (spec_read_conflicts)
__device__ __forceinline__ float spec_sim_loose(const Graph& g, uint32_t a, uint32_t b, int lane) {
  const float4* x = reinterpret_cast<const float4*>(g.vecs + (size_t)a * g.dim);
  const float4* y = reinterpret_cast<const float4*>(g.vecs + (size_t)b * g.dim);
  float acc = 0.f;
  for (uint32_t i = lane; i < g.dim / 4; i += 32) {
    const float4 u = __ldg(x + i), v = __ldg(y + i);
    const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
  return -acc;
}

// One warp: the row of read record `h` was written after the snapshot — could that have changed what the insert did?
// `buf`: `lcap` words of shared memory.  Conservative: true whenever the answer is not a provable no.
__device__ __forceinline__ bool spec_read_conflicts(const Graph& g, const SpecArgs& a, uint4 h, const uint32_t* ids, uint32_t* buf,
                                                    int lane) {
  const uint32_t kind = h.y & 7u, len = h.y >> 3;
  if (kind == kRdStrict || !a.fine) return true;
  const uint32_t *row, *ovf;
  if (h.x & 0x80000000u) {
    const uint32_t r = h.x & 0x7FFFFFFFu;
    row = g.adjU + (size_t)r * g.W, ovf = g.ovfU + r;
  } else {
    row = g.adj0 + (size_t)h.x * g.W, ovf = g.ovf0 + h.x;
  }
  __syncwarp();
  const uint32_t cur = list_load(g, row, ovf, buf, a.lcap, lane);
  if (cur == kEmpty) return true;
  if (kind == kRdLenBound) return cur > h.z + h.w;               // the cap check (core.rs:561) would not say "fits" any more
  const float thr = __uint_as_float(h.w);
  if (!(thr > -CUDART_INF_F)) return true;                       // the list was not full: any id would have been admitted
  // f32 sums of <= 65536 non-negative terms in two different orders differ by far less than 2.5e-4 of their value
  const float bar = thr - 2.5e-4f * fabsf(thr) - 1e-30f;
  const uint32_t qn = h.z;
  bool hit = false;
  for (uint32_t c = 0; c < len && !hit; c += 32) {               // ids that left the row
    const uint32_t id = c + lane < len ? __ldcg(ids + c + lane) : kEmpty;
    bool gone = id != kEmpty && id != qn;
    for (uint32_t j = 0; j < cur && gone; ++j) gone = buf[j] != id;
    uint32_t mask = __ballot_sync(kFull, gone);
    while (mask && !hit) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      hit = spec_sim_loose(g, qn, __shfl_sync(kFull, id, b), lane) > bar;
    }
  }
  for (uint32_t c = 0; c < cur && !hit; c += 32) {               // ids that came
    const uint32_t id = c + lane < cur ? buf[c + lane] : kEmpty;
    bool came = id != kEmpty && id != qn;
    for (uint32_t j = 0; j < len && came; ++j) came = __ldcg(ids + j) != id;
    uint32_t mask = __ballot_sync(kFull, came);
    while (mask && !hit) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      hit = spec_sim_loose(g, qn, __shfl_sync(kFull, id, b), lane) > bar;
    }
  }
  return hit;
}

// rows above level 0 (bit 31 of the key) and level-0 rows may have been read at different snapshots (checkpoint, K1)
__device__ __forceinline__ uint32_t spec_snap_of(uint32_t key, uint32_t snap0, uint32_t snapU) {
  return (key & 0x80000000u) ? snapU : snap0;
}

// One warp walks the read records lo + first, lo + first + step, ... < hi of a slot (32 at a time); true (warp-uniform) =
// a dependency was lost.
// `floor`: writes by inserts below it are known to be harmless (K2: the K1 of the same round checked every slot against
// everything committed before the round).
__device__ __forceinline__ bool spec_reads_conflict(const Graph& g, const SpecArgs& a, uint32_t slot, uint32_t snap0, uint32_t snapU,
                                                    uint32_t lo, uint32_t hi, uint32_t first, uint32_t step, uint32_t* buf, int lane,
                                                    uint32_t floor = 0) {
  const uint4* rdh = a.rdh + (size_t)slot * a.rmax;
  const uint32_t* rdo = a.rdo + (size_t)slot * a.rmax;
  const uint32_t* rd = a.rd + (size_t)slot * a.rcap;
  for (uint32_t i = lo + first; i < hi; i += step) {
    uint4 h = make_uint4(0, 0, 0, 0);
    uint32_t off = 0;
    bool stale = false;
    if (i + lane < hi) {
      h = __ldcg(rdh + i + lane);
      off = __ldcg(rdo + i + lane);
      stale = spec_ver(a, h.x) > max(spec_snap_of(h.x, snap0, snapU), floor);
    }
    if (__any_sync(kFull, stale && ((h.y & 7u) == kRdStrict || !a.fine))) return true;
    uint32_t mask = __ballot_sync(kFull, stale);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint4 hb = make_uint4(__shfl_sync(kFull, h.x, b), __shfl_sync(kFull, h.y, b), __shfl_sync(kFull, h.z, b), __shfl_sync(kFull, h.w, b));
      if (spec_read_conflicts(g, a, hb, rd + __shfl_sync(kFull, off, b), buf, lane)) return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------- K1

template <int EFR, int C, bool SMALL>
__global__ void __launch_bounds__(32, 1) spec_exec_kernel(Graph g, SpecArgs a) {
  constexpr int ER = SMALL ? (EFR < 2 ? EFR : 2) : EFR;
  constexpr int S = ExactStage<C>::S;
  using T = uint32_t;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const uint32_t q = a.frontier + blockIdx.x;
  const uint32_t slot = q & (a.ring - 1);
  uint32_t* hdr = a.hdr + (size_t)slot * kSpecHdrWords;
  uint64_t t_in;
  uint32_t smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (lane == 0) hdr[kSpecT0] = (uint32_t)t_in, hdr[kSpecDur] = 0, hdr[kSpecSm] = smid;

  Warp2<C, S, T> w;
  unsigned char* after = warp2_setup<C, S, T>(w, smem2, a.vis_slots, lane);
  constexpr bool kLookahead = kLookaheadInBuilders && S == 32 && RowCopy<C>::kOk;          // search_la.cuh: second stage for the next hop's rows
  LaBuf<C> lb;
  if constexpr (kLookahead) after = la_setup<C, S, T>(lb, w, after, lane);
  uint32_t* lists = reinterpret_cast<uint32_t*>(after);
  // sel[m] | old[lcap] | keep_add[lcap + W] | rem[lcap] | edit[lcap] | tmp[lcap] | wkey[wmaxe] | woff[wmaxe] | wbase[wmaxe]
  uint32_t* sel = lists;
  uint32_t* old = sel + ((a.m + 31) & ~31u);
  uint32_t* keep_add = old + a.lcap;
  uint32_t* rem = keep_add + a.lcap + g.W;
  uint32_t* edit = rem + a.lcap;
  uint32_t* tmp = edit + a.lcap;

  // Nodes behind the window (blockIdx >= count) that have upper levels only PREPARE: they run the levels above 0 — a second
  // ef_construction search, as long as the level-0 one — and leave a checkpoint, so that the insert costs one search when
  // its turn comes.  Rows above level 0 change 16x less often than level-0 rows, so the checkpoint usually survives.
  const bool far = blockIdx.x >= a.count;
  const int l = g.level[q];
  if (far && l == 0) return;
  const uint32_t state = __ldcg(hdr + kSpecState);
  bool resume = false;                                            // the part above level 0 is done and still valid
  bool relink = false;                                            // ... and so is level 0 up to a re-selection (suspended execution)
  uint32_t snapU = a.frontier, snap0 = a.frontier;
  if (state >= 1u && state <= 3u && __ldcg(hdr + kSpecNode) == q) {
    const uint32_t old_snap0 = __ldcg(hdr + kSpecSnap), flags = __ldcg(hdr + kSpecFlags), cp_reads = __ldcg(hdr + kSpecCpReads);
    const uint32_t n_reads = __ldcg(hdr + kSpecReads);
    snapU = __ldcg(hdr + kSpecSnapU);
    if (flags & kSpecWrOverflow) return;                          // unusable either way: the host runs it through EXACT
    // Warp-uniform control flow on purpose: a per-lane early exit from this loop left the warp split into groups that
    // ran the whole insert below one after the other (measured: 4.1 ms instead of 0.9 ms per execution, r2 call D).
    // A log that overflowed is only good at the head of a window, where nothing can be stale.
    // (Only the level-0 part can have overflowed when the two snapshots differ: a checkpoint is never taken with flags set.)
    const bool whole = !(flags & (kSpecRdOverflow | kSpecOpOverflow)) || old_snap0 == q;
    const bool okU = whole && !spec_reads_conflict(g, a, slot, old_snap0, snapU, 0, cp_reads, 0, 32, tmp, lane);
    if (okU && state == 2u) {
      if (far) return;                                            // prepared, still good
      resume = true;
    } else if (okU) {
      if (!spec_reads_conflict(g, a, slot, old_snap0, snapU, cp_reads, n_reads, 0, 32, tmp, lane)) {
        if (state == 1u || far) return;                           // executed (or suspended, and not wanted yet), still good
        resume = true;
        // The continuation reads at today's frontier but is validated against the old snapshot: sound (a row written in
        // between only looks stale), but a re-selected row that was written in between fails its strict check.  Fine
        // anywhere except at the head of the window, which must commit: there level 0 starts over.
        relink = blockIdx.x != 0;
        // The continuation reads its own write log, whose rows were derived from the graph as it stood at the old
        // snapshot, AND the graph as it stands now (where a sweep logs "the row in the graph" for a row it took from its
        // log, reprune_select2v): the two must be the same rows.  If any row under the write log was written since, level
        // 0 starts over.  (Found by synccheck's timing: a budget of 150 us on 32-d data committed a different graph.)
        if (relink) {
          const uint32_t n_e = __ldcg(hdr + kSpecEntries);
          const uint32_t* wk = a.wkey + (size_t)slot * a.wmaxe;
          for (uint32_t i = 0; i < n_e && relink; i += 32) {
            const uint32_t key = i + lane < n_e ? __ldcg(wk + i + lane) : kEmpty;
            relink = !__any_sync(kFull, key != kEmpty && spec_ver(a, key) > spec_snap_of(key, old_snap0, snapU));
          }
        }
        if (relink) snap0 = old_snap0;                            // what was read so far was read then
      } else {
        if (lane == 0) atomicAdd(a.ctl + kSpecDistWasted, hdr[kSpecDist] - hdr[kSpecCpDist]);
        if (far) {                                                // (a window that shrank): keep the checkpoint
          if (lane == 0) hdr[kSpecState] = 2u;
          return;
        }
        resume = true;
      }
    } else {
      if (lane == 0) atomicAdd(a.ctl + kSpecDistWasted, state == 2u ? hdr[kSpecCpDist] : hdr[kSpecDist]);
      snapU = a.frontier;
    }
    __syncwarp();
  }

  SpecLog lg;
  lg.rdh = a.rdh + (size_t)slot * a.rmax;
  lg.rdo = a.rdo + (size_t)slot * a.rmax;
  lg.rd = a.rd + (size_t)slot * a.rcap;
  lg.wdata = a.wdata + (size_t)slot * a.wcap;
  lg.okey = a.okey + (size_t)slot * a.ocap;
  lg.oval = a.oval + (size_t)slot * a.ocap;
  lg.wkey_s = tmp + a.lcap;
  lg.woff_s = lg.wkey_s + a.wmaxe;
  lg.wbase_s = lg.woff_s + a.wmaxe;
  lg.rmax = a.rmax, lg.rcap = a.rcap, lg.wcap = a.wcap, lg.wmaxe = a.wmaxe, lg.ocap = a.ocap, lg.fine = a.fine, lg.self = q;
  lg.n_reads = lg.rused = lg.n_entries = lg.used = lg.n_ops = lg.flags = 0;
  lg.cur_open = false;
  SpecSearchHook hook{g.upper_base, &lg, lane};
  uint32_t* wkey = a.wkey + (size_t)slot * a.wmaxe;
  uint32_t* woff = a.woff + (size_t)slot * a.wmaxe;
  uint32_t* wbase = a.wbase + (size_t)slot * a.wmaxe;

  CandList<EFR> L;
  CandList<ER> R;
  Counters cnt = {0, 0, 0};
  uint32_t n_reprunes = 0, t_search = 0, t_select = 0;

  int lc_first = g.meta[kMetaMaxLayer];                           // core.rs:496
  uint32_t ep = (uint32_t)g.meta[kMetaEntry];                     // core.rs:508
  uint32_t cp_reads = 0, cp_rused = 0, cp_entries = 0, cp_used = 0, cp_ops = 0, cp_ep = ep, cp_dist = 0, cp_reprunes = 0;
  if (resume) {                                                   // back to the checkpoint: logs truncated, level 0 to do
    cp_reads = __ldcg(hdr + kSpecCpReads), cp_rused = __ldcg(hdr + kSpecCpRused), cp_entries = __ldcg(hdr + kSpecCpEntries);
    cp_used = __ldcg(hdr + kSpecCpUsed), cp_ops = __ldcg(hdr + kSpecCpOps), cp_ep = __ldcg(hdr + kSpecCpEp);
    cp_dist = __ldcg(hdr + kSpecCpDist), cp_reprunes = __ldcg(hdr + kSpecCpReprunes);
    lg.n_reads = cp_reads, lg.rused = cp_rused, lg.n_entries = cp_entries, lg.used = cp_used, lg.n_ops = cp_ops;
    ep = cp_ep, cnt.n_dist = cp_dist, n_reprunes = cp_reprunes;
    if (relink) {                                                 // ... or to where the execution was suspended
      lg.n_reads = __ldcg(hdr + kSpecReads), lg.rused = __ldcg(hdr + kSpecRused), lg.n_entries = __ldcg(hdr + kSpecEntries);
      lg.used = __ldcg(hdr + kSpecUsed), lg.n_ops = __ldcg(hdr + kSpecOps);
      cnt.n_dist = __ldcg(hdr + kSpecDist), n_reprunes = __ldcg(hdr + kSpecReprunes);
      t_search = __ldcg(hdr + kSpecTSearch), t_select = __ldcg(hdr + kSpecTSelect);
    }
    for (uint32_t i = lane; i < lg.n_entries; i += 32)
      lg.wkey_s[i] = __ldcg(wkey + i), lg.woff_s[i] = __ldcg(woff + i), lg.wbase_s[i] = __ldcg(wbase + i);
    __syncwarp();
    lc_first = 0;
  }
  for (int lc = lc_first; lc >= 0; --lc) {
    if (lc == 0 && !resume) {                                     // checkpoint: everything above level 0 is done
      cp_reads = lg.n_reads, cp_rused = lg.rused, cp_entries = lg.n_entries, cp_used = lg.used, cp_ops = lg.n_ops;
      cp_ep = ep, cp_dist = cnt.n_dist, cp_reprunes = n_reprunes;
      if (far && !lg.flags) {
        __syncwarp();
        for (uint32_t i = lane; i < lg.n_entries; i += 32) wkey[i] = lg.wkey_s[i], woff[i] = lg.woff_s[i], wbase[i] = lg.wbase_s[i];
        if (lane == 0) {
          hdr[kSpecNode] = q, hdr[kSpecSnapU] = snapU, hdr[kSpecFlags] = 0;
          hdr[kSpecCpReads] = cp_reads, hdr[kSpecCpRused] = cp_rused, hdr[kSpecCpEntries] = cp_entries, hdr[kSpecCpUsed] = cp_used;
          hdr[kSpecCpOps] = cp_ops, hdr[kSpecCpEp] = cp_ep, hdr[kSpecCpDist] = cp_dist, hdr[kSpecCpReprunes] = cp_reprunes;
          uint64_t t_out;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_out));
          hdr[kSpecDur] = (uint32_t)(t_out - t_in);
          hdr[kSpecSm] = smid | 0x40000000u;
          __threadfence();
          hdr[kSpecState] = 2u;
          atomicAdd(a.ctl + kSpecPrepared, 1u);
        }
        return;
      }
    }
    const bool link = lc <= l;
    const uint32_t cap = lc == 0 ? a.cap0 : a.capU;               // core.rs:560
    uint32_t n_sel, i_first = 0;
    if (relink) {                                                 // (lc == 0) search and connect were done before the suspension
      n_sel = __ldcg(hdr + kSpecSusSel), i_first = __ldcg(hdr + kSpecSusI);
      const uint32_t* ss = a.ssel + (size_t)slot * ((a.m + 31) & ~31u);
      for (uint32_t i = lane; i < n_sel; i += 32) sel[i] = __ldcg(ss + i);
      __syncwarp();
    } else {
    load_q_from_slab<C, S, T>(w, g, q, lane);
    const uint64_t t_s0 = spec_now();
    if constexpr (kLookahead) search_layer2_la<EFR, C, S, T>(g, w, lb, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane, hook);   // :513, :524
    else search_layer2<EFR, C, S, T, RowCopy<C>::kDefault>(g, w, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane, hook);
    t_search += (uint32_t)(spec_now() - t_s0);
    float s;
    L.get(0, lane, false, ep, s);                                 // :514 / :576
    if (!link) continue;
    n_sel = min((uint32_t)L.len, a.m);                            // core.rs:531 (build.cuh header; the host sends ef_construction < m to EXACT)
#pragma unroll
    for (int r = 0; r < EFR; ++r) {
      uint32_t e = r * 32 + lane;
      if (e < n_sel) sel[e] = L.id[r] & ~kExpanded;
    }
    __syncwarp();
    view_store(g, lg, q, (uint32_t)lc, sel, n_sel, 0, lane);      // connect_neighbors (core.rs:759-774)
    for (uint32_t i = 0; i < n_sel; ++i) {
      const uint32_t r = sel[i];
      uint32_t len = view_load(g, lg, r, (uint32_t)lc, edit, a.lcap, lane);
      if (len == kEmpty || len + 1 > a.lcap) {
        if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
        continue;
      }
      const uint32_t base_len = len;
      if (list_find(edit, len, q, lane) < 0) {
        if (lane == 0) edit[len] = q;
        ++len;
        lg.op(row_key(g, r, (uint32_t)lc), q, lane);              // an operation on r, whatever r holds by then
      }
      view_store(g, lg, r, (uint32_t)lc, edit, len, base_len, lane);
    }
    }  // !relink
    for (uint32_t i = i_first; i < n_sel; ++i) {                  // shrink connections (core.rs:540-574), nearest-first
      const uint32_t e = sel[i];
      int ent;
      const uint32_t n_old = view_load(g, lg, e, (uint32_t)lc, old, a.lcap, lane, &ent);
      if (n_old == kEmpty) {
        if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
        continue;
      }
      if (n_old <= cap) {                                         // core.rs:561: fits — and keeps fitting while the row
        if (lg.fine)                                              // in the graph grows by at most cap - n_old ids
          lg.read_plain(row_key(g, e, (uint32_t)lc), kRdLenBound, ent >= 0 ? lg.wbase_s[ent] : n_old, cap - n_old, lane);
        continue;
      }
      // Time budget: a round lasts as long as its slowest execution, and the slow ones are the inserts with many
      // re-selections (~0.1 ms each; 10+ happen).  Past the budget the execution stops HERE and continues in the next K1
      // (never the head of the window: every round must be able to commit).
      if (lc == 0 && a.budget_ns && blockIdx.x != 0 && !lg.flags && spec_now() - t_in > a.budget_ns) {
        __syncwarp();
        for (uint32_t k = lane; k < lg.n_entries; k += 32) wkey[k] = lg.wkey_s[k], woff[k] = lg.woff_s[k], wbase[k] = lg.wbase_s[k];
        uint32_t* ss = a.ssel + (size_t)slot * ((a.m + 31) & ~31u);
        for (uint32_t k = lane; k < n_sel; k += 32) ss[k] = sel[k];
        if (lane == 0) {
          hdr[kSpecSnap] = snap0, hdr[kSpecSnapU] = snapU, hdr[kSpecNode] = q;
          hdr[kSpecCpReads] = cp_reads, hdr[kSpecCpRused] = cp_rused, hdr[kSpecCpEntries] = cp_entries, hdr[kSpecCpUsed] = cp_used;
          hdr[kSpecCpOps] = cp_ops, hdr[kSpecCpEp] = cp_ep, hdr[kSpecCpDist] = cp_dist, hdr[kSpecCpReprunes] = cp_reprunes;
          hdr[kSpecReads] = lg.n_reads, hdr[kSpecRused] = lg.rused, hdr[kSpecEntries] = lg.n_entries, hdr[kSpecUsed] = lg.used;
          hdr[kSpecOps] = lg.n_ops, hdr[kSpecFlags] = 0, hdr[kSpecDist] = cnt.n_dist, hdr[kSpecReprunes] = n_reprunes;
          hdr[kSpecSusI] = i, hdr[kSpecSusSel] = n_sel;
          hdr[kSpecTSearch] = t_search, hdr[kSpecTSelect] = t_select;
          hdr[kSpecDur] = (uint32_t)(spec_now() - t_in);
          hdr[kSpecSm] = smid | 0x20000000u;
          __threadfence();
          hdr[kSpecState] = 3u;
          atomicAdd(a.ctl + kSpecSuspended, 1u);
        }
        return;
      }
      if (lg.fine) lg.read_plain(row_key(g, e, (uint32_t)lc), kRdStrict, 0, 0, lane);   // re-selected: its content decides
      load_q_from_slab<C, S, T>(w, g, e, lane);
      const uint64_t t_r0 = spec_now();
      reprune_select2v<ER, C, S, T>(g, lg, w, e, (uint32_t)lc, (int)cap, old, n_old, R, cnt, lane, keep_add, tmp, a.lcap);   // :568
      t_select += (uint32_t)(spec_now() - t_r0);
      ++n_reprunes;
      uint32_t n_keep, n_add, n_rem;                              // update_node_connections (core.rs:776-822)
      reprune_delta<ER>(R, old, n_old, keep_add, rem, n_keep, n_add, n_rem, lane);
      view_store(g, lg, e, (uint32_t)lc, keep_add, n_keep + n_add, n_old, lane);
      for (uint32_t t = 0; t < n_add; ++t) {                      // :793-796 (no cap check on the other side)
        const uint32_t x = keep_add[n_keep + t];
        uint32_t len = view_load(g, lg, x, (uint32_t)lc, edit, a.lcap, lane);
        if (len == kEmpty || len + 1 > a.lcap) {
          if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
          continue;
        }
        if (list_find(edit, len, e, lane) < 0) {
          if (lane == 0) edit[len] = e;
          lg.op(row_key(g, x, (uint32_t)lc), e, lane);
          view_store(g, lg, x, (uint32_t)lc, edit, len + 1, len, lane);
        }
      }
      for (uint32_t t = 0; t < n_rem; ++t) {                      // :805-816
        const uint32_t x = rem[t];
        const uint32_t len = view_load(g, lg, x, (uint32_t)lc, edit, a.lcap, lane);
        if (len == kEmpty) continue;
        const int p = list_find(edit, len, e, lane);
        if (p < 0) continue;
        list_erase(edit, len, p, lane);
        lg.op(row_key(g, x, (uint32_t)lc), e | 0x80000000u, lane);
        view_store(g, lg, x, (uint32_t)lc, edit, len - 1, len, lane);
      }
    }
  }
  __syncwarp();
  for (uint32_t i = lane; i < lg.n_entries; i += 32) wkey[i] = lg.wkey_s[i], woff[i] = lg.woff_s[i], wbase[i] = lg.wbase_s[i];
  if (lane == 0) {
    hdr[kSpecSnap] = snap0;
    hdr[kSpecSnapU] = snapU;
    hdr[kSpecNode] = q;
    hdr[kSpecCpReads] = cp_reads, hdr[kSpecCpRused] = cp_rused, hdr[kSpecCpEntries] = cp_entries, hdr[kSpecCpUsed] = cp_used;
    hdr[kSpecCpOps] = cp_ops, hdr[kSpecCpEp] = cp_ep, hdr[kSpecCpDist] = cp_dist, hdr[kSpecCpReprunes] = cp_reprunes;
    hdr[kSpecReads] = lg.n_reads;
    hdr[kSpecEntries] = lg.n_entries;
    hdr[kSpecFlags] = lg.flags;
    hdr[kSpecOps] = lg.n_ops;
    hdr[kSpecDist] = cnt.n_dist;
    hdr[kSpecReprunes] = n_reprunes;
    uint64_t t_out;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_out));
    hdr[kSpecDur] = (uint32_t)(t_out - t_in);
    hdr[kSpecSm] = smid | 0x80000000u;
    hdr[kSpecTSearch] = t_search, hdr[kSpecTSelect] = t_select;
    __threadfence();
    hdr[kSpecState] = 1u;
    atomicAdd(a.ctl + kSpecExecuted, 1u);
  }
}

}  // namespace hnsw

// ---------------------------------------------------------------- K2
#ifdef HNSW_PLAIN_BUILD_KERNELS  // no distance arithmetic: defined once, in build_host.cu
namespace hnsw {

// One CTA commits the longest valid prefix of the window, in stream order.  Dynamic shared memory: `lcap` words per warp.
__global__ void __launch_bounds__(512) spec_commit_kernel(Graph g, SpecArgs a) {
  extern __shared__ uint32_t s_bufs[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
  uint32_t* buf = s_bufs + (size_t)warp * a.lcap;
  __shared__ uint32_t s_need;
  uint32_t committed = 0, reason = 0, dist = 0, reprunes = 0;
  for (uint32_t q = a.frontier; q < a.frontier + a.count; ++q) {
    const uint32_t slot = q & (a.ring - 1);
    uint32_t* hdr = a.hdr + (size_t)slot * kSpecHdrWords;
    const uint32_t state = __ldcg(hdr + kSpecState), node = __ldcg(hdr + kSpecNode), snap = __ldcg(hdr + kSpecSnap);
    const uint32_t n_reads = __ldcg(hdr + kSpecReads), n_entries = __ldcg(hdr + kSpecEntries), flags = __ldcg(hdr + kSpecFlags);
    const uint32_t n_ops = __ldcg(hdr + kSpecOps), snapU = __ldcg(hdr + kSpecSnapU);
    if (state != 1u || node != q) {
      reason = 1;
      break;
    }
    if (flags & kSpecWrOverflow) {
      reason = 2;
      break;
    }
    const uint32_t* wkey = a.wkey + (size_t)slot * a.wmaxe;
    const uint32_t* woff = a.woff + (size_t)slot * a.wmaxe;
    const uint32_t* wdata = a.wdata + (size_t)slot * a.wcap;
    // valid iff nothing the insert depended on was written after its snapshot (the warps share the read records)
    // (a log that overflowed is only good at the head of a window, where nothing can be stale)
    int bad = ((flags & (kSpecRdOverflow | kSpecOpOverflow)) && snap != q) ? 1 : 0;
    if (!bad) bad = spec_reads_conflict(g, a, slot, snap, snapU, 0, n_reads, (uint32_t)warp * 32, (uint32_t)warps * 32, buf, lane, a.frontier) ? 1 : 0;
    // overflow rows the commit can allocate at most (chains already in place are not counted: an upper bound; a row that
    // takes operations instead of content grows by a few ids over what it holds)
    uint32_t need = 0;
    for (uint32_t e = tid; e < n_entries; e += blockDim.x) {
      const uint32_t key = __ldcg(wkey + e);
      if (key == kEmpty) continue;
      const uint32_t len = __ldcg(wdata + __ldcg(woff + e) + 1);
      if (len > g.W) need += (len - g.W + kPoolIds - 1) / kPoolIds;
      if (spec_ver(a, key) > spec_snap_of(key, snap, snapU)) need += 2;
    }
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
    if (need) atomicAdd(&s_need, need);
    __syncthreads();
    if ((uint32_t)__ldcg(g.meta + kMetaPoolUsed) + s_need > g.pool_cap) {
      reason = 4;
      break;
    }
    const uint32_t* okey = a.okey + (size_t)slot * a.ocap;
    const uint32_t* oval = a.oval + (size_t)slot * a.ocap;
    for (uint32_t e = warp; e < n_entries; e += warps) {
      const uint32_t key = __ldcg(wkey + e);
      if (key == kEmpty) continue;
      const uint32_t* p = wdata + __ldcg(woff + e);
      uint32_t *row, *ovf;
      if (key & 0x80000000u) {
        const uint32_t r = key & 0x7FFFFFFFu;
        row = g.adjU + (size_t)r * g.W, ovf = g.ovfU + r;
      } else {
        row = g.adj0 + (size_t)key * g.W, ovf = g.ovf0 + key;
      }
      if (spec_ver(a, key) <= spec_snap_of(key, snap, snapU)) {
        list_store(g, row, ovf, p + 2, __ldcg(p + 1), lane);      // the row the insert saw is the row that is there
      } else {
        // written since the snapshot by inserts this one did not depend on: its appends / removals apply to what the row
        // holds now (add_neighbor / remove_neighbor, core.rs:137-152), in program order
        __syncwarp();
        uint32_t len = list_load(g, row, ovf, buf, a.lcap, lane);
        bool ok = len != kEmpty;
        for (uint32_t i = 0; i < n_ops && ok; i += 32) {
          uint32_t mask = __ballot_sync(kFull, i + lane < n_ops && __ldcg(okey + i + lane) == key);
          while (mask && ok) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t v = __ldcg(oval + i + b), id = v & 0x7FFFFFFFu;
            const int pos = list_find(buf, len, id, lane);
            if (v & 0x80000000u) {
              if (pos >= 0) list_erase(buf, len, pos, lane), --len;
            } else if (pos < 0) {
              if (len + 1 > a.lcap) {
                ok = false;
              } else {
                if (lane == 0) buf[len] = id;
                ++len;
              }
            }
            __syncwarp();
          }
        }
        if (ok) list_store(g, row, ovf, buf, len, lane);
        else if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(g.meta + kMetaError), (unsigned int)kErrListTooLong);
        if (lane == 0) atomicAdd(a.ctl + kSpecOpRows, 1u);
      }
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

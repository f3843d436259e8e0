// The index object behind the C ABI (host side).  Device layout: see common.cuh.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "launch.cuh"

namespace hnsw {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
};

struct HostRows {
  std::vector<uint32_t> adj0, ovf0, adjU, ovfU, pool;
};

struct Index {
  // parameters (core.rs:303-312)
  int device = 0;
  uint32_t dim = 0, m = 0, m_max = 0, m_max_0 = 0, ef_construction = 0;
  double level_mult = 0;
  int dist_mode = 0, vecw = 1, kind = 0;
  int num_sms = 148;
  size_t max_smem = 0;

  // device state
  Graph g{};
  uint64_t cap_nodes = 0, cap_upper = 0;
  cudaStream_t stream = nullptr;

  // host mirrors (core.rs:313-317)
  uint64_t n_ids = 0, node_count = 0, upper_used = 0;
  uint32_t pool_used = 0;
  int32_t entry = -1, max_layer = 0, device_error = 0;
  std::vector<int32_t> h_level;
  std::vector<uint32_t> h_upper_base;
  std::vector<uint32_t> touched;
  uint64_t rng_state = 0x9E3779B97F4A7C15ull;
  uint64_t build_stats[4] = {0, 0, 0, 0};
  // [4] wl dropped [5] re-prunes skipped [6] edges refused [7] spec rounds [8] spec executions [9] spec wasted evals
  // [10] spec exact fallbacks [11] spec largest window [12] spec rows committed as operations   (hnsw_index_build_stats_ex)
  uint64_t build_stats_ex[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};

  // options / adaptive state
  uint32_t opt_vis_slots = 0;
  int opt_ctas_per_sm = 0, opt_block = 0;
  int opt_search_impl = 0;        // 0 auto, 1 = register-staged kernel (search.cuh), 2 = TMA-staged kernel (search2.cuh)
  int opt_stage_rows = 0;         // rows per TMA stage (8 / 16 / 32), 0 = auto
  int opt_row_copy = 1;           // staged search: 1 = cp.async row copies where a row is <= 2 instructions (32-d, 128-d), 0 = bulk-async copies everywhere (search2.cuh, COPY)
  int opt_recent_ways = 1;        // DRAFT: 2 = two-way set-associative visited tags (Recent<Way2>), cp.async kinds only
  int opt_lookahead = 0;          // 1 = calls whose warps are all resident use search_knn2_la_kernel (one-hop lookahead, search_la.cuh; measured: no gain)
  int opt_search_cta = 0;         // 1 = calls with at most 2 queries per SM run one query per CTA of 4 warps (search_knn2_cta_kernel: -10 % latency;
                                  // opt-in: synccheck reports its named-barrier handshake, profiles/r2_sanitizer.md)
  int opt_recent_tag = 0;         // 0 auto (16-bit tags when every id fits), 32 = force 32-bit entries
  uint32_t opt_recent_slots = 0;  // direct-mapped visited slots of the TMA-staged kernel, 0 = auto
  uint32_t opt_build_batch = 0;
  int opt_build_impl = 0;         // 0 auto (TMA-staged K1 where a staged kernel exists), 1 = register-staged K1
  uint32_t auto_vis_ef = 0, auto_vis_slots = 0;  // adaptive visited-table size for the last-used ef
  uint32_t* h_retry_seen = nullptr;              // pinned: retry count of the previous async search

  // builder state (build_host.cu)
  uint32_t* d_stamp0 = nullptr;   // [cap_nodes] worklist de-duplication stamps (level-0 rows)
  uint32_t* d_stampU = nullptr;   // [cap_upper]
  uint32_t epoch = 0;
  uint32_t build_hint = 0;        // largest batch the running add_batch call will reach (scratch is sized once)
  uint32_t exact_vis_slots = 0;
  uint32_t* d_ver0 = nullptr;     // [cap_nodes] SPEC builder row stamps: 1 + id of the last insert that wrote the row
  uint32_t* d_verU = nullptr;     // [cap_upper]
  uint32_t opt_spec_window = 0;   // SPEC: fixed window size (0 = adaptive)
  int opt_spec_budget_us = 0;     // SPEC: an execution running longer than this stops before its next re-selection and continues in the next round; 0 = never (default: no gain once the sweeps were batched, profiles/r2_spec_build.md)
  int opt_spec_ahead = 0;         // SPEC: ids behind the window in which nodes with upper levels run those levels ahead of time; 0 = 2 x window, -1 = off
  int opt_spec_validation = 0;    // SPEC: 0 / 2 = dependency-level validation (spec.cuh), 1 = row-level (the round-2 first version; kept for A/B)
  uint32_t opt_spec_mult = 0;     // SPEC: adaptive window = mult / 10 x (inserts committed per round, running mean); 0 = 30 (3.0x: best of 1.5x .. 8x at 1M nodes, profiles/r2_spec_build.md)

  // scratch
  Scratch s_in, s_out, s_vis, s_ctl, s_build, s_stage, s_bvis, s_spec, s_list;

  ~Index();
  int use_device();
  int ensure_nodes(uint64_t n);
  int ensure_upper(uint64_t rows);
  int ensure_pool(uint64_t rows);
  int ensure_scratch(Scratch& s, size_t bytes);
  int push_meta();
  int pull_meta();
  int upload_vectors(const float* host, uint64_t first_row, uint64_t rows);
  int download_vectors(float* host, uint64_t first_row, uint64_t rows);
  int load_graph(uint64_t n, const float* vectors, const int32_t* levels, const uint64_t* row_offs,
                 const uint32_t* nbrs, int64_t entry, int32_t max_layer);
  int snapshot_rows(HostRows& h);
  void row_list(const HostRows& h, uint32_t node, uint32_t level, std::vector<uint32_t>& out) const;

  // search (search_host.cu)
  int search_device(uint64_t nq, const float* d_q, uint32_t k, uint32_t ef, uint32_t* d_ids, float* d_sims,
                    uint32_t* d_counts, uint32_t* d_stats, cudaStream_t s);
  int search_host(uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint32_t* ids, float* sims, uint32_t* counts,
                  uint32_t* stats);
  int search_level_host(const float* q, uint32_t ep, uint32_t ef, uint32_t level, uint32_t* ids, float* sims,
                        uint32_t* n_out);
  uint32_t pick_vis_slots(uint32_t ef);
  static constexpr uint32_t kCtlSlots = 8;  // 64-byte control slots for concurrently running search launches
  int search_device2(uint64_t nq, const float* d_q, uint32_t k, uint32_t ef, int efr, uint32_t* d_ids, float* d_sims,
                     uint32_t* d_counts, uint32_t* d_stats, cudaStream_t s, uint32_t ctl_slot = 0);
  int search_host_pipelined(uint64_t nq, const float* q, uint32_t k, uint32_t ef, int efr, uint32_t* ids, float* sims,
                            uint32_t* counts, const float* d_q, uint32_t* d_ids, float* d_sims, uint32_t* d_counts);
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  static constexpr size_t kPinnedStage = 128 * 1024;
  char* h_stage = nullptr;          // pinned staging buffer for small host calls
  bool last_search_staged = false;  // set by search_device: which kernel family served the last call

  // insert (build_host.cu)
  int add_batch(uint64_t count, const float* data, const int32_t* levels, int mode, uint32_t* first_id, bool want_touched);
  int add_exact(uint32_t first, uint32_t count, bool want_touched);
  int delete_node(uint32_t id);
  bool exact_staged(size_t* smem, size_t list_bytes, uint32_t* vis_slots) const;
  int add_fast(uint32_t first, uint32_t count);
  int add_spec(uint32_t first, uint32_t count);
  int fast_batch(uint32_t first, uint32_t count);
  int set_entry(int32_t entry, int32_t max_layer);
  int draw_level();
};

uint32_t next_pow2(uint64_t v);
bool kind_needs_smem_query(int kind);

}  // namespace hnsw

// the opaque handle of include/hnsw_b200.h
struct hnsw_index {
  hnsw::Index impl;
};

#define IDX_OR_FAIL(idx)                                                \
  if (!(idx)) return hnsw::fail(HNSW_ERR_INVALID, "null index handle"); \
  hnsw::Index& ix = (idx)->impl;                                        \
  {                                                                     \
    int rc_ = ix.use_device();                                          \
    if (rc_) return rc_;                                                \
  }

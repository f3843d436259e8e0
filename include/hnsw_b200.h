/*
 * hnsw_b200.h — C ABI of the B200-native HNSW engine (libhnsw_b200.so).
 *
 * This is the drop-in boundary for the hot path of zhao-lang/redis_hnsw: everything the reference's Redis
 * command handlers (src/lib.rs) and persistence conversions (src/types.rs) ask of `hnsw::Index<f32,f32>`
 * (src/hnsw/core.rs) is available here as plain-pointer `extern "C"` calls, so the reference's Rust host code
 * can bind it with an `extern "C"` block and keep HNSW.NEW / HNSW.NODE.ADD / HNSW.SEARCH
 * unchanged.  Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - Nodes are dense uint32 ids handed out in insertion order (0, 1, 2, ...).  The reference identifies nodes
 *     by name ("hnsw.{index}.{node}", lib.rs:343); the name<->id map stays in the host layer.
 *   - "sim" is the reference's similarity: NEGATIVE squared L2 in f32, bigger = closer (metrics.rs:75,80).
 *     Distances are bit-identical to the reference's AVX2+FMA path when dim % 32 == 0 and to its scalar
 *     path otherwise (metrics.rs:14-23).
 *   - All calls return an int status (HNSW_OK = 0).  On failure hnsw_last_error() holds a message whose
 *     text follows the reference's HNSWError strings (core.rs:390,408,421,479).  The library never aborts.
 *   - Caller owns every input/output buffer.  Host-pointer calls copy to/from the device internally;
 *     *_device calls take device pointers and a CUDA stream (a `cudaStream_t` passed as void*).
 *   - Calls on ONE index must be serialised by the caller (the reference runs on Redis' main thread and
 *     guards each index with try_read/try_write, lib.rs:349,474).  Different indexes are independent.
 *   - There is no CPU fallback: every compute call runs CUDA kernels on the index's device and fails with
 *     HNSW_ERR_CUDA if no device is usable.
 */
#ifndef HNSW_B200_H
#define HNSW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hnsw_index hnsw_index_t;

enum {
  HNSW_OK = 0,
  HNSW_ERR_DIM_MISMATCH = 1, /* "data dimension: {n} does not match Index"  core.rs:390,479 */
  HNSW_ERR_EXISTS = 2,       /* "Node: {name} already exists"               core.rs:408 (raised by the host layer) */
  HNSW_ERR_NOT_FOUND = 3,    /* "Node: {name} does not exist"               core.rs:421 */
  HNSW_ERR_INVALID = 4,      /* bad argument (null pointer, ef = 0, unsupported ef, ...) */
  HNSW_ERR_CUDA = 5,         /* CUDA runtime / launch failure, or no device */
  HNSW_ERR_OOM = 6           /* host or device allocation failed */
};

#define HNSW_NO_NODE 0xFFFFFFFFu

/* Build modes for hnsw_index_add_batch. */
enum {
  HNSW_BUILD_EXACT = 0, /* sequentially consistent: same graph as inserting one node at a time (reference semantics) */
  HNSW_BUILD_FAST = 1,  /* batched snapshot inserts: same per-insert algorithm, searches of one batch do not see
                           each other's edges; graph differs from the sequential one (recall-equivalent) */
  HNSW_BUILD_SPEC = 2   /* speculative-exact: windows of inserts run in parallel against the committed graph and commit
                           strictly in stream order, each only if no row it read was written since; the graph is the
                           sequential one (identical to HNSW_BUILD_EXACT and to the reference, core.rs:489-599) */
};

/* The pub fields of Index the reference's host code reads (core.rs:303-319; types.rs:62-91). */
typedef struct hnsw_params {
  uint32_t data_dim;        /* core.rs:307 */
  uint32_t m;               /* core.rs:308 */
  uint32_t m_max;           /* core.rs:309  (= m,   core.rs:335) */
  uint32_t m_max_0;         /* core.rs:310  (= 2m,  core.rs:336) */
  uint32_t ef_construction; /* core.rs:311 */
  int32_t max_layer;        /* core.rs:314 */
  double level_mult;        /* core.rs:312  (= 1/ln m, core.rs:338) */
  uint64_t node_count;      /* core.rs:313  live nodes */
  uint64_t n_ids;           /* ids handed out so far (live + deleted) */
  uint32_t enterpoint;      /* core.rs:317  HNSW_NO_NODE when the index is empty */
  int32_t device;           /* CUDA device ordinal holding the index */
} hnsw_params_t;

/* Per-query work counters of the search kernel (same definitions as the oracle's). */
typedef struct hnsw_query_stats {
  uint32_t n_dist; /* metric evaluations (core.rs:621,652) */
  uint32_t n_adj;  /* neighbour ids iterated (core.rs:646) */
  uint32_t n_hops; /* candidates expanded (core.rs:631-667) */
  uint32_t flags;  /* bit0: visited table overflowed and the query was re-run with a larger table */
} hnsw_query_stats_t;

/* ---- lifecycle ------------------------------------------------------------------------------------- */

/* Index::new(name, euclidean, data_dim, m, ef_construction)  core.rs:322-346 (call site lib.rs:153-159).
 * `device` is the CUDA ordinal (-1 = current device). */
int hnsw_index_create(uint32_t data_dim, uint32_t m, uint32_t ef_construction, int device, hnsw_index_t** out);
void hnsw_index_destroy(hnsw_index_t* idx);

/* Pre-size device storage for `n_nodes` ids (optional; storage grows on demand otherwise). */
int hnsw_index_reserve(hnsw_index_t* idx, uint64_t n_nodes);

/* Seed of the level generator used when a level of -1 is passed to add (the reference seeds from entropy,
 * core.rs:344; a fixed seed makes builds reproducible). */
int hnsw_index_seed(hnsw_index_t* idx, uint64_t seed);

/* ---- insert: Index::add_node(name, data, update_fn)  core.rs:383-412 -> insert core.rs:489-599 ------- */

/* One NODE.ADD.  `n` is the length of `data` (checked against data_dim like core.rs:389-391).
 * `level` >= 0 injects the level draw of core.rs:601-605; -1 draws it from the index's generator.
 * (The first node of an empty index ignores it: core.rs:393-405.)  `out_id` receives the new node id.
 * The set of nodes whose adjacency changed (what the reference reports through update_fn,
 * core.rs:522,535-537,570-572,580-584) is available from hnsw_index_touched until the next mutation. */
int hnsw_index_add(hnsw_index_t* idx, const float* data, uint64_t n, int32_t level, uint32_t* out_id);

/* A NODE.ADD stream of `count` vectors ([count][data_dim], row-major host memory).  `levels` may be NULL
 * (draw all) or hold one entry per vector (-1 = draw).  Ids first_id .. first_id+count-1 are assigned in order.
 * mode: HNSW_BUILD_EXACT, HNSW_BUILD_FAST or HNSW_BUILD_SPEC. */
int hnsw_index_add_batch(hnsw_index_t* idx, uint64_t count, const float* data, const int32_t* levels, int mode,
                         uint32_t* first_id);

/* Ids of the nodes touched by the last hnsw_index_add / hnsw_index_delete (replaces update_fn).
 * Writes up to `cap` ids and stores the full count in *n. */
int hnsw_index_touched(hnsw_index_t* idx, uint32_t* ids, uint64_t cap, uint64_t* n);

/* Index::delete_node(name, update_fn)  core.rs:414-475. */
int hnsw_index_delete(hnsw_index_t* idx, uint32_t id);

/* ---- search: Index::search_knn(data, k)  core.rs:477-486 -> search_knn_internal core.rs:865-892 ------ */

/* One HNSW.SEARCH.  `n` = length of `query` (core.rs:478-480).  `ef` = 0 means ef_construction, which is what
 * the reference always uses (core.rs:485); any other value is the efSearch extension.  Writes up to k results
 * nearest-first into ids/sims and the result count into *n_out (min(k, ef, reachable nodes); 0 for an empty
 * index, core.rs:481-483). */
int hnsw_index_search(hnsw_index_t* idx, const float* query, uint64_t n, uint32_t k, uint32_t ef, uint32_t* ids,
                      float* sims, uint32_t* n_out);

/* `nq` independent queries ([nq][data_dim] host memory) -> ids/sims [nq][k] (unused slots: HNSW_NO_NODE / -inf),
 * counts [nq].  `stats` may be NULL.  Copies host<->device inside the call. */
int hnsw_index_search_batch(hnsw_index_t* idx, uint64_t nq, const float* queries, uint32_t k, uint32_t ef,
                            uint32_t* ids, float* sims, uint32_t* counts, hnsw_query_stats_t* stats);

/* Same, with every buffer already in device memory of the index's device; enqueued on `stream`
 * (cudaStream_t as void*, NULL = the index's own stream) and NOT synchronised.  `stats` may be NULL. */
int hnsw_index_search_batch_device(hnsw_index_t* idx, uint64_t nq, const float* d_queries, uint32_t k, uint32_t ef,
                                   uint32_t* d_ids, float* d_sims, uint32_t* d_counts, hnsw_query_stats_t* d_stats,
                                   void* stream);

/* search_level(query, ep, ef, level)  core.rs:607-675 on its own: the whole result set (<= ef) nearest-first.
 * Used by the parity tests to check the kernel level by level. */
int hnsw_index_search_level(hnsw_index_t* idx, const float* query, uint32_t entry, uint32_t ef, uint32_t level,
                            uint32_t* ids, float* sims, uint32_t* n_out);

/* ---- metric: euclidean(v1, v2, n)  metrics.rs:14-84 ------------------------------------------------- */

/* out[i] = -||a[i] - b[i]||^2 for `rows` row pairs of length `dim` (host memory), bit-identical to the
 * reference's AVX2 path (dim % 32 == 0) or scalar path (otherwise).  `device` = CUDA ordinal or -1. */
int hnsw_l2_batch(const float* a, const float* b, uint64_t rows, uint32_t dim, float* out, int device);

/* ---- getters for the pub fields read by lib.rs / types.rs ------------------------------------------- */

int hnsw_index_params(hnsw_index_t* idx, hnsw_params_t* out);
/* Level drawn for the node (the layer set it is listed in, core.rs:596); -1 if deleted. */
int hnsw_index_node_level(hnsw_index_t* idx, uint32_t id, int32_t* level);
/* node.neighbors[level] in list order (types.rs:292-309).  Writes up to `cap` ids, full count in *n. */
int hnsw_index_node_neighbors(hnsw_index_t* idx, uint32_t id, uint32_t level, uint32_t* ids, uint64_t cap,
                              uint64_t* n);
/* node.data (types.rs:296): data_dim floats. */
int hnsw_index_node_vector(hnsw_index_t* idx, uint32_t id, float* out);

/* Adjacency lists of several (node, level) rows at once — one device gather and one copy instead of one round trip per
 * row; what a host keeps current after a mutation (the reference rewrites the record of every node reported through
 * update_fn: lib.rs:351-353, types.rs:292-309).  Row r gets min(len, stride) ids at out_ids[r * stride] and its full length
 * in out_lens[r] (a row longer than `stride` is asked for again with a larger stride).  A level the node does not have
 * gives length 0. */
int hnsw_index_rows_batch(hnsw_index_t* idx, uint64_t n_rows, const uint32_t* nodes, const uint32_t* levels, uint32_t stride,
                          uint32_t* out_ids, uint32_t* out_lens);

/* ---- whole-graph exchange (snapshot / restore; feeds the RDB records of types.rs:243-284, 410-428) ---
 * Flat graph: rows are (node, level) for level = 0..levels[node]; row index = sum_{j<node}(levels[j]+1) + level;
 * row_offs has n_rows+1 entries into nbrs; deleted nodes have levels = -1 and no rows. */
int hnsw_index_graph_sizes(hnsw_index_t* idx, uint64_t* n_ids, uint64_t* n_rows, uint64_t* n_edges);
int hnsw_index_export_graph(hnsw_index_t* idx, int32_t* levels, uint64_t* row_offs, uint32_t* nbrs, int64_t* entry,
                            int32_t* max_layer);
/* Copies the vector slab back in natural element order: [n_ids][data_dim]. */
int hnsw_index_export_vectors(hnsw_index_t* idx, float* out);
/* Replaces the index contents (make_index, lib.rs:252-315, in one pass). */
int hnsw_index_load_graph(hnsw_index_t* idx, uint64_t n_ids, const float* vectors, const int32_t* levels,
                          const uint64_t* row_offs, const uint32_t* nbrs, int64_t entry, int32_t max_layer);

/* ---- replication across GPUs (index replicated, queries sharded) -------------------------------------
 * The index's device buffers as (pointer, bytes) pairs so a host can broadcast them (NCCL) to a replica
 * created with the same parameters and reserve() size.  hnsw_index_adopt_replica refreshes the replica's
 * host-side metadata after its buffers were overwritten. */
typedef struct hnsw_device_buffer {
  void* ptr;
  uint64_t bytes;
} hnsw_device_buffer_t;
int hnsw_index_device_buffers(hnsw_index_t* idx, hnsw_device_buffer_t* out, uint32_t cap, uint32_t* n);
int hnsw_index_replica_layout(hnsw_index_t* idx, uint64_t* layout8);
int hnsw_index_prepare_replica(hnsw_index_t* idx, const uint64_t* layout8);
int hnsw_index_adopt_replica(hnsw_index_t* idx);

/* ---- tuning / diagnostics --------------------------------------------------------------------------- */

/* Named integer options: "visited_slots" (per-query visited hash slots, power of two, 0 = auto),
 * "search_ctas_per_sm", "build_batch" (nodes per batch of the FAST builder), "build_impl" (0 auto, 1 = register-staged
 * batch searches, 2 = TMA-staged), "search_impl", "stage_rows", "recent_slots", "recent_tag", "search_block", "row_copy"
 * (1 = cp.async row staging for 32-d / 128-d rows, the default; 0 = bulk-async copies for every dimension),
 * "recent_ways" (1 | 2), "lookahead" (0 | 1), "search_cta" (0 | 1: one query per CTA of four warps for small calls).
 * SPEC builder (HNSW_BUILD_SPEC; none of these changes the graph, tests/test_gpu_spec_build.py): "spec_window" (fixed
 * window, 0 = adaptive), "spec_mult" (adaptive window = value / 10 x inserts committed per round; 0 = 30),
 * "spec_validation" (0 | 2 = dependency-level, 1 = row-level), "spec_ahead" (ids behind the window in which nodes with
 * upper levels run those levels ahead of time; 0 = 2 x window, -1 = off), "spec_budget_us" (an execution running longer
 * stops before its next re-selection and continues in the next round; 0 = never).
 * Unknown names -> HNSW_ERR_INVALID. */
int hnsw_index_set_option(hnsw_index_t* idx, const char* name, int64_t value);
/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
uint64_t hnsw_launch_count(void);
/* Builder counters of the last add_batch: [0] inserts, [1] speculative conflicts re-run, [2] re-prunes,
 * [3] distance evaluations. */
int hnsw_index_build_stats(hnsw_index_t* idx, uint64_t* out4);
/* All builder counters since the index was created.  Writes up to `cap` values and stores how many exist in *n:
 * [0] inserts  [1] speculative executions thrown away (SPEC) / rows not re-selected (FAST)  [2] re-prunes
 * [3] distance evaluations  [4] FAST: over-full rows that did not fit the re-prune worklist  [5] FAST: re-prunes skipped
 * [6] FAST: edges refused because a hub row was full  [7] SPEC: commit rounds  [8] SPEC: executions (first + repeated)
 * [9] SPEC: distance evaluations of thrown-away executions  [10] SPEC: inserts sent to the one-warp EXACT kernel
 * [11] SPEC: largest window  [12] SPEC: rows committed as append / remove operations on content newer than the
 * insert's snapshot (dependency-level validation). */
int hnsw_index_build_stats_ex(hnsw_index_t* idx, uint64_t* out, uint32_t cap, uint32_t* n);

const char* hnsw_last_error(void);
const char* hnsw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HNSW_B200_H */

// Host-side mirror of the reference's `hnsw::Index<f32,f32>` (src/hnsw/core.rs:302-346) in C++, over the C ABI of
// include/hnsw_b200.h.  This is what the reference's command handlers (src/lib.rs) and persistence conversions
// (src/types.rs) see: node NAMES in, node names out, the same error strings.  The device speaks dense u32 ids handed
// out in insertion order; the name <-> id map (the reference's `nodes: HashMap<String, Node>`, core.rs:316) lives here.
// All arithmetic happens behind the C ABI on the GPU; nothing in this file computes a distance.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/hnsw_b200.h"

namespace hnswhost {

// HNSWError::{Str,String} (core.rs:24-46); what() is error_string()
struct HNSWError : std::runtime_error {
  int code;
  explicit HNSWError(const std::string& m, int c = HNSW_ERR_INVALID) : std::runtime_error(m), code(c) {}
};

// SearchResult {sim, name, data} (core.rs:48-62)
struct SearchResult {
  float sim;
  std::string name;  // last '.'-segment of the node name (core.rs:885-887)
  std::vector<float> data;
};

// One node as types.rs sees it: NodeRedis {data, neighbors: Vec<Vec<String>>} (types.rs:286-309)
struct NodeRecord {
  std::vector<float> data;
  std::vector<std::vector<std::string>> neighbors;
};

// The index as types.rs sees it: IndexRedis (types.rs:45-60)
struct IndexRecord {
  std::string name, mfunc_kind = "Euclidean";
  uint64_t data_dim = 0, m = 0, m_max = 0, m_max_0 = 0, ef_construction = 0;
  double level_mult = 0;
  uint64_t node_count = 0, max_layer = 0;
  std::vector<std::vector<std::string>> layers;  // a node is listed only in the layer of its top level (core.rs:596)
  std::vector<std::string> nodes;
  std::optional<std::string> enterpoint;
};

class NamedIndex {
 public:
  // Index::new(name, euclidean, data_dim, m, ef_construction)  core.rs:322-346
  NamedIndex(const std::string& name, uint64_t data_dim, uint64_t m, uint64_t ef_construction, int device = -1);
  ~NamedIndex();
  NamedIndex(const NamedIndex&) = delete;
  NamedIndex& operator=(const NamedIndex&) = delete;

  // make_index (lib.rs:252-315): rebuild from the persisted records in ONE pass (names -> ids -> flat graph -> upload).
  // `fetch(node_name)` returns the node's record or nullptr ("Node: {} does not exist", lib.rs:261,274).
  template <class Fetch>
  static std::unique_ptr<NamedIndex> restore(const IndexRecord& ir, Fetch fetch, int device = -1);

  // core.rs:383-412.  `touched` (optional) receives the names the reference reports through update_fn
  // (core.rs:580-584).  `level` >= 0 injects the level draw (tests); -1 draws.
  void add_node(const std::string& node_name, const float* data, size_t n, std::vector<std::string>* touched = nullptr,
                int level = -1);
  // extension: a NODE.ADD stream of `count` nodes in one call (hnsw_index_add_batch).  fast = true uses the batched
  // builder (HNSW_BUILD_FAST: same per-insert algorithm, nodes of one batch do not see each other); false is the
  // sequentially consistent stream (HNSW_BUILD_SPEC: speculative windows committed in order, the graph of one-by-one
  // NODE.ADDs).  All names are checked before anything is inserted.
  void add_nodes(const std::vector<std::string>& node_names, const float* data, size_t n, bool fast);
  // core.rs:414-475
  void delete_node(const std::string& node_name, std::vector<std::string>* touched = nullptr);
  // core.rs:477-486 (ef = 0 -> ef_construction, core.rs:485).  The reference copies each hit's vector into the result
  // (core.rs:888) but its HNSW.SEARCH reply never sends it (types.rs:436-456): `with_data` = false skips the k
  // device-to-host row copies.
  size_t clamp_k(size_t k, uint32_t ef) const;  // min(k, effective ef, live nodes): the most results a search can return
  std::vector<SearchResult> search_knn(const float* q, size_t n, size_t k, uint32_t ef = 0, bool with_data = false) const;
  // extension: nq independent queries in one device batch; result r of query i at [i][r]
  std::vector<std::vector<SearchResult>> search_knn_batch(const float* q, size_t nq, size_t n, size_t k, uint32_t ef = 0,
                                                           bool with_data = false) const;

  // pub fields (core.rs:303-319)
  const std::string& name() const { return name_; }
  hnsw_params_t params() const;
  bool contains(const std::string& node_name) const { return ids_.count(node_name) != 0; }
  size_t live_nodes() const { return ids_.size(); }
  std::optional<std::string> enterpoint() const;
  // conversions of types.rs:62-91 and :292-309
  IndexRecord to_record() const;
  NodeRecord node_record(const std::string& node_name) const;
  std::vector<std::string> node_names() const;
  // Every record at once (types.rs:243-284, 410-428 for the whole keyspace of one index).
  //
  // FORK SAFETY.  to_record(), node_record() and snapshot() never touch the device: they read a write-through host mirror
  // (vectors, adjacency lists, levels, index scalars) that every mutating call refreshes before it returns — the rows
  // the mutation touched come back in ONE gather (hnsw_index_rows_batch), a bulk load re-reads the graph once.  Redis
  // serialises keys from a fork()ed child (BGSAVE, AOF rewrite, replica sync; the persistence server event fires in
  // that child too), where the parent's CUDA context is unusable; the reference has the same property for the same
  // reason: its key values are plain host records rewritten after every mutation (lib.rs:351-365).
  struct Snapshot {
    IndexRecord index;
    std::vector<std::pair<std::string, NodeRecord>> nodes;
  };
  Snapshot snapshot() const;
  // bumped by every mutation (add_node / delete_node / restore)
  uint64_t epoch() const { return epoch_; }

  hnsw_index_t* handle() const { return h_; }

 private:
  void restore_graph(const IndexRecord& ir, const std::vector<const NodeRecord*>& recs);
  void check(int rc) const;
  std::vector<uint32_t> touched_ids() const;
  std::vector<std::string> touched_names(const std::vector<uint32_t>& ids) const;
  void refresh_rows(const std::vector<uint32_t>& ids);  // mirror <- device for the rows of these nodes (one gather)
  void refresh_all();                                   // mirror <- device for the whole graph (bulk loads)
  void resync_ids();                                    // name table <- n_ids after a failed add (dead ids stay unnamed)

  std::string name_;
  hnsw_index_t* h_ = nullptr;
  uint32_t dim_ = 0;
  std::unordered_map<std::string, uint32_t> ids_;  // live names -> id
  std::vector<std::string> names_;                 // id -> name ("" once deleted)
  std::vector<char> alive_;
  uint64_t epoch_ = 1;
  // write-through host mirror (see FORK SAFETY above)
  std::vector<std::vector<float>> hvec_;                    // id -> vector
  std::vector<std::vector<std::vector<uint32_t>>> hadj_;    // id -> level -> neighbour ids, list order
  hnsw_params_t hparams_{};                                 // index scalars after the last mutation
};

template <class Fetch>
std::unique_ptr<NamedIndex> NamedIndex::restore(const IndexRecord& ir, Fetch fetch, int device) {
  std::unique_ptr<NamedIndex> ix(new NamedIndex(ir.name, ir.data_dim, ir.m, ir.ef_construction, device));
  std::vector<const NodeRecord*> recs;
  recs.reserve(ir.nodes.size());
  for (const std::string& nn : ir.nodes) {
    const NodeRecord* r = fetch(nn);
    if (!r) throw HNSWError("Node: " + nn + " does not exist", HNSW_ERR_NOT_FOUND);  // lib.rs:261
    recs.push_back(r);
  }
  ix->restore_graph(ir, recs);
  return ix;
}

}  // namespace hnswhost

// NamedIndex: the reference's Index<f32,f32> interface (src/hnsw/core.rs) over the C ABI.  See named_index.hpp.
#include "named_index.hpp"

#include <algorithm>
#include <cmath>

namespace hnswhost {

static std::string last_error() {
  const char* e = hnsw_last_error();
  return e ? std::string(e) : std::string("unknown error");
}

void NamedIndex::check(int rc) const {
  if (rc != HNSW_OK) throw HNSWError(last_error(), rc);
}

// Rust's `{:?}` of a &str: quoted, with ", \ and control characters escaped (core.rs:408,421 format names this way)
static std::string rust_debug_str(const std::string& s) {
  std::string o = "\"";
  for (char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default: o += c;
    }
  }
  return o + "\"";
}

NamedIndex::NamedIndex(const std::string& name, uint64_t data_dim, uint64_t m, uint64_t ef_construction, int device)
    : name_(name), dim_((uint32_t)data_dim) {
  if (data_dim == 0 || data_dim > 0xFFFFFFFFull || m > 0xFFFFFFFFull || ef_construction > 0xFFFFFFFFull)
    throw HNSWError("index parameters out of range");
  int rc = hnsw_index_create((uint32_t)data_dim, (uint32_t)m, (uint32_t)ef_construction, device, &h_);
  if (rc != HNSW_OK) {
    h_ = nullptr;
    throw HNSWError(last_error(), rc);
  }
  check(hnsw_index_params(h_, &hparams_));
}

NamedIndex::~NamedIndex() {
  if (h_) hnsw_index_destroy(h_);
}

hnsw_params_t NamedIndex::params() const { return hparams_; }  // host copy, refreshed by every mutation (fork-safe)

std::optional<std::string> NamedIndex::enterpoint() const {
  if (hparams_.enterpoint == HNSW_NO_NODE || hparams_.enterpoint >= names_.size()) return std::nullopt;
  return names_[hparams_.enterpoint];
}

std::vector<uint32_t> NamedIndex::touched_ids() const {
  uint64_t n = 0;
  check(hnsw_index_touched(h_, nullptr, 0, &n));
  std::vector<uint32_t> ids(n);
  if (n) check(hnsw_index_touched(h_, ids.data(), n, &n));
  return ids;
}

std::vector<std::string> NamedIndex::touched_names(const std::vector<uint32_t>& ids) const {
  std::vector<std::string> out;
  out.reserve(ids.size());
  for (uint32_t t : ids)
    if (t < names_.size() && alive_[t]) out.push_back(names_[t]);
  return out;
}

// mirror <- device: every level of the given nodes, one gather + one copy (hnsw_index_rows_batch)
void NamedIndex::refresh_rows(const std::vector<uint32_t>& ids) {
  check(hnsw_index_params(h_, &hparams_));
  std::vector<uint32_t> nodes, levels;
  for (uint32_t id : ids) {
    if (id >= names_.size() || !alive_[id]) continue;
    int32_t lv = 0;
    check(hnsw_index_node_level(h_, id, &lv));
    if (lv < 0) continue;
    hadj_[id].resize((size_t)lv + 1);
    for (int32_t l = 0; l <= lv; ++l) nodes.push_back(id), levels.push_back((uint32_t)l);
  }
  if (nodes.empty()) return;
  uint32_t stride = 2 * hparams_.m_max_0 + 32;
  std::vector<uint32_t> lens(nodes.size()), out;
  for (;;) {
    out.resize(nodes.size() * (size_t)stride);
    check(hnsw_index_rows_batch(h_, nodes.size(), nodes.data(), levels.data(), stride, out.data(), lens.data()));
    const uint32_t longest = *std::max_element(lens.begin(), lens.end());
    if (longest <= stride) break;
    stride = longest + 32;  // degree is unbounded (core.rs:793-795): ask again with room for the longest row
  }
  for (size_t r = 0; r < nodes.size(); ++r)
    hadj_[nodes[r]][levels[r]].assign(out.begin() + (long)(r * stride), out.begin() + (long)(r * stride + lens[r]));
}

// mirror <- device for everything: one graph export + one vector export
void NamedIndex::refresh_all() {
  check(hnsw_index_params(h_, &hparams_));
  uint64_t n_ids = 0, n_rows = 0, n_edges = 0;
  check(hnsw_index_graph_sizes(h_, &n_ids, &n_rows, &n_edges));
  hvec_.assign(n_ids, {});
  hadj_.assign(n_ids, {});
  if (n_ids == 0) return;
  std::vector<int32_t> levels(n_ids);
  std::vector<uint64_t> offs(n_rows + 1);
  std::vector<uint32_t> nbrs(std::max<uint64_t>(n_edges, 1));
  int64_t entry = -1;
  int32_t max_layer = 0;
  check(hnsw_index_export_graph(h_, levels.data(), offs.data(), nbrs.data(), &entry, &max_layer));
  std::vector<float> vecs(n_ids * (size_t)dim_);
  check(hnsw_index_export_vectors(h_, vecs.data()));
  uint64_t row = 0;
  for (uint64_t i = 0; i < n_ids; ++i) {
    if (levels[i] < 0) continue;  // deleted: no rows
    hvec_[i].assign(vecs.begin() + (long)(i * (size_t)dim_), vecs.begin() + (long)((i + 1) * (size_t)dim_));
    hadj_[i].resize((size_t)levels[i] + 1);
    for (int32_t l = 0; l <= levels[i]; ++l, ++row)
      hadj_[i][(size_t)l].assign(nbrs.begin() + (long)offs[row], nbrs.begin() + (long)offs[row + 1]);
  }
}

void NamedIndex::add_node(const std::string& node_name, const float* data, size_t n, std::vector<std::string>* touched,
                          int level) {
  if (n != dim_)  // core.rs:389-391
    throw HNSWError("data dimension: " + std::to_string(n) + " does not match Index", HNSW_ERR_DIM_MISMATCH);
  // core.rs:393-409: the duplicate check is skipped for the first node of an empty index
  if (!ids_.empty() && ids_.count(node_name))
    throw HNSWError("Node: " + rust_debug_str(node_name) + " already exists", HNSW_ERR_EXISTS);
  uint32_t id = 0;
  const int rc = hnsw_index_add(h_, data, n, level, &id);
  if (rc != HNSW_OK) {
    const std::string why = last_error();
    resync_ids();  // a failed add still consumed its id (now a tombstone): keep the name table aligned with the device
    throw HNSWError(why, rc);
  }
  if (id != names_.size()) throw HNSWError("device id out of sequence");
  names_.push_back(node_name);
  alive_.push_back(1);
  ids_[node_name] = id;
  ++epoch_;
  hvec_.emplace_back(data, data + n);
  hadj_.emplace_back();
  std::vector<uint32_t> t = touched_ids();
  if (touched) *touched = touched_names(t);  // core.rs:580-584
  t.push_back(id);
  refresh_rows(t);                           // lib.rs:351-353, 361-362: every touched record and the new one are rewritten
}

void NamedIndex::add_nodes(const std::vector<std::string>& node_names, const float* data, size_t n, bool fast) {
  if (n != dim_)
    throw HNSWError("data dimension: " + std::to_string(n) + " does not match Index", HNSW_ERR_DIM_MISMATCH);
  if (node_names.empty()) return;
  {
    std::unordered_map<std::string, int> seen;
    for (const std::string& nn : node_names)
      if (ids_.count(nn) || seen[nn]++) throw HNSWError("Node: " + rust_debug_str(nn) + " already exists", HNSW_ERR_EXISTS);
  }
  uint32_t first = 0;
  const int rc = hnsw_index_add_batch(h_, node_names.size(), data, nullptr, fast ? HNSW_BUILD_FAST : HNSW_BUILD_SPEC, &first);
  if (rc != HNSW_OK && first != names_.size()) {
    const std::string why = last_error();
    resync_ids();
    throw HNSWError(why, rc);
  }
  // on a failure part-way the nodes that made it are live and named; the rest of the ids are tombstones
  for (size_t i = 0; i < node_names.size(); ++i) {
    int32_t lv = -1;
    const bool live = hnsw_index_node_level(h_, first + (uint32_t)i, &lv) == HNSW_OK && lv >= 0;
    names_.push_back(live ? node_names[i] : std::string());
    alive_.push_back(live ? 1 : 0);
    if (live) ids_[node_names[i]] = first + (uint32_t)i;
  }
  ++epoch_;
  refresh_all();  // a batch touches too many rows to name them: the mirror is re-read once
  if (rc != HNSW_OK) throw HNSWError(last_error(), rc);
}

// after a failed mutation: every id the device handed out has an entry here (dead ones unnamed)
void NamedIndex::resync_ids() {
  hnsw_params_t p{};
  if (hnsw_index_params(h_, &p) != HNSW_OK) return;
  while (names_.size() < p.n_ids) {
    names_.emplace_back();
    alive_.push_back(0);
    hvec_.emplace_back();
    hadj_.emplace_back();
  }
  hparams_ = p;
}

void NamedIndex::delete_node(const std::string& node_name, std::vector<std::string>* touched) {
  auto it = ids_.find(node_name);
  if (it == ids_.end())  // core.rs:419-422
    throw HNSWError("Node: " + rust_debug_str(node_name) + " does not exist", HNSW_ERR_NOT_FOUND);
  const uint32_t id = it->second;
  check(hnsw_index_delete(h_, id));
  ids_.erase(it);
  alive_[id] = 0;
  ++epoch_;
  std::vector<uint32_t> t = touched_ids();
  if (touched) *touched = touched_names(t);  // core.rs:443-447 (the victim itself is never reported)
  names_[id].clear();
  hvec_[id].clear();
  hvec_[id].shrink_to_fit();
  hadj_[id].clear();
  refresh_rows(t);
}

// At most min(k, ef, live nodes) results can come back (core.rs:879 pops from a result set of <= ef entries): buffers
// are sized by that, never by a user-supplied K (HNSW.SEARCH ... K 8589934592 must not allocate gigabytes).
size_t NamedIndex::clamp_k(size_t k, uint32_t ef) const {
  hnsw_params_t p = params();
  const uint64_t ef_eff = ef ? ef : p.ef_construction;
  return (size_t)std::min<uint64_t>(std::min<uint64_t>(k, ef_eff), p.node_count);
}

static std::string last_segment(const std::string& full) {  // core.rs:885-887
  size_t p = full.rfind('.');
  return p == std::string::npos ? full : full.substr(p + 1);
}

std::vector<SearchResult> NamedIndex::search_knn(const float* q, size_t n, size_t k, uint32_t ef, bool with_data) const {
  if (n != dim_)  // core.rs:478-480
    throw HNSWError("data dimension: " + std::to_string(n) + " does not match Index", HNSW_ERR_DIM_MISMATCH);
  std::vector<SearchResult> out;
  k = clamp_k(k, ef);
  if (k == 0) return out;
  std::vector<uint32_t> ids(k);
  std::vector<float> sims(k);
  uint32_t cnt = 0;
  check(hnsw_index_search(h_, q, n, (uint32_t)k, ef, ids.data(), sims.data(), &cnt));
  out.reserve(cnt);
  for (uint32_t i = 0; i < cnt; ++i) {
    if (ids[i] >= names_.size() || !alive_[ids[i]]) continue;  // never hand out an id without a name
    SearchResult r;
    r.sim = sims[i];
    r.name = last_segment(names_[ids[i]]);
    if (with_data) {
      r.data = hvec_[ids[i]];  // core.rs:888 copies the stored vector (host mirror)
    }
    out.push_back(std::move(r));
  }
  return out;
}

std::vector<std::vector<SearchResult>> NamedIndex::search_knn_batch(const float* q, size_t nq, size_t n, size_t k,
                                                                     uint32_t ef, bool with_data) const {
  if (n != dim_)
    throw HNSWError("data dimension: " + std::to_string(n) + " does not match Index", HNSW_ERR_DIM_MISMATCH);
  std::vector<std::vector<SearchResult>> out(nq);
  k = clamp_k(k, ef);
  if (k == 0 || nq == 0) return out;
  std::vector<uint32_t> ids(nq * k), counts(nq);
  std::vector<float> sims(nq * k);
  check(hnsw_index_search_batch(h_, nq, q, (uint32_t)k, ef, ids.data(), sims.data(), counts.data(), nullptr));
  for (size_t i = 0; i < nq; ++i) {
    out[i].reserve(counts[i]);
    for (uint32_t j = 0; j < counts[i]; ++j) {
      if (ids[i * k + j] >= names_.size() || !alive_[ids[i * k + j]]) continue;
      SearchResult r;
      r.sim = sims[i * k + j];
      r.name = last_segment(names_[ids[i * k + j]]);
      if (with_data) r.data = hvec_[ids[i * k + j]];
      out[i].push_back(std::move(r));
    }
  }
  return out;
}

std::vector<std::string> NamedIndex::node_names() const {
  std::vector<std::string> out;
  out.reserve(ids_.size());
  for (size_t i = 0; i < names_.size(); ++i)
    if (alive_[i]) out.push_back(names_[i]);
  return out;
}

// From<Index> for IndexRedis (types.rs:62-91) — host mirror only (fork-safe)
IndexRecord NamedIndex::to_record() const {
  const hnsw_params_t& p = hparams_;
  IndexRecord r;
  r.name = name_;
  r.data_dim = p.data_dim;
  r.m = p.m;
  r.m_max = p.m_max;
  r.m_max_0 = p.m_max_0;
  r.ef_construction = p.ef_construction;
  r.level_mult = p.level_mult;
  r.node_count = p.node_count;
  r.max_layer = (uint64_t)std::max(0, p.max_layer);
  // `layers`: one set per level 0..max_layer; a node sits in the set of its own top level (core.rs:596).  An index
  // that never held a node has no layer at all (core.rs:341).
  if (!names_.empty() || p.node_count) r.layers.resize((size_t)r.max_layer + 1);
  for (size_t i = 0; i < names_.size(); ++i) {
    if (!alive_[i] || hadj_[i].empty()) continue;
    const size_t lv = hadj_[i].size() - 1;
    if (lv >= r.layers.size()) r.layers.resize(lv + 1);
    r.layers[lv].push_back(names_[i]);
    r.nodes.push_back(names_[i]);
  }
  if (p.enterpoint != HNSW_NO_NODE && p.enterpoint < names_.size()) r.enterpoint = names_[p.enterpoint];
  return r;
}

// From<&Node> for NodeRedis (types.rs:292-309) — host mirror only (fork-safe)
NodeRecord NamedIndex::node_record(const std::string& node_name) const {
  auto it = ids_.find(node_name);
  if (it == ids_.end()) throw HNSWError("Node: " + node_name + " does not exist", HNSW_ERR_NOT_FOUND);
  const uint32_t id = it->second;
  NodeRecord r;
  r.data = hvec_[id];
  r.neighbors.reserve(hadj_[id].size());
  for (const auto& lst : hadj_[id]) {
    std::vector<std::string> layer;
    layer.reserve(lst.size());
    for (uint32_t x : lst) layer.push_back(names_[x]);
    r.neighbors.push_back(std::move(layer));
  }
  return r;
}

NamedIndex::Snapshot NamedIndex::snapshot() const {
  Snapshot s;
  s.index = to_record();
  s.nodes.reserve(ids_.size());
  for (size_t i = 0; i < names_.size(); ++i)
    if (alive_[i]) s.nodes.emplace_back(names_[i], node_record(names_[i]));
  return s;
}

// make_index (lib.rs:252-315) in one pass: ids follow the order of ir.nodes; a node's level is the layer set that
// lists it (lib.rs:289-300); its adjacency lists keep their stored order (lib.rs:275-286); the whole graph goes to the
// device with one hnsw_index_load_graph call instead of per-node pointer fix-ups.
void NamedIndex::restore_graph(const IndexRecord& ir, const std::vector<const NodeRecord*>& recs) {
  const size_t n = ir.nodes.size();
  names_ = ir.nodes;
  alive_.assign(n, 1);
  ids_.clear();
  ids_.reserve(n * 2);
  for (size_t i = 0; i < n; ++i) ids_[names_[i]] = (uint32_t)i;
  if (ids_.size() != n) throw HNSWError("duplicate node name in the index record");
  auto id_of = [&](const std::string& nn) -> uint32_t {
    auto it = ids_.find(nn);
    if (it == ids_.end()) throw HNSWError("Node: " + nn + " does not exist", HNSW_ERR_NOT_FOUND);  // lib.rs:281,295,308
    return it->second;
  };
  std::vector<int32_t> levels(n, -2);
  for (size_t l = 0; l < ir.layers.size(); ++l)
    for (const std::string& nn : ir.layers[l]) levels[id_of(nn)] = (int32_t)l;
  std::vector<float> vecs(n * (size_t)dim_);
  std::vector<uint64_t> offs;
  std::vector<uint32_t> nbrs;
  offs.push_back(0);
  for (size_t i = 0; i < n; ++i) {
    const NodeRecord& r = *recs[i];
    if (r.data.size() != dim_)
      throw HNSWError("data dimension: " + std::to_string(r.data.size()) + " does not match Index", HNSW_ERR_DIM_MISMATCH);
    std::copy(r.data.begin(), r.data.end(), vecs.begin() + i * (size_t)dim_);
    if (levels[i] == -2) levels[i] = r.neighbors.empty() ? 0 : (int32_t)r.neighbors.size() - 1;  // not listed in any layer
    if ((int32_t)r.neighbors.size() > levels[i] + 1) levels[i] = (int32_t)r.neighbors.size() - 1;
    for (int32_t l = 0; l <= levels[i]; ++l) {
      if ((size_t)l < r.neighbors.size())
        for (const std::string& nn : r.neighbors[(size_t)l]) nbrs.push_back(id_of(nn));
      offs.push_back(nbrs.size());
    }
  }
  int64_t entry = ir.enterpoint ? (int64_t)id_of(*ir.enterpoint) : -1;
  ++epoch_;
  if (nbrs.empty()) nbrs.push_back(0);
  hvec_.assign(n, {});
  hadj_.assign(n, {});
  if (n == 0) return;
  check(hnsw_index_load_graph(h_, n, vecs.data(), levels.data(), offs.data(), nbrs.data(), entry, (int32_t)ir.max_layer));
  check(hnsw_index_params(h_, &hparams_));
  for (size_t i = 0, row = 0; i < n; ++i) {  // the mirror is what was just uploaded
    hvec_[i] = recs[i]->data;
    hadj_[i].resize((size_t)levels[i] + 1);
    for (int32_t l = 0; l <= levels[i]; ++l, ++row)
      hadj_[i][(size_t)l].assign(nbrs.begin() + (long)offs[row], nbrs.begin() + (long)offs[row + 1]);
  }
}

}  // namespace hnswhost

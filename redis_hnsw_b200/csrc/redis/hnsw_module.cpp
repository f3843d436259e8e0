// libredis_hnsw_b200.so — the Redis module of the reference (src/lib.rs, src/types.rs) with the HNSW core replaced by
// the B200 engine behind include/hnsw_b200.h.
//
// Same module name ("hnsw", version 1), the same seven commands with the same flags and key spec (lib.rs:498-514), the
// same key names (`hnsw.{index}`, `hnsw.{index}.{node}`; lib.rs:137,342-343), argument defaults (M 5, EFCON 200, K 5;
// lib.rs:48,53,120), reply shapes (types.rs:122-155, 322-352, 445-456), error texts (including the `String("...")`
// rendering of HNSWError::error_string, core.rs:41-45) and the two native data types `hnswindex` / `hnswnodet` with
// their RDB encodings (types.rs:180-284, 377-428), so an RDB written by the reference loads here and vice versa.
//
// What changes underneath
//   * Index::{new, add_node, delete_node, search_knn} are NamedIndex calls -> C ABI -> CUDA kernels.
//   * The reference re-serialises the WHOLE index into its key after every mutation and rewrites every touched node
//     record (lib.rs:317-332, 351-353: O(N) per NODE.ADD).  Here key values are handles on the live index; records are
//     materialised from device state only when Redis asks for them (rdb_save, HNSW.GET, HNSW.NODE.GET).
//   * Persistence.  Redis runs rdb_save in a fork()ed child (BGSAVE, AOF rewrite, replica full sync), where the CUDA
//     context of the parent is unusable — and the persistence server event is no help: for background saves Redis fires
//     RDB_START / AOF_START from inside that child (only SYNC_RDB_START runs in the parent).  So nothing on the save path
//     may touch the device.  NamedIndex keeps a write-through host mirror of every record (vectors, adjacency lists,
//     levels, index scalars), refreshed in the parent by every mutating command from ONE gather of the rows the mutation
//     touched; rdb_save, HNSW.GET and HNSW.NODE.GET read only that mirror.  This is the reference's own contract — its
//     key values are host records rewritten after every mutation (lib.rs:351-365) — at O(touched rows) instead of O(N).
//   * Cold start (load_index -> make_index, lib.rs:229-315) rebuilds the device graph from the loaded records in one
//     pass on the first command that touches the index.
//   * Extensions (not in the reference): `EF ef` on HNSW.SEARCH (the reference always searches with ef_construction,
//     core.rs:485), HNSW.MSEARCH (a multi-query form that reaches the batched device path) and HNSW.NODE.MADD (a
//     NODE.ADD stream in one command, for bulk loads).
//   * Conscious fixes: HNSW.NODE.DEL of a missing node replies an error (the reference panics: lib.rs:384 unwrap);
//     deleting the index key by other means (DEL, FLUSHALL) also drops the cached index (the reference keeps a stale
//     entry in INDICES).
//
// No redis-server exists in this image: the module is exercised against tests/fake_redis/fake_redis_host.cpp, a small
// host that implements the same function table (tests/test_redis_module_cpu.py, tests/test_gpu_redis_module.py).
#define HNSW_REDISMODULE_MAIN
#include "redismodule_abi.h"

#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "named_index.hpp"

using hnswhost::HNSWError;
using hnswhost::IndexRecord;
using hnswhost::NamedIndex;
using hnswhost::NodeRecord;
using hnswhost::SearchResult;

namespace {

const char* const PREFIX = "hnsw";  // lib.rs:27

// ---------------------------------------------------------------- key values (types.rs)

struct IndexValue {  // value of `hnsw.{index}`  <- IndexRedis (types.rs:45-60)
  IndexRecord rec;   // as loaded from RDB; superseded by `live` (whose host mirror is always current) once that exists
  std::shared_ptr<NamedIndex> live;
};

struct NodeValue {   // value of `hnsw.{index}.{node}`  <- NodeRedis (types.rs:286-290)
  std::string name;
  NodeRecord rec;    // as loaded from RDB; superseded by the live index's mirror once the index is live
  std::weak_ptr<NamedIndex> live;
};

RedisModuleType* g_index_type = nullptr;  // HNSW_INDEX_REDIS_TYPE (types.rs:157)
RedisModuleType* g_node_type = nullptr;   // HNSW_NODE_REDIS_TYPE  (types.rs:354)
std::unordered_map<std::string, std::shared_ptr<NamedIndex>> g_indices;  // INDICES (lib.rs:32-35)
std::unordered_map<std::string, NodeValue*> g_node_values;    // node key name -> its value (owned by the keyspace)
std::unordered_map<std::string, IndexValue*> g_index_values;  // index key name -> its value

struct ReplyError {  // RedisError::String
  std::string msg;
};

// HNSWError::error_string (core.rs:41-45): `format!("{:?}", self)` of the derived Debug -> String("<escaped msg>")
std::string error_string(const std::string& msg) {
  std::string o = "String(\"";
  for (char c : msg) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default: o += c;
    }
  }
  return o + "\")";
}

// ---------------------------------------------------------------- RDB (types.rs:180-284, 377-428)

void save_str(RedisModuleIO* io, const std::string& s) { RedisModule_SaveStringBuffer(io, s.data(), s.size()); }

std::string load_str(RedisModuleIO* io) {
  size_t len = 0;
  char* p = RedisModule_LoadStringBuffer(io, &len);
  std::string s(p ? p : "", p ? len : 0);
  if (p) RedisModule_Free(p);
  return s;
}

void* index_rdb_load(RedisModuleIO* io, int encver) {  // types.rs:180-241
  if (encver != 0) return nullptr;
  std::unique_ptr<IndexValue> v(new IndexValue());
  IndexRecord& r = v->rec;
  r.name = load_str(io);
  r.mfunc_kind = load_str(io);
  r.data_dim = RedisModule_LoadUnsigned(io);
  r.m = RedisModule_LoadUnsigned(io);
  r.m_max = RedisModule_LoadUnsigned(io);
  r.m_max_0 = RedisModule_LoadUnsigned(io);
  r.ef_construction = RedisModule_LoadUnsigned(io);
  r.level_mult = RedisModule_LoadDouble(io);
  r.node_count = RedisModule_LoadUnsigned(io);
  r.max_layer = RedisModule_LoadUnsigned(io);
  uint64_t n_layers = RedisModule_LoadUnsigned(io);
  r.layers.resize(n_layers);
  for (uint64_t l = 0; l < n_layers; ++l) {
    uint64_t n = RedisModule_LoadUnsigned(io);
    r.layers[l].reserve(n);
    for (uint64_t i = 0; i < n; ++i) r.layers[l].push_back(load_str(io));
  }
  uint64_t n_nodes = RedisModule_LoadUnsigned(io);
  r.nodes.reserve(n_nodes);
  for (uint64_t i = 0; i < n_nodes; ++i) r.nodes.push_back(load_str(io));
  std::string ep = load_str(io);
  if (ep != "null") r.enterpoint = ep;  // types.rs:234-237
  return v.release();
}

void index_rdb_save(RedisModuleIO* io, void* value) {  // types.rs:243-284
  IndexValue* v = static_cast<IndexValue*>(value);
  // host data only (this runs in a fork()ed child): the live index's write-through mirror, or the loaded record
  const IndexRecord r = v->live ? v->live->to_record() : v->rec;
  save_str(io, r.name);
  save_str(io, r.mfunc_kind);
  RedisModule_SaveUnsigned(io, r.data_dim);
  RedisModule_SaveUnsigned(io, r.m);
  RedisModule_SaveUnsigned(io, r.m_max);
  RedisModule_SaveUnsigned(io, r.m_max_0);
  RedisModule_SaveUnsigned(io, r.ef_construction);
  RedisModule_SaveDouble(io, r.level_mult);
  RedisModule_SaveUnsigned(io, r.node_count);
  RedisModule_SaveUnsigned(io, r.max_layer);
  RedisModule_SaveUnsigned(io, r.layers.size());
  for (const auto& layer : r.layers) {
    RedisModule_SaveUnsigned(io, layer.size());
    for (const auto& n : layer) save_str(io, n);
  }
  RedisModule_SaveUnsigned(io, r.nodes.size());
  for (const auto& n : r.nodes) save_str(io, n);
  save_str(io, r.enterpoint ? *r.enterpoint : std::string("null"));
}

void index_free(void* value) {  // types.rs:176-178
  IndexValue* v = static_cast<IndexValue*>(value);
  for (auto it = g_index_values.begin(); it != g_index_values.end(); ++it)
    if (it->second == v) {
      g_index_values.erase(it);
      break;
    }
  if (v->live) {  // the key is gone: drop the cached index too (the reference would keep a stale INDICES entry)
    auto it = g_indices.find(v->live->name());
    if (it != g_indices.end() && it->second == v->live) g_indices.erase(it);
  }
  delete v;
}

void* node_rdb_load(RedisModuleIO* io, int encver) {  // types.rs:377-408
  if (encver != 0) return nullptr;
  std::unique_ptr<NodeValue> v(new NodeValue());
  uint64_t n = RedisModule_LoadUnsigned(io);
  v->rec.data.reserve(n);
  for (uint64_t i = 0; i < n; ++i) v->rec.data.push_back(RedisModule_LoadFloat(io));
  uint64_t n_layers = RedisModule_LoadUnsigned(io);
  v->rec.neighbors.resize(n_layers);
  for (uint64_t l = 0; l < n_layers; ++l) {
    uint64_t c = RedisModule_LoadUnsigned(io);
    v->rec.neighbors[l].reserve(c);
    for (uint64_t i = 0; i < c; ++i) v->rec.neighbors[l].push_back(load_str(io));
  }
  return v.release();
}

NodeRecord current_record(NodeValue* v) {  // From<&Node> for NodeRedis (types.rs:292-309); host mirror, never the device
  if (auto ix = v->live.lock()) {
    if (ix->contains(v->name)) return ix->node_record(v->name);
  }
  return v->rec;
}

void node_rdb_save(RedisModuleIO* io, void* value) {  // types.rs:410-428
  NodeValue* v = static_cast<NodeValue*>(value);
  const NodeRecord r = current_record(v);
  RedisModule_SaveUnsigned(io, r.data.size());
  for (float f : r.data) RedisModule_SaveFloat(io, f);
  RedisModule_SaveUnsigned(io, r.neighbors.size());
  for (const auto& layer : r.neighbors) {
    RedisModule_SaveUnsigned(io, layer.size());
    for (const auto& n : layer) save_str(io, n);
  }
}

void node_free(void* value) {  // types.rs:373-375
  NodeValue* v = static_cast<NodeValue*>(value);
  auto it = g_node_values.find(v->name);
  if (it != g_node_values.end() && it->second == v) g_node_values.erase(it);
  delete v;
}

// ---------------------------------------------------------------- keys

struct Key {
  RedisModuleCtx* ctx;
  RedisModuleKey* k;
  Key(RedisModuleCtx* c, const std::string& name, int mode) : ctx(c) {
    RedisModuleString* s = RedisModule_CreateString(c, name.data(), name.size());
    k = static_cast<RedisModuleKey*>(RedisModule_OpenKey(c, s, mode));  // NULL for a missing key opened read-only
    RedisModule_FreeString(c, s);
  }
  ~Key() {
    if (k) RedisModule_CloseKey(k);
  }
  Key(const Key&) = delete;
  Key& operator=(const Key&) = delete;

  // RedisKey::get_value (redis-module): None for an empty key, WRONGTYPE for anything that is not `type`
  template <class T>
  T* get(RedisModuleType* type) const {
    if (!k || RedisModule_KeyType(k) == REDISMODULE_KEYTYPE_EMPTY) return nullptr;
    if (RedisModule_KeyType(k) != REDISMODULE_KEYTYPE_MODULE || RedisModule_ModuleTypeGetType(k) != type)
      throw ReplyError{"WRONGTYPE Operation against a key holding the wrong kind of value"};
    return static_cast<T*>(RedisModule_ModuleTypeGetValue(k));
  }
  void set(RedisModuleType* type, void* value) const {
    if (!k || RedisModule_ModuleTypeSetValue(k, type, value) != REDISMODULE_OK) throw ReplyError{"ERR could not set the key value"};
  }
  void del() const {
    if (k) RedisModule_DeleteKey(k);
  }
};

// ---------------------------------------------------------------- argument schemas (lib.rs:37-129)

enum ArgKind { kU64, kF64Vec, kStrVec };
struct KwSpec {
  const char* name;
  ArgKind kind;
  bool required;
  uint64_t dflt;
};

struct Parsed {
  std::vector<std::string> pos;
  std::map<std::string, uint64_t> u64s;
  std::map<std::string, std::vector<double>> vecs;
  std::map<std::string, uint64_t> vec_rows;  // matrix form: leading row count
  std::map<std::string, std::vector<std::string>> strs;
};

std::string upper(std::string s) {
  for (char& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}

bool parse_u64(const std::string& s, uint64_t* out) {
  if (s.empty() || s[0] == '-') return false;
  errno = 0;
  char* end = nullptr;
  unsigned long long v = std::strtoull(s.c_str(), &end, 10);
  if (errno || !end || *end) return false;
  *out = v;
  return true;
}

bool parse_f64(const std::string& s, double* out) {
  if (s.empty()) return false;
  errno = 0;
  char* end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (!end || *end) return false;
  *out = v;
  return true;
}

// `cmd pos... [KW value | KWVEC n v1..vn]...`; keywords are matched case-insensitively, in any order.
// `matrix_kw` (optional) names a kwarg of the form `KW rows n v1..v(rows*n)`.
Parsed parse_args(const std::vector<std::string>& args, const char* cmd, size_t n_pos, const std::vector<KwSpec>& kws,
                  const char* matrix_kw = nullptr) {
  Parsed p;
  if (args.size() < 1 + n_pos) throw ReplyError{std::string("ERR wrong number of arguments for '") + cmd + "' command"};
  for (size_t i = 0; i < n_pos; ++i) p.pos.push_back(args[1 + i]);
  size_t i = 1 + n_pos;
  while (i < args.size()) {
    const std::string kw = upper(args[i]);
    const KwSpec* spec = nullptr;
    for (const KwSpec& k : kws)
      if (kw == k.name) spec = &k;
    if (!spec) throw ReplyError{"ERR unexpected argument: " + args[i]};
    ++i;
    if (spec->kind == kU64) {
      uint64_t v;
      if (i >= args.size() || !parse_u64(args[i], &v)) throw ReplyError{"ERR " + kw + " needs an unsigned integer"};
      p.u64s[kw] = v;
      ++i;
    } else if (spec->kind == kStrVec) {
      uint64_t n;
      if (i >= args.size() || !parse_u64(args[i], &n)) throw ReplyError{"ERR " + kw + " needs a count followed by that many names"};
      ++i;
      if (n > args.size() - i) throw ReplyError{"ERR " + kw + " announces more names than were given"};
      p.strs[kw] = std::vector<std::string>(args.begin() + (long)i, args.begin() + (long)(i + n));
      i += n;
    } else {
      uint64_t rows = 1, n;
      if (matrix_kw && kw == matrix_kw) {
        if (i >= args.size() || !parse_u64(args[i], &rows)) throw ReplyError{"ERR " + kw + " needs a row count"};
        ++i;
      }
      if (i >= args.size() || !parse_u64(args[i], &n)) throw ReplyError{"ERR " + kw + " needs a length followed by that many values"};
      ++i;
      if (rows > (1ull << 32) || n > (1ull << 32) || rows * n > args.size() - i)
        throw ReplyError{"ERR " + kw + " announces more values than were given"};
      std::vector<double> v(rows * n);
      for (uint64_t j = 0; j < rows * n; ++j, ++i)
        if (!parse_f64(args[i], &v[j])) throw ReplyError{"ERR " + kw + " holds a value that is not a number: " + args[i]};
      p.vecs[kw] = std::move(v);
      p.vec_rows[kw] = rows;
      p.u64s[kw + ".N"] = n;
    }
  }
  for (const KwSpec& k : kws) {
    if (k.kind == kU64 && !p.u64s.count(k.name)) {
      if (k.required) throw ReplyError{std::string("ERR missing argument: ") + k.name};
      p.u64s[k.name] = k.dflt;
    }
    if (k.kind == kF64Vec && !p.vecs.count(k.name) && k.required) throw ReplyError{std::string("ERR missing argument: ") + k.name};
    if (k.kind == kStrVec && !p.strs.count(k.name) && k.required) throw ReplyError{std::string("ERR missing argument: ") + k.name};
  }
  return p;
}

std::vector<float> narrow(const std::vector<double>& v) {  // lib.rs:345-346, 469-470: parsed as f64, narrowed to f32
  std::vector<float> f(v.size());
  for (size_t i = 0; i < v.size(); ++i) f[i] = (float)v[i];
  return f;
}

// ---------------------------------------------------------------- replies

void reply_bulk(RedisModuleCtx* ctx, const std::string& s) { RedisModule_ReplyWithStringBuffer(ctx, s.data(), s.size()); }

void reply_index(RedisModuleCtx* ctx, const IndexRecord& r) {  // From<IndexRedis> for RedisValue (types.rs:122-155)
  RedisModule_ReplyWithArray(ctx, 18);
  reply_bulk(ctx, "name"), reply_bulk(ctx, r.name);
  reply_bulk(ctx, "metric"), reply_bulk(ctx, r.mfunc_kind);
  reply_bulk(ctx, "data_dim"), RedisModule_ReplyWithLongLong(ctx, (long long)r.data_dim);
  reply_bulk(ctx, "m"), RedisModule_ReplyWithLongLong(ctx, (long long)r.m);
  reply_bulk(ctx, "ef_construction"), RedisModule_ReplyWithLongLong(ctx, (long long)r.ef_construction);
  reply_bulk(ctx, "level_mult"), RedisModule_ReplyWithDouble(ctx, r.level_mult);
  reply_bulk(ctx, "node_count"), RedisModule_ReplyWithLongLong(ctx, (long long)r.node_count);
  reply_bulk(ctx, "max_layer"), RedisModule_ReplyWithLongLong(ctx, (long long)r.max_layer);
  reply_bulk(ctx, "enterpoint");
  if (r.enterpoint) reply_bulk(ctx, *r.enterpoint);
  else RedisModule_ReplyWithNull(ctx);
}

void reply_node(RedisModuleCtx* ctx, const NodeRecord& r) {  // From<&NodeRedis> for RedisValue (types.rs:322-352)
  RedisModule_ReplyWithArray(ctx, 4);
  reply_bulk(ctx, "data");
  RedisModule_ReplyWithArray(ctx, (long)r.data.size());
  for (float f : r.data) RedisModule_ReplyWithDouble(ctx, (double)f);
  reply_bulk(ctx, "neighbors");
  RedisModule_ReplyWithArray(ctx, (long)r.neighbors.size());
  for (const auto& layer : r.neighbors) {
    RedisModule_ReplyWithArray(ctx, (long)layer.size());
    for (const auto& n : layer) reply_bulk(ctx, n);
  }
}

void reply_results(RedisModuleCtx* ctx, const std::vector<SearchResult>& res) {  // lib.rs:485-492, types.rs:445-456
  RedisModule_ReplyWithArray(ctx, (long)res.size() + 1);
  RedisModule_ReplyWithLongLong(ctx, (long long)res.size());
  for (const SearchResult& r : res) {
    RedisModule_ReplyWithArray(ctx, 4);
    reply_bulk(ctx, "similarity");
    RedisModule_ReplyWithDouble(ctx, (double)r.sim);  // f32 widened to f64 (types.rs:439)
    reply_bulk(ctx, "name");
    reply_bulk(ctx, r.name);
  }
}

// ---------------------------------------------------------------- index cache (lib.rs:229-332)

// load_index (lib.rs:229-250) + make_index (lib.rs:252-315)
std::shared_ptr<NamedIndex> load_index(RedisModuleCtx* ctx, const std::string& index_name) {
  auto it = g_indices.find(index_name);
  if (it != g_indices.end()) return it->second;
  Key key(ctx, index_name, REDISMODULE_READ);
  IndexValue* iv = key.get<IndexValue>(g_index_type);
  if (!iv) throw ReplyError{"Index: " + index_name + " does not exist"};  // lib.rs:241
  if (!iv->live) {
    std::vector<NodeValue*> values;
    try {
      std::unique_ptr<NamedIndex> ix = NamedIndex::restore(iv->rec, [&](const std::string& node_name) -> const NodeRecord* {
        Key nk(ctx, node_name, REDISMODULE_READ);
        NodeValue* nv = nk.get<NodeValue>(g_node_type);
        if (!nv) return nullptr;
        values.push_back(nv);
        return &nv->rec;
      });
      iv->live = std::shared_ptr<NamedIndex>(ix.release());
    } catch (const HNSWError& e) {
      throw ReplyError{e.what()};
    }
    for (size_t i = 0; i < values.size(); ++i) {  // node keys become handles on the live index
      values[i]->name = iv->rec.nodes[i];
      values[i]->live = iv->live;
      g_node_values[values[i]->name] = values[i];
    }
  }
  g_index_values[index_name] = iv;
  g_indices[index_name] = iv->live;
  return iv->live;
}

// update_index (lib.rs:317-332): the reference re-serialises the whole index here; the key already holds a handle on
// the live index, so only the existence check survives.
void update_index(RedisModuleCtx* ctx, const std::string& index_name, const std::shared_ptr<NamedIndex>& ix) {
  Key key(ctx, index_name, REDISMODULE_READ | REDISMODULE_WRITE);
  IndexValue* iv = key.get<IndexValue>(g_index_type);
  if (!iv) throw ReplyError{"Index: " + index_name + " does not exist"};
  iv->live = ix;
  g_index_values[index_name] = iv;
}

// write_node (lib.rs:446-460)
void write_node(RedisModuleCtx* ctx, const std::string& node_name, const std::shared_ptr<NamedIndex>& ix) {
  Key key(ctx, node_name, REDISMODULE_READ | REDISMODULE_WRITE);
  NodeValue* nv = key.get<NodeValue>(g_node_type);
  if (nv) {
    nv->name = node_name;
    nv->live = ix;
    nv->rec = NodeRecord();
  } else {
    std::unique_ptr<NodeValue> v(new NodeValue());
    v->name = node_name;
    v->live = ix;
    key.set(g_node_type, v.get());
    nv = v.release();
  }
  g_node_values[node_name] = nv;
}

// delete_node_redis (lib.rs:409-423)
void delete_node_redis(RedisModuleCtx* ctx, const std::string& node_name) {
  Key key(ctx, node_name, REDISMODULE_READ | REDISMODULE_WRITE);
  if (!key.get<NodeValue>(g_node_type)) throw ReplyError{"Node: " + node_name + " does not exist"};
  key.del();
}

// ---------------------------------------------------------------- command handlers (lib.rs:131-496)

void cmd_new(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:131-171
  Parsed p = parse_args(args, "hnsw.new", 1, {{"DIM", kU64, true, 0}, {"M", kU64, false, 5}, {"EFCON", kU64, false, 200}});
  const std::string index_name = std::string(PREFIX) + "." + p.pos[0];
  Key key(ctx, index_name, REDISMODULE_READ | REDISMODULE_WRITE);
  if (key.get<IndexValue>(g_index_type)) throw ReplyError{"Index: " + index_name + " already exists"};  // lib.rs:146-149
  std::shared_ptr<NamedIndex> ix;
  try {
    ix = std::make_shared<NamedIndex>(index_name, p.u64s["DIM"], p.u64s["M"], p.u64s["EFCON"]);
  } catch (const HNSWError& e) {
    throw ReplyError{std::string("ERR ") + e.what()};
  }
  std::unique_ptr<IndexValue> v(new IndexValue());
  v->live = ix;
  v->rec = ix->to_record();
  key.set(g_index_type, v.get());
  g_index_values[index_name] = v.release();
  g_indices[index_name] = ix;  // lib.rs:163-166
  RedisModule_ReplyWithSimpleString(ctx, "OK");
}

void cmd_get(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:173-190
  Parsed p = parse_args(args, "hnsw.get", 1, {});
  auto ix = load_index(ctx, std::string(PREFIX) + "." + p.pos[0]);
  reply_index(ctx, ix->to_record());
}

void cmd_del(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:192-227
  Parsed p = parse_args(args, "hnsw.del", 1, {});
  const std::string index_name = std::string(PREFIX) + "." + p.pos[0];
  auto ix = load_index(ctx, index_name);
  g_indices.erase(index_name);
  for (const std::string& n : ix->node_names()) delete_node_redis(ctx, n);  // lib.rs:208-210
  Key key(ctx, index_name, REDISMODULE_READ | REDISMODULE_WRITE);
  if (!key.get<IndexValue>(g_index_type)) throw ReplyError{"Index: " + p.pos[0] + " does not exist"};
  key.del();
  RedisModule_ReplyWithLongLong(ctx, 1);  // lib.rs:226 (integer 1, not "OK")
}

void cmd_node_add(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:334-368
  Parsed p = parse_args(args, "hnsw.node.add", 2, {{"DATA", kF64Vec, true, 0}});
  const std::string index_name = std::string(PREFIX) + "." + p.pos[0];
  const std::string node_name = index_name + "." + p.pos[1];
  std::vector<float> data = narrow(p.vecs["DATA"]);
  auto ix = load_index(ctx, index_name);
  try {
    ix->add_node(node_name, data.data(), data.size());  // lib.rs:356-358
  } catch (const HNSWError& e) {
    throw ReplyError{error_string(e.what())};
  }
  // lib.rs:351-353, 361-362: the reference rewrites the record of every touched node and of the new node.  Touched nodes
  // hold handles on the live index, whose host mirror add_node has already refreshed; only the new node's key is created.
  write_node(ctx, node_name, ix);
  update_index(ctx, index_name, ix);  // lib.rs:365
  RedisModule_ReplyWithSimpleString(ctx, "OK");
}

void cmd_node_del(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:370-407
  Parsed p = parse_args(args, "hnsw.node.del", 2, {});
  const std::string index_name = std::string(PREFIX) + "." + p.pos[0];
  const std::string node_name = index_name + "." + p.pos[1];
  auto ix = load_index(ctx, index_name);
  try {
    ix->delete_node(node_name);  // lib.rs:397-399 (a missing node is an error here; the reference panics at lib.rs:384)
  } catch (const HNSWError& e) {
    throw ReplyError{error_string(e.what())};
  }
  delete_node_redis(ctx, node_name);   // lib.rs:401
  update_index(ctx, index_name, ix);   // lib.rs:404
  RedisModule_ReplyWithLongLong(ctx, 1);  // lib.rs:406
}

void cmd_node_get(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:425-444
  Parsed p = parse_args(args, "hnsw.node.get", 2, {});
  const std::string node_name = std::string(PREFIX) + "." + p.pos[0] + "." + p.pos[1];
  Key key(ctx, node_name, REDISMODULE_READ);
  NodeValue* nv = key.get<NodeValue>(g_node_type);
  if (!nv) throw ReplyError{"Node: " + node_name + " does not exist"};
  try {
    reply_node(ctx, current_record(nv));
  } catch (const HNSWError& e) {
    throw ReplyError{std::string("ERR ") + e.what()};
  }
}

void cmd_search(RedisModuleCtx* ctx, const std::vector<std::string>& args) {  // lib.rs:462-496
  Parsed p = parse_args(args, "hnsw.search", 1, {{"K", kU64, false, 5}, {"QUERY", kF64Vec, true, 0}, {"EF", kU64, false, 0}});
  std::vector<float> q = narrow(p.vecs["QUERY"]);
  auto ix = load_index(ctx, std::string(PREFIX) + "." + p.pos[0]);
  std::vector<SearchResult> res;
  try {
    res = ix->search_knn(q.data(), q.size(), p.u64s["K"], (uint32_t)p.u64s["EF"]);  // lib.rs:484
  } catch (const HNSWError& e) {
    throw ReplyError{error_string(e.what())};
  }
  reply_results(ctx, res);
}

// extension: HNSW.MSEARCH {index} [K k] [EF ef] QUERIES {nq} {dim} {nq*dim values} -> array of nq HNSW.SEARCH replies
void cmd_msearch(RedisModuleCtx* ctx, const std::vector<std::string>& args) {
  Parsed p = parse_args(args, "hnsw.msearch", 1, {{"K", kU64, false, 5}, {"QUERIES", kF64Vec, true, 0}, {"EF", kU64, false, 0}},
                        "QUERIES");
  std::vector<float> q = narrow(p.vecs["QUERIES"]);
  const uint64_t nq = p.vec_rows["QUERIES"], dim = p.u64s["QUERIES.N"];
  auto ix = load_index(ctx, std::string(PREFIX) + "." + p.pos[0]);
  std::vector<std::vector<SearchResult>> res;
  try {
    res = ix->search_knn_batch(q.data(), nq, dim, p.u64s["K"], (uint32_t)p.u64s["EF"]);
  } catch (const HNSWError& e) {
    throw ReplyError{error_string(e.what())};
  }
  RedisModule_ReplyWithArray(ctx, (long)res.size());
  for (const auto& r : res) reply_results(ctx, r);
}

// extension: HNSW.NODE.MADD {index} [FAST 0|1] NODES {n} {name1..namen} DATA {n} {dim} {n*dim values}
// a NODE.ADD stream in one command (bulk load).  FAST 0 (default) builds exactly the graph the same stream of NODE.ADD
// commands builds (the device's speculative-exact builder, csrc/spec.cuh); FAST 1 uses the batched builder, whose graph
// has the same quality but is not the reference's.  Replies the number of nodes added.
void cmd_node_madd(RedisModuleCtx* ctx, const std::vector<std::string>& args) {
  Parsed p = parse_args(args, "hnsw.node.madd", 1,
                        {{"FAST", kU64, false, 0}, {"NODES", kStrVec, true, 0}, {"DATA", kF64Vec, true, 0}}, "DATA");
  const std::string index_name = std::string(PREFIX) + "." + p.pos[0];
  const std::vector<std::string>& suffixes = p.strs["NODES"];
  if (p.vec_rows["DATA"] != suffixes.size()) throw ReplyError{"ERR NODES and DATA announce different counts"};
  std::vector<std::string> names;
  names.reserve(suffixes.size());
  for (const std::string& sfx : suffixes) names.push_back(index_name + "." + sfx);
  std::vector<float> data = narrow(p.vecs["DATA"]);
  auto ix = load_index(ctx, index_name);
  try {
    ix->add_nodes(names, data.data(), p.u64s["DATA.N"], p.u64s["FAST"] != 0);
  } catch (const HNSWError& e) {
    throw ReplyError{error_string(e.what())};
  }
  for (const std::string& nn : names) write_node(ctx, nn, ix);
  update_index(ctx, index_name, ix);
  RedisModule_ReplyWithLongLong(ctx, (long long)names.size());
}

typedef void (*Handler)(RedisModuleCtx*, const std::vector<std::string>&);

template <Handler H>
int command(RedisModuleCtx* ctx, RedisModuleString** argv, int argc) {
  RedisModule_AutoMemory(ctx);  // ctx.auto_memory() (lib.rs:132 ...)
  try {
    std::vector<std::string> args;
    args.reserve((size_t)argc);
    for (int i = 0; i < argc; ++i) {
      size_t len = 0;
      const char* s = RedisModule_StringPtrLen(argv[i], &len);
      args.emplace_back(s, len);
    }
    H(ctx, args);
  } catch (const ReplyError& e) {
    RedisModule_ReplyWithError(ctx, e.msg.c_str());
  } catch (const std::exception& e) {  // never unwind into the host
    RedisModule_ReplyWithError(ctx, (std::string("ERR ") + e.what()).c_str());
  }
  return REDISMODULE_OK;
}

}  // namespace

extern "C" int RedisModule_OnLoad(RedisModuleCtx* ctx, RedisModuleString** argv, int argc) {  // redis_module! (lib.rs:498-514)
  (void)argv;
  (void)argc;
  if (RedisModule_Init(ctx, "hnsw", 1, REDISMODULE_APIVER_1) != REDISMODULE_OK) return REDISMODULE_ERR;
  static RedisModuleTypeMethods index_methods = {REDISMODULE_TYPE_METHOD_VERSION, index_rdb_load, index_rdb_save, nullptr,
                                                 nullptr, nullptr, index_free};
  static RedisModuleTypeMethods node_methods = {REDISMODULE_TYPE_METHOD_VERSION, node_rdb_load, node_rdb_save, nullptr,
                                                nullptr, nullptr, node_free};
  g_index_type = RedisModule_CreateDataType(ctx, "hnswindex", 0, &index_methods);  // types.rs:157-174
  g_node_type = RedisModule_CreateDataType(ctx, "hnswnodet", 0, &node_methods);    // types.rs:354-371
  if (!g_index_type || !g_node_type) return REDISMODULE_ERR;
  struct {
    const char* name;
    RedisModuleCmdFunc fn;
    const char* flags;
  } cmds[] = {
      {"hnsw.new", command<cmd_new>, "write"},           {"hnsw.get", command<cmd_get>, "readonly"},
      {"hnsw.del", command<cmd_del>, "write"},           {"hnsw.search", command<cmd_search>, "readonly"},
      {"hnsw.node.add", command<cmd_node_add>, "write"}, {"hnsw.node.get", command<cmd_node_get>, "readonly"},
      {"hnsw.node.del", command<cmd_node_del>, "write"}, {"hnsw.msearch", command<cmd_msearch>, "readonly"},
      {"hnsw.node.madd", command<cmd_node_madd>, "write"},
  };
  for (const auto& c : cmds)
    if (RedisModule_CreateCommand(ctx, c.name, c.fn, c.flags, 0, 0, 0) != REDISMODULE_OK) return REDISMODULE_ERR;
  return REDISMODULE_OK;  // no server-event subscription: nothing on the persistence path needs the device (see the header)
}

// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// A single-file CPU restatement of the reference's HNSW hot path (zhao-lang/redis_hnsw @ v0.2.1),
// used as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs.  Nothing under redis_hnsw_b200/ may link, import or call this file.
//
// Parity pinning: the reference is Rust and cannot be compiled in this environment (no rustc/cargo),
// so this restatement is pinned against the reference's own known-answer tests
// (src/hnsw/metrics_tests.rs:4-33 and src/hnsw/core_tests.rs:7-81, see tests/test_oracle_kat.py) and
// reviewed line by line against the cited reference lines below.  Second pin: tests/pyref_hnsw.py is an independent
// pure-Python transliteration of core.rs (with Rust's BinaryHeap sift rules); tests/test_pyref_cross_check.py requires
// both restatements to agree on graphs (list order included), update_fn sets, deletes and search results, on
// continuous data and on grid data where almost every comparison is a tie.
//
// What is restated (reference file:line):
//   metrics.rs:14-23   euclidean()        -> orc_euclidean  (AVX2 path iff dim % 32 == 0, else scalar)
//   metrics.rs:48-77   sim_func_avx_euc   -> sim_avx_order  (4x8-lane FMA accumulators, fixed hsum tree)
//   metrics.rs:79-84   sim_func_euc       -> sim_scalar     (left fold, separate mul / add roundings)
//   core.rs:322-346    Index::new         -> Oracle ctor    (m_max = m, m_max_0 = 2m, level_mult = 1/ln m)
//   core.rs:383-412    add_node           -> Oracle::add
//   core.rs:414-475    delete_node        -> Oracle::del
//   core.rs:477-486    search_knn         -> orc_search (ef passed explicitly; reference uses ef_construction)
//   core.rs:489-599    insert             -> Oracle::insert
//   core.rs:601-605    gen_random_level   -> orc_level_from_u (the uniform draw u is injected)
//   core.rs:607-675    search_level       -> Oracle::search_level
//   core.rs:677-757    select_neighbors   -> Oracle::select_neighbors
//   core.rs:759-774    connect_neighbors  -> Oracle::connect_neighbors
//   core.rs:776-822    update_node_connections -> Oracle::update_node_connections
//   core.rs:824-863    delete_node_from_neighbors -> Oracle::delete_from_neighbors
//   core.rs:865-892    search_knn_internal -> Oracle::search_knn_internal
//   core.rs:127-152    _Node::{push_levels,add_neighbor,rm_neighbor} -> add_nb / rm_nb
// Third-party behaviour that leaks into result ORDER and is therefore emulated here:
//   Rust std::collections::BinaryHeap (push = sift_up with strict `>`; pop = swap-last +
//   sift_down_to_bottom + sift_up; iteration / clone / into_vec expose the raw array order),
//   std::cmp::Reverse, ordered-float 1.0.2 OrderedFloat (total order, NaN greatest, -0.0 == 0.0).
// Nodes are dense u32 ids in insertion order (the reference keys them by name; the name<->id map
// lives in the host layer).  Deleted ids become tombstones and are never reused.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Stats {
  uint64_t n_dist = 0;  // metric calls
  uint64_t n_adj = 0;   // neighbour ids iterated in search_level (core.rs:646)
  uint64_t n_hops = 0;  // candidates expanded in search_level
  uint64_t n_ties = 0;  // heap comparisons between different nodes with equal sim
};

thread_local Stats* g_stats = nullptr;

// Ties that can change a RESULT (not just an order): two different nodes with the same sim on either side of a cut —
// the m-th / (m+1)-th candidate of select_neighbors, or the two worst members of `w` at an eviction in search_level.
// There the reference's outcome is decided by BinaryHeap internals; parity fixtures are chosen so that this stays 0
// (tests/golden/make_graph_fingerprint.py).  Diagnostic only: no behaviour depends on it.
std::atomic<uint64_t> g_cut_ties{0};     // select_neighbors: m-th and (m+1)-th candidate tie
std::atomic<uint64_t> g_order_ties{0};   // select_neighbors: two SELECTED candidates tie (the set is fixed, their order in the list is not)
std::atomic<uint64_t> g_evict_ties{0};   // search_level: the two worst members of w tie at an eviction (almost always harmless:
                                         // both sit at the far edge of an early, wide w and are evicted shortly after)

// ---------------------------------------------------------------- metric (metrics.rs)

// metrics.rs:79-84 — strict left fold, mul and add rounded separately (no contraction).
float sim_scalar(const float* a, const float* b, size_t n) {
  volatile float acc = 0.0f;  // volatile: forbid re-association / contraction by the compiler
  for (size_t i = 0; i < n; ++i) {
    volatile float d = a[i] - b[i];
    volatile float p = d * d;
    acc = acc + p;
  }
  return -acc;
}

// metrics.rs:48-77 restated lane by lane with scalar fmaf: accumulator a in 0..3, AVX lane j in 0..7,
// element index i = 32*c + 8*a + j.  Used to cross-check the intrinsic version and on hosts without AVX2.
float sim_avx_order_portable(const float* x, const float* y, size_t n) {
  float acc[4][8];
  for (int a = 0; a < 4; ++a)
    for (int j = 0; j < 8; ++j) acc[a][j] = 0.0f;
  for (size_t i = 0; i < n; i += 32)
    for (int a = 0; a < 4; ++a)
      for (int j = 0; j < 8; ++j) {
        float d = x[i + 8 * a + j] - y[i + 8 * a + j];
        acc[a][j] = std::fmaf(d, d, acc[a][j]);
      }
  float L[8];
  for (int j = 0; j < 8; ++j) L[j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);  // :71-74
  float S[4];
  for (int j = 0; j < 4; ++j) S[j] = L[j] + L[j + 4];  // hsum256_ps_avx :37-39
  // hsum_ps_sse3 :27-31: sums = v + movehdup(v) -> (S0+S1, ., S2+S3, .); then lane0 + lane2
  float r = (S[0] + S[1]) + (S[2] + S[3]);
  return -r;
}

#if defined(__x86_64__)
__attribute__((target("avx2,fma"))) float sim_avx_order_intrin(const float* a, const float* b, size_t n) {
  __m256 e1 = _mm256_setzero_ps(), e2 = _mm256_setzero_ps(), e3 = _mm256_setzero_ps(), e4 = _mm256_setzero_ps();
  for (size_t i = 0; i < n; i += 32) {
    __m256 v1 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
    e1 = _mm256_fmadd_ps(v1, v1, e1);
    __m256 v2 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8));
    e2 = _mm256_fmadd_ps(v2, v2, e2);
    __m256 v3 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16));
    e3 = _mm256_fmadd_ps(v3, v3, e3);
    __m256 v4 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24));
    e4 = _mm256_fmadd_ps(v4, v4, e4);
  }
  __m256 v = _mm256_add_ps(_mm256_add_ps(e1, e2), _mm256_add_ps(e3, e4));
  __m128 lo = _mm256_castps256_ps128(v);
  __m128 hi = _mm256_extractf128_ps(v, 1);
  lo = _mm_add_ps(lo, hi);
  __m128 shuf = _mm_movehdup_ps(lo);
  __m128 sums = _mm_add_ps(lo, shuf);
  shuf = _mm_movehl_ps(shuf, sums);
  sums = _mm_add_ss(sums, shuf);
  return -_mm_cvtss_f32(sums);
}
#endif

bool g_have_avx2 = false;
struct CpuInit {
  CpuInit() {
#if defined(__x86_64__)
    __builtin_cpu_init();
    g_have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
#endif
  }
} g_cpu_init;

inline float sim_avx_order(const float* a, const float* b, size_t n) {
#if defined(__x86_64__)
  if (g_have_avx2) return sim_avx_order_intrin(a, b, n);
#endif
  return sim_avx_order_portable(a, b, n);
}

// metrics.rs:14-23.  The reference takes the AVX path iff the host has AVX2 and len % 32 == 0; the
// oracle always follows the AVX *ordering* for len % 32 == 0 (any x86-64 the reference is deployed on
// today has AVX2), which is also the ordering the device kernels reproduce.
inline float euclidean(const float* a, const float* b, size_t n) {
  if (n % 32 == 0) return sim_avx_order(a, b, n);
  return sim_scalar(a, b, n);
}

// ---------------------------------------------------------------- OrderedFloat + BinaryHeap emulation

struct Pair {
  float sim;
  uint32_t id;
};

// ordered-float 1.0.2 `Ord for OrderedFloat`: total order, NaN greatest and equal to itself, -0.0 == 0.0.
inline int of_cmp(float a, float b) {
  if (a < b) return -1;
  if (a > b) return 1;
  if (a == b) return 0;
  bool an = a != a, bn = b != b;
  if (an && bn) return 0;
  return an ? 1 : -1;
}

inline int pair_cmp(const Pair& a, const Pair& b) {
  int c = of_cmp(a.sim, b.sim);
  if (c == 0 && a.id != b.id && g_stats) g_stats->n_ties++;
  return c;
}

// Rust std BinaryHeap<T> (array-backed max-heap).  REV = true models BinaryHeap<Reverse<T>>.
template <bool REV>
struct RustHeap {
  std::vector<Pair> data;

  static bool le(const Pair& a, const Pair& b) { return REV ? pair_cmp(b, a) <= 0 : pair_cmp(a, b) <= 0; }

  bool empty() const { return data.empty(); }
  size_t len() const { return data.size(); }
  const Pair& peek() const { return data[0]; }

  void sift_up(size_t start, size_t pos) {
    Pair elem = data[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elem, data[parent])) break;  // `if hole.element() <= hole.get(parent) { break }`
      data[pos] = data[parent];
      pos = parent;
    }
    data[pos] = elem;
  }

  void sift_down_to_bottom(size_t pos) {
    size_t end = data.size();
    size_t start = pos;
    Pair elem = data[pos];
    size_t child = 2 * pos + 1;
    while (child < end) {
      size_t right = child + 1;
      if (right < end && le(data[child], data[right])) child = right;  // right if left <= right
      data[pos] = data[child];
      pos = child;
      child = 2 * pos + 1;
    }
    data[pos] = elem;
    sift_up(start, pos);
  }

  void push(Pair p) {
    size_t old = data.size();
    data.push_back(p);
    sift_up(0, old);
  }

  Pair pop() {
    Pair item = data.back();
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }
};

using MaxHeap = RustHeap<false>;
using MinHeap = RustHeap<true>;

constexpr uint32_t NONE = 0xFFFFFFFFu;

// ---------------------------------------------------------------- the index

struct Oracle {
  int dim, m, m_max, m_max_0, ef_construction;
  double level_mult;
  uint64_t node_count = 0;
  int max_layer = 0;
  uint32_t enterpoint = NONE;

  std::vector<float> vecs;                               // [n][dim]
  std::vector<int32_t> level;                            // drawn level per node (-1 = deleted)
  std::vector<std::vector<std::vector<uint32_t>>> nbrs;  // [node][level] ordered adjacency lists
  std::vector<uint32_t> touched;                         // nodes reported through update_fn by the last mutation

  // frozen CSR snapshot for the timed read-only search path (same lists, same order)
  bool frozen = false;
  std::vector<uint64_t> f_off0;
  std::vector<uint32_t> f_nbr0;

  Oracle(int dim_, int m_, int efc) : dim(dim_), m(m_), m_max(m_), m_max_0(2 * m_), ef_construction(efc) {
    level_mult = 1.0 / std::log(1.0 * (double)m_);  // core.rs:338
  }

  const float* vec(uint32_t id) const { return vecs.data() + (size_t)id * dim; }
  uint32_t n_ids() const { return (uint32_t)level.size(); }

  float sim(const float* a, const float* b) const {
    if (g_stats) g_stats->n_dist++;
    return euclidean(a, b, (size_t)dim);
  }

  // core.rs:127-135
  void push_levels(uint32_t id, int lvl) {
    auto& nb = nbrs[id];
    while ((int)nb.size() < lvl + 1) nb.emplace_back();
  }
  // core.rs:137-143
  void add_nb(uint32_t id, int lvl, uint32_t other) {
    push_levels(id, lvl);
    auto& l = nbrs[id][lvl];
    if (std::find(l.begin(), l.end(), other) == l.end()) l.push_back(other);
  }
  // core.rs:145-152 (the reference panics if absent; we report it)
  bool rm_nb(uint32_t id, int lvl, uint32_t other) {
    auto& l = nbrs[id][lvl];
    auto it = std::find(l.begin(), l.end(), other);
    if (it == l.end()) return false;
    l.erase(it);
    return true;
  }

  const uint32_t* nb_list(uint32_t id, int lvl, size_t* n) const {
    if (frozen && lvl == 0) {
      *n = (size_t)(f_off0[id + 1] - f_off0[id]);
      return f_nbr0.data() + f_off0[id];
    }
    const auto& nb = nbrs[id];
    if ((int)nb.size() <= lvl) {  // reference: push_levels creates the empty list (core.rs:642)
      *n = 0;
      return nullptr;
    }
    *n = nb[lvl].size();
    return nb[lvl].data();
  }

  void freeze() {
    uint32_t n = n_ids();
    f_off0.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) f_off0[i + 1] = f_off0[i] + (nbrs[i].empty() ? 0 : nbrs[i][0].size());
    f_nbr0.resize(f_off0[n]);
    for (uint32_t i = 0; i < n; ++i)
      if (!nbrs[i].empty()) std::copy(nbrs[i][0].begin(), nbrs[i][0].end(), f_nbr0.begin() + f_off0[i]);
    frozen = true;
  }
  void thaw() { frozen = false; }

  // visited set: epoch-stamped array, one per searching thread
  struct Visited {
    std::vector<uint32_t> stamp;
    uint32_t epoch = 0;
    void begin(size_t n) {
      if (stamp.size() < n) stamp.resize(n, 0);
      if (++epoch == 0) {
        std::fill(stamp.begin(), stamp.end(), 0);
        epoch = 1;
      }
    }
    bool test_and_set(uint32_t id) {
      if (stamp[id] == epoch) return true;
      stamp[id] = epoch;
      return false;
    }
    bool test(uint32_t id) const { return stamp[id] == epoch; }
    void set(uint32_t id) { stamp[id] = epoch; }
  };

  Visited mut_v;  // scratch for the single-threaded mutation / single-query paths

  // core.rs:607-675.  Returns `res` (max-heap) built by pushing w's raw array in order (:670-674).
  MaxHeap search_level(const float* q, uint32_t ep, size_t ef, int lvl, Visited& v) const {
    v.begin(n_ids());
    v.set(ep);                                   // :617
    float qsim = sim(q, vec(ep));                // :621
    MaxHeap c;                                   // :625
    MinHeap w;                                   // :626
    c.data.reserve(ef + 64);
    w.data.reserve(ef + 1);
    c.push({qsim, ep});
    w.push({qsim, ep});
    while (!c.empty()) {                         // :630
      Pair cp = c.pop();                         // :631
      const Pair& fp = w.peek();                 // :632
      int brk = of_cmp(cp.sim, fp.sim);
      if (brk == 0 && cp.id != fp.id && g_stats) g_stats->n_ties++;
      if (brk < 0) break;                        // :635 strict
      if (g_stats) g_stats->n_hops++;
      size_t nn;
      const uint32_t* lst = nb_list(cp.id, lvl, &nn);  // :642-645
      for (size_t i = 0; i < nn; ++i) {          // :646 list order
        uint32_t nb = lst[i];
        if (g_stats) g_stats->n_adj++;
        if (!v.test(nb)) {                       // :648
          v.set(nb);                             // :649
          float worst = w.peek().sim;            // :651
          float e = sim(q, vec(nb));             // :652-656
          int adm = of_cmp(e, worst);
          if (adm == 0 && w.len() >= ef && g_stats) g_stats->n_ties++;
          if (adm > 0 || w.len() < ef) {         // :657
            c.push({e, nb});                     // :659
            w.push({e, nb});                     // :660
            if (w.len() > ef) {                  // :662-664
              Pair gone = w.pop();
              if (of_cmp(gone.sim, w.peek().sim) == 0 && gone.id != w.peek().id) {
                g_evict_ties++;
                if (getenv("ORC_TIE_DEBUG")) fprintf(stderr, "evict tie: ef=%zu lvl=%d gone=%u kept=%u sim=%g\n", ef, lvl, gone.id, w.peek().id, gone.sim);
              }
            }
          }
        }
      }
    }
    MaxHeap res;                                 // :670
    for (const Pair& p : w.data) res.push(p);    // :671-673 raw array order
    return res;
  }

  // core.rs:677-757 with extend_candidates = keep_pruned_connections = true (the only values ever passed)
  MaxHeap select_neighbors(uint32_t query, const MaxHeap& c, size_t mm, int lc, uint32_t ignored, Visited& v) const {
    MaxHeap r;
    MaxHeap w = c;   // :685 verbatim array copy
    MaxHeap wd;      // :686
    {
      MaxHeap ccopy = c;  // :690
      v.begin(n_ids());
      while (!ccopy.empty()) v.set(ccopy.pop().id);  // :693-696
      ccopy = c;                                       // :698
      const float* qv = vec(query);
      while (!ccopy.empty()) {
        Pair ep = ccopy.pop();                         // :700 nearest first
        size_t nn;
        const uint32_t* lst = nb_list(ep.id, lc, &nn);
        for (size_t i = 0; i < nn; ++i) {              // :702
          uint32_t en = lst[i];
          if (en == query || (ignored != NONE && en == ignored)) continue;  // :704-708
          if (!v.test(en)) {                           // :710
            float s = sim(qv, vec(en));                // :711-715
            w.push({s, en});                           // :717
            v.set(en);                                 // :718
          }
        }
      }
    }
    while (!w.empty() && r.len() < mm) {               // :724
      Pair e = w.pop();
      if (e.id == query || (ignored != NONE && e.id == ignored)) continue;  // :728-731
      if (r.empty() || pair_cmp(e, r.peek()) > 0) r.push(e);               // :733-734
      else wd.push(e);                                                      // :736
    }
    while (!wd.empty() && r.len() < mm) {              // :742
      Pair p = wd.pop();
      if (p.id == query || (ignored != NONE && p.id == ignored)) continue;
      r.push(p);                                       // :752
    }
    if (r.len() == mm && mm > 0) {                     // diagnostic: a tie across the cut (see g_cut_ties)
      float worst = r.data[0].sim;
      for (const Pair& p : r.data) worst = std::min(worst, p.sim);
      bool tie = false;
      for (const MaxHeap* rest : {&wd, &w})
        for (const Pair& p : rest->data)
          if (p.id != query && p.id != ignored && of_cmp(p.sim, worst) == 0) tie = true;
      {  // equal sims among the selected: the result SET is pinned, the list order of the tied pair is a heap accident
        std::vector<float> ss;
        for (const Pair& p : r.data) ss.push_back(p.sim);
        std::sort(ss.begin(), ss.end());
        for (size_t i = 1; i < ss.size(); ++i)
          if (of_cmp(ss[i], ss[i - 1]) == 0) g_order_ties++;
      }
      if (tie) {
        g_cut_ties++;
        if (getenv("ORC_TIE_DEBUG")) fprintf(stderr, "select tie: query=%u mm=%zu lc=%d worst=%g |w|=%zu |wd|=%zu\n", query, mm, lc, worst, w.len(), wd.len());
      }
    }
    return r;
  }

  // core.rs:759-774
  void connect_neighbors(uint32_t query, const MaxHeap& neighbors, int lvl) {
    MaxHeap nb = neighbors;
    while (!nb.empty()) {
      Pair p = nb.pop();
      add_nb(query, lvl, p.id);
      add_nb(p.id, lvl, query);
    }
  }

  // core.rs:776-822; appends every updated node to `upd`
  void update_node_connections(uint32_t node, const MaxHeap& new_nb, const MaxHeap& old_nb, int lvl, uint32_t ignored,
                               std::vector<uint32_t>& upd) {
    MaxHeap newconn = new_nb;
    std::vector<Pair> rmconn = old_nb.data;  // :785 into_vec = raw array
    upd.push_back(node);
    while (!newconn.empty()) {               // :790
      Pair np = newconn.pop();
      add_nb(node, lvl, np.id);              // :793
      add_nb(np.id, lvl, node);              // :794-795 (no cap check on the other side)
      upd.push_back(np.id);
      for (size_t i = 0; i < rmconn.size(); ++i)  // :799-801 first match, order-preserving remove
        if (rmconn[i].id == np.id) {
          rmconn.erase(rmconn.begin() + i);
          break;
        }
    }
    while (!rmconn.empty()) {                // :805 pop from the back
      Pair rp = rmconn.back();
      rmconn.pop_back();
      rm_nb(node, lvl, rp.id);               // :808
      if (ignored != NONE && rp.id == ignored) continue;  // :810-813
      rm_nb(rp.id, lvl, node);               // :815
      upd.push_back(rp.id);
    }
  }

  static int level_from_u(double u, double level_mult) {
    double x = -std::log(u) * level_mult;  // core.rs:604; `as usize` saturates
    if (!(x < 2147483647.0)) return 2147483647;
    if (x < 0) return 0;
    return (int)x;
  }

  // core.rs:383-412 + 489-599.  `lvl` is the injected level draw (ignored for the first node, which
  // draws nothing: core.rs:393-405).
  uint32_t add(const float* data, int lvl) {
    thaw();
    uint32_t id = n_ids();
    vecs.insert(vecs.end(), data, data + dim);
    nbrs.emplace_back();
    touched.clear();
    if (node_count == 0) {                  // :393-405
      level.push_back(0);
      enterpoint = id;
      max_layer = 0;   // `layers = [{node}]`; max_layer untouched (0 on a fresh index; delete keeps it >= 0)
      node_count = 1;
      return id;
    }
    level.push_back(lvl);
    insert(id, lvl);
    return id;
  }

  void insert(uint32_t query, int l) {
    Visited& v = mut_v;
    int l_max = max_layer;                  // :496
    node_count += 1;                        // :505
    const float* data = vec(query);
    uint32_t ep = enterpoint;               // :508
    int lc = l_max;
    while (lc > l) {                        // :512
      MaxHeap w = search_level(data, ep, 1, lc, v);
      ep = w.pop().id;                      // :514
      if (lc == 0) break;
      lc -= 1;
    }
    std::vector<uint32_t> upd;
    for (int lc2 = std::min(l_max, l); lc2 >= 0; --lc2) {  // :523
      MaxHeap w = search_level(data, ep, (size_t)ef_construction, lc2, v);  // :524
      MaxHeap neighbors = select_neighbors(query, w, (size_t)m, lc2, NONE, v);  // :531
      connect_neighbors(query, neighbors, lc2);                                   // :532
      for (const Pair& p : neighbors.data) upd.push_back(p.id);                   // :535-537
      while (!neighbors.empty()) {          // :540
        Pair epair = neighbors.pop();
        uint32_t e = epair.id;
        MaxHeap econn;                      // :544-558
        {
          const auto& en = nbrs[e][lc2];
          const float* ev = vec(e);
          for (uint32_t n : en) econn.push({sim(ev, vec(n)), n});
        }
        size_t cap = (lc2 == 0) ? (size_t)m_max_0 : (size_t)m_max;  // :560
        if (econn.len() > cap) {            // :561
          MaxHeap enew = select_neighbors(e, econn, cap, lc2, NONE, v);    // :568
          update_node_connections(e, enew, econn, lc2, NONE, upd);         // :569
        }
      }
      ep = w.peek().id;                     // :576
    }
    std::sort(upd.begin(), upd.end());
    upd.erase(std::unique(upd.begin(), upd.end()), upd.end());
    touched = upd;                          // :580-584 (update_fn calls; set semantics)
    if (l > l_max) {                        // :587-593
      max_layer = l;
      enterpoint = query;
    }
  }

  // core.rs:824-863
  void delete_from_neighbors(uint32_t node, int lc, std::vector<uint32_t>& upd, Visited& v) {
    std::vector<uint32_t> lst = nbrs[node][lc];  // the victim's own list is not mutated while iterating
    for (uint32_t n : lst) {
      MaxHeap nconn;
      const float* nv = vec(n);
      for (uint32_t nn : nbrs[n][lc]) nconn.push({sim(nv, vec(nn)), nn});  // :838-844
      size_t cap = (lc == 0) ? (size_t)m_max_0 : (size_t)m_max;
      MaxHeap nnew = select_neighbors(n, nconn, cap, lc, node, v);          // :853
      upd.push_back(n);
      update_node_connections(n, nnew, nconn, lc, node, upd);               // :856
    }
  }

  // core.rs:414-475.  Returns false if the id is unknown / already deleted.
  // NOTE: the reference picks the replacement enterpoint as "first element of a HashSet iterator"
  // (core.rs:453), which is not deterministic; the oracle picks the smallest id of that layer.
  bool del(uint32_t node) {
    thaw();
    if (node >= n_ids() || level[node] < 0) return false;
    touched.clear();
    node_count -= 1;
    int node_level = level[node];
    level[node] = -1;  // removes it from `nodes` and from `layers[node_level]`
    std::vector<uint32_t> upd;
    Visited& v = mut_v;
    for (int lc = 0; lc < (int)nbrs[node].size(); ++lc) delete_from_neighbors(node, lc, upd, v);  // :434-440
    std::sort(upd.begin(), upd.end());
    upd.erase(std::unique(upd.begin(), upd.end()), upd.end());
    touched = upd;
    if (enterpoint == node) {  // :449-472
      uint32_t new_ep = NONE;
      for (int lc = max_layer; lc >= 0; --lc) {
        uint32_t first = NONE;
        for (uint32_t i = 0; i < n_ids(); ++i)
          if (level[i] == lc) {
            first = i;
            break;
          }
        if (first != NONE) {
          new_ep = first;
          break;
        }
        if (max_layer > 0) max_layer -= 1;  // :460-463
      }
      enterpoint = new_ep;
    }
    (void)node_level;
    nbrs[node].clear();
    return true;
  }

  // core.rs:865-892
  size_t search_knn_internal(const float* q, size_t k, size_t ef, uint32_t* ids, float* sims, Visited& v) const {
    uint32_t ep = enterpoint;
    int lc = max_layer;
    while (lc > 0) {                        // :870
      MaxHeap w = search_level(q, ep, 1, lc, v);
      ep = w.peek().id;                     // :872
      lc -= 1;
    }
    MaxHeap w = search_level(q, ep, ef, 0, v);  // :876
    size_t n = 0;
    while (n < k && !w.empty()) {           // :879
      Pair c = w.pop();
      ids[n] = c.id;
      sims[n] = c.sim;
      ++n;
    }
    return n;
  }
};

}  // namespace

// ---------------------------------------------------------------- C ABI (ctypes; tests and bench baseline only)

extern "C" {

void* orc_create(int dim, int m, int ef_construction) { return new Oracle(dim, m, ef_construction); }
void orc_destroy(void* h) { delete (Oracle*)h; }

float orc_euclidean(const float* a, const float* b, uint64_t n) { return euclidean(a, b, (size_t)n); }
float orc_sim_avx(const float* a, const float* b, uint64_t n) { return sim_avx_order(a, b, (size_t)n); }
float orc_sim_avx_portable(const float* a, const float* b, uint64_t n) { return sim_avx_order_portable(a, b, (size_t)n); }
float orc_sim_scalar(const float* a, const float* b, uint64_t n) { return sim_scalar(a, b, (size_t)n); }
int orc_have_avx2() { return g_have_avx2 ? 1 : 0; }

// out[i] = euclidean(a[i], b[i]) for n row pairs of length dim
void orc_euclidean_batch(const float* a, const float* b, uint64_t n, uint64_t dim, float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = euclidean(a + i * dim, b + i * dim, (size_t)dim);
}

uint64_t orc_cut_ties() { return g_cut_ties.load(); }
uint64_t orc_evict_ties() { return g_evict_ties.load(); }
uint64_t orc_order_ties() { return g_order_ties.load(); }
void orc_cut_ties_reset() { g_cut_ties = 0, g_evict_ties = 0, g_order_ties = 0; }

int orc_level_from_u(double u, int m) { return Oracle::level_from_u(u, 1.0 / std::log((double)m)); }

// returns the new node id, or -1 on dim mismatch (core.rs:389-391 is checked by the caller through `dim`)
int64_t orc_add(void* h, const float* v, int dim, int level, uint64_t* stats4) {
  Oracle* o = (Oracle*)h;
  if (dim != o->dim) return -1;
  Stats st;
  g_stats = stats4 ? &st : nullptr;
  uint32_t id = o->add(v, level);
  g_stats = nullptr;
  if (stats4) {
    stats4[0] = st.n_dist;
    stats4[1] = st.n_adj;
    stats4[2] = st.n_hops;
    stats4[3] = st.n_ties;
  }
  return (int64_t)id;
}

// bulk NODE.ADD stream; levels[i] is the injected level of node i (levels[0] of an empty index is ignored)
void orc_add_batch(void* h, uint64_t n, const float* vecs, const int32_t* levels, uint64_t* stats4) {
  Oracle* o = (Oracle*)h;
  Stats st;
  g_stats = stats4 ? &st : nullptr;
  for (uint64_t i = 0; i < n; ++i) o->add(vecs + i * (size_t)o->dim, levels[i]);
  g_stats = nullptr;
  if (stats4) {
    stats4[0] = st.n_dist;
    stats4[1] = st.n_adj;
    stats4[2] = st.n_hops;
    stats4[3] = st.n_ties;
  }
}

int orc_delete(void* h, uint32_t id) { return ((Oracle*)h)->del(id) ? 0 : 1; }

uint64_t orc_touched(void* h, uint32_t* out, uint64_t cap) {
  Oracle* o = (Oracle*)h;
  uint64_t n = o->touched.size();
  for (uint64_t i = 0; i < n && i < cap; ++i) out[i] = o->touched[i];
  return n;
}

// search_knn (core.rs:477-486) with explicit ef.  Returns the number of results (0 for an empty index).
// stats4 = {n_dist, n_adj, n_hops, n_ties} for this query (may be NULL).
int orc_search(void* h, const float* q, int k, int ef, uint32_t* ids, float* sims, uint64_t* stats4) {
  Oracle* o = (Oracle*)h;
  if (o->enterpoint == NONE || o->node_count == 0) return 0;
  Stats st;
  g_stats = stats4 ? &st : nullptr;
  int n = (int)o->search_knn_internal(q, (size_t)k, (size_t)ef, ids, sims, o->mut_v);
  g_stats = nullptr;
  if (stats4) {
    stats4[0] = st.n_dist;
    stats4[1] = st.n_adj;
    stats4[2] = st.n_hops;
    stats4[3] = st.n_ties;
  }
  return n;
}

// Batch of independent queries over a frozen (read-only) graph, optionally on several host threads
// (disjoint query shards).  ids/sims are [nq][k], counts [nq], stats [nq][4] (may be NULL).
// Returns elapsed seconds of the search loop itself (steady_clock).
double orc_search_batch(void* h, uint64_t nq, const float* Q, int k, int ef, uint32_t* ids, float* sims,
                        uint32_t* counts, uint64_t* stats, int n_threads) {
  Oracle* o = (Oracle*)h;
  if (o->enterpoint == NONE || o->node_count == 0) {
    for (uint64_t i = 0; i < nq; ++i) counts[i] = 0;
    return 0.0;
  }
  if (!o->frozen) o->freeze();
  if (n_threads < 1) n_threads = 1;
  auto work = [&](uint64_t lo, uint64_t hi) {
    Oracle::Visited v;
    Stats st;
    for (uint64_t i = lo; i < hi; ++i) {
      if (stats) {
        st = Stats();
        g_stats = &st;
      }
      counts[i] = (uint32_t)o->search_knn_internal(Q + i * (size_t)o->dim, (size_t)k, (size_t)ef, ids + i * (size_t)k,
                                                   sims + i * (size_t)k, v);
      if (stats) {
        g_stats = nullptr;
        stats[i * 4 + 0] = st.n_dist;
        stats[i * 4 + 1] = st.n_adj;
        stats[i * 4 + 2] = st.n_hops;
        stats[i * 4 + 3] = st.n_ties;
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) {
    work(0, nq);
  } else {
    std::vector<std::thread> th;
    uint64_t per = (nq + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
      uint64_t lo = std::min(nq, (uint64_t)t * per), hi = std::min(nq, lo + per);
      if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// search_level on one level from a given entry point; returns the whole result set nearest-first
// (popping the returned max-heap).  Used to check the device search_layer kernel level by level.
int orc_search_level(void* h, const float* q, uint32_t ep, int ef, int level, uint32_t* ids, float* sims, int cap) {
  Oracle* o = (Oracle*)h;
  MaxHeap w = o->search_level(q, ep, (size_t)ef, level, o->mut_v);
  int n = 0;
  while (!w.empty() && n < cap) {
    Pair p = w.pop();
    ids[n] = p.id;
    sims[n] = p.sim;
    ++n;
  }
  return n;
}

// ---- getters (the pub fields lib.rs / types.rs read: core.rs:303-319)
void orc_params(void* h, int64_t* out8) {
  Oracle* o = (Oracle*)h;
  out8[0] = o->dim;
  out8[1] = o->m;
  out8[2] = o->m_max;
  out8[3] = o->m_max_0;
  out8[4] = o->ef_construction;
  out8[5] = (int64_t)o->node_count;
  out8[6] = o->max_layer;
  out8[7] = (o->enterpoint == NONE) ? -1 : (int64_t)o->enterpoint;
}
double orc_level_mult(void* h) { return ((Oracle*)h)->level_mult; }
uint64_t orc_n_ids(void* h) { return ((Oracle*)h)->n_ids(); }
int orc_node_level(void* h, uint32_t id) { return ((Oracle*)h)->level[id]; }
int orc_node_n_levels(void* h, uint32_t id) { return (int)((Oracle*)h)->nbrs[id].size(); }
uint64_t orc_node_neighbors(void* h, uint32_t id, int lvl, uint32_t* out, uint64_t cap) {
  Oracle* o = (Oracle*)h;
  if ((int)o->nbrs[id].size() <= lvl) return 0;
  const auto& l = o->nbrs[id][lvl];
  for (uint64_t i = 0; i < l.size() && i < cap; ++i) out[i] = l[i];
  return l.size();
}
void orc_node_vector(void* h, uint32_t id, float* out) {
  Oracle* o = (Oracle*)h;
  std::memcpy(out, o->vec(id), sizeof(float) * (size_t)o->dim);
}

// ---- flat graph exchange: rows are (node, level) for level in 0..levels[node], row index =
// sum_{j<node}(levels[j]+1) + level; row_offs has n_rows+1 entries; deleted nodes have levels = -1 and no rows.
void orc_graph_sizes(void* h, uint64_t* n_ids, uint64_t* n_rows, uint64_t* n_edges) {
  Oracle* o = (Oracle*)h;
  uint64_t rows = 0, edges = 0;
  for (uint32_t i = 0; i < o->n_ids(); ++i) {
    if (o->level[i] < 0) continue;
    rows += (uint64_t)o->level[i] + 1;
    for (int l = 0; l <= o->level[i] && l < (int)o->nbrs[i].size(); ++l) edges += o->nbrs[i][l].size();
  }
  *n_ids = o->n_ids();
  *n_rows = rows;
  *n_edges = edges;
}

void orc_export(void* h, int32_t* levels, uint64_t* row_offs, uint32_t* nbr, int64_t* entry, int32_t* max_layer) {
  Oracle* o = (Oracle*)h;
  uint64_t r = 0, e = 0;
  row_offs[0] = 0;
  for (uint32_t i = 0; i < o->n_ids(); ++i) {
    levels[i] = o->level[i];
    if (o->level[i] < 0) continue;
    for (int l = 0; l <= o->level[i]; ++l) {
      if (l < (int)o->nbrs[i].size())
        for (uint32_t x : o->nbrs[i][l]) nbr[e++] = x;
      row_offs[++r] = e;
    }
  }
  *entry = (o->enterpoint == NONE) ? -1 : (int64_t)o->enterpoint;
  *max_layer = o->max_layer;
}

// replace the oracle's contents with a graph built elsewhere (e.g. exported from the device index)
void orc_import(void* h, uint64_t n, const float* vecs, const int32_t* levels, const uint64_t* row_offs,
                const uint32_t* nbr, int64_t entry, int32_t max_layer) {
  Oracle* o = (Oracle*)h;
  o->thaw();
  o->vecs.assign(vecs, vecs + n * (size_t)o->dim);
  o->level.assign(levels, levels + n);
  o->nbrs.assign(n, {});
  uint64_t r = 0, cnt = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (levels[i] < 0) continue;
    cnt++;
    o->nbrs[i].resize((size_t)levels[i] + 1);
    for (int l = 0; l <= levels[i]; ++l) {
      o->nbrs[i][l].assign(nbr + row_offs[r], nbr + row_offs[r + 1]);
      ++r;
    }
  }
  o->node_count = cnt;
  o->enterpoint = entry < 0 ? NONE : (uint32_t)entry;
  o->max_layer = max_layer;
}

}  // extern "C"

// Design probe for the speculative-exact batched builder: how long are the dependency chains
// between the inserts of a window?  CPU only, set semantics of core.rs:489-599 on tie-free data (the algorithm the
// device kernels run: search -> top-m -> mirrored link -> re-selection of over-full rows).
//
// Every insert of a window is executed in stream order while its read set and write set are logged:
//   reads   search rows (row, threshold at expansion time), re-selection sweep rows (row, e, cap-th sim), strict rows
//           (rows whose whole content matters: read-modify-write targets, the re-selected row itself)
//   writes  (row, ids added, ids removed)
// Insert i depends on an earlier insert j of the window when a write of j meets a read of i
//   coarse : on the same row
//   fine   : ... and a changed id could enter the list the read fed (sim above the recorded threshold), or the read is strict
// Output per window size B: inserts with at least one dependency, longest chain (= Jacobi sweeps needed), length of
// the conflict-free prefix.
//
// Modes (environment):
//   (none)          dependency statistics per window size, for row-level / fine / op-log validation (the original probe)
//   SIM_VERIFY=B    SOUNDNESS of the device's validation rules (spec.cuh): every insert of a window of B is executed twice,
//                   against the snapshot at the window start and in stream order; whenever a rule set calls the speculative
//                   execution valid, its writes must BE the sequential ones.  Prints accepted / violations (must be 0) for
//                   row-level | fine | fine + operations, and which read kind caused the first conflict.
//                   SIM_WINDOWS=n windows per checkpoint.  SIM_RETRO=1 adds the retroactive search threshold (see search_level).
//                   SIM_UNSOUND=1 drops the strict-read rule from the third rule set: the replay must then find violations
//                   (tests/test_spec_rules_cpu.py uses it as the negative control).
//   SIM_PIPE=1      event simulation of persistent warps with a ticket and in-order self-commit (no rounds): inserts/s for
//                   W = 8..128 warps, with and without early re-execution, row-level vs fine + operations.
//                   SIM_PIPE_COMMIT_US=c cost of one commit (default 15), SIM_PIPE_QUICK=1 only the interesting corner.
//
// usage: sim_spec_build vecs.bin n dim m efc seed checkpoints...      (vecs.bin = raw f32 [n][dim])
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <random>
#include <unordered_map>
#include <vector>

static int DIM, M, EFC;
static std::vector<float> V;
static std::vector<int> LEVEL;
typedef std::vector<std::vector<std::vector<uint32_t>>> GraphT;
static GraphT NB_true, NB_snap;                              // [node][level]; the snapshot is what speculative runs see
static GraphT* GP = &NB_true;
#define NB (*GP)
// content of every row an insert writes, as it was BEFORE the insert touched it (verify mode: undo + diffs)
static std::unordered_map<uint64_t, std::vector<uint32_t>>* OLD = nullptr;
static int max_layer = 0;
static uint32_t entry = 0;

static inline float sim(uint32_t a, uint32_t b) {
  const float *x = &V[(size_t)a * DIM], *y = &V[(size_t)b * DIM];
  float acc = 0;
  for (int i = 0; i < DIM; ++i) {
    float d = x[i] - y[i];
    acc += d * d;
  }
  return -acc;
}

struct Read {
  uint64_t row;
  int kind;        // 0 strict (a row that is re-selected: its content decides), 1 search (query = the insert's node),
                   // 2 sweep (query = e), 3 append target (read-modify-write: only "does the append cross the cap" matters),
                   // 4 remove target (only the presence of the removed id matters)
  uint32_t qnode;  // node whose vector the changed ids are compared with
  float thr;       // -inf: list not full
};
struct Write {
  uint64_t row;
  std::vector<uint32_t> diff;
  int op;  // 0 replace, 1 append, 2 remove
};
struct Log {
  std::vector<Read> reads;
  std::vector<Write> writes;
};
static Log* LOG = nullptr;
static inline uint64_t key(uint32_t node, int lvl) { return ((uint64_t)lvl << 32) | node; }
static inline void remember(uint32_t node, int lvl) {
  if (OLD && !OLD->count(key(node, lvl))) (*OLD)[key(node, lvl)] = NB[node][lvl];
}

static std::vector<uint32_t> stamp;
static uint32_t epoch = 0;

typedef std::pair<float, uint32_t> P;

static bool RETRO = false;   // SIM_RETRO=1: search reads get the retroactive threshold (see search_level)
static std::vector<P> search_level(uint32_t q, uint32_t ep, int ef, int lvl) {
  const size_t first_read = LOG ? LOG->reads.size() : 0;
  std::vector<float> csim;                                        // sim of the candidate whose row read i is
  ++epoch;
  stamp[ep] = epoch;
  std::priority_queue<P> c;                                      // nearest first
  std::priority_queue<P, std::vector<P>, std::greater<P>> w;    // worst first
  float s = sim(q, ep);
  c.push({s, ep});
  w.push({s, ep});
  while (!c.empty()) {
    P cp = c.top();
    c.pop();
    if (cp.first < w.top().first) break;
    const auto& nb = NB[cp.second];
    if ((int)nb.size() <= lvl) continue;
    if (LOG) LOG->reads.push_back({key(cp.second, lvl), 1, q, (int)w.size() >= ef ? w.top().first : -INFINITY}), csim.push_back(cp.first);
    for (uint32_t n : nb[lvl]) {
      if (stamp[n] == epoch) continue;
      stamp[n] = epoch;
      float e = sim(q, n);
      if (e > w.top().first || (int)w.size() < ef) {
        c.push({e, n});
        w.push({e, n});
        if ((int)w.size() > ef) w.pop();
      }
    }
  }
  // Retroactive threshold: an id that shows up in (or vanishes from) the row of read i is harmless when it is farther than
  // every candidate expanded AFTER read i and farther than the final worst of a full list: it would sit in the list for
  // a while, never be the nearest unexpanded candidate, and be gone at the end; everything nearer evolves the same.
  if (LOG && RETRO && (int)w.size() >= ef) {
    float t = w.top().first;
    for (size_t i = csim.size(); i-- > 0;) {
      Read& rd = LOG->reads[first_read + i];
      const float below = std::nextafter(t, -INFINITY);
      if (below > rd.thr) rd.thr = below;
      t = std::min(t, csim[i]);
    }
  }
  std::vector<P> out;
  while (!w.empty()) out.push_back(w.top()), w.pop();
  std::reverse(out.begin(), out.end());
  return out;
}

static void log_write(uint32_t node, int lvl, std::vector<uint32_t> diff, int op = 0) {
  if (LOG) LOG->writes.push_back({key(node, lvl), std::move(diff), op});
}
static void log_strict(uint32_t node, int lvl, int kind = 0) {
  if (LOG) LOG->reads.push_back({key(node, lvl), kind, 0, 0});
}

static void add_nb(uint32_t a, int lvl, uint32_t b) {
  auto& l = NB[a][lvl];
  if (std::find(l.begin(), l.end(), b) == l.end()) remember(a, lvl), l.push_back(b), log_write(a, lvl, {b}, 1);
}
static void rm_nb(uint32_t a, int lvl, uint32_t b) {
  auto& l = NB[a][lvl];
  auto it = std::find(l.begin(), l.end(), b);
  if (it != l.end()) remember(a, lvl), l.erase(it), log_write(a, lvl, {b}, 2);
}

static uint64_t n_reprunes = 0;
static uint64_t KH[8];

static void reprune(uint32_t e, int lvl, int cap) {
  std::vector<uint32_t> old = NB[e][lvl];
  log_strict(e, lvl, 0);  // the row being re-selected: its whole content decides the outcome
  ++epoch;
  stamp[e] = epoch;
  std::vector<P> cand;
  for (uint32_t n : old)
    if (stamp[n] != epoch) stamp[n] = epoch, cand.push_back({sim(e, n), n});
  size_t first_sweep_read = LOG ? LOG->reads.size() : 0;
  for (uint32_t n : old) {
    if (LOG) LOG->reads.push_back({key(n, lvl), 2, e, 0});
    for (uint32_t x : NB[n][lvl])
      if (stamp[x] != epoch) stamp[x] = epoch, cand.push_back({sim(e, x), x});
  }
  std::sort(cand.begin(), cand.end(), std::greater<P>());
  if ((int)cand.size() > cap) cand.resize(cap);
  float thr = (int)cand.size() >= cap ? cand.back().first : -INFINITY;
  if (LOG)
    for (size_t i = first_sweep_read; i < LOG->reads.size(); ++i) LOG->reads[i].thr = thr;
  std::vector<uint32_t> sel;
  for (auto& p : cand) sel.push_back(p.second);
  std::vector<uint32_t> keep, add, rem;
  for (uint32_t n : old) (std::find(sel.begin(), sel.end(), n) != sel.end() ? keep : rem).push_back(n);
  for (uint32_t n : sel)
    if (std::find(old.begin(), old.end(), n) == old.end()) add.push_back(n);
  std::vector<uint32_t> nl = keep;
  nl.insert(nl.end(), add.begin(), add.end());
  remember(e, lvl);
  NB[e][lvl] = nl;
  std::vector<uint32_t> d = add;
  d.insert(d.end(), rem.begin(), rem.end());
  log_write(e, lvl, d);
  for (uint32_t x : add) log_strict(x, lvl, 3), add_nb(x, lvl, e);
  for (uint32_t x : rem) log_strict(x, lvl, 4), rm_nb(x, lvl, e);
  ++n_reprunes;
}

static void insert(uint32_t q) {
  int l = LEVEL[q];
  NB[q].resize(l + 1);
  int l_max = max_layer;
  uint32_t ep = entry;
  for (int lc = l_max; lc >= 0; --lc) {
    bool link = lc <= l;
    auto w = search_level(q, ep, link ? EFC : 1, lc);
    ep = w[0].second;
    if (!link) continue;
    int cap = lc == 0 ? 2 * M : M;
    size_t n_sel = std::min<size_t>(w.size(), M);
    std::vector<uint32_t> sel;
    for (size_t i = 0; i < n_sel; ++i) sel.push_back(w[i].second);
    remember(q, lc);
    NB[q][lc] = sel;
    log_write(q, lc, sel);
    for (uint32_t r : sel) add_nb(r, lc, q);                          // an operation on r, whatever r holds
    for (uint32_t e : sel) {
      if ((int)NB[e][lc].size() > cap) {
        reprune(e, lc, cap);
      } else if (LOG) {                                               // 6: the cap check (core.rs:561) said "fits": the row may
        const size_t base = OLD && OLD->count(key(e, lc)) ? (*OLD)[key(e, lc)].size() : 0;   // grow by `thr` ids before it says otherwise
        LOG->reads.push_back({key(e, lc), 6, (uint32_t)base, (float)(cap - (int)NB[e][lc].size())});
      }
    }
  }
  if (l > l_max) max_layer = l, entry = q;
}


// ---------------------------------------------------------------- verify mode: is the validation criterion SOUND?
// SIM_VERIFY=B: for windows of B inserts, every insert is executed twice — speculatively against the graph as it stood at
// the window start (what the device's K1 sees), and in stream order (the truth).  For each criterion: whenever it calls the
// speculative execution valid, the speculative writes must BE the true writes.  Prints violations (must be 0) and how many
// inserts each criterion accepts.
struct Run {
  Log log;
  std::unordered_map<uint64_t, std::vector<uint32_t>> before, after;
};

static void run_insert(uint32_t q, GraphT* g, Run& r, bool undo) {
  GP = g;
  LOG = &r.log;
  OLD = &r.before;
  const int ml = max_layer;
  const uint32_t en = entry;
  insert(q);
  for (auto& kv : r.before) r.after[kv.first] = NB[(uint32_t)kv.first][kv.first >> 32];
  if (undo) {
    for (auto& kv : r.before) NB[(uint32_t)kv.first][kv.first >> 32] = kv.second;
    NB[q].clear();
    max_layer = ml, entry = en;
  }
  LOG = nullptr;
  OLD = nullptr;
  GP = &NB_true;
}

static int verify(size_t n, const std::vector<size_t>& cps, int B) {
  stamp.assign(n, 0);
  NB[0].resize(1);
  size_t next = 1;
  for (size_t cp : cps) {
    for (; next < cp && next < n; ++next) insert((uint32_t)next);
    size_t accepted[3] = {0, 0, 0}, violations[3] = {0, 0, 0}, total = 0, prefix_sum[3] = {0, 0, 0};
    const int windows = getenv("SIM_WINDOWS") ? atoi(getenv("SIM_WINDOWS")) : 12;
    for (int wdx = 0; wdx < windows && next + B < n; ++wdx) {
      NB_snap = NB_true;
      std::vector<Run> truth;
      std::vector<uint32_t> ids;
      bool open[3] = {true, true, true};
      size_t pre[3] = {0, 0, 0};
      while ((int)ids.size() < B && next < n) {
        const uint32_t q = (uint32_t)next++;
        if (LEVEL[q] > max_layer) {   // raises max_layer: ends the window on the device; here it simply runs alone
          insert(q);
          NB_snap = NB_true;
          truth.clear();
          ids.clear();
          for (int c = 0; c < 3; ++c) open[c] = true, pre[c] = 0;
          continue;
        }
        Run spec, tru;
        NB_snap[q].clear();
        run_insert(q, &NB_snap, spec, true);
        run_insert(q, &NB_true, tru, false);
        // conflicts of the speculative reads with the true writes of the earlier inserts of the window
        bool conflict[3] = {false, false, false};   // coarse | fine | fine + op-log
        for (const Read& rd : spec.log.reads)
          for (size_t j = 0; j < truth.size(); ++j) {
            auto bj = truth[j].before.find(rd.row);
            if (bj == truth[j].before.end()) continue;
            const auto& before = bj->second;
            const auto& after = truth[j].after[rd.row];
            std::vector<uint32_t> changed;
            for (uint32_t x : after) if (std::find(before.begin(), before.end(), x) == before.end()) changed.push_back(x);
            const size_t n_added = changed.size();
            for (uint32_t x : before) if (std::find(after.begin(), after.end(), x) == after.end()) changed.push_back(x);
            if (changed.empty()) continue;
            conflict[0] = true;
            bool fine_hit = true, op_hit = true;
            if (rd.kind == 1 || rd.kind == 2) {
              fine_hit = rd.thr == -INFINITY;
              for (size_t c = 0; c < changed.size() && !fine_hit; ++c) {
                if (changed[c] == rd.qnode) continue;
                const float sv = sim(rd.qnode, changed[c]);
                fine_hit = rd.kind == 1 ? sv > rd.thr : sv >= rd.thr;
              }
              op_hit = fine_hit;
            } else if (rd.kind == 3 || rd.kind == 4) {
              op_hit = false;                       // append / remove of one id on a row others only appended to or removed from
            } else if (rd.kind == 6) {              // the cap check must still say "fits" on the row as it really is
              const size_t true_len = tru.before.count(rd.row) ? tru.before[rd.row].size() : after.size();
              op_hit = true_len > (size_t)rd.qnode + (size_t)rd.thr;
              (void)n_added;
            }
            if (getenv("SIM_UNSOUND") && rd.kind == 0) op_hit = false;   // negative control: drop the strict rule
            if (fine_hit) conflict[1] = true;
            if (op_hit && !conflict[2]) ++KH[rd.kind & 7];
            if (op_hit) conflict[2] = true;
          }
        // what did the two executions write?
        bool same_content = spec.after.size() == tru.after.size();
        for (auto& kv : spec.after)
          if (same_content && (!tru.after.count(kv.first) || tru.after[kv.first] != kv.second)) same_content = false;
        // op-log equality: the same ids added and removed on every row (the base row may differ), same content where a row
        // was REPLACED (the insert's own rows, re-selected rows)
        bool same_ops = spec.after.size() == tru.after.size();
        for (auto& kv : spec.after) {
          if (!same_ops) break;
          if (!tru.after.count(kv.first)) { same_ops = false; break; }
          auto diff = [](const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
            std::vector<uint32_t> d;
            for (uint32_t x : b) if (std::find(a.begin(), a.end(), x) == a.end()) d.push_back(x);
            for (uint32_t x : a) if (std::find(b.begin(), b.end(), x) == b.end()) d.push_back(x | 0x80000000u);
            std::sort(d.begin(), d.end());
            return d;
          };
          if (diff(spec.before[kv.first], kv.second) != diff(tru.before[kv.first], tru.after[kv.first])) same_ops = false;
        }
        bool replaced_same = true;
        for (const Write& w : tru.log.writes)
          if (w.op == 0 && (!spec.after.count(w.row) || spec.after[w.row] != tru.after[w.row])) replaced_same = false;
        for (const Write& w : spec.log.writes)
          if (w.op == 0 && (!tru.after.count(w.row) || spec.after[w.row] != tru.after[w.row])) replaced_same = false;
        ++total;
        for (int c = 0; c < 3; ++c) {
          if (conflict[c]) { open[c] = false; continue; }
          ++accepted[c];
          if (open[c]) ++pre[c];
          const bool ok = c < 2 ? same_content : (same_ops && replaced_same);
          if (!ok) {
            ++violations[c];
            if (violations[c] <= 3) printf("  VIOLATION criterion %d at insert %u (window position %zu)\n", c, q, ids.size());
          }
        }
        truth.push_back(std::move(tru));
        ids.push_back(q);
      }
      for (int c = 0; c < 3; ++c) prefix_sum[c] += pre[c];
    }
    printf("   first conflicting read of fine+oplog by kind: strict %llu search %llu sweep %llu lenbound %llu\n", (unsigned long long)KH[0],
           (unsigned long long)KH[1], (unsigned long long)KH[2], (unsigned long long)KH[6]);
    for (auto& k : KH) k = 0;
    printf("N=%zu B=%d inserts=%zu | accepted coarse %zu fine %zu fine+oplog %zu | violations %zu %zu %zu | mean prefix %.1f %.1f %.1f\n", cp, B,
           total, accepted[0], accepted[1], accepted[2], violations[0], violations[1], violations[2], (double)prefix_sum[0] / windows,
           (double)prefix_sum[1] / windows, (double)prefix_sum[2] / windows);
    fflush(stdout);
  }
  return 0;
}

// ---------------------------------------------------------------- pipe mode: in-order commit WITHOUT rounds
// SIM_PIPE=1: event simulation of W persistent warps.  A warp takes the next insert, executes it against the graph as it
// stands (snapshot = commit frontier at that moment, T = 0.9 ms x reads / 355), waits until it is the oldest, validates,
// commits (c = 15 us) or re-executes.  early = re-execute as soon as a commit invalidates a finished execution instead
// of waiting to become the head.  Reports inserts/s for coarse and fine+op-log validation.
static bool run_conflicts(const Run& spec, const Run& tru, int crit) {
  for (const Read& rd : spec.log.reads) {
    auto bj = tru.before.find(rd.row);
    if (bj == tru.before.end()) continue;
    const auto& before = bj->second;
    const auto& after = tru.after.at(rd.row);
    std::vector<uint32_t> changed;
    for (uint32_t x : after) if (std::find(before.begin(), before.end(), x) == before.end()) changed.push_back(x);
    for (uint32_t x : before) if (std::find(after.begin(), after.end(), x) == after.end()) changed.push_back(x);
    if (changed.empty()) continue;
    if (crit == 0) return true;
    if (rd.kind == 1 || rd.kind == 2) {
      if (rd.thr == -INFINITY) return true;
      for (uint32_t x : changed) {
        if (x == rd.qnode) continue;
        const float sv = sim(rd.qnode, x);
        if (rd.kind == 1 ? sv > rd.thr : sv >= rd.thr) return true;
      }
    } else if (rd.kind == 3 || rd.kind == 4) {
      continue;
    } else if (rd.kind == 6) {
      if (after.size() > (size_t)rd.qnode + (size_t)rd.thr) return true;
    } else {
      return true;
    }
  }
  return false;
}

struct Flight {
  uint32_t q, snap;
  double finish;
  bool finished, invalid;
  Run spec;
};

static int pipe_sim(size_t n, const std::vector<size_t>& cps) {
  stamp.assign(n, 0);
  NB[0].resize(1);
  size_t next = 1;
  const double Tbase = 0.9e-3, c_commit = (getenv("SIM_PIPE_COMMIT_US") ? atof(getenv("SIM_PIPE_COMMIT_US")) : 15.0) * 1e-6;
  const bool quick = getenv("SIM_PIPE_QUICK") != nullptr;   // only fine + operations with early re-execution, W = 16..64
  const int per_cfg = getenv("SIM_PIPE_N") ? atoi(getenv("SIM_PIPE_N")) : 1500;
  for (size_t cp : cps) {
    for (; next < cp && next < n; ++next) insert((uint32_t)next);
    for (int crit : {0, 2})
      for (int early : {0, 1})
        for (int W : {8, 16, 32, 64, 128}) {
          if (quick && (W < 16 || W > 64 || crit != 2 || early != 1)) continue;
          if (next + per_cfg + 200 >= n) return 0;
          const size_t first = next, last = next + per_cfg;
          std::vector<Flight*> fl;          // in flight, ordered by q
          double now = 0;
          uint64_t execs = 0;
          auto start = [&](Flight* f) {
            f->spec = Run();
            f->snap = (uint32_t)next;       // the commit frontier
            run_insert(f->q, &NB_true, f->spec, true);
            f->finish = now + Tbase * (double)f->spec.log.reads.size() / 355.0;
            f->finished = false, f->invalid = false;
            ++execs;
          };
          size_t ticket = next;
          while (next < last) {
            while ((int)fl.size() < W && ticket < last) {   // free warps take tickets
              Flight* f = new Flight();
              f->q = (uint32_t)ticket++;
              fl.push_back(f);
              start(f);
            }
            // next event: the earliest finish of a running execution, or the head's commit if it is ready
            Flight* head = fl.front();
            if (head->finished && !head->invalid) {
              now += c_commit;
              Run tru;
              const bool raises = LEVEL[head->q] > max_layer;
              run_insert(head->q, &NB_true, tru, false);
              ++next;
              fl.erase(fl.begin());
              for (Flight* f : fl)
                if (!f->invalid && (raises || run_conflicts(f->spec, tru, crit))) {
                  f->invalid = true;
                  if (f->finished && early) start(f);
                }
              delete head;
              continue;
            }
            if (head->finished && head->invalid) {   // (only without early re-execution)
              start(head);
              continue;
            }
            Flight* e = nullptr;
            for (Flight* f : fl)
              if (!f->finished && (!e || f->finish < e->finish)) e = f;
            now = std::max(now, e->finish);
            e->finished = true;
            if (e->invalid) start(e);                // validation at the end of the execution: stale, run again
          }
          for (Flight* f : fl) delete f;   // (none: the loop ends when everything is committed)
          printf("N=%zu %s early=%d W=%3d : %7.0f inserts/s  executions/insert %.2f\n", first, crit ? "fine+oplog" : "coarse    ", early, W,
                 (double)(last - first) / now, (double)execs / (double)(last - first));
          fflush(stdout);
        }
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 8) return 1;
  const char* path = argv[1];
  size_t n = atol(argv[2]);
  DIM = atoi(argv[3]), M = atoi(argv[4]), EFC = atoi(argv[5]);
  int seed = atoi(argv[6]);
  std::vector<size_t> cps;
  for (int i = 7; i < argc; ++i) cps.push_back(atol(argv[i]));
  V.resize(n * DIM);
  FILE* f = fopen(path, "rb");
  if (!f || fread(V.data(), 4, n * DIM, f) != n * DIM) return 2;
  fclose(f);
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(0, 1);
  LEVEL.resize(n);
  for (size_t i = 0; i < n; ++i) {
    double u = U(rng);
    if (u == 0) u = 1e-300;
    LEVEL[i] = (int)std::floor(-std::log(u) / std::log((double)M));
  }
  LEVEL[0] = 0;
  NB.resize(n);
  RETRO = getenv("SIM_RETRO") != nullptr;
  if (getenv("SIM_PIPE")) return pipe_sim(n, cps);
  if (getenv("SIM_VERIFY")) return verify(n, cps, atoi(getenv("SIM_VERIFY")));
  stamp.assign(n, 0);
  NB[0].resize(1);
  const size_t WMAX = 2048;
  size_t next = 1;
  for (size_t cp : cps) {
    for (; next < cp && next < n; ++next) insert((uint32_t)next);
    if (next + WMAX > n) break;
    // a window never contains a node that raises max_layer (it runs alone): skip those here
    std::vector<Log> logs;
    std::vector<uint32_t> ids;
    uint64_t r0 = n_reprunes;
    while (ids.size() < WMAX && next < n) {
      if (LEVEL[next] > max_layer) {
        insert((uint32_t)next++);
        continue;
      }
      logs.emplace_back();
      LOG = &logs.back();
      insert((uint32_t)next);
      LOG = nullptr;
      ids.push_back((uint32_t)next++);
    }
    double avg_r = 0, avg_w = 0;
    for (auto& L : logs) avg_r += L.reads.size(), avg_w += L.writes.size();
    printf("N=%zu window=%zu reads/insert=%.1f writes/insert=%.1f reprunes/insert=%.2f\n", cp, ids.size(), avg_r / ids.size(),
           avg_w / ids.size(), (double)(n_reprunes - r0) / ids.size());
    for (int mode = 0; mode < 4; ++mode)
      for (size_t B : {16, 64, 256, 1024}) {
        const int fine = mode & 1, oplog = mode >> 1;
        // average over the disjoint windows of size B inside the logged stretch
        double dep_frac = 0, depth_sum = 0, prefix_sum = 0, sweeps_cost = 0, sprefix_sum = 0, redo_sum = 0;
        size_t nw = ids.size() / B;
        for (size_t wdx = 0; wdx < nw; ++wdx) {
          std::unordered_map<uint64_t, std::vector<std::pair<int, const Write*>>> wr;
          std::vector<int> depth(B, 0);
          size_t n_dep = 0, prefix = B, sprefix = B, redo_in_sprefix = 0;
          int maxd = 0;
          for (size_t i = 0; i < B; ++i) {
            const Log& L = logs[wdx * B + i];
            int d = 0;
            bool search_dep = false, commit_dep = false;
            for (const Read& r : L.reads) {
              auto it = wr.find(r.row);
              if (it == wr.end()) continue;
              for (auto& jw : it->second) {
                bool hit = true;
                if (oplog && (r.kind == 3 || r.kind == 5) && jw.second->op == 1) hit = false;   // two appends commute (no cap crossing: see header)
                else if (oplog && r.kind == 4 && jw.second->op != 0) hit = false;   // remove of one id vs append / remove of another
                else if (oplog && (r.kind == 3 || r.kind == 4 || r.kind == 5)) hit = true;
                else if (fine && (r.kind == 1 || r.kind == 2)) {
                  hit = false;
                  if (r.thr == -INFINITY) hit = true;
                  else
                    for (uint32_t z : jw.second->diff)
                      if (z != r.qnode && sim(r.qnode, z) > r.thr) hit = true;
                }
                if (hit) d = std::max(d, depth[jw.first] + 1);
                if (hit) (r.kind == 1 ? search_dep : commit_dep) = true;

              }
            }
            depth[i] = d;
            if (d > 0) {
              ++n_dep;
              if (prefix == B) prefix = i;
            }
            if (sprefix == B) {
              if (search_dep) sprefix = i;
              else if (commit_dep) ++redo_in_sprefix;
            }
            maxd = std::max(maxd, d);
            for (const Write& w : L.writes) wr[w.row].push_back({(int)i, &w});
          }
          dep_frac += (double)n_dep / B;
          depth_sum += maxd + 1;
          prefix_sum += prefix;
          sprefix_sum += sprefix;
          redo_sum += redo_in_sprefix;
          // executions if every insert at chain depth d runs d+1 times (upper bound of the Jacobi re-executions)
          double ex = 0;
          for (size_t i = 0; i < B; ++i) ex += depth[i] + 1;
          sweeps_cost += ex / B;
        }
        if (nw)
          printf("  %s%s B=%4zu  dependent=%.3f  inserts/sweep=%.1f  prefix=%.1f | search-only prefix=%.1f with %.1f link-phase redos inside\n",
                 fine ? "fine  " : "coarse", oplog ? "+oplog" : "      ", B, dep_frac / nw, B / (depth_sum / nw), prefix_sum / nw, sprefix_sum / nw,
                 redo_sum / nw);
      }
    fflush(stdout);
  }
  return 0;
}

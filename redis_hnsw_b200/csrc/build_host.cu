// Mutation entry points of the C ABI (insert / delete / replication).  Kernels: build.cuh.
#define HNSW_PLAIN_BUILD_KERNELS
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/hnsw_b200.h"
#include "index.hpp"
#include "spec_launch.cuh"

namespace hnsw {

// K2 / K4 of the batched builder do no distance arithmetic: one instantiation serves every index.
static cudaError_t launch_plain(void (*k)(Graph, FastArgs), const LaunchCfg& c, const Graph& g, const FastArgs& a) {
  cudaError_t e = set_smem(k, c.smem);
  if (e != cudaSuccess) return e;
  g_launches++;
  k<<<c.grid, c.block, c.smem, c.stream>>>(g, a);
  return cudaGetLastError();
}

// level = floor(-ln(u) * level_mult), u ~ U[0,1)  (core.rs:601-605); the reference saturates at usize::MAX for
// u = 0, here the level is clamped to kMaxLevel
static constexpr int kMaxLevel = 30;
int Index::draw_level() {
  rng_state += 0x9E3779B97F4A7C15ull;  // splitmix64
  uint64_t z = rng_state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  double x = -std::log(u) * level_mult;
  if (!(x < (double)kMaxLevel)) return kMaxLevel;
  return x < 0 ? 0 : (int)x;
}

int Index::set_entry(int32_t entry_, int32_t max_layer_) {
  entry = entry_;
  max_layer = max_layer_;
  int32_t v[2] = {entry_, max_layer_};
  static_assert(kMetaEntry == 0 && kMetaMaxLayer == 1, "meta layout");
  cudaError_t e = cudaMemcpyAsync(g.meta, v, sizeof(v), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "set_entry");
  return HNSW_OK;
}

// Ids a list edit buffer holds.  Degree is unbounded in the reference (core.rs:793-795 appends back-edges without a
// cap check) and the batched builder applies a whole batch of such appends before a hub row is re-selected: rows of
// 230+ ids were measured at 70K x 768-d, M=32 (cap 64), so the buffers are sized well past the cap.
static uint32_t list_capacity(uint32_t W) { return std::max<uint32_t>(512, 16 * W); }

// list registers: the candidate list must hold ef_construction entries (search) and m_max_0 entries (re-selection)
static int build_efr(const Index& ix) {
  const int a = efr_for(ix.ef_construction), b = efr_for(ix.m_max_0);
  return (a && b) ? std::max(a, b) : 0;  // 0: one of the two does not fit a list class (m_max_0 > 1024 must fail loudly)
}

// ---------------------------------------------------------------- EXACT

// The one-warp EXACT kernels have a TMA-staged flavour (build2.cuh) for the dimensions with a staged search kernel:
// stage of ExactStage<C>::S rows + 8192 32-bit visited tags + the list buffers.  Returns true (and the launch's shared
// memory size / visited slots) when it applies.
bool Index::exact_staged(size_t* smem, size_t list_bytes, uint32_t* vis_slots) const {
  if (kind_needs_smem_query(kind) || opt_build_impl == 1) return false;
  if (build_efr(*this) == kEfrMem) return false;  // lists beyond the register classes: register-staged kernels only
  const int S = dim <= 128 ? 32 : 8;  // ExactStage<C>::S
  const uint32_t slots = 8192;
  // 32-d / 128-d rows: a second stage for the one-hop lookahead of the insert's searches (search_la.cuh)
  const size_t need = warp2_smem_bytes(dim, S, slots, 4) + ((kLookaheadInBuilders && (kind == kKindR1 || kind == kKindR4)) ? la_smem_bytes(dim) : 0) + list_bytes;
  if (need > max_smem) return false;
  *smem = need;
  *vis_slots = slots;
  return true;
}

int Index::add_exact(uint32_t first, uint32_t count, bool want_touched) {
  if (count == 0) return HNSW_OK;
  const int efr = build_efr(*this);
  if (!efr) return fail(HNSW_ERR_INVALID, "m too large for the builder (2m <= 65536)");
  if (!exact_vis_slots) exact_vis_slots = next_pow2(std::max<uint64_t>(1u << 16, (uint64_t)ef_construction * 256));
  const uint32_t lcap = list_capacity(g.W);
  const uint32_t touched_cap = want_touched ? 1u << 16 : 0;
  int rc = ensure_scratch(s_bvis, (size_t)exact_vis_slots * 4);
  if (rc) return rc;
  if ((rc = ensure_scratch(s_ctl, 64 + (size_t)kCtlWords * 4 + (size_t)touched_cap * 4))) return rc;
  uint32_t* ctl = (uint32_t*)s_ctl.p;
  cudaError_t e = cudaMemsetAsync(ctl, 0, (size_t)kCtlWords * 4, stream);
  if (e != cudaSuccess) return cuda_fail(e, "exact ctl memset");
  ExactArgs a{};
  if (efr == kEfrMem) {
    a.list_cap = (std::max(ef_construction, m_max_0) + 31) & ~31u;
    if ((rc = ensure_scratch(s_list, (size_t)2 * a.list_cap * 4))) return rc;
    a.list_mem = (uint32_t*)s_list.p;
  }
  a.first = first;
  a.count = count;
  a.m = m;
  a.cap0 = m_max_0;
  a.capU = m_max;
  a.efc = ef_construction;
  a.lcap = lcap;
  a.vis_slots = exact_vis_slots;
  a.vis = (uint32_t*)s_bvis.p;
  a.ctl = ctl;
  a.touched = want_touched ? ctl + kCtlWords : nullptr;
  a.touched_cap = touched_cap;
  size_t smem = ((size_t)((m + 31) & ~31u) + 4 * (size_t)lcap + g.W) * 4 + (kind_needs_smem_query(kind) ? (size_t)dim * 4 : 0);
  int kern = kKernExact;
  if (exact_staged(&smem, ((size_t)((m + 31) & ~31u) + 4 * (size_t)lcap + g.W) * 4, &a.vis_slots))
    kern = m_max_0 <= 64 ? kKernExact2Small : kKernExact2;  // Small: re-selection lists of <= 64 entries in 2 registers
  LaunchCfg c{1, 32, smem, stream};
  e = run(kind, kern, efr, c, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "insert_exact launch");
  uint32_t h[kCtlWords];
  e = cudaMemcpyAsync(h, ctl, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "insert_exact");
  if ((rc = pull_meta())) return rc;
  build_stats[0] += h[kCtlProgress];
  build_stats[2] += h[kCtlReprunes];
  build_stats[3] += h[kCtlDistEvals];
  if (want_touched) {
    uint32_t n = std::min(h[kCtlTouched], touched_cap);
    touched.resize(n);
    if (n) {
      e = cudaMemcpyAsync(touched.data(), a.touched, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) return cuda_fail(e, "touched D2H");
    }
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
  }
  if (device_error || h[kCtlProgress] != count) {
    int err = device_error;
    device_error = 0;
    push_meta();
    return fail(HNSW_ERR_INVALID, "exact insert stopped after %u of %u nodes (device error flags 0x%x: %s)", h[kCtlProgress],
                count, err,
                (err & kErrVisitedOverflow) ? "visited table overflow" : (err & kErrPoolExhausted) ? "overflow-row pool exhausted"
                                                                                                    : "adjacency list too long");
  }
  return HNSW_OK;
}

// ---------------------------------------------------------------- FAST

struct FastPlan {
  uint32_t slots = 0, block = 0, grid = 0;
  bool vis_smem = false;
  size_t smem = 0;
};

// HNSW_BUILD_TRACE=1: per-phase wall time of every batch that takes longer than HNSW_BUILD_TRACE_MS (default 20)
struct BatchTrace {
  bool on;
  double t0, last, ms[6] = {0, 0, 0, 0, 0, 0};
  cudaStream_t s;
  static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  explicit BatchTrace(cudaStream_t st) : on(std::getenv("HNSW_BUILD_TRACE") != nullptr), s(st) { t0 = last = on ? now() : 0; }
  void mark(int phase) {
    if (!on) return;
    cudaStreamSynchronize(s);
    double t = now();
    ms[phase] += t - last;
    last = t;
  }
  void report(uint32_t first, uint32_t count) {
    if (!on) return;
    const char* lim = std::getenv("HNSW_BUILD_TRACE_MS");
    if (now() - t0 < (lim ? std::atof(lim) : 20.0)) return;
    std::fprintf(stderr, "[build] batch first=%u count=%u setup=%.1f K1=%.1f K2=%.1f K3=%.1f K4=%.1f finish=%.1f ms\n", first, count,
                 ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
  }
};

int Index::fast_batch(uint32_t first, uint32_t count) {
  BatchTrace tr(stream);
  const int efr = build_efr(*this);
  const uint32_t W = g.W, lcap = list_capacity(W);
  // link tasks: (node, level) for level = 0 .. min(level(node), max_layer)
  std::vector<uint32_t> task_base(count), task_node, task_level;
  for (uint32_t b = 0; b < count; ++b) {
    task_base[b] = (uint32_t)task_node.size();
    int top = std::min(h_level[first + b], max_layer);
    for (int lc = 0; lc <= top; ++lc) task_node.push_back(first + b), task_level.push_back((uint32_t)lc);
  }
  const uint32_t n_tasks = (uint32_t)task_node.size();
  const uint32_t wl_cap = std::min<uint64_t>((uint64_t)n_tasks * m, 10ull * count + 1024);

  // visited tables of K1 (same policy as the search path: shared memory when it fits)
  const uint32_t slots = next_pow2(std::max<uint64_t>(1024, (uint64_t)ef_construction * 32));
  const size_t qs = kind_needs_smem_query(kind) ? (size_t)dim * 4 : 0;
  int block = 256;
  while (block > 32 && (size_t)(block / 32) * ((size_t)slots * 4 + qs) > max_smem) block /= 2;
  const bool vis_smem = (size_t)(block / 32) * ((size_t)slots * 4 + qs) <= max_smem;
  if (!vis_smem) block = 128;
  const int warps = block / 32;
  const size_t smem1 = vis_smem ? (size_t)warps * ((size_t)slots * 4 + qs) : (size_t)warps * qs;
  const int id1 = vis_smem ? kKernBuildSearchSmem : kKernBuildSearchGlobal;
  int occ = occupancy(kind, id1, efr, block, smem1);
  if (occ < 1) return fail(HNSW_ERR_CUDA, "build search kernel cannot be resident (block %d, smem %zu)", block, smem1);
  const int grid1 = (int)std::min<uint64_t>((uint64_t)num_sms * occ, (count + warps - 1) / warps);
  const uint32_t big_slots = next_pow2(std::max<uint64_t>(1u << 16, (uint64_t)slots * 8));
  const int block1b = 128, warps1b = 4;
  const int grid1b = (int)std::min<uint64_t>((uint64_t)num_sms, (count + warps1b - 1) / warps1b);
  const size_t vis1 = vis_smem ? 0 : (size_t)grid1 * warps * slots * 4;
  const size_t vis1b = (size_t)grid1b * warps1b * big_slots * 4;
  int rc = HNSW_OK;

  // scratch layout
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += al(bytes);
    return o;
  };
  const size_t o_ctl = take((size_t)kCtlWords * 4), o_tb = take((size_t)count * 4), o_tn = take((size_t)n_tasks * 4),
               o_tl = take((size_t)n_tasks * 4), o_sel = take((size_t)n_tasks * m * 4), o_cnt = take((size_t)n_tasks * 4),
               o_retry = take((size_t)count * 4), o_wn = take((size_t)wl_cap * 4), o_wlv = take((size_t)wl_cap * 4),
               o_wlen = take((size_t)wl_cap * 8), o_wold = take((size_t)wl_cap * lcap * 4),
               o_wnew = take((size_t)wl_cap * W * 4);
  {
    // Allocate once for the largest batch this add_batch call will reach (cudaMalloc / cudaFree of scratch that grows
    // with every batch of the ramp cost 30-800 ms each on the test box: profiles/r1d_build.md), lay out for this one.
    const size_t hc = std::max<size_t>(count, build_hint), ht = hc + hc / 2 + 64, hw = 10 * hc + 1024;
    const size_t want = al((size_t)kCtlWords * 4) + 2 * al(hc * 4) + 3 * al(ht * 4) + al(ht * m * 4) + 2 * al(hw * 4) + al(hw * 8) +
                        al(hw * (size_t)lcap * 4) + al(hw * (size_t)W * 4);
    if ((rc = ensure_scratch(s_build, std::max(off, want)))) return rc;
  }
  char* base = (char*)s_build.p;
  cudaError_t e = cudaMemsetAsync(base + o_ctl, 0, (size_t)kCtlWords * 4, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_tb, task_base.data(), (size_t)count * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_tn, task_node.data(), (size_t)n_tasks * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_tl, task_level.data(), (size_t)n_tasks * 4, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "build batch upload");

  FastArgs a{};
  a.first = first;
  a.n_new = count;
  a.n_tasks = n_tasks;
  a.task_base = (const uint32_t*)(base + o_tb);
  a.task_node = (const uint32_t*)(base + o_tn);
  a.task_level = (const uint32_t*)(base + o_tl);
  a.sel_ids = (uint32_t*)(base + o_sel);
  a.sel_cnt = (uint32_t*)(base + o_cnt);
  a.m = m;
  a.cap0 = m_max_0;
  a.capU = m_max;
  a.efc = ef_construction;
  a.lcap = lcap;
  a.ctl = (uint32_t*)(base + o_ctl);
  a.vis_slots = slots;
  a.vis_global = (uint32_t*)s_bvis.p;
  a.retry_list = (uint32_t*)(base + o_retry);
  a.retry_pass = 0;
  a.epoch = ++epoch;
  a.stamp0 = d_stamp0;
  a.stampU = d_stampU;
  a.wl_cap = wl_cap;
  a.wl_node = (uint32_t*)(base + o_wn);
  a.wl_level = (uint32_t*)(base + o_wlv);
  a.wl_len = (uint32_t*)(base + o_wlen);
  a.wl_old = (uint32_t*)(base + o_wold);
  a.wl_new = (uint32_t*)(base + o_wnew);

  tr.mark(0);
  // K1
  bool staged = !kind_needs_smem_query(kind) && opt_build_impl != 1;
  if (staged) {
    // TMA-staged searches with a lossy visited table (build2.cuh): no retry pass, many more warps per SM
    const int S = dim == 32 ? 32 : (dim <= 128 ? 8 : 4);  // BuildStage<C>::S
    uint32_t vslots = next_pow2(std::max<uint64_t>(1024, (uint64_t)ef_construction * 16));
    int vbits = 0;
    while ((1u << vbits) < vslots) ++vbits;
    const bool tag16 = vbits + 15 < 32 && n_ids <= (1ull << (vbits + 15));
    const size_t per_warp = warp2_smem_bytes(dim, S, vslots, tag16 ? 2 : 4);
    int blk = 64;
    while (blk > 32 && (size_t)(blk / 32) * per_warp > max_smem) blk /= 2;
    const int w2 = blk / 32;
    const size_t smem2 = (size_t)w2 * per_warp;
    const int id2 = kKernBuildSearch2 + (tag16 ? 1 : 0);
    int occ2 = smem2 <= max_smem ? occupancy(kind, id2, efr, blk, smem2) : 0;
    if (occ2 < 1) {
      staged = false;
    } else {
      FastArgs a2 = a;
      a2.vis_slots = vslots;
      LaunchCfg c2{(int)std::min<uint64_t>((uint64_t)num_sms * occ2, (count + w2 - 1) / w2), blk, smem2, stream};
      e = run(kind, id2, efr, c2, g, &a2);
      if (e != cudaSuccess) return cuda_fail(e, "build_search2 launch");
    }
  }
  if (!staged) {  // register-staged searches with an exact visited set (+ retry pass with large global-memory tables)
    if ((rc = ensure_scratch(s_bvis, vis1 + vis1b))) return rc;
    a.vis_global = (uint32_t*)s_bvis.p;
    LaunchCfg c1{grid1, block, smem1, stream};
    e = run(kind, id1, efr, c1, g, &a);
    if (e != cudaSuccess) return cuda_fail(e, "build_search launch");
    FastArgs a1b = a;
    a1b.retry_pass = 1;
    a1b.vis_slots = big_slots;
    a1b.vis_global = (uint32_t*)((char*)s_bvis.p + vis1);
    LaunchCfg c1b{grid1b, block1b, (size_t)warps1b * qs, stream};
    e = run(kind, kKernBuildSearchGlobal, efr, c1b, g, &a1b);
    if (e != cudaSuccess) return cuda_fail(e, "build_search retry launch");
  }
  tr.mark(1);
  // K2
  {
    const int blk = 256, w = blk / 32;
    LaunchCfg c{(int)std::min<uint64_t>((uint64_t)num_sms * 4, (n_tasks + w - 1) / w), blk, (size_t)w * lcap * 4, stream};
    e = launch_plain(build_link_kernel, c, g, a);
    if (e != cudaSuccess) return cuda_fail(e, "build_link launch");
  }
  tr.mark(2);
  // K3
  bool staged3 = staged;
  if (staged3) {  // re-selection on the staged machinery (build_reprune2_kernel)
    const int S = dim == 32 ? 32 : (dim <= 128 ? 8 : 4);  // BuildStage<C>::S
    const uint32_t vslots = next_pow2(std::max<uint64_t>(1024, (uint64_t)m_max_0 * 64));
    int vbits = 0;
    while ((1u << vbits) < vslots) ++vbits;
    const bool tag16 = vbits + 15 < 32 && n_ids <= (1ull << (vbits + 15));
    const size_t per_warp = warp2_smem_bytes(dim, S, vslots, tag16 ? 2 : 4) + ((size_t)lcap + 64) * 4;
    int blk = 64;
    while (blk > 32 && (size_t)(blk / 32) * per_warp > max_smem) blk /= 2;
    const int w3 = blk / 32;
    const size_t smem3 = (size_t)w3 * per_warp;
    const int efr3 = efr_for(m_max_0);
    const int id3 = kKernBuildReprune2 + (tag16 ? 1 : 0);
    const int occ3 = smem3 <= max_smem ? occupancy(kind, id3, efr3, blk, smem3) : 0;
    if (occ3 < 1) {
      staged3 = false;
    } else {
      FastArgs a3 = a;
      a3.vis_slots = vslots;
      LaunchCfg c{(int)std::min<uint64_t>((uint64_t)num_sms * occ3, (wl_cap + w3 - 1) / w3), blk, smem3, stream};
      e = run(kind, id3, efr3, c, g, &a3);
      if (e != cudaSuccess) return cuda_fail(e, "build_reprune2 launch");
    }
  }
  if (!staged3) {
    const uint32_t cap = m_max_0;
    // lossy direct-mapped table (expand_chunk_lossy): any size is correct, a larger one saves re-evaluations
    const uint32_t rslots = next_pow2(std::max<uint64_t>(1024, (uint64_t)cap * 64));
    FastArgs a3 = a;
    a3.vis_slots = rslots;
    int blk = 256;
    while (blk > 32 && (size_t)(blk / 32) * ((size_t)(rslots + lcap) * 4 + qs) > max_smem) blk /= 2;
    const int w = blk / 32;
    const size_t smem3 = (size_t)w * ((size_t)(rslots + lcap) * 4 + qs);
    if (smem3 > max_smem) return fail(HNSW_ERR_INVALID, "m too large for the re-selection kernel");
    const int efr3 = efr_for(m_max_0);  // the re-selected list holds at most m_max_0 entries: a short register list
    int occ3 = occupancy(kind, kKernBuildReprune, efr3, blk, smem3);
    if (occ3 < 1) return fail(HNSW_ERR_CUDA, "re-selection kernel cannot be resident");
    LaunchCfg c{(int)std::min<uint64_t>((uint64_t)num_sms * occ3, (wl_cap + w - 1) / w), blk, smem3, stream};
    e = run(kind, kKernBuildReprune, efr3, c, g, &a3);
    if (e != cudaSuccess) return cuda_fail(e, "build_reprune launch");
  }
  tr.mark(3);
  // K4
  {
    const int blk = 128, w = blk / 32;
    const size_t smem4 = (size_t)w * (2 * (size_t)lcap + W) * 4;
    LaunchCfg c{(int)std::min<uint64_t>((uint64_t)num_sms * 4, (wl_cap + w - 1) / w), blk, smem4, stream};
    e = launch_plain(build_apply_kernel, c, g, a);
    if (e != cudaSuccess) return cuda_fail(e, "build_apply launch");
  }
  uint32_t h[kCtlWords];
  e = cudaMemcpyAsync(h, a.ctl, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "build batch");
  tr.mark(4);
  if ((rc = pull_meta())) return rc;
  build_stats[0] += count;
  build_stats[1] += h[kCtlWlDropped] + h[kCtlSkipped];
  build_stats_ex[4 - 4] += h[kCtlWlDropped];
  build_stats_ex[5 - 4] += h[kCtlSkipped];
  build_stats_ex[6 - 4] += h[kCtlRefused];
  build_stats[2] += h[kCtlReprunes];
  build_stats[3] += h[kCtlDistEvals];
  tr.mark(5);
  tr.report(first, count);
  if (device_error) {
    int err = device_error;
    device_error = 0;
    push_meta();
    if (err & (kErrPoolExhausted | kErrVisitedOverflow))
      return fail(HNSW_ERR_INVALID, "batched insert failed (device error flags 0x%x)", err);
    // kErrListTooLong: a hub row outgrew the edit buffer; the append was skipped (edge missing one way) -- report
    return fail(HNSW_ERR_INVALID, "adjacency list longer than %u ids during batched insert", lcap);
  }
  return HNSW_OK;
}

int Index::add_fast(uint32_t first, uint32_t count) {
  const uint32_t bmax = opt_build_batch ? opt_build_batch : 8192;
  build_hint = (uint32_t)std::min<uint64_t>(bmax, std::max<uint64_t>(1, (node_count + count) / 32));
  uint32_t pos = 0;
  int rc;
  while (pos < count) {
    // batch size ramps with the graph: nodes of one batch do not see each other
    uint32_t B = (uint32_t)std::min<uint64_t>(bmax, std::max<uint64_t>(1, node_count / 32));
    B = std::min(B, count - pos);
    // a node that raises max_layer becomes the enterpoint (core.rs:587-593): it goes alone
    uint32_t cut = B;
    for (uint32_t i = 0; i < B; ++i)
      if (h_level[first + pos + i] > max_layer) {
        cut = i;
        break;
      }
    const bool solo_top = cut == 0;
    if (solo_top) B = 1;
    else B = cut;
    // overflow rows this batch can allocate at most: one per append
    uint64_t appends = (uint64_t)B * (m_max_0 + 8) * 4;
    if ((rc = ensure_pool((uint64_t)pool_used + appends))) return rc;
    if ((rc = fast_batch(first + pos, B))) return rc;
    node_count += B;
    if (solo_top && (rc = set_entry((int32_t)(first + pos), h_level[first + pos]))) return rc;
    pos += B;
  }
  return HNSW_OK;
}


// ---------------------------------------------------------------- SPEC (spec.cuh)

static cudaError_t run_spec(int kind, int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occ_only,
                            int* occ) {
  switch (kind) {
    case kKindR1: return run_spec_r1(efr, small, c, g, a, occ_only, occ);
    case kKindR4: return run_spec_r4(efr, small, c, g, a, occ_only, occ);
    case kKindR24: return run_spec_r24(efr, small, c, g, a, occ_only, occ);
  }
  return cudaErrorInvalidValue;
}

// Speculative-exact NODE.ADD stream: ids first .. first+count-1, committed strictly in order (see spec.cuh).
int Index::add_spec(uint32_t first, uint32_t count) {
  if (count == 0) return HNSW_OK;
  const int efr = build_efr(*this);
  if (!efr) return fail(HNSW_ERR_INVALID, "m too large for the builder (2m <= 65536)");
  const uint32_t lcap = list_capacity(g.W);
  const bool staged_kind = kind == kKindR1 || kind == kKindR4 || kind == kKindR24;
  const int S = dim <= 128 ? 32 : 8;  // ExactStage<C>::S
  const bool fine = opt_spec_validation != 1;
  // per slot: read records, words of row ids behind them (~350 reads of ~25 ids per insert; twice that with upper levels),
  // operations, write-log entries and words
  const uint32_t vis_slots = 8192, wmaxe = 256, rmax = 2048, rcap = fine ? 49152 : 32, ocap = 512, wcap = 8192, ring = 1024;
  const size_t list_words = (size_t)((m + 31) & ~31u) + 5 * (size_t)lcap + g.W + 3 * (size_t)wmaxe;
  const size_t smem = warp2_smem_bytes(dim, S, vis_slots, 4) + ((kLookaheadInBuilders && (kind == kKindR1 || kind == kKindR4)) ? la_smem_bytes(dim) : 0) +
                      list_words * 4;
  const bool small = m_max_0 <= 64;
  int occ = 0;
  if (staged_kind && smem <= max_smem) {
    LaunchCfg c0{1, 32, smem, stream};
    SpecArgs a0{};
    run_spec(kind, efr, small, c0, g, a0, true, &occ);
  }
  if (occ < 1) {  // dimensions without a staged kernel (or lists too long for shared memory): the one-warp EXACT stream
    int rc = add_exact(first, count, false);
    if (!rc) node_count += count;
    return rc;
  }
  const uint32_t resident = (uint32_t)std::min<uint64_t>(ring, (uint64_t)occ * num_sms);

  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_ctl = 0, o_hdr = al(kSpecCtlWords * 4), o_rdh = o_hdr + al((size_t)ring * kSpecHdrWords * 4),
               o_rdo = o_rdh + al((size_t)ring * rmax * 16), o_rd = o_rdo + al((size_t)ring * rmax * 4),
               o_okey = o_rd + al((size_t)ring * rcap * 4), o_oval = o_okey + al((size_t)ring * ocap * 4),
               o_wkey = o_oval + al((size_t)ring * ocap * 4), o_woff = o_wkey + al((size_t)ring * wmaxe * 4),
               o_wbase = o_woff + al((size_t)ring * wmaxe * 4), o_ssel = o_wbase + al((size_t)ring * wmaxe * 4),
               o_wdata = o_ssel + al((size_t)ring * ((m + 31) & ~31u) * 4), total = o_wdata + al((size_t)ring * wcap * 4);
  // K2: one list buffer per warp
  const int k2_warps = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(200 * 1024) / ((size_t)lcap * 4)));
  const size_t k2_smem = (size_t)k2_warps * lcap * 4;
  if (k2_smem > 48 * 1024) cudaFuncSetAttribute(spec_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2_smem);
  int rc = ensure_scratch(s_spec, total);
  if (rc) return rc;
  char* base = (char*)s_spec.p;
  cudaError_t e = cudaMemsetAsync(base, 0, o_rdh, stream);  // control words + slot headers
  if (e != cudaSuccess) return cuda_fail(e, "spec memset");
  SpecArgs a{};
  a.ring = ring;
  a.m = m, a.cap0 = m_max_0, a.capU = m_max, a.efc = ef_construction, a.lcap = lcap, a.vis_slots = vis_slots;
  a.rcap = rcap, a.wcap = wcap, a.wmaxe = wmaxe, a.rmax = rmax, a.ocap = ocap, a.fine = fine ? 1u : 0u;
  a.ctl = (uint32_t*)(base + o_ctl);
  a.hdr = (uint32_t*)(base + o_hdr);
  a.rdh = (uint4*)(base + o_rdh);
  a.rdo = (uint32_t*)(base + o_rdo);
  a.rd = (uint32_t*)(base + o_rd);
  a.okey = (uint32_t*)(base + o_okey);
  a.oval = (uint32_t*)(base + o_oval);
  a.wkey = (uint32_t*)(base + o_wkey);
  a.woff = (uint32_t*)(base + o_woff);
  a.wbase = (uint32_t*)(base + o_wbase);
  a.ssel = (uint32_t*)(base + o_ssel);
  a.budget_ns = opt_spec_budget_us > 0 ? (uint32_t)opt_spec_budget_us * 1000u : 0u;
  a.wdata = (uint32_t*)(base + o_wdata);

  uint32_t f = first;
  const uint32_t end = first + count;
  double ema = 4.0;  // committed inserts per round
  uint32_t h[kSpecCtlWords];
  uint32_t prev_exec = 0, prev_dist = 0, prev_repr = 0, prev_waste = 0, prev_oprows = 0, idle_rounds = 0;
  const bool trace = std::getenv("HNSW_BUILD_TRACE") != nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double k1_ms = 0, k2_ms = 0, host_ms = 0;
  if (trace)
    for (auto& x : ev) cudaEventCreate(&x);
  auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  while (f < end) {
    const double t_round = trace ? now_ms() : 0;
    uint32_t B = opt_spec_window ? opt_spec_window : (uint32_t)std::max(8.0, 0.1 * (opt_spec_mult ? opt_spec_mult : 30) * ema);
    B = std::min(std::min(B, resident), end - f);
    // a node that raises max_layer becomes the enterpoint of everything after it (core.rs:587-593): the window ends there,
    // and so does the stretch behind the window in which nodes with upper levels prepare (K1: checkpoints)
    const uint32_t ahead = opt_spec_ahead < 0 ? 0 : (opt_spec_ahead ? (uint32_t)opt_spec_ahead : 2 * B);
    uint32_t wend = f + B, pend = (uint32_t)std::min<uint64_t>((uint64_t)f + B + ahead, std::min<uint64_t>(end, (uint64_t)f + ring));
    for (uint32_t q = f; q < pend; ++q)
      if (h_level[q] > max_layer) {
        wend = std::min(wend, q + 1);
        pend = q + 1;
        break;
      }
    B = wend - f;
    if ((rc = ensure_pool((uint64_t)pool_used + (uint64_t)B * 64 + 4096))) return rc;
    a.frontier = f;
    a.count = B;
    a.ver0 = d_ver0;
    a.verU = d_verU;
    LaunchCfg c1{(int)(pend - f), 32, smem, stream};
    g_launches++;
    if (trace) cudaEventRecord(ev[0], stream);
    e = run_spec(kind, efr, small, c1, g, a, false, nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "spec_exec launch");
    if (trace) cudaEventRecord(ev[1], stream);
    g_launches++;
    spec_commit_kernel<<<1, 32 * k2_warps, k2_smem, stream>>>(g, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "spec_commit launch");
    if (trace) cudaEventRecord(ev[2], stream);
    e = cudaMemcpyAsync(h, a.ctl, sizeof(h), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "spec round");
    if (trace) {
      float a_ms = 0, b_ms = 0;
      cudaEventElapsedTime(&a_ms, ev[0], ev[1]);
      cudaEventElapsedTime(&b_ms, ev[1], ev[2]);
      k1_ms += a_ms, k2_ms += b_ms, host_ms += now_ms() - t_round;
    }
    const uint32_t committed = h[kSpecCommitted], reason = h[kSpecReason];
    pool_used = h[kSpecPoolUsed];
    max_layer = (int32_t)h[kSpecMaxLayer];
    entry = (int32_t)h[kSpecEntry];
    device_error = (int32_t)h[kSpecError];
    build_stats[0] += committed;
    // executions that were thrown away = executions - commits (a running balance: a round can commit inserts executed earlier)
    build_stats[1] = (uint64_t)((int64_t)build_stats[1] + (int64_t)(h[kSpecExecuted] - prev_exec) - (int64_t)committed);
    build_stats[2] += h[kSpecReprunesDone] - prev_repr;
    build_stats[3] += h[kSpecDistEvals] - prev_dist;
    build_stats_ex[7 - 4] += 1;
    build_stats_ex[8 - 4] += h[kSpecExecuted] - prev_exec;
    build_stats_ex[9 - 4] += h[kSpecDistWasted] - prev_waste;
    build_stats_ex[11 - 4] = std::max<uint64_t>(build_stats_ex[11 - 4], B);
    build_stats_ex[12 - 4] += h[kSpecOpRows] - prev_oprows;
    prev_oprows = h[kSpecOpRows];
    prev_exec = h[kSpecExecuted], prev_dist = h[kSpecDistEvals], prev_repr = h[kSpecReprunesDone], prev_waste = h[kSpecDistWasted];
    if (trace && (build_stats_ex[7 - 4] % 128) == 0) {
      std::fprintf(stderr, "[spec] f=%u window=%u committed=%u reason=%u ema=%.1f | last 128 rounds: K1 %.3f ms, K2 %.3f ms, round (host clock) %.3f ms\n",
                   f, B, committed, reason, ema, k1_ms / 128, k2_ms / 128, host_ms / 128);
      k1_ms = k2_ms = host_ms = 0;
      // where and for how long every warp of this round's K1 ran (slot headers still hold the round's diagnostics)
      std::vector<uint32_t> hh((size_t)ring * kSpecHdrWords);
      if (cudaMemcpy(hh.data(), a.hdr, hh.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
        uint32_t t_min = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < B; ++i) t_min = std::min(t_min, hh[(size_t)((a.frontier + i) & (ring - 1)) * kSpecHdrWords + kSpecT0]);
        std::fprintf(stderr, "[spec]   K1 warps (slot: sm, start us, run us, x = executed):");
        for (uint32_t i = 0; i < B; ++i) {
          const uint32_t* w = &hh[(size_t)((a.frontier + i) & (ring - 1)) * kSpecHdrWords];
          std::fprintf(stderr, " %u:%u,%.0f,%.0f%s", i, w[kSpecSm] & 0xFFFFu, (w[kSpecT0] - t_min) / 1e3, w[kSpecDur] / 1e3,
                       (w[kSpecSm] & 0x80000000u) ? "x" : "");
          if (w[kSpecSm] & 0x80000000u)   // executed: level, ns in searches, ns in re-selections, re-selections
            std::fprintf(stderr, "[L%d s%.0f r%.0f n%u]", h_level[a.frontier + i], w[kSpecTSearch] / 1e3, w[kSpecTSelect] / 1e3, w[kSpecReprunes]);
        }
        std::fprintf(stderr, "\n");
      }
    }
    f += committed;
    node_count += committed;
    ema = 0.8 * ema + 0.2 * committed;
    if (device_error) {
      int err = device_error;
      device_error = 0;
      push_meta();
      return fail(HNSW_ERR_INVALID, "speculative insert stopped at node %u (device error flags 0x%x: %s)", f, err,
                  (err & kErrPoolExhausted) ? "overflow-row pool exhausted" : "adjacency list too long");
    }
    if (reason == 2) {  // the head insert's write log overflowed (a very large re-selection): it goes through the EXACT kernel
      if ((rc = ensure_pool((uint64_t)pool_used + (uint64_t)(m_max_0 + 8) * 2 + 4096))) return rc;
      if ((rc = add_exact(f, 1, false))) return rc;
      node_count += 1;
      f += 1;
      build_stats_ex[10 - 4] += 1;
      // the EXACT kernel does not stamp the rows it writes: every kept log is void
      e = cudaMemsetAsync(a.hdr, 0, (size_t)ring * kSpecHdrWords * 4, stream);
      if (e != cudaSuccess) return cuda_fail(e, "spec header reset");
    } else if (reason == 4) {
      if ((rc = ensure_pool((uint64_t)g.pool_cap * 2))) return rc;
    } else if (committed == 0) {
      // an invalid head was reset by K2 and runs from scratch next round, where it cannot fail: one idle round is legal
      if (++idle_rounds > 3) return fail(HNSW_ERR_CUDA, "speculative builder made no progress at node %u (reason %u)", f, reason);
    }
    if (committed) idle_rounds = 0;
  }
  if (trace)
    for (auto& x : ev) cudaEventDestroy(x);
  return HNSW_OK;
}

// ---------------------------------------------------------------- delete_node (core.rs:414-475)

int Index::delete_node(uint32_t id) {
  if (id >= n_ids || h_level[id] < 0) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", id);  // core.rs:421
  const int efr = efr_for(m_max_0);  // delete only re-selects lists of at most m_max_0 entries
  if (!efr) return fail(HNSW_ERR_INVALID, "m too large for the builder (2m <= 65536)");
  int rc = pull_meta();  // the device owns pool_used
  if (rc) return rc;
  const uint32_t lcap = list_capacity(g.W);
  // every re-selection may append this node to up to `cap` other rows (core.rs:793-795)
  if ((rc = ensure_pool((uint64_t)pool_used + (uint64_t)(h_level[id] + 1) * lcap * 2 + 1024))) return rc;
  if (!exact_vis_slots) exact_vis_slots = next_pow2(std::max<uint64_t>(1u << 16, (uint64_t)ef_construction * 256));
  const uint32_t touched_cap = 1u << 16;
  if ((rc = ensure_scratch(s_bvis, (size_t)exact_vis_slots * 4))) return rc;
  if ((rc = ensure_scratch(s_ctl, 64 + (size_t)kCtlWords * 4 + (size_t)touched_cap * 4))) return rc;
  uint32_t* ctl = (uint32_t*)s_ctl.p;
  cudaError_t e = cudaMemsetAsync(ctl, 0, (size_t)kCtlWords * 4, stream);
  if (e != cudaSuccess) return cuda_fail(e, "delete ctl memset");
  ExactArgs a{};
  if (efr == kEfrMem) {
    a.list_cap = (m_max_0 + 31) & ~31u;
    if ((rc = ensure_scratch(s_list, (size_t)2 * a.list_cap * 4))) return rc;
    a.list_mem = (uint32_t*)s_list.p;
  }
  a.first = id;
  a.count = 1;
  a.m = m;
  a.cap0 = m_max_0;
  a.capU = m_max;
  a.efc = ef_construction;
  a.lcap = lcap;
  a.vis_slots = exact_vis_slots;
  a.vis = (uint32_t*)s_bvis.p;
  a.ctl = ctl;
  a.touched = ctl + kCtlWords;
  a.touched_cap = touched_cap;
  size_t smem = (5 * (size_t)lcap + g.W) * 4 + (kind_needs_smem_query(kind) ? (size_t)dim * 4 : 0);
  int kern = kKernDelete;
  if (exact_staged(&smem, (5 * (size_t)lcap + g.W) * 4, &a.vis_slots)) kern = kKernDelete2;
  LaunchCfg c{1, 32, smem, stream};
  e = run(kind, kern, efr, c, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "delete_exact launch");
  uint32_t h[kCtlWords];
  e = cudaMemcpyAsync(h, ctl, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "delete_exact");
  if ((rc = pull_meta())) return rc;
  build_stats[2] += h[kCtlReprunes];
  build_stats[3] += h[kCtlDistEvals];
  uint32_t nt = std::min(h[kCtlTouched], touched_cap);
  touched.resize(nt);
  if (nt) {
    e = cudaMemcpyAsync(touched.data(), a.touched, (size_t)nt * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "touched D2H");
  }
  std::sort(touched.begin(), touched.end());
  touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
  if (device_error || h[kCtlProgress] != 1) {
    int err = device_error;
    device_error = 0;
    push_meta();
    return fail(HNSW_ERR_INVALID, "delete of node %u stopped part-way (device error flags 0x%x: %s)", id, err,
                (err & kErrVisitedOverflow) ? "visited table overflow" : (err & kErrPoolExhausted) ? "overflow-row pool exhausted"
                                                                                                    : "adjacency list too long");
  }
  h_level[id] = -1;  // leaves `nodes` and its layer set (core.rs:419, 426-430)
  node_count -= 1;   // core.rs:424
  if (entry == (int32_t)id) {  // core.rs:449-472
    // The reference takes "the first element of a HashSet iterator" of the highest non-empty layer set
    // (core.rs:453: not deterministic); like the oracle we take the smallest id listed in that layer.  A node is
    // listed only in the layer of its own top level (core.rs:596).
    int32_t new_ep = -1, ml = max_layer;
    for (int lc = max_layer; lc >= 0; --lc) {
      int64_t first = -1;
      for (uint64_t i = 0; i < n_ids; ++i)
        if (h_level[i] == lc) {
          first = (int64_t)i;
          break;
        }
      if (first >= 0) {
        new_ep = (int32_t)first;
        break;
      }
      if (ml > 0) ml -= 1;  // core.rs:460-463
    }
    if ((rc = set_entry(new_ep, ml))) return rc;
  }
  return HNSW_OK;
}

// ---------------------------------------------------------------- add_node (core.rs:383-412)

int Index::add_batch(uint64_t count, const float* data, const int32_t* levels, int mode, uint32_t* first_id,
                     bool want_touched) {
  if (count == 0) return HNSW_OK;
  if (!data) return fail(HNSW_ERR_INVALID, "null data");
  if (mode != HNSW_BUILD_EXACT && mode != HNSW_BUILD_FAST && mode != HNSW_BUILD_SPEC) return fail(HNSW_ERR_INVALID, "unknown build mode %d", mode);
  if (n_ids + count >= 0x7FFFFFFFull) return fail(HNSW_ERR_INVALID, "too many nodes");
  if (!build_efr(*this)) return fail(HNSW_ERR_INVALID, "m too large for the builder (2m <= 65536)");
  int rc = pull_meta();  // the device owns pool_used
  if (rc) return rc;
  const uint32_t first = (uint32_t)n_ids;
  // levels (core.rs:495, 601-605); the first node of an empty index draws nothing and sits on level 0 (core.rs:393-405)
  std::vector<int32_t> lv(count);
  uint64_t new_upper = 0;
  for (uint64_t i = 0; i < count; ++i) {
    int32_t l = (levels && levels[i] >= 0) ? std::min<int32_t>(levels[i], kMaxLevel) : (int32_t)draw_level();
    if (node_count == 0 && i == 0) l = 0;
    lv[i] = l;
    new_upper += (uint64_t)l;
  }
  if ((rc = ensure_nodes(n_ids + count))) return rc;
  if ((rc = ensure_upper(upper_used + new_upper))) return rc;
  std::vector<uint32_t> ub(count, kEmpty);
  uint64_t u = upper_used;
  for (uint64_t i = 0; i < count; ++i)
    if (lv[i] > 0) {
      ub[i] = (uint32_t)u;
      u += (uint64_t)lv[i];
    }
  if ((rc = upload_vectors(data, first, count))) return rc;
  cudaError_t e = cudaMemcpyAsync(g.level + first, lv.data(), count * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g.upper_base + first, ub.data(), count * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "level upload");
  h_level.insert(h_level.end(), lv.begin(), lv.end());
  h_upper_base.insert(h_upper_base.end(), ub.begin(), ub.end());
  n_ids += count;
  upper_used = u;
  touched.clear();
  if (first_id) *first_id = first;

  uint32_t start = 0;
  if (node_count == 0) {  // core.rs:393-405
    if ((rc = set_entry((int32_t)first, 0))) return rc;
    node_count = 1;
    start = 1;
  }
  const uint32_t rest = (uint32_t)count - start;
  const uint64_t live_before = node_count;
  // ef_construction < m: select_neighbors at core.rs:531 is a genuine 2-hop sweep (build.cuh header), which only the
  // EXACT kernels compute; the batched builder would link ef_construction instead of m neighbours per node
  if (mode == HNSW_BUILD_EXACT || ef_construction < m || build_efr(*this) == kEfrMem) {
    // overflow rows: the stream can allocate one per append in the worst case, but rows over W ids are rare (1-3 % of the
    // nodes): reserve for a bounded burst and let long streams run in pieces
    for (uint32_t pos = 0; pos < rest && !rc; pos += 65536) {
      const uint32_t piece = std::min<uint32_t>(65536, rest - pos);
      if ((rc = pull_meta())) break;  // the device owns pool_used
      if ((rc = ensure_pool((uint64_t)pool_used + (uint64_t)piece * (m_max_0 + 8) * 2 + 4096))) break;
      const uint64_t seen = build_stats[0];
      rc = add_exact(first + start + pos, piece, want_touched);
      node_count += rc ? (build_stats[0] - seen) : piece;   // on failure: the inserts the kernel completed
    }
  } else if (mode == HNSW_BUILD_SPEC) {
    rc = add_spec(first + start, rest);
  } else {
    rc = add_fast(first + start, rest);
  }
  if (rc) {
    // The ids were handed out before the kernels ran.  Nodes that were not (completely) inserted must not stay live: they
    // become tombstones (level -1) on host and device, so that no later call treats them as members of the graph and the
    // host layers can resynchronise their name tables from n_ids (orphan ids after a failed add).
    const uint64_t first_bad = std::min<uint64_t>(n_ids, (uint64_t)first + start + (node_count - live_before));
    for (uint64_t i = first_bad; i < n_ids; ++i) h_level[i] = -1;
    if (first_bad < n_ids) cudaMemsetAsync(g.level + first_bad, 0xFF, (n_ids - first_bad) * 4, stream);
    cudaStreamSynchronize(stream);
    const std::string keep = g_last_error;
    pull_meta();
    g_last_error = keep;
  }
  return rc;
}

}  // namespace hnsw

using namespace hnsw;

extern "C" {

int hnsw_index_add(hnsw_index_t* idx, const float* data, uint64_t n, int32_t level, uint32_t* out_id) {
  IDX_OR_FAIL(idx)
  if (n != ix.dim) return fail(HNSW_ERR_DIM_MISMATCH, "data dimension: %llu does not match Index", (unsigned long long)n);  // core.rs:390
  return ix.add_batch(1, data, &level, HNSW_BUILD_EXACT, out_id, true);
}

int hnsw_index_add_batch(hnsw_index_t* idx, uint64_t count, const float* data, const int32_t* levels, int mode,
                         uint32_t* first_id) {
  IDX_OR_FAIL(idx)
  return ix.add_batch(count, data, levels, mode, first_id, false);
}

int hnsw_index_touched(hnsw_index_t* idx, uint32_t* ids, uint64_t cap, uint64_t* n) {
  IDX_OR_FAIL(idx)
  if (n) *n = ix.touched.size();
  for (uint64_t i = 0; i < ix.touched.size() && i < cap; ++i) ids[i] = ix.touched[i];
  return HNSW_OK;
}

int hnsw_index_delete(hnsw_index_t* idx, uint32_t id) {
  IDX_OR_FAIL(idx)
  return ix.delete_node(id);
}

int hnsw_index_build_stats(hnsw_index_t* idx, uint64_t* out4) {
  IDX_OR_FAIL(idx)
  for (int i = 0; i < 4; ++i) out4[i] = ix.build_stats[i];
  return HNSW_OK;
}

int hnsw_index_build_stats_ex(hnsw_index_t* idx, uint64_t* out, uint32_t cap, uint32_t* n) {
  IDX_OR_FAIL(idx)
  if (n) *n = 13;
  for (uint32_t i = 0; i < 13 && i < cap; ++i) out[i] = i < 4 ? ix.build_stats[i] : ix.build_stats_ex[i - 4];
  return HNSW_OK;
}

int hnsw_index_device_buffers(hnsw_index_t* idx, hnsw_device_buffer_t* out, uint32_t cap, uint32_t* n) {
  IDX_OR_FAIL(idx)
  const uint32_t W = ix.g.W;
  hnsw_device_buffer_t b[9] = {
      {ix.g.vecs, ix.cap_nodes * (uint64_t)ix.dim * 4}, {ix.g.adj0, ix.cap_nodes * (uint64_t)W * 4},
      {ix.g.ovf0, ix.cap_nodes * 4},                    {ix.g.upper_base, ix.cap_nodes * 4},
      {ix.g.level, ix.cap_nodes * 4},                   {ix.g.adjU, ix.cap_upper * (uint64_t)W * 4},
      {ix.g.ovfU, ix.cap_upper * 4},                    {ix.g.pool, (uint64_t)ix.g.pool_cap * 128},
      {ix.g.meta, sizeof(int32_t) * kMetaCount}};
  if (n) *n = 9;
  for (uint32_t i = 0; i < 9 && i < cap; ++i) out[i] = b[i];
  return HNSW_OK;
}

// layout8 = {cap_nodes, cap_upper, pool_cap, n_ids, node_count, upper_used, 0, 0}
int hnsw_index_replica_layout(hnsw_index_t* idx, uint64_t* layout8) {
  IDX_OR_FAIL(idx)
  layout8[0] = ix.cap_nodes;
  layout8[1] = ix.cap_upper;
  layout8[2] = ix.g.pool_cap;
  layout8[3] = ix.n_ids;
  layout8[4] = ix.node_count;
  layout8[5] = ix.upper_used;
  layout8[6] = layout8[7] = 0;
  return HNSW_OK;
}

int hnsw_index_prepare_replica(hnsw_index_t* idx, const uint64_t* layout8) {
  IDX_OR_FAIL(idx)
  if (ix.cap_nodes > layout8[0] || ix.cap_upper > layout8[1] || ix.g.pool_cap > layout8[2])
    return fail(HNSW_ERR_INVALID, "replica already holds larger buffers than the source; create a fresh index");
  // allocate exactly the source capacities so the buffers can be overwritten byte for byte
  int rc;
  if (ix.cap_nodes < layout8[0]) {
    ix.cap_nodes = 0;  // forget the small initial allocation: grow_buf copies min(old,new)=0 bytes
    void* olds[] = {ix.g.vecs, ix.g.adj0, ix.g.ovf0, ix.g.upper_base, ix.g.level, ix.d_stamp0};
    for (void* p : olds) cudaFree(p);
    ix.g.vecs = nullptr, ix.g.adj0 = nullptr, ix.g.ovf0 = nullptr, ix.g.upper_base = nullptr, ix.g.level = nullptr;
    ix.d_stamp0 = nullptr;
    if ((rc = ix.ensure_nodes(layout8[0]))) return rc;
  }
  if (ix.cap_upper < layout8[1]) {
    cudaFree(ix.g.adjU), cudaFree(ix.g.ovfU), cudaFree(ix.d_stampU);
    ix.g.adjU = nullptr, ix.g.ovfU = nullptr, ix.d_stampU = nullptr, ix.cap_upper = 0;
    if ((rc = ix.ensure_upper(layout8[1]))) return rc;
  }
  if (ix.g.pool_cap < layout8[2]) {
    cudaFree(ix.g.pool);
    ix.g.pool = nullptr, ix.g.pool_cap = 0;
    if ((rc = ix.ensure_pool(layout8[2]))) return rc;
  }
  if (ix.cap_nodes != layout8[0] || ix.cap_upper != layout8[1] || ix.g.pool_cap != layout8[2])
    return fail(HNSW_ERR_INVALID, "replica capacities do not match the source layout");
  ix.n_ids = layout8[3];
  ix.node_count = layout8[4];
  ix.upper_used = layout8[5];
  return HNSW_OK;
}

// after the device buffers were overwritten with the source's: rebuild the host mirrors from device state
int hnsw_index_adopt_replica(hnsw_index_t* idx) {
  IDX_OR_FAIL(idx)
  int rc = ix.pull_meta();
  if (rc) return rc;
  ix.h_level.resize(ix.n_ids);
  ix.h_upper_base.resize(ix.n_ids);
  cudaError_t e = cudaSuccess;
  if (ix.n_ids) {
    e = cudaMemcpyAsync(ix.h_level.data(), ix.g.level, ix.n_ids * 4, cudaMemcpyDeviceToHost, ix.stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ix.h_upper_base.data(), ix.g.upper_base, ix.n_ids * 4, cudaMemcpyDeviceToHost, ix.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix.stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "adopt_replica");
  ix.touched.clear();
  return HNSW_OK;
}

}  // extern "C"

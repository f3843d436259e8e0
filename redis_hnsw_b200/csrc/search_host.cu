// Host-side dispatch of the search kernels (search.cuh) and the search entry points of the C ABI.
#include <algorithm>
#include <cstring>
#include <cmath>

#include "../../include/hnsw_b200.h"
#include "index.hpp"

namespace hnsw {

uint32_t next_pow2(uint64_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

bool kind_needs_smem_query(int kind) { return kind == kKindGeneric || kind == kKindScalar; }

cudaError_t run(int kind, int id, int efr, const LaunchCfg& c, const Graph& g, const void* args) {
  g_launches++;
  KernelArgs ka{&g, args};
  switch (kind) {
    case kKindR1: return run_kind_r1(id, efr, c, ka, false, nullptr);
    case kKindR4: return run_kind_r4(id, efr, c, ka, false, nullptr);
    case kKindR24: return run_kind_r24(id, efr, c, ka, false, nullptr);
    case kKindGeneric: return run_kind_generic(id, efr, c, ka, false, nullptr);
    default: return run_kind_scalar(id, efr, c, ka, false, nullptr);
  }
}

int occupancy(int kind, int id, int efr, int block, size_t smem) {
  LaunchCfg c{1, block, smem, nullptr};
  KernelArgs ka{nullptr, nullptr};
  int occ = 0;
  switch (kind) {
    case kKindR1: run_kind_r1(id, efr, c, ka, true, &occ); break;
    case kKindR4: run_kind_r4(id, efr, c, ka, true, &occ); break;
    case kKindR24: run_kind_r24(id, efr, c, ka, true, &occ); break;
    case kKindGeneric: run_kind_generic(id, efr, c, ka, true, &occ); break;
    default: run_kind_scalar(id, efr, c, ka, true, &occ); break;
  }
  return occ;
}

// Per-query visited-table slots for the shared-memory pass.  Starts at 32 slots per unit of ef and doubles when
// the previous batch at this ef sent more than 1 % of its queries to the retry pass.
uint32_t Index::pick_vis_slots(uint32_t ef) {
  if (opt_vis_slots) return opt_vis_slots;
  if (auto_vis_ef != ef) {
    auto_vis_ef = ef;
    auto_vis_slots = next_pow2(std::max<uint64_t>(1024, (uint64_t)ef * 32));
    if (h_retry_seen) h_retry_seen[0] = h_retry_seen[1] = 0;
  } else if (h_retry_seen && h_retry_seen[1] > 0) {
    if ((uint64_t)h_retry_seen[0] * 100 > (uint64_t)h_retry_seen[1] && auto_vis_slots < (1u << 20)) auto_vis_slots *= 2;
    h_retry_seen[0] = h_retry_seen[1] = 0;
  }
  return auto_vis_slots;
}

int Index::search_device(uint64_t nq, const float* d_q, uint32_t k, uint32_t ef, uint32_t* d_ids, float* d_sims,
                         uint32_t* d_counts, uint32_t* d_stats, cudaStream_t s) {
  if (nq == 0) return HNSW_OK;
  if (nq > 0x7FFFFFFFull) return fail(HNSW_ERR_INVALID, "too many queries in one batch");
  if (!d_q || !d_ids || !d_sims || !d_counts) return fail(HNSW_ERR_INVALID, "null buffer");
  if (k == 0) return fail(HNSW_ERR_INVALID, "k must be > 0");
  if (ef == 0) ef = ef_construction;  // core.rs:485
  const int efr = efr_for(ef);
  if (!efr) return fail(HNSW_ERR_INVALID, "ef = %u is not supported (1..%u)", ef, kMaxMemEf);
  const bool staged_kind = kind == kKindR1 || kind == kKindR4 || kind == kKindR24;
  // ef beyond the register classes: the memory-backed list exists in the register-staged kernel family only
  last_search_staged = efr != kEfrMem && staged_kind && opt_search_impl != 1 && (opt_search_impl == 2 || !d_stats);
  if (last_search_staged) return search_device2(nq, d_q, k, ef, efr, d_ids, d_sims, d_counts, d_stats, s);
  if (!h_retry_seen) {
    cudaError_t e = cudaHostAlloc((void**)&h_retry_seen, 16, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(e, "pinned alloc");
    h_retry_seen[0] = h_retry_seen[1] = 0;
  }

  const uint32_t slots = pick_vis_slots(ef);
  const size_t qs = kind_needs_smem_query(kind) ? (size_t)dim * 4 : 0;
  int block = opt_block ? opt_block : 256;
  while (block > 32 && (size_t)(block / 32) * ((size_t)slots * 4 + qs) > max_smem) block /= 2;
  bool vis_smem = (size_t)(block / 32) * ((size_t)slots * 4 + qs) <= max_smem;
  if (!vis_smem) block = opt_block ? opt_block : 256;
  int warps = block / 32;
  if ((size_t)warps * qs > max_smem) return fail(HNSW_ERR_INVALID, "dimension too large for the query staging buffer");
  size_t smem = vis_smem ? (size_t)warps * ((size_t)slots * 4 + qs) : (size_t)warps * qs;
  int occ = occupancy(kind, vis_smem ? kKernSearchSmem : kKernSearchGlobal, efr, block, smem);
  if (occ < 1) return fail(HNSW_ERR_CUDA, "search kernel cannot be resident (block %d, smem %zu)", block, smem);
  if (opt_ctas_per_sm > 0) occ = std::min(occ, opt_ctas_per_sm);
  if (efr == kEfrMem) occ = 1;  // every warp carries 8 * ef bytes of list and a large visited table in global memory
  int grid = (int)std::min<uint64_t>((uint64_t)num_sms * occ, (nq + warps - 1) / warps);
  const uint32_t list_cap = efr == kEfrMem ? ((ef + 31) & ~31u) : 0;

  // retry pass: global-memory tables, 4x the slots (at least 16K), one CTA of 4 warps per SM
  const uint32_t big_slots = next_pow2(std::max<uint64_t>(16384, (uint64_t)slots * 4));
  const int block2 = 128, warps2 = 4;
  const int grid2 = (int)std::min<uint64_t>((uint64_t)num_sms, (nq + warps2 - 1) / warps2);
  size_t vis1_bytes = vis_smem ? 0 : (size_t)grid * warps * slots * 4;
  size_t vis2_bytes = (size_t)grid2 * warps2 * big_slots * 4;
  int rc = ensure_scratch(s_vis, vis1_bytes + vis2_bytes);
  if (rc) return rc;
  if ((rc = ensure_scratch(s_ctl, 64 + (size_t)nq * 4))) return rc;
  if (list_cap && (rc = ensure_scratch(s_list, (size_t)std::max(grid * warps, grid2 * warps2) * 2 * list_cap * 4))) return rc;
  uint32_t* ctl = (uint32_t*)s_ctl.p;
  cudaError_t e = cudaMemsetAsync(ctl, 0, 64, s);
  if (e != cudaSuccess) return cuda_fail(e, "search ctl memset");

  SearchArgs a{};
  a.list_mem = (uint32_t*)s_list.p;
  a.list_cap = list_cap;
  a.queries = d_q;
  a.nq = (uint32_t)nq;
  a.k = k;
  a.ef = ef;
  a.ids = d_ids;
  a.sims = d_sims;
  a.counts = d_counts;
  a.stats = d_stats;
  a.work_counter = ctl + 0;
  a.retry_list = ctl + 16;
  a.retry_count = ctl + 2;
  a.vis_slots = slots;
  a.vis_global = (uint32_t*)s_vis.p;
  a.retry_pass = 0;
  LaunchCfg c{grid, block, smem, s};
  e = run(kind, vis_smem ? kKernSearchSmem : kKernSearchGlobal, efr, c, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "search_knn launch");

  a.work_counter = ctl + 1;
  a.vis_slots = big_slots;
  a.vis_global = (uint32_t*)((char*)s_vis.p + vis1_bytes);
  a.retry_pass = 1;
  LaunchCfg c2{grid2, block2, (size_t)warps2 * qs, s};
  e = run(kind, kKernSearchGlobal, efr, c2, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "search_knn retry launch");

  // feedback for the adaptive table size (read at the next call; harmless if it has not landed yet)
  e = cudaMemcpyAsync(h_retry_seen, ctl + 2, 4, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_fail(e, "retry feedback");
  h_retry_seen[1] = (uint32_t)nq;
  return HNSW_OK;
}

// TMA-staged kernel (search2.cuh): no retry pass (its visited table cannot overflow)
int Index::search_device2(uint64_t nq, const float* d_q, uint32_t k, uint32_t ef, int efr, uint32_t* d_ids, float* d_sims,
                          uint32_t* d_counts, uint32_t* d_stats, cudaStream_t s, uint32_t ctl_slot) {
  int S = opt_stage_rows;
  if (!S) {
    // about 4 KB of rows in flight per warp: on a B200 the kernel is occupancy-hungry (measured, profiles/), many
    // warps with a small stage each beat few warps with a deep one
    uint32_t rows = 4096u / (dim * 4);
    S = rows >= 32 ? 32 : (rows >= 16 ? 16 : (rows >= 8 ? 8 : 4));
  }
  const uint32_t slots = opt_recent_slots ? opt_recent_slots : 1024;
  bool use_la = false;
  if (!opt_stage_rows) {
    // latency mode: when every query of the call is resident at once anyway (a slice of a sharded batch, one HNSW.SEARCH)
    // nothing is gained by keeping the stage small, and a stage that takes a whole adjacency chunk in one round shortens
    // every hop (one HNSW.SEARCH: 437 -> 330 us on a 1M x 128 graph).  Take the deepest stage that still lets all the
    // call's warps sit on the SMs together (about 200 KB of shared memory per SM for the warps' stages and tag tables).
    const uint64_t warps_per_sm = (nq + (uint64_t)num_sms - 1) / (uint64_t)num_sms;
    const bool cp_kind = opt_row_copy == 1 && (kind == kKindR4 || kind == kKindR1);
    for (int cand = 32; cand >= S; cand /= 2) {
      // a 32-row stage of cp.async rows also gets the second stage of the lookahead kernel (search_la.cuh)
      const bool la = cand == 32 && cp_kind && opt_lookahead;
      const size_t pw = warp2_smem_bytes(dim, cand, slots, 4) + (la ? la_smem_bytes(dim) : 0);
      if ((size_t)dim * 4 * cand <= 32768 && pw <= max_smem / 2 && warps_per_sm * pw <= 200u * 1024u) {
        S = cand;
        use_la = la;
        break;
      }
    }
  }
  int slot_bits = 0;
  while ((1u << slot_bits) < slots) ++slot_bits;
  // 16-bit tags identify an id exactly only below 2^(log2(slots) + 15)
  const bool tag16 = opt_recent_tag != 32 && slot_bits + 15 < 32 && n_ids <= (1ull << (slot_bits + 15));
  // few queries (one HNSW.SEARCH, a handful per SM): one query per CTA of 4 warps that share the row copies, partial sums
  // and reduction of every hop (search2.cuh, COPY == 2): 211 -> 185 us for one query, 349 -> 303 us for 148 (profiles/r2_experiments.md)
  if (opt_search_cta && !d_stats && (kind == kKindR4 || kind == kKindR1) && nq <= 2ull * (uint64_t)num_sms) {
    const size_t smem_cta = warp2_smem_bytes(dim, 32, slots, tag16 ? 2 : 4) + 256;
    if (smem_cta <= max_smem) {
      int rc0 = ensure_scratch(s_ctl, std::max<size_t>(64 + (size_t)nq * 4, 64 * kCtlSlots));
      if (rc0) return rc0;
      SearchArgs ac{};
      ac.queries = d_q;
      ac.nq = (uint32_t)nq;
      ac.k = k;
      ac.ef = ef;
      ac.ids = d_ids;
      ac.sims = d_sims;
      ac.counts = d_counts;
      ac.stats = d_stats;
      ac.vis_slots = slots;
      LaunchCfg cc{(int)nq, 128, smem_cta, s};
      cudaError_t ec = run(kind, kKernSearch2Cta + (tag16 ? 1 : 0), efr, cc, g, &ac);
      if (ec != cudaSuccess) return cuda_fail(ec, "search_knn2_cta launch");
      return HNSW_OK;
    }
  }
  const size_t per_warp = warp2_smem_bytes(dim, S, slots, tag16 ? 2 : 4) + (use_la ? la_smem_bytes(dim) : 0);
  int block = opt_block ? std::min(opt_block, 128) : 64;  // search_knn2_kernel is bounded at 128 threads per CTA
  while (block > 32 && (size_t)(block / 32) * per_warp > max_smem) block /= 2;
  const int warps = block / 32;
  const size_t smem = (size_t)warps * per_warp;
  if (smem > max_smem) return fail(HNSW_ERR_INVALID, "dimension too large for the staged search kernel");
  // rows of up to two cp.async instructions (32-d, 128-d) are copied with cp.async, longer ones with bulk-async copies
  const bool cp = opt_row_copy == 1 && (kind == kKindR4 || kind == kKindR1);
  int id = search2_id(S, tag16) + (cp ? kKernSearch2Cp - kKernSearch2 : 0);
  if (use_la) id = kKernSearch2La + (tag16 ? 1 : 0);
  uint32_t table_slots = slots;
  if (!use_la && opt_recent_ways == 2 && cp && tag16 && slots >= 128 && n_ids <= (1ull << (slot_bits - 1 + 15))) {
    // DRAFT: the same bytes as `slots` 16-bit tags, organised as slots / 2 two-way sets (one 32-bit word per set)
    table_slots = slots / 2;
    id = kKernSearch2W2 + (S == 4 ? 0 : S == 8 ? 1 : S == 16 ? 2 : 3);
  }
  int occ = occupancy(kind, id, efr, block, smem);
  if (occ < 1) return fail(HNSW_ERR_CUDA, "staged search kernel cannot be resident (block %d, smem %zu)", block, smem);
  if (opt_ctas_per_sm > 0) occ = std::min(occ, opt_ctas_per_sm);
  const int grid = (int)std::min<uint64_t>((uint64_t)num_sms * occ, (nq + warps - 1) / warps);
  // control words: one 64-byte slot per concurrently running launch (search_host pipelines chunks on two streams)
  int rc = ensure_scratch(s_ctl, std::max<size_t>(64 + (size_t)nq * 4, 64 * kCtlSlots));
  if (rc) return rc;
  uint32_t* ctl = (uint32_t*)s_ctl.p + (size_t)(ctl_slot % kCtlSlots) * 16;
  cudaError_t e = cudaMemsetAsync(ctl, 0, 64, s);
  if (e != cudaSuccess) return cuda_fail(e, "search ctl memset");
  SearchArgs a{};
  a.queries = d_q;
  a.nq = (uint32_t)nq;
  a.k = k;
  a.ef = ef;
  a.ids = d_ids;
  a.sims = d_sims;
  a.counts = d_counts;
  a.stats = d_stats;
  a.work_counter = ctl + 0;
  a.retry_count = ctl + 2;
  a.vis_slots = table_slots;
  LaunchCfg c{grid, block, smem, s};
  e = run(kind, id, efr, c, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "search_knn2 launch");
  return HNSW_OK;
}

// Host-buffer batches on the staged kernel are pipelined: the batch is cut into chunks that alternate between two
// streams, so the H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the kernel of chunk i (pinned host
// memory; pageable memory still works, the copies then serialise).
int Index::search_host_pipelined(uint64_t nq, const float* q, uint32_t k, uint32_t ef, int efr, uint32_t* ids, float* sims,
                                 uint32_t* counts, const float* d_q, uint32_t* d_ids, float* d_sims, uint32_t* d_counts) {
  for (int i = 0; i < 2; ++i)
    if (!aux_stream[i]) {
      cudaError_t e = cudaStreamCreateWithFlags(&aux_stream[i], cudaStreamNonBlocking);
      if (e != cudaSuccess) return cuda_fail(e, "aux stream");
    }
  // slot 0 belongs to the un-pipelined / device-pointer path (which may still be in flight on the caller's stream)
  const uint64_t n_chunks = std::min<uint64_t>(kCtlSlots - 1, std::max<uint64_t>(2, nq / 16384));
  const uint64_t per = (nq + n_chunks - 1) / n_chunks;
  // size the control scratch before anything is in flight (ensure_scratch may reallocate)
  int rc = ensure_scratch(s_ctl, std::max<size_t>(64 + (size_t)nq * 4, 64 * kCtlSlots));
  if (rc) return rc;
  cudaError_t e = cudaSuccess;
  for (uint64_t c = 0, lo = 0; lo < nq; ++c, lo += per) {
    const uint64_t n = std::min(per, nq - lo);
    cudaStream_t st = aux_stream[c & 1];
    e = cudaMemcpyAsync((float*)d_q + lo * dim, q + lo * dim, n * dim * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "query H2D");
    if ((rc = search_device2(n, d_q + lo * dim, k, ef, efr, d_ids + lo * k, d_sims + lo * k, d_counts + lo, nullptr, st,
                             (uint32_t)c + 1)))
      return rc;
    e = cudaMemcpyAsync(ids + lo * k, d_ids + lo * k, n * k * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sims + lo * k, d_sims + lo * k, n * k * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts + lo, d_counts + lo, n * 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "result D2H");
  }
  for (int i = 0; i < 2; ++i) {
    e = cudaStreamSynchronize(aux_stream[i]);
    if (e != cudaSuccess) return cuda_fail(e, "search_batch");
  }
  return HNSW_OK;
}

int Index::search_host(uint64_t nq, const float* q, uint32_t k, uint32_t ef, uint32_t* ids, float* sims,
                       uint32_t* counts, uint32_t* stats) {
  if (nq == 0) return HNSW_OK;
  if (!q || !ids || !sims || !counts) return fail(HNSW_ERR_INVALID, "null buffer");
  if (k == 0) return fail(HNSW_ERR_INVALID, "k must be > 0");
  size_t qb = (size_t)nq * dim * 4, ib = (size_t)nq * k * 4, cb = (size_t)nq * 4, sb = stats ? (size_t)nq * 16 : 0;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  int rc = ensure_scratch(s_in, qb);
  if (rc) return rc;
  if ((rc = ensure_scratch(s_out, al(ib) * 2 + al(cb) + al(sb)))) return rc;
  char* o = (char*)s_out.p;
  uint32_t* d_ids = (uint32_t*)o;
  float* d_sims = (float*)(o + al(ib));
  uint32_t* d_counts = (uint32_t*)(o + 2 * al(ib));
  uint32_t* d_stats = stats ? (uint32_t*)(o + 2 * al(ib) + al(cb)) : nullptr;
  {
    const uint32_t ef_eff = ef ? ef : ef_construction;
    const int efr = efr_for(ef_eff);
    const bool staged_kind = kind == kKindR1 || kind == kKindR4 || kind == kKindR24;
    if (!stats && nq >= 8192 && efr && efr != kEfrMem && staged_kind && opt_search_impl != 1 && nq <= 0x7FFFFFFFull) {
      return search_host_pipelined(nq, q, k, ef_eff, efr, ids, sims, counts, (const float*)s_in.p, d_ids, d_sims, d_counts);
    }
  }
  // Small calls (one HNSW.SEARCH): go through a pinned staging buffer so that the query upload is one truly
  // asynchronous copy and the three result arrays come back in ONE device-to-host copy.
  const size_t out_span = 2 * al(ib) + cb;
  if (!stats && qb <= kPinnedStage / 2 && out_span <= kPinnedStage / 2) {
    if (!h_stage) {
      cudaError_t e0 = cudaHostAlloc((void**)&h_stage, kPinnedStage, cudaHostAllocDefault);
      if (e0 != cudaSuccess) return cuda_fail(e0, "pinned staging alloc");
    }
    std::memcpy(h_stage, q, qb);
    cudaError_t e1 = cudaMemcpyAsync(s_in.p, h_stage, qb, cudaMemcpyHostToDevice, stream);
    if (e1 != cudaSuccess) return cuda_fail(e1, "query H2D");
    if ((rc = search_device(nq, (const float*)s_in.p, k, ef, d_ids, d_sims, d_counts, nullptr, stream))) return rc;
    char* back = h_stage + kPinnedStage / 2;
    e1 = cudaMemcpyAsync(back, o, out_span, cudaMemcpyDeviceToHost, stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(stream);
    if (e1 != cudaSuccess) return cuda_fail(e1, "search_batch");
    std::memcpy(ids, back, ib);
    std::memcpy(sims, back + al(ib), ib);
    std::memcpy(counts, back + 2 * al(ib), cb);
    if (last_search_staged) return HNSW_OK;
    if ((rc = pull_meta())) return rc;
    if (device_error & kErrVisitedOverflow) {
      push_meta();
      return fail(HNSW_ERR_INVALID, "visited table overflow even in the retry pass; raise the visited_slots option");
    }
    return HNSW_OK;
  }
  cudaError_t e = cudaMemcpyAsync(s_in.p, q, qb, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "query H2D");
  if ((rc = search_device(nq, (const float*)s_in.p, k, ef, d_ids, d_sims, d_counts, d_stats, stream))) return rc;
  e = cudaMemcpyAsync(ids, d_ids, ib, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sims, d_sims, ib, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_counts, cb, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess && stats) e = cudaMemcpyAsync(stats, d_stats, sb, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "search_batch");
  if (last_search_staged) return HNSW_OK;  // the staged kernel cannot overflow its visited table: no flag to fetch
  if ((rc = pull_meta())) return rc;
  if (device_error & kErrVisitedOverflow) {
    push_meta();
    return fail(HNSW_ERR_INVALID, "visited table overflow even in the retry pass; raise the visited_slots option");
  }
  return HNSW_OK;
}

int Index::search_level_host(const float* q, uint32_t ep, uint32_t ef, uint32_t level, uint32_t* ids, float* sims,
                             uint32_t* n_out) {
  const int efr = efr_for(ef);
  if (!efr) return fail(HNSW_ERR_INVALID, "ef = %u is not supported (1..%u)", ef, kMaxMemEf);
  if (ep >= n_ids || h_level[ep] < 0) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", ep);
  const uint32_t slots = next_pow2(std::max<uint64_t>(65536, (uint64_t)ef * 512));
  int rc = ensure_scratch(s_vis, (size_t)slots * 4);
  if (rc) return rc;
  if ((rc = ensure_scratch(s_in, (size_t)dim * 4))) return rc;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  if ((rc = ensure_scratch(s_out, al((size_t)ef * 4) * 2 + 256))) return rc;
  char* o = (char*)s_out.p;
  const uint32_t list_cap = efr == kEfrMem ? ((ef + 31) & ~31u) : 0;
  if (list_cap && (rc = ensure_scratch(s_list, (size_t)2 * list_cap * 4))) return rc;
  LevelArgs a{};
  a.list_mem = (uint32_t*)s_list.p;
  a.list_cap = list_cap;
  a.query = (const float*)s_in.p;
  a.entry = ep;
  a.ef = ef;
  a.level = level;
  a.ids = (uint32_t*)o;
  a.sims = (float*)(o + al((size_t)ef * 4));
  a.n_out = (uint32_t*)(o + 2 * al((size_t)ef * 4));
  a.vis_slots = slots;
  a.vis_global = (uint32_t*)s_vis.p;
  cudaError_t e = cudaMemcpyAsync(s_in.p, q, (size_t)dim * 4, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "query H2D");
  LaunchCfg c{1, 32, kind_needs_smem_query(kind) ? (size_t)dim * 4 : 0, stream};
  e = run(kind, kKernLevel, efr, c, g, &a);
  if (e != cudaSuccess) return cuda_fail(e, "search_level launch");
  uint32_t n = 0;
  e = cudaMemcpyAsync(&n, a.n_out, 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "search_level");
  if (n == 0xFFFFFFFFu) return fail(HNSW_ERR_INVALID, "visited table overflow in search_level");
  e = cudaMemcpyAsync(ids, a.ids, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sims, a.sims, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "search_level D2H");
  *n_out = n;
  return HNSW_OK;
}

}  // namespace hnsw

using namespace hnsw;

extern "C" {

int hnsw_index_search_batch_device(hnsw_index_t* idx, uint64_t nq, const float* d_queries, uint32_t k, uint32_t ef,
                                   uint32_t* d_ids, float* d_sims, uint32_t* d_counts, hnsw_query_stats_t* d_stats,
                                   void* stream) {
  IDX_OR_FAIL(idx)
  return ix.search_device(nq, d_queries, k, ef, d_ids, d_sims, d_counts, (uint32_t*)d_stats,
                          stream ? (cudaStream_t)stream : ix.stream);
}

int hnsw_index_search_batch(hnsw_index_t* idx, uint64_t nq, const float* queries, uint32_t k, uint32_t ef, uint32_t* ids,
                            float* sims, uint32_t* counts, hnsw_query_stats_t* stats) {
  IDX_OR_FAIL(idx)
  return ix.search_host(nq, queries, k, ef, ids, sims, counts, (uint32_t*)stats);
}

int hnsw_index_search(hnsw_index_t* idx, const float* query, uint64_t n, uint32_t k, uint32_t ef, uint32_t* ids,
                      float* sims, uint32_t* n_out) {
  IDX_OR_FAIL(idx)
  if (n != ix.dim) return fail(HNSW_ERR_DIM_MISMATCH, "data dimension: %llu does not match Index", (unsigned long long)n);  // core.rs:479
  if (!n_out) return fail(HNSW_ERR_INVALID, "null n_out");
  if (ix.entry < 0 || ix.node_count == 0) {  // core.rs:481-483
    *n_out = 0;
    return HNSW_OK;
  }
  uint32_t cnt = 0;
  int rc = ix.search_host(1, query, k, ef, ids, sims, &cnt, nullptr);
  if (rc) return rc;
  *n_out = cnt;
  return HNSW_OK;
}

int hnsw_index_search_level(hnsw_index_t* idx, const float* query, uint32_t entry, uint32_t ef, uint32_t level,
                            uint32_t* ids, float* sims, uint32_t* n_out) {
  IDX_OR_FAIL(idx)
  if (!query || !ids || !sims || !n_out) return fail(HNSW_ERR_INVALID, "null buffer");
  return ix.search_level_host(query, entry, ef, level, ids, sims, n_out);
}

}  // extern "C"

// K1 of the batched builder, second generation: the per-insert searches of a batch (core.rs:511-531) on the
// TMA-staged machinery of search2.cuh — bulk-async row staging with an mbarrier per warp, transposed bit-exact
// reduction, lossy exact-tag visited table — instead of the register-staged kernel with an exact visited set.
//
// Why: with ef_construction = 200 the exact set needs 8192 slots x 4 B = 32 KB per warp, which leaves 4 warps per SM
// (measured: 12 % of the HBM roofline for K1, 81 % of the build's kernel time; profiles/r1d_build_launches.md).  The
// lossy table (4096 16-bit tags = 8 KB) plus an S-row stage fits 16+ warps per SM.  The result of a search is the same
// top-ef set either way (search2.cuh explains why forgetting a visited id is harmless), so `sel` is unchanged.
#pragma once
#include "build.cuh"
#include "search2.cuh"

namespace hnsw {

// rows per stage for the builder's searches (one choice per dimension keeps the number of instantiations down)
template <int C>
struct BuildStage {
  static constexpr int S = (C == 1) ? 32 : ((C <= 4) ? 8 : 4);
};

template <int EFR, int C, class T>
__global__ void __launch_bounds__(256) build_search2_kernel(Graph g, FastArgs a) {
  constexpr int S = BuildStage<C>::S;
  constexpr int V = RowRegs<C>::V;
  extern __shared__ __align__(128) unsigned char smem2[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  unsigned char* base = smem2 + (size_t)warp * warp2_smem_bytes(32 * C, S, a.vis_slots, sizeof(T));
  Warp2<C, S, T> w;
  w.stage = reinterpret_cast<const float4*>(base);
  w.stage_s = smem_u32(base);
  const uint32_t tab_bytes = (a.vis_slots * (uint32_t)sizeof(T) + 127u) & ~127u;
  w.seen.tab = reinterpret_cast<T*>(base + (size_t)S * C * 128);
  w.seen.n16 = tab_bytes / 16;
  w.seen.bits = 31 - __clz(a.vis_slots);
  w.ids = reinterpret_cast<uint32_t*>(base + (size_t)S * C * 128 + tab_bytes);
  w.bar = smem_u32(w.ids + 32);
  w.parity = 0;
  if (lane == 0) mbar_init(w.bar, 1);
  __syncwarp();

  CandList<EFR> L;
  Counters cnt = {0, 0, 0};
  const int l_max = g.meta[kMetaMaxLayer];
  const uint32_t entry = (uint32_t)g.meta[kMetaEntry];
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.ctl + kCtlWorkSearch, 1u);
    wi = __shfl_sync(kFull, wi, 0);
    if (wi >= a.n_new) break;
    const uint32_t q = a.first + wi;
    const int l = g.level[q];
    const uint32_t tb = a.task_base[wi];
    {  // the query is the node's own (lane-permuted) slab row: chunk c of lane t sits at ((c / V) * 32 + t) * V + c % V
      const float* row = g.vecs + (size_t)q * (32 * C);
#pragma unroll
      for (int c = 0; c < C; ++c) w.q[c] = row[((c / V) * 32 + lane) * V + (c % V)];
    }
    uint32_t ep = entry;
    for (int lc = l_max; lc >= 0; --lc) {
      const bool link = lc <= l;
      search_layer2<EFR, C, S, T>(g, w, ep, link ? (int)a.efc : 1, (uint32_t)lc, L, cnt, lane);  // core.rs:513, :524
      float s;
      L.get(0, lane, false, ep, s);                                                              // core.rs:514, :576
      if (!link) continue;
      const uint32_t n_sel = min((uint32_t)L.len, a.m);                                          // core.rs:531 (build.cuh header)
      uint32_t* out = a.sel_ids + (size_t)(tb + lc) * a.m;
#pragma unroll
      for (int r = 0; r < EFR; ++r) {
        uint32_t e = r * 32 + lane;
        if (e < n_sel) out[e] = L.id[r] & ~kExpanded;
      }
      if (lane == 0) a.sel_cnt[tb + lc] = n_sel;
    }
  }
  if (lane == 0) atomicAdd(a.ctl + kCtlDistEvals, cnt.n_dist);
}

}  // namespace hnsw

// Launch plumbing shared by the per-distance-mode translation units (one TU per Dist keeps nvcc parallel).
#pragma once
#include "build.cuh"
#include "search.cuh"
#include "search2.cuh"
#include "build2.cuh"
#include "search_la.cuh"

namespace hnsw {

enum DistKind : int { kKindR1 = 0, kKindR4 = 1, kKindR24 = 2, kKindGeneric = 3, kKindScalar = 4, kKindCount = 5 };

struct LaunchCfg {
  int grid, block;
  size_t smem;
  cudaStream_t stream;
};

// the memory-backed list class: ef / ef_construction / 2m up to kMaxMemEf entries, register-staged kernels only
constexpr int kEfrMem = 64;
constexpr uint32_t kMaxMemEf = 65536;

// list registers per lane for a given ef (0 = unsupported, kEfrMem = memory-backed list)
inline int efr_for(uint32_t ef) {
  if (ef == 0) return 0;
  if (ef <= 32) return 1;
  if (ef <= 64) return 2;
  if (ef <= 128) return 4;
  if (ef <= 256) return 8;
  if (ef <= 512) return 16;
  if (ef <= 1024) return 32;
  if (ef <= kMaxMemEf) return kEfrMem;  // beyond the register classes: the list lives in memory (CandList<0>, search.cuh)
  return 0;
}

// which kernel of a kind
enum KernelId : int {
  kKernSearchSmem = 0,   // search_knn_kernel, visited table in shared memory
  kKernSearchGlobal = 1, // search_knn_kernel, visited table in global memory
  kKernLevel = 2,        // search_level_kernel
  kKernBuildSearchSmem = 3,
  kKernBuildSearchGlobal = 4,
  kKernBuildReprune = 5,
  kKernExact = 6,
  kKernDelete = 7,
  // search_knn2_kernel (TMA-staged rows): id = kKernSearch2 + 2 * log2(S / 4) + (16-bit visited tags ? 1 : 0)
  kKernSearch2 = 16,
};
// build_search2_kernel (TMA-staged K1 of the batched builder): id = kKernBuildSearch2 + (16-bit visited tags ? 1 : 0)
constexpr int kKernSearch2Cp = 24;  // search_knn2_kernel with cp.async row copies (RowCopy<C>::kOk): + 2 * log2(S / 4) + (16-bit tags ? 1 : 0)
constexpr int kKernBuildSearch2 = 32;
constexpr int kKernSearch2W2 = 52;   // DRAFT search_knn2_kernel, cp.async rows, 2-way visited sets: + log2(S / 4)
constexpr int kKernSearch2La = 56;   // search_knn2_la_kernel (32-row stage, cp.async rows, one-hop lookahead): + (16-bit tags ? 1 : 0)
constexpr int kKernSearch2Cta = 48;  // DRAFT search_knn2_cta_kernel (one query per CTA of 4 warps): + (16-bit tags ? 1 : 0)
constexpr int kKernExact2 = 40;   // insert_exact2_kernel / delete_exact2_kernel (TMA-staged, build2.cuh)
constexpr int kKernDelete2 = 41;
constexpr int kKernExact2Small = 42;
constexpr int kKernBuildReprune2 = 44;  // build_reprune2_kernel: + (16-bit visited tags ? 1 : 0)  // same, the re-selection list is CandList<min(EFR, 2)> (m_max_0 <= 64)
inline int search2_id(int S, bool tag16) { return kKernSearch2 + 2 * (S == 4 ? 0 : S == 8 ? 1 : S == 16 ? 2 : 3) + (tag16 ? 1 : 0); }

// kernel arguments are passed type-erased so that one entry point per kind serves every kernel
struct KernelArgs {
  const Graph* g;
  const void* a;  // SearchArgs / LevelArgs / FastArgs / ExactArgs
};

template <class K>
inline cudaError_t set_smem(K k, size_t smem) {
  if (smem > 48 * 1024) return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cudaSuccess;
}

template <class K, class A>
inline cudaError_t launch_k(K k, const LaunchCfg& c, const Graph& g, const A& a) {
  cudaError_t e = set_smem(k, c.smem);
  if (e != cudaSuccess) return e;
  k<<<c.grid, c.block, c.smem, c.stream>>>(g, a);
  return cudaGetLastError();
}

template <class K>
inline int occ_k(K k, int block, size_t smem) {
  if (set_smem(k, smem) != cudaSuccess) return 0;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, block, smem) != cudaSuccess) return 0;
  return n;
}

// launch (occupancy_only = false) or query resident CTAs per SM (occupancy_only = true, returned as a
// non-negative int through *occ)
template <int EFR, class Dist>
cudaError_t run_kernel(int id, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ) {
#define HNSW_RUN(KERNEL, ARGT)                                             \
  {                                                                        \
    auto k = KERNEL;                                                       \
    if (occupancy_only) {                                                  \
      *occ = occ_k(k, c.block, c.smem);                                    \
      return cudaSuccess;                                                  \
    }                                                                      \
    return launch_k(k, c, *ka.g, *static_cast<const ARGT*>(ka.a));         \
  }
  switch (id) {
    case kKernSearchSmem: HNSW_RUN((search_knn_kernel<EFR, Dist, true>), SearchArgs)
    case kKernSearchGlobal: HNSW_RUN((search_knn_kernel<EFR, Dist, false>), SearchArgs)
    case kKernLevel: HNSW_RUN((search_level_kernel<EFR, Dist>), LevelArgs)
    case kKernExact: HNSW_RUN((insert_exact_kernel<EFR, Dist>), ExactArgs)
    case kKernDelete: HNSW_RUN((delete_exact_kernel<EFR, Dist>), ExactArgs)
  }
  if constexpr (EFR > 0) {  // the batched builder has no memory-backed class: such requests run the EXACT stream
    switch (id) {
      case kKernBuildSearchSmem: HNSW_RUN((build_search_kernel<EFR, Dist, true>), FastArgs)
      case kKernBuildSearchGlobal: HNSW_RUN((build_search_kernel<EFR, Dist, false>), FastArgs)
      case kKernBuildReprune: HNSW_RUN((build_reprune_kernel<EFR, Dist>), FastArgs)
    }
  }
  if constexpr (Dist::kStaged && EFR > 0) {
    switch (id) {
      case kKernSearch2 + 0: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 4, uint32_t>), SearchArgs)
      case kKernSearch2 + 1: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 4, uint16_t>), SearchArgs)
      case kKernSearch2 + 2: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 8, uint32_t>), SearchArgs)
      case kKernSearch2 + 3: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 8, uint16_t>), SearchArgs)
      case kKernSearch2 + 4: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 16, uint32_t>), SearchArgs)
      case kKernSearch2 + 5: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 16, uint16_t>), SearchArgs)
      case kKernSearch2 + 6: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 32, uint32_t>), SearchArgs)
      case kKernSearch2 + 7: HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 32, uint16_t>), SearchArgs)
      case kKernSearch2Cp + 0: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 4, uint32_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 1: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 4, uint16_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 2: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 8, uint32_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 3: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 8, uint16_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 4: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 16, uint32_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 5: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 16, uint16_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 6: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 32, uint32_t, 1>), SearchArgs) break;
      case kKernSearch2Cp + 7: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 32, uint16_t, 1>), SearchArgs) break;
      case kKernSearch2W2 + 0: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 4, Way2, 1>), SearchArgs) break;
      case kKernSearch2W2 + 1: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 8, Way2, 1>), SearchArgs) break;
      case kKernSearch2W2 + 2: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 16, Way2, 1>), SearchArgs) break;
      case kKernSearch2W2 + 3: if constexpr (RowCopy<Dist::C>::kOk) HNSW_RUN((search_knn2_kernel<EFR, Dist::C, 32, Way2, 1>), SearchArgs) break;
      case kKernSearch2La + 0: if constexpr (RowCopy<Dist::C>::kOk && EFR <= 2) HNSW_RUN((search_knn2_la_kernel<EFR, Dist::C, uint32_t>), SearchArgs) break;
      case kKernSearch2La + 1: if constexpr (RowCopy<Dist::C>::kOk && EFR <= 2) HNSW_RUN((search_knn2_la_kernel<EFR, Dist::C, uint16_t>), SearchArgs) break;
      case kKernSearch2Cta + 0: HNSW_RUN((search_knn2_cta_kernel<EFR, Dist::C, uint32_t>), SearchArgs)
      case kKernSearch2Cta + 1: HNSW_RUN((search_knn2_cta_kernel<EFR, Dist::C, uint16_t>), SearchArgs)
      case kKernBuildSearch2 + 0: HNSW_RUN((build_search2_kernel<EFR, Dist::C, uint32_t>), FastArgs)
      case kKernBuildSearch2 + 1: HNSW_RUN((build_search2_kernel<EFR, Dist::C, uint16_t>), FastArgs)
      case kKernBuildReprune2 + 0: HNSW_RUN((build_reprune2_kernel<EFR, Dist::C, uint32_t>), FastArgs)
      case kKernBuildReprune2 + 1: HNSW_RUN((build_reprune2_kernel<EFR, Dist::C, uint16_t>), FastArgs)
      case kKernExact2: HNSW_RUN((insert_exact2_kernel<EFR, Dist::C, false>), ExactArgs)
      case kKernExact2Small: HNSW_RUN((insert_exact2_kernel<EFR, Dist::C, true>), ExactArgs)
      case kKernDelete2: HNSW_RUN((delete_exact2_kernel<EFR, Dist::C>), ExactArgs)
    }
  }
#undef HNSW_RUN
  return cudaErrorInvalidValue;
}

// PART splits the list classes of one distance kind over translation units (compile time: every class instantiates every
// kernel): 0 = all classes, 1 = EFR 1 / 2 / 4, 2 = EFR 8 / 16, 3 = EFR 32
template <class Dist, int PART = 0>
cudaError_t run_kind(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ) {
  if constexpr (PART == 0 || PART == 1) {
    switch (efr) {
      case 1: return run_kernel<1, Dist>(id, c, ka, occupancy_only, occ);
      case 2: return run_kernel<2, Dist>(id, c, ka, occupancy_only, occ);
      case 4: return run_kernel<4, Dist>(id, c, ka, occupancy_only, occ);
    }
  }
  if constexpr (PART == 0 || PART == 2) {
    switch (efr) {
      case 8: return run_kernel<8, Dist>(id, c, ka, occupancy_only, occ);
      case 16: return run_kernel<16, Dist>(id, c, ka, occupancy_only, occ);
    }
  }
  if constexpr (PART == 0 || PART == 3) {
    if (efr == 32) return run_kernel<32, Dist>(id, c, ka, occupancy_only, occ);
    if (efr == kEfrMem) return run_kernel<0, Dist>(id, c, ka, occupancy_only, occ);
  }
  return cudaErrorInvalidValue;
}

// entry points defined in kernels_<kind>.cu
cudaError_t run_kind_r1(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ);
cudaError_t run_kind_r4(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ);
cudaError_t run_kind_r24(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ);
cudaError_t run_kind_generic(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ);
cudaError_t run_kind_scalar(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only, int* occ);

#define HNSW_DEFINE_KIND(NAME, DIST)                                                                             \
  namespace hnsw {                                                                                               \
  cudaError_t run_kind_##NAME(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only,    \
                              int* occ) {                                                                        \
    return run_kind<DIST>(id, efr, c, ka, occupancy_only, occ);                                                  \
  }                                                                                                              \
  }
// one part of a kind (see run_kind): defines run_kind_<NAME>_p<PART>
#define HNSW_DEFINE_KIND_PART(NAME, DIST, PART)                                                                  \
  namespace hnsw {                                                                                               \
  cudaError_t run_kind_##NAME##_p##PART(int id, int efr, const LaunchCfg& c, const KernelArgs& ka,               \
                                        bool occupancy_only, int* occ) {                                        \
    return run_kind<DIST, PART>(id, efr, c, ka, occupancy_only, occ);                                            \
  }                                                                                                              \
  }
// the kind's entry point when its classes are split over three translation units
#define HNSW_DECLARE_KIND_PARTS(NAME)                                                                            \
  namespace hnsw {                                                                                               \
  cudaError_t run_kind_##NAME##_p1(int, int, const LaunchCfg&, const KernelArgs&, bool, int*);                  \
  cudaError_t run_kind_##NAME##_p2(int, int, const LaunchCfg&, const KernelArgs&, bool, int*);                  \
  cudaError_t run_kind_##NAME##_p3(int, int, const LaunchCfg&, const KernelArgs&, bool, int*);                  \
  cudaError_t run_kind_##NAME(int id, int efr, const LaunchCfg& c, const KernelArgs& ka, bool occupancy_only,    \
                              int* occ) {                                                                        \
    if (efr <= 4) return run_kind_##NAME##_p1(id, efr, c, ka, occupancy_only, occ);                              \
    if (efr <= 16) return run_kind_##NAME##_p2(id, efr, c, ka, occupancy_only, occ);                             \
    return run_kind_##NAME##_p3(id, efr, c, ka, occupancy_only, occ);                                            \
  }                                                                                                              \
  }

// dispatch on the index's distance kind (search_host.cu)
cudaError_t run(int kind, int id, int efr, const LaunchCfg& c, const Graph& g, const void* args);
int occupancy(int kind, int id, int efr, int block, size_t smem);

}  // namespace hnsw

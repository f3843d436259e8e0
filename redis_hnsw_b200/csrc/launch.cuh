// Launch plumbing shared by the per-distance-mode translation units (one TU per Dist keeps nvcc parallel).
#pragma once
#include "search.cuh"

namespace hnsw {

enum DistKind : int { kKindR1 = 0, kKindR4 = 1, kKindR24 = 2, kKindGeneric = 3, kKindScalar = 4, kKindCount = 5 };

struct LaunchCfg {
  int grid, block;
  size_t smem;
  cudaStream_t stream;
};

// list registers per lane for a given ef (0 = unsupported)
inline int efr_for(uint32_t ef) {
  if (ef == 0) return 0;
  if (ef <= 32) return 1;
  if (ef <= 64) return 2;
  if (ef <= 128) return 4;
  if (ef <= 256) return 8;
  if (ef <= 512) return 16;
  return 0;
}

template <int EFR, class Dist, bool VS>
cudaError_t launch_search_one(const LaunchCfg& c, const Graph& g, const SearchArgs& a) {
  auto k = search_knn_kernel<EFR, Dist, VS>;
  if (c.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
  }
  k<<<c.grid, c.block, c.smem, c.stream>>>(g, a);
  return cudaGetLastError();
}

template <class Dist>
cudaError_t launch_search(int efr, bool vis_smem, const LaunchCfg& c, const Graph& g, const SearchArgs& a) {
#define HNSW_CASE(E)                                                       \
  case E:                                                                  \
    return vis_smem ? launch_search_one<E, Dist, true>(c, g, a) : launch_search_one<E, Dist, false>(c, g, a);
  switch (efr) {
    HNSW_CASE(1)
    HNSW_CASE(2)
    HNSW_CASE(4)
    HNSW_CASE(8)
    HNSW_CASE(16)
  }
#undef HNSW_CASE
  return cudaErrorInvalidValue;
}

template <class Dist>
cudaError_t launch_level(int efr, const LaunchCfg& c, const Graph& g, const LevelArgs& a) {
#define HNSW_CASE(E)                                                      \
  case E:                                                                 \
    search_level_kernel<E, Dist><<<1, 32, c.smem, c.stream>>>(g, a);      \
    return cudaGetLastError();
  switch (efr) {
    HNSW_CASE(1)
    HNSW_CASE(2)
    HNSW_CASE(4)
    HNSW_CASE(8)
    HNSW_CASE(16)
  }
#undef HNSW_CASE
  return cudaErrorInvalidValue;
}

// resident CTAs per SM the search kernel can reach with `block` threads and `smem` dynamic bytes
template <class Dist>
int occupancy_search(int efr, bool vis_smem, int block, size_t smem) {
  int n = 0;
#define HNSW_CASE(E)                                                                                              \
  case E:                                                                                                         \
    if (vis_smem) {                                                                                               \
      if (smem > 48 * 1024)                                                                                       \
        cudaFuncSetAttribute(search_knn_kernel<E, Dist, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             (int)smem);                                                                          \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, search_knn_kernel<E, Dist, true>, block, smem);           \
    } else {                                                                                                      \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, search_knn_kernel<E, Dist, false>, block, smem);          \
    }                                                                                                             \
    break;
  switch (efr) {
    HNSW_CASE(1)
    HNSW_CASE(2)
    HNSW_CASE(4)
    HNSW_CASE(8)
    HNSW_CASE(16)
  }
#undef HNSW_CASE
  return n;
}

// entry points defined in kernels_<kind>.cu
#define HNSW_DECL_KIND(NAME)                                                                                   \
  cudaError_t launch_search_##NAME(int efr, bool vis_smem, const LaunchCfg& c, const Graph& g, const SearchArgs& a); \
  cudaError_t launch_level_##NAME(int efr, const LaunchCfg& c, const Graph& g, const LevelArgs& a);            \
  int occupancy_search_##NAME(int efr, bool vis_smem, int block, size_t smem);
HNSW_DECL_KIND(r1)
HNSW_DECL_KIND(r4)
HNSW_DECL_KIND(r24)
HNSW_DECL_KIND(generic)
HNSW_DECL_KIND(scalar)
#undef HNSW_DECL_KIND

#define HNSW_DEFINE_KIND(NAME, DIST)                                                                            \
  namespace hnsw {                                                                                              \
  cudaError_t launch_search_##NAME(int efr, bool vis_smem, const LaunchCfg& c, const Graph& g, const SearchArgs& a) { \
    return launch_search<DIST>(efr, vis_smem, c, g, a);                                                         \
  }                                                                                                             \
  cudaError_t launch_level_##NAME(int efr, const LaunchCfg& c, const Graph& g, const LevelArgs& a) {            \
    return launch_level<DIST>(efr, c, g, a);                                                                    \
  }                                                                                                             \
  int occupancy_search_##NAME(int efr, bool vis_smem, int block, size_t smem) {                                 \
    return occupancy_search<DIST>(efr, vis_smem, block, smem);                                                  \
  }                                                                                                             \
  }

}  // namespace hnsw

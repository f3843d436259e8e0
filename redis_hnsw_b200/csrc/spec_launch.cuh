// Launch plumbing of the SPEC builder's K1 (spec.cuh): one translation unit per vector dimension class, like kernels_*.cu.
#pragma once
#include "launch.cuh"
#include "spec.cuh"

namespace hnsw {

template <int EFR, int C>
cudaError_t run_spec_efr(bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, int* occ) {
  if (small) {
    auto k = spec_exec_kernel<EFR, C, true>;
    if (occupancy_only) {
      *occ = occ_k(k, c.block, c.smem);
      return cudaSuccess;
    }
    return launch_k(k, c, g, a);
  }
  auto k = spec_exec_kernel<EFR, C, false>;
  if (occupancy_only) {
    *occ = occ_k(k, c.block, c.smem);
    return cudaSuccess;
  }
  return launch_k(k, c, g, a);
}

template <int C, int PART>
cudaError_t run_spec_c(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, int* occ) {
  if constexpr (PART == 1) {
    switch (efr) {
      case 1: return run_spec_efr<1, C>(small, c, g, a, occupancy_only, occ);
      case 2: return run_spec_efr<2, C>(small, c, g, a, occupancy_only, occ);
      case 4: return run_spec_efr<4, C>(small, c, g, a, occupancy_only, occ);
      case 8: return run_spec_efr<8, C>(small, c, g, a, occupancy_only, occ);
    }
  } else {
    switch (efr) {
      case 16: return run_spec_efr<16, C>(small, c, g, a, occupancy_only, occ);
      case 32: return run_spec_efr<32, C>(small, c, g, a, occupancy_only, occ);
    }
  }
  return cudaErrorInvalidValue;
}

// entry points defined in kernels_spec_<kind>.cu
cudaError_t run_spec_r1(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, int* occ);
cudaError_t run_spec_r4(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, int* occ);
cudaError_t run_spec_r24(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, int* occ);

#define HNSW_DEFINE_SPEC_KIND_PART(NAME, CVAL, PART)                                                                        \
  namespace hnsw {                                                                                                          \
  cudaError_t run_spec_##NAME##_p##PART(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a,         \
                                        bool occupancy_only, int* occ) {                                                   \
    return run_spec_c<CVAL, PART>(efr, small, c, g, a, occupancy_only, occ);                                                \
  }                                                                                                                         \
  }
#define HNSW_DECLARE_SPEC_KIND_PARTS(NAME)                                                                                  \
  namespace hnsw {                                                                                                          \
  cudaError_t run_spec_##NAME##_p1(int, bool, const LaunchCfg&, const Graph&, const SpecArgs&, bool, int*);                \
  cudaError_t run_spec_##NAME##_p2(int, bool, const LaunchCfg&, const Graph&, const SpecArgs&, bool, int*);                \
  cudaError_t run_spec_##NAME(int efr, bool small, const LaunchCfg& c, const Graph& g, const SpecArgs& a, bool occupancy_only, \
                              int* occ) {                                                                                   \
    return efr <= 8 ? run_spec_##NAME##_p1(efr, small, c, g, a, occupancy_only, occ)                                        \
                    : run_spec_##NAME##_p2(efr, small, c, g, a, occupancy_only, occ);                                       \
  }                                                                                                                         \
  }

}  // namespace hnsw

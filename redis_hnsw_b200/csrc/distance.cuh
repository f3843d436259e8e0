// Bit-exact restatement of the reference metric (src/hnsw/metrics.rs) for a 32-lane warp.
//
// AVX path (dim % 32 == 0, metrics.rs:48-77): the reference keeps 4 accumulators x 8 AVX lanes; element
// i = 32c + 8a + j goes to accumulator a, lane j, folded with FMA in increasing c.  Warp lane t = 8a + j owns
// exactly those elements (i % 32 == t), folds them with fmaf in increasing c, and the horizontal sum
//   (e1+e2)+(e3+e4)            metrics.rs:71-74   -> xor 8, xor 16
//   lo128 + hi128              metrics.rs:37-39   -> xor 4
//   (s0+s1)+(s2+s3)            metrics.rs:25-32   -> xor 1, xor 2
// is the butterfly below.  fp32 add is commutative, so every lane ends with the reference's exact bits.
//
// Slab layout for that path ("lane-permuted"): C = dim/32 chunks, V = 4|2|1 the widest vector dividing C.
// Element (chunk c, lane t) is stored at word ((c / V) * 32 + t) * V + (c % V), so lane t fetches its V chunks
// of group g with ONE V-wide load at vector index g*32 + t and a warp-wide load instruction covers 128*V
// contiguous bytes (512 B for dim = 128: the whole row in a single LDG.128 per lane).
//
// Scalar path (dim % 32 != 0, metrics.rs:79-84): a strict left fold with separately rounded mul and add,
// inherently sequential, so ONE lane folds one row (natural element order in the slab).
#pragma once
#include "common.cuh"

namespace hnsw {

__host__ __device__ __forceinline__ int dist_vec_width(uint32_t dim) {
  uint32_t C = dim / 32;
  return (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1);
}

// word offset of natural element i inside a lane-permuted row
__host__ __device__ __forceinline__ uint32_t permuted_pos(uint32_t i, int V) {
  uint32_t c = i >> 5, t = i & 31;
  return ((c / V) * 32 + t) * V + (c % V);
}

__device__ __forceinline__ float warp_hsum_avx_order(float acc) {
  acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 8));
  acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 16));
  acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 1));
  acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 2));
  return -acc;  // metrics.rs:75
}

// ---- compile-time C (query chunks live in registers: q[c] = query[32c + lane])

template <int C>
struct RowRegs {
  static constexpr int V = (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1);
  float x[C];
};

template <int C>
__device__ __forceinline__ void load_row_regs(const float* __restrict__ row, int lane, RowRegs<C>& r) {
  constexpr int V = RowRegs<C>::V;
#pragma unroll
  for (int g = 0; g < C / V; ++g) {
    if constexpr (V == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row) + g * 32 + lane);
      r.x[4 * g + 0] = v.x, r.x[4 * g + 1] = v.y, r.x[4 * g + 2] = v.z, r.x[4 * g + 3] = v.w;
    } else if constexpr (V == 2) {
      float2 v = __ldg(reinterpret_cast<const float2*>(row) + g * 32 + lane);
      r.x[2 * g + 0] = v.x, r.x[2 * g + 1] = v.y;
    } else {
      r.x[g] = __ldg(row + g * 32 + lane);
    }
  }
}

// per-lane partial (before the butterfly): sequential FMA over chunks, first step is fma(d,d,0)
template <int C>
__device__ __forceinline__ float lane_partial(const float (&q)[C], const RowRegs<C>& r) {
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float d = __fsub_rn(q[c], r.x[c]);
    acc = __fmaf_rn(d, d, acc);
  }
  return acc;
}

// ---- runtime C (query chunks in shared memory, permuted like the rows)

__device__ __forceinline__ float lane_partial_generic(const float* __restrict__ qperm, const float* __restrict__ row,
                                                      uint32_t C, int V, int lane) {
  float acc = 0.0f;
  if (V == 4) {
    for (uint32_t g = 0; g < C / 4; ++g) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row) + g * 32 + lane);
      float4 q = reinterpret_cast<const float4*>(qperm)[g * 32 + lane];
      float d;
      d = __fsub_rn(q.x, v.x), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q.y, v.y), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q.z, v.z), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q.w, v.w), acc = __fmaf_rn(d, d, acc);
    }
  } else if (V == 2) {
    for (uint32_t g = 0; g < C / 2; ++g) {
      float2 v = __ldg(reinterpret_cast<const float2*>(row) + g * 32 + lane);
      float2 q = reinterpret_cast<const float2*>(qperm)[g * 32 + lane];
      float d;
      d = __fsub_rn(q.x, v.x), acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(q.y, v.y), acc = __fmaf_rn(d, d, acc);
    }
  } else {
    for (uint32_t g = 0; g < C; ++g) {
      float d = __fsub_rn(qperm[g * 32 + lane], __ldg(row + g * 32 + lane));
      acc = __fmaf_rn(d, d, acc);
    }
  }
  return acc;
}

// ---- scalar path: one lane folds one row (metrics.rs:79-84); q and row in natural order
__device__ __forceinline__ float scalar_sim(const float* __restrict__ q, const float* __restrict__ row, uint32_t dim) {
  float acc = 0.0f;
  for (uint32_t i = 0; i < dim; ++i) {
    float d = __fsub_rn(q[i], __ldg(row + i));
    acc = __fadd_rn(acc, __fmul_rn(d, d));
  }
  return -acc;
}

}  // namespace hnsw

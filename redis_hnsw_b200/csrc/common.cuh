// Shared definitions of the B200 HNSW engine: device graph layout, row format, small warp helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hnsw {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;     // empty adjacency slot / "no node" / "no row"
constexpr uint32_t kExpanded = 0x80000000u;  // flag bit in candidate-list ids (node ids are < 2^31)
constexpr uint32_t kFull = 0xFFFFFFFFu;      // full warp mask
constexpr int kPoolIds = 31;                 // ids per overflow row; word 31 links to the next row

// How a vector row is laid out in the slab and reduced (see distance.cuh).
enum DistMode : int {
  kDistScalar = 0,  // dim % 32 != 0: natural order, reference scalar fold (metrics.rs:79-84)
  kDistAvx = 1,     // dim % 32 == 0: lane-permuted order, reference AVX2+FMA order (metrics.rs:48-77)
};

// Device-resident scalars of an index (one int32 array so kernels and the host share one copy).
enum Meta : int {
  kMetaEntry = 0,      // enterpoint id or -1          (core.rs:317)
  kMetaMaxLayer = 1,   // max_layer                    (core.rs:314)
  kMetaPoolUsed = 2,   // overflow rows handed out
  kMetaError = 3,      // sticky device-side error flags
  kMetaCount = 8,
};
constexpr int kErrPoolExhausted = 1;
constexpr int kErrVisitedOverflow = 2;
constexpr int kErrListTooLong = 4;  // an adjacency list outgrew the builder's edit buffer

// Read-only (for search) / mutable (for build) view of the graph in HBM.
//
//   vecs       [n][dim] f32   vector slab; rows are 4*dim bytes (128 B multiples when dim % 32 == 0)
//   adj0       [n][W]   u32   level-0 adjacency rows, W = round_up(m_max_0, 32) words, kEmpty-padded, compact
//   ovf0       [n]      u32   first overflow row of the level-0 list, or kEmpty
//   level      [n]      i32   level drawn for the node (-1 = deleted)
//   upper_base [n]      u32   first upper row of the node (levels 1..level), or kEmpty for level-0-only nodes
//   adjU       [nU][W]  u32   upper-level rows: row = upper_base[node] + (level - 1)
//   ovfU       [nU]     u32
//   locks      [2^k]    u32   hashed row locks, only used by the batched builder
//   pool       [P][32]  u32   overflow rows: 31 ids + link.  The reference does not bound a node's degree
//                             (core.rs:793-795 adds back-edges without a cap check), so lists can outgrow W.
struct Graph {
  float* vecs;
  uint32_t* adj0;
  uint32_t* ovf0;
  uint32_t* upper_base;
  int32_t* level;  // [n] level drawn for the node (rows exist for 0..level); -1 = deleted
  uint32_t* adjU;
  uint32_t* ovfU;
  uint32_t* pool;
  int32_t* meta;
  uint32_t* locks;  // hashed row locks of the batched builder (build.cuh)
  int lock_shift;   // 32 - log2(number of locks)
  uint32_t W;
  uint32_t dim;
  uint32_t pool_cap;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Row base pointers for (node, level).
__device__ __forceinline__ uint32_t* row_ptr(const Graph& g, uint32_t node, uint32_t level, uint32_t** ovf) {
  if (level == 0) {
    *ovf = g.ovf0 + node;
    return g.adj0 + (size_t)node * g.W;
  }
  uint32_t base = g.upper_base[node];
  if (base == kEmpty || (int32_t)level > g.level[node]) {  // node has no such level (reference: push_levels makes an empty list, core.rs:642)
    *ovf = nullptr;
    return nullptr;
  }
  uint32_t r = base + (level - 1);
  *ovf = g.ovfU + r;
  return g.adjU + (size_t)r * g.W;
}

}  // namespace hnsw

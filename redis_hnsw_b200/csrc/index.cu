// Host side of libhnsw_b200.so: the index object, device memory management, layout kernels and the C ABI
// declared in include/hnsw_b200.h.  The hot kernels live in search.cuh (search) and build.cuh (insert/delete).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/hnsw_b200.h"
#include "index.hpp"

namespace hnsw {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(e == cudaErrorMemoryAllocation ? HNSW_ERR_OOM : HNSW_ERR_CUDA, "CUDA error in %s: %s", what,
              cudaGetErrorString(e));
}

// ---------------------------------------------------------------- small layout kernels

// natural [rows][dim] -> slab rows (lane-permuted when mode == kDistAvx), rows first_row..first_row+rows-1
__global__ void pack_rows_kernel(const float* __restrict__ src, float* __restrict__ slab, uint64_t first_row,
                                 uint64_t rows, uint32_t dim, int mode, int V) {
  uint64_t total = rows * dim;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / dim;
    uint32_t e = (uint32_t)(i % dim);
    uint32_t pos = (mode == kDistAvx) ? permuted_pos(e, V) : e;
    slab[(first_row + r) * dim + pos] = src[i];
  }
}

__global__ void unpack_rows_kernel(const float* __restrict__ slab, float* __restrict__ dst, uint64_t first_row,
                                   uint64_t rows, uint32_t dim, int mode, int V) {
  uint64_t total = rows * dim;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / dim;
    uint32_t e = (uint32_t)(i % dim);
    uint32_t pos = (mode == kDistAvx) ? permuted_pos(e, V) : e;
    dst[i] = slab[(first_row + r) * dim + pos];
  }
}

// metrics.rs:14-84 for independent row pairs (natural order in memory): one warp per pair on the AVX path,
// one lane per pair on the scalar path.
__global__ void l2_batch_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t rows, uint32_t dim,
                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  if (dim % 32 == 0) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < rows; r += nwarps) {
      const float* x = a + r * dim;
      const float* y = b + r * dim;
      float acc = 0.f;
      for (uint32_t i = lane; i < dim; i += 32) {  // lane t owns i % 32 == t, increasing chunk order
        float d = __fsub_rn(x[i], y[i]);
        acc = __fmaf_rn(d, d, acc);
      }
      float s = warp_hsum_avx_order(acc);
      if (lane == 0) out[r] = s;
    }
  } else {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < rows; r += (uint64_t)gridDim.x * blockDim.x)
      out[r] = scalar_sim(a + r * dim, b + r * dim, dim);
  }
}

// one warp per requested (node, level) row: fixed row + overflow chain -> out_ids[r * stride ..], full length -> out_lens[r]
__global__ void gather_rows_kernel(Graph g, uint32_t n_rows, const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ levels,
                                   uint32_t stride, uint32_t* __restrict__ out_ids, uint32_t* __restrict__ out_lens) {
  const int lane = threadIdx.x & 31;
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  uint32_t* ovf;
  const uint32_t* row = row_ptr(g, nodes[r], levels[r], &ovf);
  uint32_t len = 0;
  if (row) {
    uint32_t* out = out_ids + (size_t)r * stride;
    bool more = true;
    for (uint32_t w = 0; w < g.W / 32 && more; ++w) {
      const uint32_t nb = row[w * 32 + lane];
      const uint32_t cnt = __popc(__ballot_sync(kFull, nb != kEmpty));
      if (nb != kEmpty && len + lane < stride) out[len + lane] = nb;   // rows are compact: valid ids form a prefix
      len += cnt;
      more = cnt == 32;
    }
    uint32_t link = more ? *ovf : kEmpty;
    while (link != kEmpty) {
      const uint32_t nb = g.pool[(size_t)link * 32 + lane];
      const uint32_t next = __shfl_sync(kFull, nb, 31);
      const bool valid = lane < kPoolIds && nb != kEmpty;
      const uint32_t cnt = __popc(__ballot_sync(kFull, valid));
      if (valid && len + lane < stride) out[len + lane] = nb;
      len += cnt;
      link = cnt == (uint32_t)kPoolIds ? next : kEmpty;
    }
  }
  if (lane == 0) out_lens[r] = len;
}

// ---------------------------------------------------------------- Index: memory

static cudaError_t grow_buf(void** p, size_t old_bytes, size_t new_bytes, int fill, cudaStream_t s) {
  void* n = nullptr;
  cudaError_t e = cudaMalloc(&n, new_bytes ? new_bytes : 16);
  if (e != cudaSuccess) return e;
  if (fill >= 0) {
    e = cudaMemsetAsync(n, fill, new_bytes, s);
    if (e != cudaSuccess) return e;
  }
  if (*p && old_bytes) {
    e = cudaMemcpyAsync(n, *p, std::min(old_bytes, new_bytes), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
  }
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  if (*p) cudaFree(*p);
  *p = n;
  return cudaSuccess;
}

int Index::use_device() {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  return HNSW_OK;
}

int Index::ensure_nodes(uint64_t n) {
  if (n <= cap_nodes) return HNSW_OK;
  uint64_t nc = std::max<uint64_t>(n, std::max<uint64_t>(1024, cap_nodes * 2));
  cudaError_t e;
#define GROW(ptr, per, fill)                                                                              \
  e = grow_buf((void**)&ptr, (size_t)cap_nodes * (per), (size_t)nc * (per), fill, stream);                \
  if (e != cudaSuccess) return cuda_fail(e, "grow " #ptr);
  GROW(g.vecs, (size_t)dim * 4, -1)
  GROW(g.adj0, (size_t)g.W * 4, 0xFF)
  GROW(g.ovf0, 4, 0xFF)
  GROW(g.upper_base, 4, 0xFF)
  GROW(g.level, 4, 0xFF)
  GROW(d_stamp0, 4, 0)
  GROW(d_ver0, 4, 0)
#undef GROW
  cap_nodes = nc;
  return HNSW_OK;
}

int Index::ensure_upper(uint64_t rows) {
  if (rows <= cap_upper) return HNSW_OK;
  uint64_t nc = std::max<uint64_t>(rows, std::max<uint64_t>(1024, cap_upper * 2));
  cudaError_t e = grow_buf((void**)&g.adjU, (size_t)cap_upper * g.W * 4, (size_t)nc * g.W * 4, 0xFF, stream);
  if (e != cudaSuccess) return cuda_fail(e, "grow adjU");
  e = grow_buf((void**)&g.ovfU, (size_t)cap_upper * 4, (size_t)nc * 4, 0xFF, stream);
  if (e != cudaSuccess) return cuda_fail(e, "grow ovfU");
  e = grow_buf((void**)&d_stampU, (size_t)cap_upper * 4, (size_t)nc * 4, 0, stream);
  if (e != cudaSuccess) return cuda_fail(e, "grow stampU");
  e = grow_buf((void**)&d_verU, (size_t)cap_upper * 4, (size_t)nc * 4, 0, stream);
  if (e != cudaSuccess) return cuda_fail(e, "grow verU");
  cap_upper = nc;
  return HNSW_OK;
}

int Index::ensure_pool(uint64_t rows) {
  if (rows <= g.pool_cap) return HNSW_OK;
  uint64_t nc = std::max<uint64_t>(rows, std::max<uint64_t>(4096, (uint64_t)g.pool_cap * 2));
  if (nc >= 0x7FFFFFFFull) return fail(HNSW_ERR_OOM, "overflow-row pool too large");
  cudaError_t e = grow_buf((void**)&g.pool, (size_t)g.pool_cap * 32 * 4, (size_t)nc * 32 * 4, 0xFF, stream);
  if (e != cudaSuccess) return cuda_fail(e, "grow pool");
  g.pool_cap = (uint32_t)nc;
  return HNSW_OK;
}

int Index::ensure_scratch(Scratch& s, size_t bytes) {
  if (bytes <= s.bytes) return HNSW_OK;
  if (s.p) cudaFree(s.p);
  s.p = nullptr;
  s.bytes = 0;
  size_t nb = std::max(bytes, (size_t)4096);
  cudaError_t e = cudaMalloc(&s.p, nb);
  if (e != cudaSuccess) return cuda_fail(e, "scratch alloc");
  s.bytes = nb;
  return HNSW_OK;
}

int Index::push_meta() {
  int32_t m[kMetaCount] = {0};
  m[kMetaEntry] = entry;
  m[kMetaMaxLayer] = max_layer;
  m[kMetaPoolUsed] = (int32_t)pool_used;
  m[kMetaError] = 0;
  cudaError_t e = cudaMemcpyAsync(g.meta, m, sizeof(m), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "push meta");
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "push meta sync");
  return HNSW_OK;
}

int Index::pull_meta() {
  int32_t m[kMetaCount];
  cudaError_t e = cudaMemcpyAsync(m, g.meta, sizeof(m), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pull meta");
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "pull meta sync");
  entry = m[kMetaEntry];
  max_layer = m[kMetaMaxLayer];
  pool_used = (uint32_t)m[kMetaPoolUsed];
  device_error = m[kMetaError];
  return HNSW_OK;
}

Index::~Index() {
  if (cudaSetDevice(device) != cudaSuccess) return;
  void* ptrs[] = {g.vecs, g.adj0, g.ovf0, g.upper_base, g.level, g.adjU, g.ovfU, g.pool, g.meta, g.locks,
                  d_stamp0, d_stampU, d_ver0, d_verU, s_in.p, s_out.p, s_vis.p, s_ctl.p, s_build.p, s_stage.p, s_bvis.p, s_spec.p, s_list.p};
  if (h_retry_seen) cudaFreeHost(h_retry_seen);
  if (h_stage) cudaFreeHost(h_stage);
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (stream) cudaStreamDestroy(stream);
  for (cudaStream_t a : aux_stream)
    if (a) cudaStreamDestroy(a);
}

// ---------------------------------------------------------------- Index: vectors

// upload `rows` natural-order host vectors into slab rows [first_row, first_row+rows)
int Index::upload_vectors(const float* host, uint64_t first_row, uint64_t rows) {
  if (rows == 0) return HNSW_OK;
  const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)dim * 4));
  for (uint64_t r0 = 0; r0 < rows; r0 += chunk) {
    uint64_t n = std::min(chunk, rows - r0);
    int rc = ensure_scratch(s_stage, (size_t)n * dim * 4);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(s_stage.p, host + r0 * dim, (size_t)n * dim * 4, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "vector H2D");
    int grid = (int)std::min<uint64_t>(148 * 8, (n * dim + 255) / 256);
    pack_rows_kernel<<<grid, 256, 0, stream>>>((const float*)s_stage.p, g.vecs, first_row + r0, n, dim, dist_mode, vecw);
    g_launches++;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "pack_rows");
  }
  return HNSW_OK;
}

int Index::download_vectors(float* host, uint64_t first_row, uint64_t rows) {
  if (rows == 0) return HNSW_OK;
  const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)dim * 4));
  for (uint64_t r0 = 0; r0 < rows; r0 += chunk) {
    uint64_t n = std::min(chunk, rows - r0);
    int rc = ensure_scratch(s_stage, (size_t)n * dim * 4);
    if (rc) return rc;
    int grid = (int)std::min<uint64_t>(148 * 8, (n * dim + 255) / 256);
    unpack_rows_kernel<<<grid, 256, 0, stream>>>(g.vecs, (float*)s_stage.p, first_row + r0, n, dim, dist_mode, vecw);
    g_launches++;
    cudaError_t e = cudaMemcpyAsync(host + r0 * dim, s_stage.p, (size_t)n * dim * 4, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return cuda_fail(e, "vector D2H");
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "unpack_rows");
  }
  return HNSW_OK;
}

// ---------------------------------------------------------------- Index: whole-graph load / export

int Index::load_graph(uint64_t n, const float* vectors, const int32_t* levels, const uint64_t* row_offs,
                      const uint32_t* nbrs, int64_t entry_, int32_t max_layer_) {
  int rc = use_device();
  if (rc) return rc;
  if (n >= 0x7FFFFFFFull) return fail(HNSW_ERR_INVALID, "too many nodes");
  // host-side row images
  const uint32_t W = g.W;
  std::vector<uint32_t> adj0((size_t)n * W, kEmpty), ovf0(n, kEmpty), ubase(n, kEmpty);
  std::vector<int32_t> lvl(levels, levels + n);
  uint64_t n_upper = 0, live = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (levels[i] >= 0) {
      live++;
      if (levels[i] > 0) {
        ubase[i] = (uint32_t)n_upper;
        n_upper += (uint64_t)levels[i];
      }
    }
  std::vector<uint32_t> adjU((size_t)std::max<uint64_t>(n_upper, 1) * W, kEmpty), ovfU(std::max<uint64_t>(n_upper, 1), kEmpty);
  std::vector<uint32_t> pool;
  auto fill_row = [&](uint32_t* row, uint32_t* ovf, const uint32_t* src, uint64_t cnt) {
    uint64_t k = std::min<uint64_t>(cnt, W);
    std::copy(src, src + k, row);
    int64_t prev = -1;  // previous overflow row of this chain
    while (k < cnt) {   // chain overflow rows of 31 ids + link
      uint32_t pr = (uint32_t)(pool.size() / 32);
      pool.resize(pool.size() + 32, kEmpty);
      if (prev < 0) *ovf = pr;
      else pool[(size_t)prev * 32 + 31] = pr;
      uint64_t t = std::min<uint64_t>(cnt - k, kPoolIds);
      std::copy(src + k, src + k + t, pool.begin() + (size_t)pr * 32);
      k += t;
      prev = pr;
    }
  };
  uint64_t r = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (levels[i] < 0) continue;
    for (int l = 0; l <= levels[i]; ++l, ++r) {
      const uint32_t* src = nbrs + row_offs[r];
      uint64_t cnt = row_offs[r + 1] - row_offs[r];
      for (uint64_t t = 0; t < cnt; ++t)
        if (src[t] >= n || levels[src[t]] < l) return fail(HNSW_ERR_INVALID, "graph row (%llu,%d) references a bad node",
                                                            (unsigned long long)i, l);
      if (l == 0) fill_row(adj0.data() + i * W, &ovf0[i], src, cnt);
      else {
        uint64_t ur = ubase[i] + (uint64_t)(l - 1);
        fill_row(adjU.data() + ur * W, &ovfU[ur], src, cnt);
      }
    }
  }
  if (entry_ >= 0 && ((uint64_t)entry_ >= n || levels[entry_] < 0)) return fail(HNSW_ERR_INVALID, "bad enterpoint");
  // the descent starts at (enterpoint, max_layer): a record whose max_layer exceeds the enterpoint's own level (a corrupt
  // RDB) would make the kernels read upper rows that do not exist
  if (max_layer_ < 0 || (entry_ >= 0 && max_layer_ > levels[entry_]))
    return fail(HNSW_ERR_INVALID, "max_layer %d does not match the enterpoint's level", max_layer_);

  if ((rc = ensure_nodes(std::max<uint64_t>(n, 1)))) return rc;
  if ((rc = ensure_upper(std::max<uint64_t>(n_upper, 1)))) return rc;
  if ((rc = ensure_pool(pool.size() / 32 + 4096))) return rc;
  cudaError_t e;
  // reset everything beyond the loaded prefix as well (the index may have held a larger graph)
  e = cudaMemsetAsync(g.adj0, 0xFF, (size_t)cap_nodes * W * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.ovf0, 0xFF, (size_t)cap_nodes * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.upper_base, 0xFF, (size_t)cap_nodes * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.level, 0xFF, (size_t)cap_nodes * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.adjU, 0xFF, (size_t)cap_upper * W * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.ovfU, 0xFF, (size_t)cap_upper * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.pool, 0xFF, (size_t)g.pool_cap * 32 * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_ver0, 0, (size_t)cap_nodes * 4, stream);   // ids restart: stamps of the old graph are void
  if (e == cudaSuccess) e = cudaMemsetAsync(d_verU, 0, (size_t)cap_upper * 4, stream);
#define UP(dst, vec, cnt)                                                                                 \
  if (e == cudaSuccess && (cnt) != 0)                                                                     \
    e = cudaMemcpyAsync(dst, vec.data(), (size_t)(cnt) * sizeof(vec[0]), cudaMemcpyHostToDevice, stream);
  UP(g.adj0, adj0, n * W)
  UP(g.ovf0, ovf0, n)
  UP(g.upper_base, ubase, n)
  UP(g.level, lvl, n)
  UP(g.adjU, adjU, n_upper * W)
  UP(g.ovfU, ovfU, n_upper)
  UP(g.pool, pool, pool.size())
#undef UP
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "load_graph upload");
  if ((rc = upload_vectors(vectors, 0, n))) return rc;

  h_level.assign(levels, levels + n);
  h_upper_base = ubase;
  n_ids = n;
  node_count = live;
  upper_used = n_upper;
  pool_used = (uint32_t)(pool.size() / 32);
  entry = (int32_t)entry_;
  max_layer = max_layer_;
  touched.clear();
  return push_meta();
}

// download the adjacency into host row images
int Index::snapshot_rows(HostRows& h) {
  int rc = use_device();
  if (rc) return rc;
  if ((rc = pull_meta())) return rc;
  const uint32_t W = g.W;
  h.adj0.resize((size_t)n_ids * W);
  h.ovf0.resize(n_ids);
  h.adjU.resize((size_t)upper_used * W);
  h.ovfU.resize(upper_used);
  h.pool.resize((size_t)pool_used * 32);
  cudaError_t e = cudaSuccess;
#define DOWN(vec, src)                                                                                      \
  if (e == cudaSuccess && !vec.empty())                                                                     \
    e = cudaMemcpyAsync(vec.data(), src, vec.size() * sizeof(vec[0]), cudaMemcpyDeviceToHost, stream);
  DOWN(h.adj0, g.adj0)
  DOWN(h.ovf0, g.ovf0)
  DOWN(h.adjU, g.adjU)
  DOWN(h.ovfU, g.ovfU)
  DOWN(h.pool, g.pool)
#undef DOWN
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cuda_fail(e, "snapshot_rows");
  return HNSW_OK;
}

void Index::row_list(const HostRows& h, uint32_t node, uint32_t level, std::vector<uint32_t>& out) const {
  out.clear();
  const uint32_t W = g.W;
  const uint32_t* row;
  uint32_t link;
  if (level == 0) {
    row = h.adj0.data() + (size_t)node * W;
    link = h.ovf0[node];
  } else {
    uint32_t ur = h_upper_base[node] + (level - 1);
    row = h.adjU.data() + (size_t)ur * W;
    link = h.ovfU[ur];
  }
  for (uint32_t i = 0; i < W; ++i) {
    if (row[i] == kEmpty) return;
    out.push_back(row[i]);
  }
  while (link != kEmpty && (size_t)link * 32 + 31 < h.pool.size()) {
    const uint32_t* pr = h.pool.data() + (size_t)link * 32;
    for (int i = 0; i < kPoolIds; ++i) {
      if (pr[i] == kEmpty) return;
      out.push_back(pr[i]);
    }
    link = pr[31];
  }
}

}  // namespace hnsw

// ================================================================ C ABI

using namespace hnsw;

extern "C" {

const char* hnsw_last_error(void) { return g_last_error.c_str(); }
const char* hnsw_version(void) { return "hnsw_b200 0.1 (sm_100a)"; }
uint64_t hnsw_launch_count(void) { return g_launches.load(); }

int hnsw_index_create(uint32_t data_dim, uint32_t m, uint32_t ef_construction, int device, hnsw_index_t** out) {
  if (!out) return fail(HNSW_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (data_dim == 0 || m == 0 || ef_construction == 0) return fail(HNSW_ERR_INVALID, "dim, m and ef_construction must be > 0");
  if (efr_for(ef_construction) == 0) return fail(HNSW_ERR_INVALID, "ef_construction > %u is not supported", kMaxMemEf);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(HNSW_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(e));
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  }
  if (device >= ndev) return fail(HNSW_ERR_INVALID, "device %d out of range", device);
  hnsw_index* h = new (std::nothrow) hnsw_index();
  if (!h) return fail(HNSW_ERR_OOM, "host allocation failed");
  Index& ix = h->impl;
  ix.device = device;
  ix.dim = data_dim;
  ix.m = m;
  ix.m_max = m;                                   // core.rs:335
  ix.m_max_0 = 2 * m;                             // core.rs:336
  ix.ef_construction = ef_construction;
  ix.level_mult = 1.0 / std::log(1.0 * (double)m);  // core.rs:338
  ix.dist_mode = (data_dim % 32 == 0) ? kDistAvx : kDistScalar;
  ix.vecw = (ix.dist_mode == kDistAvx) ? dist_vec_width(data_dim) : 1;
  ix.kind = (ix.dist_mode == kDistScalar) ? kKindScalar
            : (data_dim == 32)            ? kKindR1
            : (data_dim == 128)           ? kKindR4
            : (data_dim == 768)           ? kKindR24
                                          : kKindGeneric;
  ix.g = Graph{};
  ix.g.W = ((ix.m_max_0 + 31) / 32) * 32;
  ix.g.dim = data_dim;
  {  // StdRng::from_entropy() (core.rs:344): every index draws its own level sequence; hnsw_index_seed pins it for tests
    std::random_device rd;
    ix.rng_state = ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ 0x9E3779B97F4A7C15ull;
    // test hook for hosts that cannot call hnsw_index_seed (the Redis module in a fresh process): a pinned level sequence
    if (const char* s = std::getenv("HNSW_LEVEL_SEED")) ix.rng_state = std::strtoull(s, nullptr, 10) ^ 0x9E3779B97F4A7C15ull;
  }
  int rc = ix.use_device();
  if (!rc) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaGetDeviceProperties");
    else {
      ix.num_sms = prop.multiProcessorCount;
      ix.max_smem = (size_t)prop.sharedMemPerBlockOptin;
    }
  }
  if (!rc) {
    e = cudaStreamCreateWithFlags(&ix.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamCreate");
  }
  if (!rc) {
    e = cudaMalloc((void**)&ix.g.meta, sizeof(int32_t) * kMetaCount);
    if (e != cudaSuccess) rc = cuda_fail(e, "meta alloc");
  }
  if (!rc) {
    const int lock_bits = 20;
    e = cudaMalloc((void**)&ix.g.locks, sizeof(uint32_t) << lock_bits);
    if (e == cudaSuccess) e = cudaMemset(ix.g.locks, 0, sizeof(uint32_t) << lock_bits);
    if (e != cudaSuccess) rc = cuda_fail(e, "lock table alloc");
    ix.g.lock_shift = 32 - lock_bits;
  }
  if (!rc) rc = ix.ensure_nodes(1024);
  if (!rc) rc = ix.ensure_upper(1024);
  if (!rc) rc = ix.ensure_pool(4096);
  if (!rc) rc = ix.push_meta();
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return HNSW_OK;
}

void hnsw_index_destroy(hnsw_index_t* idx) { delete idx; }

int hnsw_index_reserve(hnsw_index_t* idx, uint64_t n_nodes) {
  IDX_OR_FAIL(idx)
  int rc = ix.ensure_nodes(n_nodes);
  if (rc) return rc;
  uint64_t exp_upper = (uint64_t)((double)n_nodes / std::max(1.0, (double)ix.m - 1.0) * 1.5) + 1024;
  if ((rc = ix.ensure_upper(exp_upper))) return rc;
  return ix.ensure_pool(n_nodes / 8 + 4096);
}

int hnsw_index_seed(hnsw_index_t* idx, uint64_t seed) {
  IDX_OR_FAIL(idx)
  ix.rng_state = seed ? seed : 0x9E3779B97F4A7C15ull;
  return HNSW_OK;
}

int hnsw_index_params(hnsw_index_t* idx, hnsw_params_t* out) {
  IDX_OR_FAIL(idx)
  if (!out) return fail(HNSW_ERR_INVALID, "null out pointer");
  out->data_dim = ix.dim;
  out->m = ix.m;
  out->m_max = ix.m_max;
  out->m_max_0 = ix.m_max_0;
  out->ef_construction = ix.ef_construction;
  out->max_layer = ix.max_layer;
  out->level_mult = ix.level_mult;
  out->node_count = ix.node_count;
  out->n_ids = ix.n_ids;
  out->enterpoint = ix.entry < 0 ? HNSW_NO_NODE : (uint32_t)ix.entry;
  out->device = ix.device;
  return HNSW_OK;
}

int hnsw_index_node_level(hnsw_index_t* idx, uint32_t id, int32_t* level) {
  IDX_OR_FAIL(idx)
  if (id >= ix.n_ids) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", id);
  *level = ix.h_level[id];
  return HNSW_OK;
}

int hnsw_index_node_vector(hnsw_index_t* idx, uint32_t id, float* out) {
  IDX_OR_FAIL(idx)
  if (id >= ix.n_ids || ix.h_level[id] < 0) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", id);
  return ix.download_vectors(out, id, 1);
}

int hnsw_index_node_neighbors(hnsw_index_t* idx, uint32_t id, uint32_t level, uint32_t* ids, uint64_t cap, uint64_t* n) {
  IDX_OR_FAIL(idx)
  if (id >= ix.n_ids || ix.h_level[id] < 0) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", id);
  if (n) *n = 0;
  if ((int32_t)level > ix.h_level[id]) return HNSW_OK;  // no such list (reference: neighbors.len() <= level)
  // walk the row and its overflow chain with small copies
  const uint32_t W = ix.g.W;
  std::vector<uint32_t> row(W), out;
  uint32_t link = kEmpty;
  const uint32_t* drow;
  const uint32_t* dovf;
  if (level == 0) {
    drow = ix.g.adj0 + (size_t)id * W;
    dovf = ix.g.ovf0 + id;
  } else {
    uint32_t ur = ix.h_upper_base[id] + (level - 1);
    drow = ix.g.adjU + (size_t)ur * W;
    dovf = ix.g.ovfU + ur;
  }
  cudaError_t e = cudaMemcpyAsync(row.data(), drow, W * 4, cudaMemcpyDeviceToHost, ix.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&link, dovf, 4, cudaMemcpyDeviceToHost, ix.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix.stream);
  if (e != cudaSuccess) return cuda_fail(e, "node_neighbors");
  bool done = false;
  for (uint32_t i = 0; i < W && !done; ++i) {
    if (row[i] == kEmpty) done = true;
    else out.push_back(row[i]);
  }
  while (!done && link != kEmpty) {
    uint32_t pr[32];
    e = cudaMemcpyAsync(pr, ix.g.pool + (size_t)link * 32, 128, cudaMemcpyDeviceToHost, ix.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix.stream);
    if (e != cudaSuccess) return cuda_fail(e, "node_neighbors pool");
    for (int i = 0; i < kPoolIds && !done; ++i) {
      if (pr[i] == kEmpty) done = true;
      else out.push_back(pr[i]);
    }
    link = pr[31];
  }
  if (n) *n = out.size();
  for (uint64_t i = 0; i < out.size() && i < cap; ++i) ids[i] = out[i];
  return HNSW_OK;
}

int hnsw_index_rows_batch(hnsw_index_t* idx, uint64_t n_rows, const uint32_t* nodes, const uint32_t* levels, uint32_t stride,
                          uint32_t* out_ids, uint32_t* out_lens) {
  IDX_OR_FAIL(idx)
  if (n_rows == 0) return HNSW_OK;
  if (!nodes || !levels || !out_ids || !out_lens || stride == 0 || n_rows > 0x7FFFFFFFull) return fail(HNSW_ERR_INVALID, "bad arguments");
  for (uint64_t i = 0; i < n_rows; ++i)
    if (nodes[i] >= ix.n_ids || ix.h_level[nodes[i]] < 0) return fail(HNSW_ERR_NOT_FOUND, "Node: %u does not exist", nodes[i]);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t in_b = al(n_rows * 4), out_b = al(n_rows * (size_t)stride * 4);
  int rc = ix.ensure_scratch(ix.s_stage, 3 * in_b + out_b);
  if (rc) return rc;
  char* base = (char*)ix.s_stage.p;
  uint32_t *d_nodes = (uint32_t*)base, *d_levels = (uint32_t*)(base + in_b), *d_lens = (uint32_t*)(base + 2 * in_b),
           *d_ids = (uint32_t*)(base + 3 * in_b);
  cudaError_t e = cudaMemcpyAsync(d_nodes, nodes, n_rows * 4, cudaMemcpyHostToDevice, ix.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_levels, levels, n_rows * 4, cudaMemcpyHostToDevice, ix.stream);
  if (e != cudaSuccess) return cuda_fail(e, "rows_batch H2D");
  gather_rows_kernel<<<(unsigned)((n_rows * 32 + 127) / 128), 128, 0, ix.stream>>>(ix.g, (uint32_t)n_rows, d_nodes, d_levels, stride, d_ids,
                                                                                  d_lens);
  g_launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_lens, d_lens, n_rows * 4, cudaMemcpyDeviceToHost, ix.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_ids, d_ids, n_rows * (size_t)stride * 4, cudaMemcpyDeviceToHost, ix.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix.stream);
  if (e != cudaSuccess) return cuda_fail(e, "rows_batch");
  return HNSW_OK;
}

int hnsw_index_graph_sizes(hnsw_index_t* idx, uint64_t* n_ids, uint64_t* n_rows, uint64_t* n_edges) {
  IDX_OR_FAIL(idx)
  HostRows h;
  int rc = ix.snapshot_rows(h);
  if (rc) return rc;
  uint64_t rows = 0, edges = 0;
  std::vector<uint32_t> lst;
  for (uint64_t i = 0; i < ix.n_ids; ++i) {
    if (ix.h_level[i] < 0) continue;
    for (int l = 0; l <= ix.h_level[i]; ++l) {
      ix.row_list(h, (uint32_t)i, (uint32_t)l, lst);
      edges += lst.size();
      rows++;
    }
  }
  *n_ids = ix.n_ids;
  *n_rows = rows;
  *n_edges = edges;
  return HNSW_OK;
}

int hnsw_index_export_graph(hnsw_index_t* idx, int32_t* levels, uint64_t* row_offs, uint32_t* nbrs, int64_t* entry,
                            int32_t* max_layer) {
  IDX_OR_FAIL(idx)
  HostRows h;
  int rc = ix.snapshot_rows(h);
  if (rc) return rc;
  uint64_t r = 0, e = 0;
  row_offs[0] = 0;
  std::vector<uint32_t> lst;
  for (uint64_t i = 0; i < ix.n_ids; ++i) {
    levels[i] = ix.h_level[i];
    if (ix.h_level[i] < 0) continue;
    for (int l = 0; l <= ix.h_level[i]; ++l) {
      ix.row_list(h, (uint32_t)i, (uint32_t)l, lst);
      for (uint32_t x : lst) nbrs[e++] = x;
      row_offs[++r] = e;
    }
  }
  *entry = ix.entry;
  *max_layer = ix.max_layer;
  return HNSW_OK;
}

int hnsw_index_export_vectors(hnsw_index_t* idx, float* out) {
  IDX_OR_FAIL(idx)
  return ix.download_vectors(out, 0, ix.n_ids);
}

int hnsw_index_load_graph(hnsw_index_t* idx, uint64_t n_ids, const float* vectors, const int32_t* levels,
                          const uint64_t* row_offs, const uint32_t* nbrs, int64_t entry, int32_t max_layer) {
  IDX_OR_FAIL(idx)
  if (n_ids && (!vectors || !levels || !row_offs)) return fail(HNSW_ERR_INVALID, "null graph arrays");
  return ix.load_graph(n_ids, vectors, levels, row_offs, nbrs, entry, max_layer);
}

int hnsw_index_set_option(hnsw_index_t* idx, const char* name, int64_t value) {
  IDX_OR_FAIL(idx)
  if (!name) return fail(HNSW_ERR_INVALID, "null option name");
  std::string n(name);
  if (n == "visited_slots") {
    if (value != 0 && (value < 256 || (value & (value - 1)))) return fail(HNSW_ERR_INVALID, "visited_slots must be 0 or a power of two >= 256");
    ix.opt_vis_slots = (uint32_t)value;
    ix.auto_vis_ef = 0;
  } else if (n == "search_ctas_per_sm") {
    ix.opt_ctas_per_sm = (int)value;
  } else if (n == "search_block") {
    if (value != 0 && (value < 32 || value > 256 || value % 32)) return fail(HNSW_ERR_INVALID, "search_block must be 32..256, multiple of 32");
    ix.opt_block = (int)value;
  } else if (n == "search_impl") {
    if (value < 0 || value > 2) return fail(HNSW_ERR_INVALID, "search_impl must be 0 (auto), 1 or 2");
    ix.opt_search_impl = (int)value;
  } else if (n == "stage_rows") {
    if (value != 0 && value != 4 && value != 8 && value != 16 && value != 32) return fail(HNSW_ERR_INVALID, "stage_rows must be 0, 4, 8, 16 or 32");
    ix.opt_stage_rows = (int)value;
  } else if (n == "recent_slots") {
    if (value != 0 && (value < 64 || (value & (value - 1)))) return fail(HNSW_ERR_INVALID, "recent_slots must be 0 or a power of two >= 64");
    ix.opt_recent_slots = (uint32_t)value;
  } else if (n == "recent_tag") {
    if (value != 0 && value != 32) return fail(HNSW_ERR_INVALID, "recent_tag must be 0 (auto) or 32");
    ix.opt_recent_tag = (int)value;
  } else if (n == "row_copy") {
    if (value < 0 || value > 1) return fail(HNSW_ERR_INVALID, "row_copy must be 0 (bulk-async copies everywhere) or 1 (cp.async for 32-d / 128-d rows)");
    ix.opt_row_copy = (int)value;
  } else if (n == "recent_ways") {
    if (value != 1 && value != 2) return fail(HNSW_ERR_INVALID, "recent_ways must be 1 or 2");
    ix.opt_recent_ways = (int)value;
  } else if (n == "lookahead") {
    if (value < 0 || value > 1) return fail(HNSW_ERR_INVALID, "lookahead must be 0 or 1");
    ix.opt_lookahead = (int)value;
  } else if (n == "search_cta") {
    if (value < 0 || value > 1) return fail(HNSW_ERR_INVALID, "search_cta must be 0 or 1");
    ix.opt_search_cta = (int)value;
  } else if (n == "build_batch") {
    if (value < 1) return fail(HNSW_ERR_INVALID, "build_batch must be >= 1");
    ix.opt_build_batch = (uint32_t)value;
  } else if (n == "spec_window") {
    if (value < 0 || value > 1024) return fail(HNSW_ERR_INVALID, "spec_window must be 0 (adaptive) .. 1024");
    ix.opt_spec_window = (uint32_t)value;
  } else if (n == "spec_mult") {
    if (value < 0 || value > 1000) return fail(HNSW_ERR_INVALID, "spec_mult must be 0 (default 30 = 3.0x) .. 1000");
    ix.opt_spec_mult = (uint32_t)value;
  } else if (n == "spec_budget_us") {
    if (value < 0 || value > 100000) return fail(HNSW_ERR_INVALID, "spec_budget_us must be 0 (never suspend; default) .. 100000");
    ix.opt_spec_budget_us = (int)value;
  } else if (n == "spec_ahead") {
    if (value < -1 || value > 512) return fail(HNSW_ERR_INVALID, "spec_ahead must be -1 (off), 0 (default: 2 x window) .. 512");
    ix.opt_spec_ahead = (int)value;
  } else if (n == "spec_validation") {
    if (value < 0 || value > 2) return fail(HNSW_ERR_INVALID, "spec_validation must be 0 (default), 1 (row-level) or 2 (dependency-level)");
    ix.opt_spec_validation = (int)value;
  } else if (n == "build_impl") {
    if (value < 0 || value > 2) return fail(HNSW_ERR_INVALID, "build_impl must be 0 (auto), 1 (register-staged) or 2 (TMA-staged)");
    ix.opt_build_impl = (int)value;
  } else {
    return fail(HNSW_ERR_INVALID, "unknown option %s", name);
  }
  return HNSW_OK;
}

int hnsw_l2_batch(const float* a, const float* b, uint64_t rows, uint32_t dim, float* out, int device) {
  if (rows == 0) return HNSW_OK;
  if (!a || !b || !out || dim == 0) return fail(HNSW_ERR_INVALID, "bad arguments");
  cudaError_t e;
  if (device >= 0 && (e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  float *da = nullptr, *db = nullptr, *dout = nullptr;
  size_t bytes = (size_t)rows * dim * 4;
  int rc = HNSW_OK;
  if ((e = cudaMalloc(&da, bytes)) != cudaSuccess || (e = cudaMalloc(&db, bytes)) != cudaSuccess ||
      (e = cudaMalloc(&dout, rows * 4)) != cudaSuccess)
    rc = cuda_fail(e, "l2_batch alloc");
  if (!rc) {
    e = cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      uint64_t threads = (dim % 32 == 0) ? rows * 32 : rows;
      int grid = (int)std::min<uint64_t>(148 * 16, (threads + 255) / 256);
      l2_batch_kernel<<<grid, 256>>>(da, db, rows, dim, dout);
      g_launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, rows * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "l2_batch");
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return rc;
}

}  // extern "C"

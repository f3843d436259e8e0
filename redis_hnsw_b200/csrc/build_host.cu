// Mutation entry points of the C ABI (insert / delete / replication).  Kernels: build.cuh.
#include "../../include/hnsw_b200.h"
#include "index.hpp"

using namespace hnsw;

extern "C" {

int hnsw_index_add(hnsw_index_t* idx, const float* data, uint64_t n, int32_t level, uint32_t* out_id) {
  IDX_OR_FAIL(idx)
  (void)data, (void)n, (void)level, (void)out_id;
  return fail(HNSW_ERR_INVALID, "hnsw_index_add: not implemented yet");
}

int hnsw_index_add_batch(hnsw_index_t* idx, uint64_t count, const float* data, const int32_t* levels, int mode,
                         uint32_t* first_id) {
  IDX_OR_FAIL(idx)
  (void)count, (void)data, (void)levels, (void)mode, (void)first_id;
  return fail(HNSW_ERR_INVALID, "hnsw_index_add_batch: not implemented yet");
}

int hnsw_index_touched(hnsw_index_t* idx, uint32_t* ids, uint64_t cap, uint64_t* n) {
  IDX_OR_FAIL(idx)
  if (n) *n = ix.touched.size();
  for (uint64_t i = 0; i < ix.touched.size() && i < cap; ++i) ids[i] = ix.touched[i];
  return HNSW_OK;
}

int hnsw_index_delete(hnsw_index_t* idx, uint32_t id) {
  IDX_OR_FAIL(idx)
  (void)id;
  return fail(HNSW_ERR_INVALID, "hnsw_index_delete: not implemented yet");
}

int hnsw_index_build_stats(hnsw_index_t* idx, uint64_t* out4) {
  IDX_OR_FAIL(idx)
  for (int i = 0; i < 4; ++i) out4[i] = ix.build_stats[i];
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
    void* olds[] = {ix.g.vecs, ix.g.adj0, ix.g.ovf0, ix.g.upper_base, ix.g.level};
    for (void* p : olds) cudaFree(p);
    ix.g.vecs = nullptr, ix.g.adj0 = nullptr, ix.g.ovf0 = nullptr, ix.g.upper_base = nullptr, ix.g.level = nullptr;
    if ((rc = ix.ensure_nodes(layout8[0]))) return rc;
  }
  if (ix.cap_upper < layout8[1]) {
    cudaFree(ix.g.adjU), cudaFree(ix.g.ovfU);
    ix.g.adjU = nullptr, ix.g.ovfU = nullptr, ix.cap_upper = 0;
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

// blocks.cu — trainer-side block construction: COO of one message-flow block -> CSC.
//
// The reference trainer rebuilds a DGL block from the COO edge list on every step
// (training_backend/legion_graphsage.py:66-79: create_unitgraph_from_coo(2, num_src, num_dst, src, dst, 'coo',
// row_sorted=True) although the list is not sorted) and DGL converts it to CSC for the SpMM of every layer.
// lg_block_csc produces that CSC once, on the trainer's stream, straight from the CUDA-IPC buffers:
//   indptr[d] .. indptr[d+1]  = the in-edges of destination d (batch-local index), in COO order (stable)
//   indices[k]                = batch-local source of the k-th in-edge
//   eids[k]                   = its position in the COO (DGL edge id), optional
// Stable order makes the result a pure function of the COO, so it is checked bit-exactly against the oracle.
// Sort: CUB radix sort over the ceil(log2(num_dst)) significant key bits (library call, like the ranking sort of the
// cache build); the boundary search and the source gather are the kernels below.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

using namespace lg;

namespace {

__global__ void __launch_bounds__(256) iota_kernel(int32_t* __restrict__ v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (int32_t)i;
}

// sorted_dst[k] ascending: indptr[d] = first k with sorted_dst[k] >= d; also indices[k] = src[eid[k]]
__global__ void __launch_bounds__(256) csc_finish_kernel(const int32_t* __restrict__ sorted_dst, const int32_t* __restrict__ eid,
                                                         const int32_t* __restrict__ src, int64_t n_edges, int32_t num_dst,
                                                         int32_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_edges; k += stride) {
    // destinations in (prev, cur] start at k (prev = -1 before the first edge, cur = num_dst after the last)
    const int32_t prev = (k == 0) ? -1 : sorted_dst[k - 1];
    int32_t cur = (k == n_edges) ? num_dst : sorted_dst[k];
    if (cur > num_dst) cur = num_dst;
    for (int32_t d = prev + 1; d <= cur; d++) indptr[d] = (int32_t)k;
    if (k < n_edges) indices[k] = src[eid[k]];
  }
}

}  // namespace

extern "C" int lg_block_csc_workspace(int64_t max_edges, int64_t* bytes) {
  LG_REQUIRE(bytes && max_edges >= 0 && max_edges < (1ll << 31), "lg_block_csc_workspace: bad argument");
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)max_edges, 0, 31, (cudaStream_t)0);
  // [iota eid_in][sorted dst][eid_out when the caller passes no eids][cub temp]
  const int64_t a = (max_edges * 4 + 255) & ~255ll;
  *bytes = 3 * a + (int64_t)cub_bytes + 256;
  return 0;
}

extern "C" int lg_block_csc(lg_stream_t stream, const int32_t* agg_src, const int32_t* agg_dst, int64_t n_edges,
                            int32_t num_dst, int32_t* indptr, int32_t* indices, int32_t* eids, void* workspace,
                            int64_t workspace_bytes) {
  LG_REQUIRE(indptr && num_dst >= 0, "lg_block_csc: null indptr / negative num_dst");
  LG_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31), "lg_block_csc: n_edges %lld", (long long)n_edges);
  LG_REQUIRE(n_edges == 0 || (agg_src && agg_dst && indices), "lg_block_csc: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t need = 0;
  int rc = lg_block_csc_workspace(n_edges, &need);
  if (rc) return rc;
  LG_REQUIRE(n_edges == 0 || (workspace && workspace_bytes >= need), "lg_block_csc: workspace of %lld bytes, need %lld",
             (long long)workspace_bytes, (long long)need);
  const int64_t a = (n_edges * 4 + 255) & ~255ll;
  int32_t* iota = (int32_t*)workspace;
  int32_t* sorted_dst = (int32_t*)((char*)workspace + a);
  int32_t* eid_out = eids ? eids : (int32_t*)((char*)workspace + 2 * a);
  void* cub_tmp = (char*)workspace + 3 * a;
  if (n_edges > 0) {
    int grid = (int)((n_edges + 255) / 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    iota_kernel<<<grid, 256, 0, st>>>(iota, n_edges);
    LG_LAUNCH_OK();
    int end_bit = 1;
    while (end_bit < 31 && (1ll << end_bit) < (int64_t)num_dst) end_bit++;
    size_t cub_bytes = (size_t)(workspace_bytes - 3 * a);
    LG_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, agg_dst, sorted_dst, iota, eid_out, (int)n_edges, 0, end_bit, st));
  }
  int grid = (int)((n_edges + 1 + 255) / 256);
  if (grid > kSMs * 8) grid = kSMs * 8;
  csc_finish_kernel<<<grid, 256, 0, st>>>(sorted_dst, eid_out, agg_src, n_edges, num_dst, indptr, indices);
  LG_LAUNCH_OK();
  return 0;
}

// cache_build.cu — unified-cache construction (init-time): hotness aggregation, ranking,
// interleaved placement, shard fill.  Restates cache/cache.cu:360-443 (CandidateSelection),
// :71-136 (InitializeMap/Insert), :553-611 (FillUp), storage/graph_storage.cu:76-111 (GraphCache)
// with dense int32 directories instead of three bucketed-cuckoo hash maps per GPU, 16-byte row
// copies instead of one thread per float, and a warp per adjacency list instead of one thread.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace lg;

namespace {
constexpr int kBlock = 256;
inline int grid_for(int64_t n, int per_thread = 1) {
  int64_t g = (n + (int64_t)kBlock * per_thread - 1) / ((int64_t)kBlock * per_thread);
  int64_t cap = (int64_t)kSMs * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

__global__ void accumulate_kernel(u64* __restrict__ agg, const u64* __restrict__ part, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    agg[i] += part[i];  // aggregate_access, cache/cache_impl.cuh:72-76
}
__global__ void iota_kernel(int32_t* v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (int32_t)i;  // init_cache_order, cache/cache_impl.cuh:79-83
}
__global__ void fill_kernel(int32_t* v, int32_t x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = x;
}
// InitPair / InitIndexPair+InitOffsetPair followed by the hash insert, as one directory scatter.
// Hybrid placement (rep > 0, features only): the rep hottest ranks live on EVERY GPU of the clique at row r and resolve to
// the reader's own part (`self`); the ranks after them are interleaved over the parts like the reference's, below row rep.
__global__ void place_kernel(const int32_t* __restrict__ order, int32_t cap, int32_t kg, int32_t part_base,
                             int64_t num_nodes, int32_t* __restrict__ directory, int32_t rep, int32_t self) {
  const int64_t total = (int64_t)rep + (int64_t)(cap - rep) * kg;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total && r < num_nodes;
       r += (int64_t)gridDim.x * blockDim.x) {
    int32_t part, row;
    if (r < rep) {
      part = self;
      row = (int32_t)r;
    } else {
      const int64_t q = r - rep;
      part = (int32_t)(q % kg);
      row = rep + (int32_t)(q / kg);
    }
    directory[order[r]] = (part + part_base) * cap + row;
  }
}
// FeatFillUp: one warp per row, 16-byte chunks when the row allows it
__global__ void fill_feature_kernel(const int32_t* __restrict__ order, int32_t cap, int32_t kg, int32_t j,
                                    int32_t dim, int64_t num_nodes, const float* __restrict__ backing,
                                    float* __restrict__ shard, int vec4, int32_t rep) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < cap; r += n_warps) {
    const int64_t rank = r < rep ? r : rep + (r - rep) * kg + j;  // replicated head, interleaved tail
    float* d = shard + r * dim;
    if (rank >= num_nodes) {
      for (int c = lane; c < dim; c += 32) d[c] = 0.f;
      continue;
    }
    const float* s = backing + (int64_t)order[rank] * dim;
    if (vec4) {
      for (int c = lane; c < (dim >> 2); c += 32) reinterpret_cast<float4*>(d)[c] = reinterpret_cast<const float4*>(s)[c];
    } else {
      for (int c = lane; c < dim; c += 32) d[c] = s[c];
    }
  }
}
// GetNeighborCount (storage/graph_storage_impl.cuh:33-39)
__global__ void topo_count_kernel(const int32_t* __restrict__ order, int32_t cap, int32_t kg, int32_t j,
                                  int64_t num_nodes, const int64_t* __restrict__ indptr, int64_t* __restrict__ counts) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < cap; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rank = r * kg + j;
    int64_t c = 0;
    if (rank < num_nodes) {
      int32_t id = order[rank];
      c = indptr[id + 1] - indptr[id];
    }
    counts[r] = c;
  }
}
// TopoFillUp (storage/graph_storage_impl.cuh:41-53): a warp copies one adjacency list
__global__ void topo_fill_kernel(const int32_t* __restrict__ order, int32_t cap, int32_t kg, int32_t j,
                                 int64_t num_nodes, const int64_t* __restrict__ indptr,
                                 const int32_t* __restrict__ indices, const int64_t* __restrict__ shard_indptr,
                                 int32_t* __restrict__ shard_indices) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < cap; r += n_warps) {
    const int64_t rank = r * kg + j;
    if (rank >= num_nodes) continue;
    const int32_t id = order[rank];
    const int64_t s0 = indptr[id], n = indptr[id + 1] - s0, d0 = shard_indptr[r];
    for (int64_t k = lane; k < n; k += 32) shard_indices[d0 + k] = indices[s0 + k];
  }
}
}  // namespace

extern "C" int lg_hotness_accumulate(lg_stream_t stream, unsigned long long* agg, const unsigned long long* part,
                                     int64_t n) {
  LG_REQUIRE(agg && part && n >= 0, "lg_hotness_accumulate: bad argument");
  if (n == 0) return 0;
  accumulate_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>((u64*)agg, (const u64*)part, n);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_hotness_rank(lg_stream_t stream, const unsigned long long* hotness, int64_t n, int32_t* order,
                               unsigned long long* sorted_hotness, void* tmp, int64_t* tmp_bytes) {
  LG_REQUIRE(tmp_bytes && n > 0 && n < (1ll << 31), "lg_hotness_rank: bad argument");
  // workspace layout: [iota int32 n][keys_out u64 n (if caller passes no sorted_hotness)][cub temp]
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending((void*)nullptr, cub_bytes, (const u64*)nullptr, (u64*)nullptr,
                                            (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  const int64_t a = ((n * 4 + 255) / 256) * 256, b = ((n * 8 + 255) / 256) * 256;
  const int64_t need = a + b + (int64_t)cub_bytes;
  if (!tmp) {
    *tmp_bytes = need;
    return 0;
  }
  LG_REQUIRE(*tmp_bytes >= need, "lg_hotness_rank: workspace %lld < %lld", (long long)*tmp_bytes, (long long)need);
  LG_REQUIRE(hotness && order, "lg_hotness_rank: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* iota = (int32_t*)tmp;
  u64* keys_out = sorted_hotness ? (u64*)sorted_hotness : (u64*)((char*)tmp + a);
  void* cub_tmp = (char*)tmp + a + b;
  iota_kernel<<<grid_for(n), kBlock, 0, st>>>(iota, n);
  LG_LAUNCH_OK();
  // LSD radix sort is stable: equal hotness keeps ascending vertex id (the tie rule of the oracle)
  LG_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp, cub_bytes, (const u64*)hotness, keys_out, iota, order,
                                                    (int)n, 0, 64, st));
  return 0;
}

extern "C" int lg_fill_i32(lg_stream_t stream, int32_t* dst, int32_t value, int64_t n) {
  LG_REQUIRE(dst && n >= 0, "lg_fill_i32: bad argument");
  if (n == 0) return 0;
  fill_kernel<<<grid_for(n, 4), kBlock, 0, (cudaStream_t)stream>>>(dst, value, n);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_place_features(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                                 int64_t num_nodes, int32_t* directory) {
  LG_REQUIRE(order && directory && cap > 0 && kg > 0, "lg_place_features: bad argument");
  LG_REQUIRE((int64_t)cap * kg < (1ll << 31), "lg_place_features: cap*kg overflows int32");
  place_kernel<<<grid_for((int64_t)cap * kg), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, 0, num_nodes, directory, 0, 0);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_place_features_hybrid(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t rep,
                                        int32_t j, int64_t num_nodes, int32_t* directory) {
  LG_REQUIRE(order && directory && cap > 0 && kg > 0 && rep >= 0 && rep <= cap && j >= 0 && j < kg,
             "lg_place_features_hybrid: bad argument");
  LG_REQUIRE((int64_t)cap * kg < (1ll << 31), "lg_place_features_hybrid: cap*kg overflows int32");
  place_kernel<<<grid_for((int64_t)cap * kg), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, 0, num_nodes, directory, rep, j);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_place_topology(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t ki,
                                 int64_t num_nodes, int32_t* directory) {
  LG_REQUIRE(order && directory && cap > 0 && kg > 0 && ki >= 0, "lg_place_topology: bad argument");
  LG_REQUIRE((int64_t)cap * kg * (ki + 1) < (1ll << 31), "lg_place_topology: packed location overflows int32");
  place_kernel<<<grid_for((int64_t)cap * kg), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, ki * kg, num_nodes,
                                                                                 directory, 0, 0);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_fill_feature_shard(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t j,
                                     int32_t dim, int64_t num_nodes, const float* backing, float* shard) {
  LG_REQUIRE(order && backing && shard && cap > 0 && kg > 0 && j >= 0 && j < kg && dim > 0,
             "lg_fill_feature_shard: bad argument");
  int vec4 = (dim % 4 == 0) && (((uintptr_t)backing & 15) == 0) && (((uintptr_t)shard & 15) == 0);
  fill_feature_kernel<<<grid_for((int64_t)cap * 32), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, j, dim, num_nodes,
                                                                                      backing, shard, vec4, 0);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_fill_feature_shard_hybrid(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t rep,
                                            int32_t j, int32_t dim, int64_t num_nodes, const float* backing, float* shard) {
  LG_REQUIRE(order && backing && shard && cap > 0 && kg > 0 && rep >= 0 && rep <= cap && j >= 0 && j < kg && dim > 0,
             "lg_fill_feature_shard_hybrid: bad argument");
  int vec4 = (dim % 4 == 0) && (((uintptr_t)backing & 15) == 0) && (((uintptr_t)shard & 15) == 0);
  fill_feature_kernel<<<grid_for((int64_t)cap * 32), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, j, dim, num_nodes,
                                                                                      backing, shard, vec4, rep);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_topo_shard_indptr(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t j,
                                    int64_t num_nodes, const int64_t* indptr, int64_t* shard_indptr) {
  LG_REQUIRE(order && indptr && shard_indptr && cap > 0 && kg > 0 && j >= 0 && j < kg,
             "lg_topo_shard_indptr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // counts land in shard_indptr[1..cap]; an in-place inclusive scan turns them into offsets
  LG_CUDA(cudaMemsetAsync(shard_indptr, 0, sizeof(int64_t), st));
  topo_count_kernel<<<grid_for(cap), kBlock, 0, st>>>(order, cap, kg, j, num_nodes, indptr, shard_indptr + 1);
  LG_LAUNCH_OK();
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum((void*)nullptr, bytes, shard_indptr + 1, shard_indptr + 1, cap, st);
  void* tmp = nullptr;
  LG_CUDA(cudaMallocAsync(&tmp, bytes, st));
  LG_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, shard_indptr + 1, shard_indptr + 1, cap, st));
  LG_CUDA(cudaFreeAsync(tmp, st));
  return 0;
}

extern "C" int lg_topo_shard_fill(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t j,
                                  int64_t num_nodes, const int64_t* indptr, const int32_t* indices,
                                  const int64_t* shard_indptr, int32_t* shard_indices) {
  LG_REQUIRE(order && indptr && indices && shard_indptr && shard_indices && cap > 0 && kg > 0,
             "lg_topo_shard_fill: bad argument");
  topo_fill_kernel<<<grid_for((int64_t)cap * 32), kBlock, 0, (cudaStream_t)stream>>>(order, cap, kg, j, num_nodes, indptr,
                                                                                    indices, shard_indptr, shard_indices);
  LG_LAUNCH_OK();
  return 0;
}

// CostModel — host arithmetic of cache/cache.cu:445-551 (float accumulators, 1 % alpha sweep,
// first maximum, +1 on both capacities).  Not a copy of the oracle: same published algorithm.
static int cost_model_impl(const unsigned long long* sorted_node_hotness,
                           const unsigned long long* sorted_edge_hotness, const int32_t* topo_order,
                           const int64_t* indptr, int64_t num_nodes, int32_t dim, int64_t cache_bytes, int32_t kg,
                           uint64_t topo_trans, uint64_t feat_trans, int32_t* node_capacity, int32_t* edge_capacity,
                           double* alpha, bool saturate) {
  LG_REQUIRE(sorted_node_hotness && sorted_edge_hotness && topo_order && indptr && node_capacity && edge_capacity,
             "lg_cost_model: null argument");
  LG_REQUIRE(num_nodes > 0 && dim > 0 && cache_bytes > 0 && kg > 0, "lg_cost_model: bad size");
  const int64_t n = num_nodes;
  std::vector<uint64_t> node_prefix(n), edge_prefix(n), mem_prefix(n);
  uint64_t sa = 0, sb = 0, sc = 0;
  for (int64_t i = 0; i < n; i++) {
    sa += sorted_node_hotness[i];
    sb += sorted_edge_hotness[i];
    const int32_t id = topo_order[i];
    sc += sizeof(int64_t) + sizeof(int32_t) * (uint64_t)(indptr[id + 1] - indptr[id]);
    node_prefix[i] = sa;
    edge_prefix[i] = sb;
    mem_prefix[i] = sc;
  }
  const int64_t total_mem = cache_bytes * kg;
  int64_t step = (int64_t)((double)total_mem * 0.01);
  if (step <= 0) step = 1;
  const int64_t steps = (total_mem - 1) / step + 1;
  std::vector<float> t_topo(steps + 1, 0.f), t_feat(steps + 1, 0.f), c_topo(steps + 1, 0.f), c_feat(steps + 1, 0.f),
      t_total(steps + 1, 0.f);
  const int64_t row_bytes = (int64_t)dim * (int64_t)sizeof(float);
  int64_t k = 0;
  for (int64_t mem = 0; mem < total_mem; mem += step, k++) {
    int64_t n_feat64 = ((uint64_t)mem > (uint64_t)n * (uint64_t)row_bytes) ? n : (k + 1) * (step / row_bytes);
    if (saturate && n_feat64 > n) n_feat64 = n;
    const int32_t n_feat = (int32_t)n_feat64;
    int32_t n_topo;
    if ((uint64_t)mem > mem_prefix[n - 1])
      n_topo = (int32_t)n;
    else
      n_topo = (int32_t)(std::lower_bound(mem_prefix.begin(), mem_prefix.end(), (uint64_t)mem) - mem_prefix.begin());
    if (n_topo < n || saturate) {
      const uint64_t pre = n_topo > 0 ? edge_prefix[n_topo - 1] : 0;
      t_topo[k] = (float)(topo_trans * 1.0 / (double)edge_prefix[n - 1] * (double)pre);
      c_topo[k] = (float)(n_topo / kg);
    }
    if (n_feat < n || saturate) {
      const uint64_t pre = n_feat > 0 ? node_prefix[n_feat - 1] : 0;
      t_feat[k] = (float)(feat_trans * 1.0 / (double)node_prefix[n - 1] * (double)pre);
      c_feat[k] = (float)(n_feat / kg);
    }
  }
  for (int64_t s = 1; s < steps; s++) t_total[s] = t_topo[s] + t_feat[steps - 1 - s];
  const int64_t best = std::max_element(t_total.begin(), t_total.end()) - t_total.begin();
  if (alpha) *alpha = (double)best * 0.01;
  *node_capacity = (int32_t)(c_feat[steps - 1 - best] + 1);
  *edge_capacity = (int32_t)(c_topo[best] + 1);
  return 0;
}

extern "C" int lg_cost_model(const unsigned long long* sorted_node_hotness,
                             const unsigned long long* sorted_edge_hotness, const int32_t* topo_order,
                             const int64_t* indptr, int64_t num_nodes, int32_t dim, int64_t cache_bytes, int32_t kg,
                             uint64_t topo_trans, uint64_t feat_trans, int32_t* node_capacity, int32_t* edge_capacity,
                             double* alpha) {
  return cost_model_impl(sorted_node_hotness, sorted_edge_hotness, topo_order, indptr, num_nodes, dim, cache_bytes, kg,
                         topo_trans, feat_trans, node_capacity, edge_capacity, alpha, false);
}

extern "C" int lg_cost_model_saturating(const unsigned long long* sorted_node_hotness,
                                        const unsigned long long* sorted_edge_hotness, const int32_t* topo_order,
                                        const int64_t* indptr, int64_t num_nodes, int32_t dim, int64_t cache_bytes,
                                        int32_t kg, uint64_t topo_trans, uint64_t feat_trans, int32_t* node_capacity,
                                        int32_t* edge_capacity, double* alpha) {
  return cost_model_impl(sorted_node_hotness, sorted_edge_hotness, topo_order, indptr, num_nodes, dim, cache_bytes, kg,
                         topo_trans, feat_trans, node_capacity, edge_capacity, alpha, true);
}

// synth.cu — device-side synthetic dataset generator (see include/legion_b200_synth.h).
#include <cub/device/device_scan.cuh>

#include "../../include/legion_b200_synth.h"
#include "common.cuh"

using namespace lg;

namespace {
__host__ __device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ u64 hash2(u64 seed, u64 a) { return mix64(seed ^ mix64(a + 0x9E3779B97F4A7C15ull)); }
__device__ __forceinline__ double unit(u64 h) { return __dmul_rn((double)(h >> 11), 1.1102230246251565e-16); }  // 2^-53

__global__ void degree_kernel(int64_t n, double dmin, int32_t dmax, u64 seed, int64_t* __restrict__ out) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    double x = __dsub_rn(1.0, unit(hash2(seed, (u64)v)));
    long long d = (long long)__ddiv_rn(dmin, __dsqrt_rn(x));
    if (d > dmax) d = dmax;
    out[v] = d;
  }
}
__global__ void indices_kernel(int64_t n, const int64_t* __restrict__ indptr, u64 seed, int32_t* __restrict__ indices) {
  // one warp per vertex: lanes stride over the adjacency list (coalesced stores)
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const u64 s2 = seed ^ 0xA5A5A5A55A5A5A5Aull;
  for (int64_t v = warp; v < n; v += n_warps) {
    const int64_t b = indptr[v], e = indptr[v + 1];
    for (int64_t k = lane; k < e - b; k += 32) {
      double u = unit(hash2(s2, ((u64)v << 21) + (u64)k));
      double t = __dmul_rn(__dmul_rn(u, u), u);
      long long r = (long long)__dmul_rn(t, (double)n);
      if (r >= n) r = n - 1;
      indices[b + k] = (int32_t)(((u64)r * 2654435761ull + 12345ull) % (u64)n);
    }
  }
}
__global__ void features_kernel(int64_t row0, int64_t rows, int32_t dim, u64 seed, float* __restrict__ out) {
  const u64 s3 = seed ^ 0xFEA7FEA7FEA7FEA7ull;
  const int64_t total = rows * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t bits = (uint32_t)hash2(s3, (u64)(row0 * dim + i)) & 0xBFFFFFFFu;
    out[i] = __uint_as_float(bits);
  }
}
// rows addressed through an id list: out[r, c] = feat(idmap(r), c)
__global__ void feature_rows_kernel(const int32_t* __restrict__ ids, int64_t stride, int64_t offset, int64_t id_count,
                                    int64_t rows, int32_t dim, u64 seed, float* __restrict__ out, int64_t rep = 0) {
  const u64 s3 = seed ^ 0xFEA7FEA7FEA7FEA7ull;
  const int64_t total = rows * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int32_t c = (int32_t)(i - r * dim);
    const int64_t k = r < rep ? r : rep + (r - rep) * stride + offset;  // rep: replicated head of a hybrid shard
    const int32_t v = (k < id_count) ? ids[k] : -1;
    uint32_t bits = 0;
    if (v >= 0) bits = (uint32_t)hash2(s3, (u64)((int64_t)v * dim + c)) & 0xBFFFFFFFu;
    out[i] = __uint_as_float(bits);
  }
}
__global__ void labels_kernel(int64_t n, int32_t classes, int32_t* __restrict__ out) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    out[v] = (int32_t)(v % classes);
}
inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256, cap = (int64_t)kSMs * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}
}  // namespace

extern "C" int lg_synth_indptr(void* stream, int64_t n, double dmin, int32_t dmax, uint64_t seed, int64_t* indptr) {
  LG_REQUIRE(indptr && n > 0 && n < (1ll << 31) && dmin > 0 && dmax > 0 && dmax < (1 << 21), "lg_synth_indptr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  degree_kernel<<<grid_for(n), 256, 0, st>>>(n, dmin, dmax, seed, indptr);
  LG_LAUNCH_OK();
  // exclusive scan over n+1 entries (the last input is ignored): indptr[n] = E
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, bytes, indptr, indptr, n + 1, st);
  void* tmp = nullptr;
  LG_CUDA(cudaMallocAsync(&tmp, bytes, st));
  LG_CUDA(cudaMemsetAsync(indptr + n, 0, sizeof(int64_t), st));
  LG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, indptr, indptr, n + 1, st));
  LG_CUDA(cudaFreeAsync(tmp, st));
  return 0;
}
extern "C" int lg_synth_indices(void* stream, int64_t n, const int64_t* indptr, uint64_t seed, int32_t* indices) {
  LG_REQUIRE(indptr && indices && n > 0, "lg_synth_indices: bad argument");
  indices_kernel<<<grid_for(n * 32), 256, 0, (cudaStream_t)stream>>>(n, indptr, seed, indices);
  LG_LAUNCH_OK();
  return 0;
}
extern "C" int lg_synth_features(void* stream, int64_t row0, int64_t rows, int32_t dim, uint64_t seed, float* out) {
  LG_REQUIRE(out && rows >= 0 && dim > 0, "lg_synth_features: bad argument");
  if (rows == 0) return 0;
  features_kernel<<<grid_for(rows * dim), 256, 0, (cudaStream_t)stream>>>(row0, rows, dim, seed, out);
  LG_LAUNCH_OK();
  return 0;
}
extern "C" int lg_synth_labels(void* stream, int64_t n, int32_t classes, int32_t* labels) {
  LG_REQUIRE(labels && n > 0 && classes > 0, "lg_synth_labels: bad argument");
  labels_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, classes, labels);
  LG_LAUNCH_OK();
  return 0;
}
extern "C" int lg_synth_feature_rows(void* stream, const int32_t* ids, int64_t n, int32_t dim, uint64_t seed, float* out) {
  LG_REQUIRE(ids && out && n >= 0 && dim > 0, "lg_synth_feature_rows: bad argument");
  if (n == 0) return 0;
  feature_rows_kernel<<<grid_for(n * dim), 256, 0, (cudaStream_t)stream>>>(ids, 1, 0, n, n, dim, seed, out);
  LG_LAUNCH_OK();
  return 0;
}
extern "C" int lg_synth_feature_shard(void* stream, const int32_t* order, int64_t cap, int32_t kg, int32_t j, int32_t dim,
                                      int64_t num_nodes, uint64_t seed, float* shard) {
  LG_REQUIRE(order && shard && cap > 0 && kg > 0 && j >= 0 && j < kg && dim > 0, "lg_synth_feature_shard: bad argument");
  feature_rows_kernel<<<grid_for(cap * dim), 256, 0, (cudaStream_t)stream>>>(order, kg, j, num_nodes, cap, dim, seed, shard);
  LG_LAUNCH_OK();
  return 0;
}
extern "C" int lg_synth_feature_shard_hybrid(void* stream, const int32_t* order, int64_t cap, int32_t kg, int64_t rep, int32_t j,
                                             int32_t dim, int64_t num_nodes, uint64_t seed, float* shard) {
  LG_REQUIRE(order && shard && cap > 0 && kg > 0 && rep >= 0 && rep <= cap && j >= 0 && j < kg && dim > 0,
             "lg_synth_feature_shard_hybrid: bad argument");
  feature_rows_kernel<<<grid_for(cap * dim), 256, 0, (cudaStream_t)stream>>>(order, kg, j, num_nodes, cap, dim, seed, shard, rep);
  LG_LAUNCH_OK();
  return 0;
}

// ref_ops.cu — the REFERENCE's own device kernels, compiled from the sources where they lie under
// /root/reference (nothing is copied into this repository) and exposed through plain-pointer launchers
// so that tests can run them on the GPU box beside liblegion_b200.so.  Output: oracle/_ref/libref_ops.so
// (git-ignored, travels with gpurun).  TEST INFRASTRUCTURE ONLY.
//
// The three translation units below carry every kernel of the hot path:
//   engine/operator_impl.cu : batch_generate, counter_update, random_sample, pre_sample, construct_graph, ClearPosMap
//   cache/cache.cu (+ cache_impl.cuh) : InitPair, InitIndexPair, InitOffsetPair, FeatFillUp, HotnessMeasure,
//                                       multiGPU_feat_cache_lookup, aggregate_access + bght::bcht maps
//   engine/memorypool.cu : MemoryPool ctor (needed to link the host wrappers of operator_impl.cu)
//   cache/cache.cu also carries the HOST code of UnifiedCache::CandidateSelection / CostModel (:360-551), driven below
//   through the class's own methods; its private state (hotness arrays, capacities) is reached by compiling the
//   reference headers with `private` spelled `public` — test infrastructure, nothing of this ships.
#include <algorithm>
#include <cstdint>
#include <iostream>
#include <vector>
#define private public
#include "engine/memorypool.cu"
#include "cache/cache.cu"
#undef cudaCheckError
#include "engine/operator_impl.cu"
#undef private

#include <cstdint>

#define REF_OK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" {

// op 0: memsets + batch_generate + counter_update(op 0)  — the launch sequence of BatchGenerate(), engine/operator_impl.cu:151-165
int ref_batch_generate(cudaStream_t st, int32_t* batch_ids, int32_t* labels, int32_t batch_size, int32_t counter,
                       int32_t* all_ids, int32_t* all_labels, int32_t total_cap, int32_t* position_map,
                       uint32_t* accessed_map, int32_t total_node_num, int32_t* node_counter, int32_t* edge_counter,
                       int32_t hop_num) {
  cudaMemsetAsync(accessed_map, 0, int64_t(int64_t((total_node_num / 32) + 1) * int64_t(sizeof(uint32_t))), st);
  cudaMemsetAsync(node_counter, 0, 16 * sizeof(int32_t), st);
  cudaMemsetAsync(edge_counter, 0, 16 * sizeof(int32_t), st);
  int32_t size = ((batch_size * (counter + 1)) >= total_cap) ? (total_cap - batch_size * counter) : batch_size;
  dim3 bg_block((size - 1) / OP_THREAD_NUM + 1, 1), bg_thread(OP_THREAD_NUM, 1);
  batch_generate<<<bg_block, bg_thread, 0, st>>>(batch_ids, labels, size, counter, all_ids, all_labels, total_cap,
                                                 position_map, accessed_map);
  counter_update<<<1, 1, 0, st>>>(node_counter, edge_counter, 0, size, hop_num);
  REF_OK();
  return 0;
}

int ref_counter_update(cudaStream_t st, int32_t* node_counter, int32_t* edge_counter, int32_t op_id) {
  counter_update<<<1, 1, 0, st>>>(node_counter, edge_counter, op_id, 0, 0);
  REF_OK();
  return 0;
}

// ops 3h: random_sample + construct_graph + counter_update, launch configuration of RandomSample() (:436-491).
// csr tables: P+1 slots as in GraphStorage; part/offset arrays = what FindTopo would have produced.
int ref_random_sample(cudaStream_t st, int32_t* sampled_ids, int32_t op_id, int64_t** csr_node_index,
                      int32_t** csr_dst_node_ids, char* partition_index, int32_t* partition_offset, int32_t count,
                      int32_t partition_count, int32_t* agg_src_ids, int32_t* agg_dst_ids, int32_t* agg_src_off,
                      int32_t* agg_dst_off, uint32_t* accessed_map, int32_t* position_map, int32_t* node_counter,
                      int32_t* edge_counter) {
  dim3 block_num(16, 1), thread_num(OP_THREAD_NUM, 1);
  random_sample<<<block_num, thread_num, 0, st>>>(sampled_ids, op_id, csr_node_index, csr_dst_node_ids, partition_index,
                                                  partition_offset, count, partition_count, agg_src_ids, agg_dst_ids,
                                                  accessed_map, position_map, node_counter, edge_counter, 0);
  construct_graph<<<block_num, thread_num, 0, st>>>(agg_src_ids, agg_dst_ids, agg_src_off, agg_dst_off, position_map,
                                                    edge_counter, node_counter, op_id, 0);
  counter_update<<<1, 1, 0, st>>>(node_counter, edge_counter, op_id, 0, 0);
  REF_OK();
  return 0;
}

// presampling variant (:474) — host CSR pointers, edge_access_time
int ref_pre_sample(cudaStream_t st, int32_t* sampled_ids, int32_t op_id, int64_t* csr_node_index,
                   int32_t* csr_dst_node_ids, int32_t count, int32_t* agg_src_ids, int32_t* agg_dst_ids,
                   int32_t* agg_src_off, int32_t* agg_dst_off, uint32_t* accessed_map, int32_t* position_map,
                   int32_t* node_counter, int32_t* edge_counter, unsigned long long* edge_access_time) {
  dim3 block_num(16, 1), thread_num(OP_THREAD_NUM, 1);
  pre_sample<<<block_num, thread_num, 0, st>>>(sampled_ids, op_id, csr_node_index, csr_dst_node_ids, nullptr, nullptr, count,
                                               0, agg_src_ids, agg_dst_ids, accessed_map, position_map, node_counter,
                                               edge_counter, 0, edge_access_time);
  construct_graph<<<block_num, thread_num, 0, st>>>(agg_src_ids, agg_dst_ids, agg_src_off, agg_dst_off, position_map,
                                                    edge_counter, node_counter, op_id, 0);
  counter_update<<<1, 1, 0, st>>>(node_counter, edge_counter, op_id, 0, 0);
  REF_OK();
  return 0;
}

int ref_hotness_measure(cudaStream_t st, int32_t* ids, int32_t* node_counter, unsigned long long* access_map) {
  HotnessMeasure<<<dim3(32, 1), dim3(1024, 1), 0, st>>>(ids, node_counter, access_map);
  REF_OK();
  return 0;
}

// placement pairs exactly as PreSCCacheController::Insert builds them (cache/cache.cu:94-124); pairs are {int32 key, value}
int ref_init_pair(cudaStream_t st, int32_t* pair_kv /*[2*cap*expand]*/, int32_t* QF, int32_t cap, int32_t expand, int32_t Kg) {
  InitPair<<<dim3(80, 1), dim3(1024, 1), 0, st>>>((pair_type*)pair_kv, QF, cap, expand, Kg);
  REF_OK();
  return 0;
}
int ref_pair_sizes(int32_t* out3) {
  out3[0] = (int32_t)sizeof(pair_type);
  out3[1] = (int32_t)sizeof(index_pair_type);
  out3[2] = (int32_t)sizeof(offset_pair_type);
  return 0;
}
int ref_init_topo_pairs(cudaStream_t st, void* index_pair, void* offset_pair, int32_t* QT, int32_t cap, int32_t expand,
                        int32_t Kg, int32_t Ki) {
  InitIndexPair<<<dim3(80, 1), dim3(1024, 1), 0, st>>>((index_pair_type*)index_pair, QT, cap, expand, Kg, Ki);
  InitOffsetPair<<<dim3(80, 1), dim3(1024, 1), 0, st>>>((offset_pair_type*)offset_pair, QT, cap, expand, Kg);
  REF_OK();
  return 0;
}
int ref_feat_fill_up(cudaStream_t st, int32_t cap, int32_t dim, float* feature_cache, float* cpu_float_feature, int32_t* QF,
                     int32_t Kg, int32_t Ki) {
  FeatFillUp<<<128, 1024, 0, st>>>(cap, dim, feature_cache, cpu_float_feature, QF, Kg, Ki);
  REF_OK();
  return 0;
}

// the gather, launch configuration of FeatCacheLookup (cache/cache.cu:729-741)
int ref_feat_cache_lookup(cudaStream_t st, float* cpu_float_features, float** gpu_float_feature, int32_t dim,
                          int32_t* sampled_ids, int32_t* cache_index, int32_t cache_capacity, int32_t* node_counter,
                          float* dst, int32_t total_num_nodes, int32_t op_id) {
  multiGPU_feat_cache_lookup<<<dim3(32, 1), dim3(1024, 1), 0, st>>>(cpu_float_features, gpu_float_feature, dim, sampled_ids,
                                                                    cache_index, cache_capacity, node_counter, dst,
                                                                    total_num_nodes, 0, op_id);
  REF_OK();
  return 0;
}

// the reference's lookup structure: a bght::bcht<int32,int32> built from pairs, queried for keys (find => value or -2)
int ref_bcht_build_and_find(cudaStream_t st, int32_t* pair_kv, int32_t n_pairs, int64_t capacity, int32_t* keys,
                            int32_t n_keys, int32_t* out_values) {
  auto* map = new bght::bcht<int32_t, int32_t>(capacity, CACHEMISS_FLAG, CACHEMISS_FLAG);
  pair_type* p = (pair_type*)pair_kv;
  bool ok = map->insert(p, p + n_pairs, st);
  map->find(keys, keys + n_keys, out_values, st);
  cudaStreamSynchronize(st);
  delete map;
  REF_OK();
  return ok ? 0 : -1;
}

// ---- UnifiedCache::CandidateSelection + CostModel (cache/cache.cu:360-551) on one device ----
namespace {
struct StubFeature : public FeatureStorage {
  int32_t n, dim;
  void Build(BuildInfo*, int) override {}
  void Finalize() override {}
  int32_t* GetTrainingSetIds(int32_t) const override { return nullptr; }
  int32_t* GetValidationSetIds(int32_t) const override { return nullptr; }
  int32_t* GetTestingSetIds(int32_t) const override { return nullptr; }
  int32_t* GetTrainingLabels(int32_t) const override { return nullptr; }
  int32_t* GetValidationLabels(int32_t) const override { return nullptr; }
  int32_t* GetTestingLabels(int32_t) const override { return nullptr; }
  int32_t TrainingSetSize(int32_t) const override { return 0; }
  int32_t ValidationSetSize(int32_t) const override { return 0; }
  int32_t TestingSetSize(int32_t) const override { return 0; }
  int32_t TotalNodeNum() const override { return n; }
  float* GetAllFloatFeature() const override { return nullptr; }
  int32_t GetFloatFeatureLen() const override { return dim; }
  void IOSubmit(int32_t*, int32_t*, int32_t*, float*, int32_t, int32_t, cudaStream_t) override {}
  void IOComplete() override {}
};
struct StubGraph : public GraphStorage {
  int64_t* indptr;  // device-readable (GetEdgeMem dereferences it in a kernel, cache/cache_impl.cuh:63-69)
  void Build(BuildInfo*) override {}
  void GraphCache(int32_t*, int32_t, int32_t, int32_t) override {}
  void Finalize() override {}
  int32_t GetPartitionCount() const override { return 1; }
  int64_t** GetCSRNodeIndex(int32_t) const override { return nullptr; }
  int32_t** GetCSRNodeMatrix(int32_t) const override { return nullptr; }
  int64_t* GetCSRNodeIndexCPU() const override { return indptr; }
  int32_t* GetCSRNodeMatrixCPU() const override { return nullptr; }
  int64_t Src_Size(int32_t) const override { return 0; }
  int64_t Dst_Size(int32_t) const override { return 0; }
  char* PartitionIndex(int32_t) const override { return nullptr; }
  int32_t* PartitionOffset(int32_t) const override { return nullptr; }
};
}  // namespace

// Kg controllers, all on the current device; node_hot/edge_hot: Kg host arrays of N counters each; max_ids: Kg values.
// Outputs: per-GPU capacities exactly as UnifiedCache stores them (node_capacity_[0], edge_capacity_[0]); optionally the
// reference's own ranking (QF/QT, int32[N]) and sorted aggregates (AF/AT, u64[N]) copied to the host.
int ref_cost_model(int32_t N, int32_t dim, int64_t cache_memory, int32_t Kg, int32_t train_step,
                   const unsigned long long* const* node_hot, const unsigned long long* const* edge_hot,
                   const int32_t* max_ids, const int64_t* indptr_host, uint64_t topo_counter0, uint64_t topo_counter1,
                   int32_t* node_capacity, int32_t* edge_capacity, int32_t* QF, int32_t* QT, unsigned long long* AF,
                   unsigned long long* AT) {
  int mode = 0;
  while ((1 << mode) < Kg) mode++;
  if ((1 << mode) != Kg || mode > 3) return -2;
  int dev = 0;
  cudaGetDevice(&dev);
  UnifiedCache uc;
  uc.Initialize(cache_memory, dim, train_step, Kg, 0, 0);
  const size_t nb = (size_t)N * sizeof(unsigned long long);
  for (int j = 0; j < Kg; j++) {
    uc.cache_controller_[j]->Initialize(dev, N);  // every controller on this device (the box may have one GPU)
    cudaMemcpy(uc.cache_controller_[j]->GetNodeAccessedMap(), node_hot[j], nb, cudaMemcpyHostToDevice);
    cudaMemcpy(uc.cache_controller_[j]->GetEdgeAccessedMap(), edge_hot[j], nb, cudaMemcpyHostToDevice);
    static_cast<PreSCCacheController*>(uc.cache_controller_[j])->max_ids_ = max_ids[j];
  }
  StubFeature feat;
  feat.n = N;
  feat.dim = dim;
  StubGraph graph;
  cudaMallocManaged(&graph.indptr, (size_t)(N + 1) * sizeof(int64_t));
  memcpy(graph.indptr, indptr_host, (size_t)(N + 1) * sizeof(int64_t));
  uc.CandidateSelection(mode, &feat, &graph);
  if (uc.Kg_ != Kg || uc.Kc_ != 1) return -3;
  std::vector<uint64_t> counters = {topo_counter0, topo_counter1};
  uc.CostModel(mode, &feat, &graph, counters, train_step);
  cudaDeviceSynchronize();
  *node_capacity = uc.node_capacity_[0];
  *edge_capacity = uc.edge_capacity_[0];
  if (QF) cudaMemcpy(QF, uc.QF_[0], (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (QT) cudaMemcpy(QT, uc.QT_[0], (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (AF) cudaMemcpy(AF, uc.AF_[0], nb, cudaMemcpyDeviceToHost);
  if (AT) cudaMemcpy(AT, uc.AT_[0], nb, cudaMemcpyDeviceToHost);
  cudaFree(graph.indptr);
  REF_OK();
  return 0;
}

}  // extern "C"

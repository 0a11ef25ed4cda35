// cache.h — UnifiedCache host orchestration (reference: cache/cache.cuh:66-177, cache/cache.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "storage.h"

class UnifiedCache {
 public:
  void Initialize(int64_t cache_memory, int32_t float_feature_len, int32_t train_step, int32_t device_count);
  void InitializeCacheController(int32_t dev_id, int32_t total_num_nodes);  // hotness arrays on that GPU
  unsigned long long* GetNodeAccessedMap(int32_t dev) { return node_access_[dev]; }
  unsigned long long* GetEdgeAccessedMap(int32_t dev) { return edge_access_[dev]; }
  int32_t* MaxIdsDevice(int32_t dev) { return max_ids_dev_[dev]; }
  int32_t MaxIdNum(int32_t dev);
  void CandidateSelection(int cache_agg_mode, FeatureStorage* feature, GraphStorage* graph);
  void CostModel(int cache_agg_mode, FeatureStorage* feature, GraphStorage* graph, std::vector<uint64_t>& counters,
                 int32_t train_step);
  void FillUp(int cache_agg_mode, FeatureStorage* feature, GraphStorage* graph);
  lg_feature_cache* FeatureCache(int32_t dev) { return &fcache_[dev]; }
  int32_t LocalPart(int32_t dev) const { return dev % Kg_; }
  bool IsPresc() const { return is_presc_; }
  int32_t NodeCapacity(int clique) const { return node_capacity_[clique]; }
  int32_t EdgeCapacity(int clique) const { return edge_capacity_[clique]; }
  unsigned long long* TierRows(int32_t dev) { return tier_rows_[dev]; }
  void ReadTierRows(int32_t dev, unsigned long long out[3]);  // synchronous copy of the device counters

 private:
  int64_t cache_memory_ = 0;
  int32_t float_feature_len_ = 0, train_step_ = 0, device_count_ = 0, total_num_nodes_ = 0;
  int32_t Kc_ = 1, Kg_ = 1;
  bool is_presc_ = true;
  std::vector<unsigned long long*> node_access_, edge_access_, tier_rows_;
  std::vector<int32_t*> max_ids_dev_;
  std::vector<int32_t*> QF_, QT_;                      // per clique, on the clique's first GPU
  std::vector<unsigned long long*> AF_, AT_;           // sorted aggregated hotness per clique
  std::vector<int32_t> node_capacity_, edge_capacity_; // per clique, rows per GPU
  std::vector<lg_feature_cache> fcache_;
};

// storage.h — StorageManagement / GraphStorage / FeatureStorage of the sampling server
// (reference: storage/storage_management.cuh:10, graph_storage.cuh:7-24, feature_storage.cuh:6-34).
#pragma once
#include <string>
#include <vector>

#include "buildinfo.h"
#include "ipc_service.h"

// per-GPU device copies of the seed sets (FeatureStorage::Build, storage/feature_storage.cu:18-90)
class FeatureStorage {
 public:
  void Build(BuildInfo* info);
  void Finalize();
  int32_t* GetTrainingSetIds(int d) const { return train_ids_[d]; }
  int32_t* GetTrainingLabels(int d) const { return train_labels_[d]; }
  int32_t* GetValidationSetIds(int d) const { return valid_ids_[d]; }
  int32_t* GetValidationLabels(int d) const { return valid_labels_[d]; }
  int32_t* GetTestingSetIds(int d) const { return test_ids_[d]; }
  int32_t* GetTestingLabels(int d) const { return test_labels_[d]; }
  int32_t TrainingSetSize(int d) const { return train_num_[d]; }
  int32_t ValidationSetSize(int d) const { return valid_num_[d]; }
  int32_t TestingSetSize(int d) const { return test_num_[d]; }
  int32_t TotalNodeNum() const { return total_num_nodes_; }
  int32_t GetFloatFeatureLen() const { return float_feature_len_; }
  float* GetAllFloatFeature() const { return float_feature_dev_; }  // UVA alias of the pinned matrix

 private:
  std::vector<int32_t*> train_ids_, train_labels_, valid_ids_, valid_labels_, test_ids_, test_labels_;
  std::vector<int32_t> train_num_, valid_num_, test_num_;
  int32_t total_num_nodes_ = 0, float_feature_len_ = 0, partition_count_ = 0;
  float* float_feature_dev_ = nullptr;
  float* float_feature_host_ = nullptr;
};

// per-GPU topology descriptor: slot P = full host CSR through UVA (storage/graph_storage.cu:60-62)
class GraphStorage {
 public:
  void Build(BuildInfo* info);
  void Finalize();
  int32_t GetPartitionCount() const { return partition_count_; }
  lg_topology* Topology(int dev) { return &topo_[dev]; }
  const int64_t* GetCSRNodeIndexCPU() const { return indptr_dev_; }
  const int32_t* GetCSRNodeMatrixCPU() const { return indices_dev_; }
  const int64_t* HostIndptr() const { return indptr_host_; }

 private:
  int32_t partition_count_ = 0;
  std::vector<lg_topology> topo_;
  int64_t *indptr_host_ = nullptr, *indptr_dev_ = nullptr;
  int32_t *indices_host_ = nullptr, *indices_dev_ = nullptr;
};

class UnifiedCache;

class StorageManagement {
 public:
  void Initialze(int32_t partition_count, int32_t in_memory_mode, const std::vector<int>& fanout);
  GraphStorage* GetGraph() { return graph_; }
  FeatureStorage* GetFeature() { return feature_; }
  UnifiedCache* GetCache() { return cache_; }
  IPCEnv* GetIPCEnv() { return env_; }
  BuildInfo* GetInfo() { return info_; }

 private:
  void ReadMetaFIle(BuildInfo* info);
  void LoadGraph(BuildInfo* info);
  void LoadFeature(BuildInfo* info);
  std::string dataset_path_;
  int32_t raw_batch_size_ = 0, float_feature_len_ = 0, epoch_ = 0;
  int64_t node_num_ = 0, edge_num_ = 0, cache_memory_ = 0;
  int64_t training_set_num_ = 0, validation_set_num_ = 0, testing_set_num_ = 0;
  GraphStorage* graph_ = nullptr;
  FeatureStorage* feature_ = nullptr;
  UnifiedCache* cache_ = nullptr;
  IPCEnv* env_ = nullptr;
  BuildInfo* info_ = nullptr;
};

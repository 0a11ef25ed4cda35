// buildinfo.h — what StorageManagement hands to the storages, the IPC env and the cache
// (reference: sampling_server/src/include/buildinfo.h:6-72, without the unused NVMe/BaM fields).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct BuildInfo {
  int32_t partition_count = 1;
  // host-pinned (UVA) full CSR + features
  int64_t* csr_node_index = nullptr;    // int64[N+1]  (edge_src)
  int32_t* csr_dst_node_ids = nullptr;  // int32[E]    (edge_dst)
  int64_t* csr_node_index_dev = nullptr;  // device alias of the mapped host memory
  int32_t* csr_dst_node_ids_dev = nullptr;
  float* host_float_feature = nullptr;
  float* host_float_feature_dev = nullptr;
  int32_t float_feature_len = 0;
  int32_t total_num_nodes = 0;
  int64_t total_edge_num = 0;
  int64_t cache_memory = 0;
  int32_t raw_batch_size = 0;
  int32_t epoch = 0;
  std::vector<int> fanout;
  std::vector<std::vector<int32_t>> training_set_ids, training_labels;
  std::vector<std::vector<int32_t>> validation_set_ids, validation_labels;
  std::vector<std::vector<int32_t>> testing_set_ids, testing_labels;
  std::vector<int32_t> training_set_num, validation_set_num, testing_set_num;
};

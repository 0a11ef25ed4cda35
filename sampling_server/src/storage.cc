// storage.cc — meta_config, dataset files, seed split, per-GPU storages.
// Contract: reference storage/storage_management.cu:29-269 (file names, dtypes, split rules, echo
// lines).  Changes: files are copied into pinned memory by parallel memcpy instead of an
// element-wise loop (storage_management_impl.cuh:85-121), and `features` / `labels` ARE loaded
// (the reference leaves its loads commented out, :162,:164 — features would be uninitialised).
#include "storage.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "cache.h"

namespace {
// returns bytes copied, -1 if the file does not exist
int64_t LoadFile(const std::string& path, void* dst, int64_t max_bytes) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return -1;
  struct stat st;
  fstat(fd, &st);
  int64_t bytes = st.st_size < max_bytes ? (int64_t)st.st_size : max_bytes;
  if (bytes > 0) {
    const char* src = (const char*)mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (src == MAP_FAILED) {
      std::cout << "cannot mmap file: " << path << std::endl;
      std::exit(EXIT_FAILURE);
    }
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (bytes < (1 << 22)) nt = 1;
    std::vector<std::thread> th;
    int64_t chunk = (bytes + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
      int64_t lo = t * chunk, hi = lo + chunk < bytes ? lo + chunk : bytes;
      if (lo >= hi) break;
      th.emplace_back([=] { std::memcpy((char*)dst + lo, src + lo, hi - lo); });
    }
    for (auto& x : th) x.join();
    munmap((void*)src, bytes);
  }
  close(fd);
  return bytes;
}
void LoadIds(const std::string& path, std::vector<int32_t>& v) {
  if (v.empty()) return;
  if (LoadFile(path, v.data(), (int64_t)v.size() * 4) < 0) {
    std::cout << "cannout open file: " << path << std::endl;
    std::exit(EXIT_FAILURE);
  }
}
int32_t* ToDevice(const std::vector<int32_t>& v) {
  void* p = nullptr;
  LGCHECK(lg_device_alloc(&p, (int64_t)v.size() * 4));
  if (!v.empty()) LGCHECK(lg_memcpy_h2d(p, v.data(), (int64_t)v.size() * 4, nullptr));
  LGCHECK(lg_stream_synchronize(nullptr));
  return (int32_t*)p;
}
}  // namespace

void StorageManagement::ReadMetaFIle(BuildInfo* info) {
  std::ifstream meta("./meta_config");  // cwd-relative, like the reference (:32)
  if (!meta.is_open()) {
    std::cout << "unable to open meta config file\n";
    std::exit(EXIT_FAILURE);
  }
  std::string line;
  getline(meta, line);
  std::istringstream iss(line);
  iss >> dataset_path_ >> raw_batch_size_ >> node_num_ >> edge_num_ >> float_feature_len_ >> training_set_num_ >>
      validation_set_num_ >> testing_set_num_ >> cache_memory_ >> epoch_;
  std::cout << "Dataset path:       " << dataset_path_ << "\n";
  std::cout << "Raw Batchsize:      " << raw_batch_size_ << "\n";
  std::cout << "Graph nodes num:    " << node_num_ << "\n";
  std::cout << "Graph edges num:    " << edge_num_ << "\n";
  std::cout << "Feature dim:        " << float_feature_len_ << "\n";
  std::cout << "Training set num:   " << training_set_num_ << "\n";
  std::cout << "Validation set num: " << validation_set_num_ << "\n";
  std::cout << "Testing set num:    " << testing_set_num_ << "\n";
  std::cout << "Cache memory:       " << cache_memory_ << "\n";
  std::cout << "Train epoch:        " << epoch_ << "\n";
  // backward-compatible extension: fields 11.. = fan-out per hop (the reference parses --fanout in
  // legion_server.py:120 but never forwards it; sampling_server/src/main.cu:9-11 hard-codes 25,10)
  int f;
  std::vector<int> fo;
  while (iss >> f) fo.push_back(f);
  if (!fo.empty()) info->fanout = fo;
  info->raw_batch_size = raw_batch_size_;
  info->epoch = epoch_;
  info->cache_memory = cache_memory_;
}

void StorageManagement::LoadGraph(BuildInfo* info) {
  info->total_edge_num = edge_num_;
  void *h = nullptr, *d = nullptr;
  LGCHECK(lg_host_alloc_mapped(&h, &d, (node_num_ + 1) * 8));
  info->csr_node_index = (int64_t*)h;
  info->csr_node_index_dev = (int64_t*)d;
  LGCHECK(lg_host_alloc_mapped(&h, &d, edge_num_ * 4));
  info->csr_dst_node_ids = (int32_t*)h;
  info->csr_dst_node_ids_dev = (int32_t*)d;
  if (LoadFile(dataset_path_ + "edge_src", info->csr_node_index, (node_num_ + 1) * 8) < 0 ||
      LoadFile(dataset_path_ + "edge_dst", info->csr_dst_node_ids, edge_num_ * 4) < 0) {
    std::cout << "cannout open file: " << dataset_path_ << "edge_src / edge_dst" << std::endl;
    std::exit(EXIT_FAILURE);
  }
}

void StorageManagement::LoadFeature(BuildInfo* info) {
  int32_t P = info->partition_count;
  int32_t nf = float_feature_len_;
  std::vector<int32_t> training_ids(training_set_num_), validation_ids(validation_set_num_), testing_ids(testing_set_num_);
  std::vector<int32_t> all_labels(node_num_, 0), partition(node_num_);
  LoadIds(dataset_path_ + "trainingset", training_ids);
  LoadIds(dataset_path_ + "validationset", validation_ids);
  LoadIds(dataset_path_ + "testingset", testing_ids);
  void *h = nullptr, *d = nullptr;
  LGCHECK(lg_host_alloc_mapped(&h, &d, node_num_ * nf * 4));
  if (LoadFile(dataset_path_ + "features", h, node_num_ * nf * 4) < 0) {
    std::cout << "features file absent: feature matrix zero-filled\n";
    std::memset(h, 0, (size_t)(node_num_ * nf * 4));
  }
  if (LoadFile(dataset_path_ + "labels", all_labels.data(), node_num_ * 4) < 0)
    std::cout << "labels file absent: labels zero-filled\n";
  bool has_partition = LoadFile(dataset_path_ + "partition", partition.data(), node_num_ * 4) >= 0;
  std::cout << "Finish Reading All Files\n";

  info->training_set_ids.resize(P);
  info->validation_set_ids.resize(P);
  info->testing_set_ids.resize(P);
  info->training_labels.resize(P);
  info->validation_labels.resize(P);
  info->testing_labels.resize(P);
  int count = 0;
  for (int32_t tid : training_ids) {  // :171-184
    int32_t part = has_partition ? partition[tid] : tid % P;
    if (part < P && part >= 0) {
      info->training_set_ids[part].push_back(tid);
      count++;
    }
  }
  std::cout << "training set count " << count << "\n";
  for (int32_t tid : validation_ids) info->validation_set_ids[tid % P].push_back(tid);  // :187-194
  for (int32_t tid : testing_ids) info->testing_set_ids[tid % P].push_back(tid);        // :196-203
  for (int32_t p = 0; p < P; p++) {
    for (int32_t id : info->training_set_ids[p]) info->training_labels[p].push_back(all_labels[id]);
    for (int32_t id : info->validation_set_ids[p]) info->validation_labels[p].push_back(all_labels[id]);
    for (int32_t id : info->testing_set_ids[p]) info->testing_labels[p].push_back(all_labels[id]);
    info->training_set_num.push_back((int32_t)info->training_set_ids[p].size());
    info->validation_set_num.push_back((int32_t)info->validation_set_ids[p].size());
    info->testing_set_num.push_back((int32_t)info->testing_set_ids[p].size());
  }
  info->host_float_feature = (float*)h;
  info->host_float_feature_dev = (float*)d;
  info->float_feature_len = nf;
  info->total_num_nodes = (int32_t)node_num_;
}

void StorageManagement::Initialze(int32_t partition_count, int32_t, const std::vector<int>& fanout) {
  info_ = new BuildInfo();
  info_->fanout = fanout;
  LGCHECK(lg_enable_peer_access(partition_count));  // EnableP2PAccess :5-23
  info_->partition_count = partition_count;
  ReadMetaFIle(info_);
  LoadGraph(info_);
  LoadFeature(info_);
  env_ = NewIPCEnv(partition_count);
  env_->Coordinate(info_);
  feature_ = new FeatureStorage();
  feature_->Build(info_);
  graph_ = new GraphStorage();
  graph_->Build(info_);
  cache_ = new UnifiedCache();
  LGCHECK(lg_set_device(0));
  cache_->Initialize(cache_memory_, float_feature_len_, env_->GetTrainStep(), partition_count);
  std::cout << "Storage Initialized\n";
}

void FeatureStorage::Build(BuildInfo* info) {
  partition_count_ = info->partition_count;
  total_num_nodes_ = info->total_num_nodes;
  float_feature_len_ = info->float_feature_len;
  float_feature_dev_ = info->host_float_feature_dev;
  float_feature_host_ = info->host_float_feature;
  for (int32_t p = 0; p < partition_count_; p++) {
    LGCHECK(lg_set_device(p));
    train_num_.push_back(info->training_set_num[p]);
    valid_num_.push_back(info->validation_set_num[p]);
    test_num_.push_back(info->testing_set_num[p]);
    train_ids_.push_back(ToDevice(info->training_set_ids[p]));
    valid_ids_.push_back(ToDevice(info->validation_set_ids[p]));
    test_ids_.push_back(ToDevice(info->testing_set_ids[p]));
    train_labels_.push_back(ToDevice(info->training_labels[p]));
    valid_labels_.push_back(ToDevice(info->validation_labels[p]));
    test_labels_.push_back(ToDevice(info->testing_labels[p]));
  }
}

void FeatureStorage::Finalize() {
  lg_host_free(float_feature_host_);
  for (int32_t p = 0; p < partition_count_; p++) {
    lg_set_device(p);
    for (int32_t* q : {train_ids_[p], valid_ids_[p], test_ids_[p], train_labels_[p], valid_labels_[p], test_labels_[p]})
      lg_device_free(q);
  }
}

void GraphStorage::Build(BuildInfo* info) {
  partition_count_ = info->partition_count;
  indptr_host_ = info->csr_node_index;
  indices_host_ = info->csr_dst_node_ids;
  indptr_dev_ = info->csr_node_index_dev;
  indices_dev_ = info->csr_dst_node_ids_dev;
  topo_.resize(partition_count_);
  for (auto& t : topo_) {
    std::memset(&t, 0, sizeof(t));
    t.n_parts = 0;  // until GraphCache fills the shards, every row is read from the host CSR
    t.num_nodes = info->total_num_nodes;
    t.indptr[0] = indptr_dev_;
    t.indices[0] = indices_dev_;
  }
}

void GraphStorage::Finalize() {
  lg_host_free(indptr_host_);
  lg_host_free(indices_host_);
}

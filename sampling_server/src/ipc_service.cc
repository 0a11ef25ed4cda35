// ipc_service.cc — shm segment, CUDA-IPC buffers, named semaphores, step schedule.
// Contract: SURVEY 8b / reference engine/ipc_service.cu:33-351.  Differences, all on the safe side:
// stale semaphores of a crashed run are unlinked before sem_open (the reference inherits their
// counts, detail_parameter_settings/README.md:48) and buffers come from the C ABI (plain cudaMalloc,
// so cudaIpcGetMemHandle works with the trainer's legacy cudaIpcOpenMemHandle).
#include "ipc_service.h"

#include "../../include/legion_b200_ext.h"

#include <fcntl.h>
#include <semaphore.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace {
const char kShmName[] = "simpleIPCshm";

class CUDAIPCEnv : public IPCEnv {
 public:
  explicit CUDAIPCEnv(int32_t device_count) : device_count_(device_count) {
    // sharedMemoryCreate semantics (engine/helper_multiprocess.cu:48-68): O_CREAT, 0777, ftruncate, mmap
    fd_ = shm_open(kShmName, O_RDWR | O_CREAT, 0777);
    if (fd_ < 0 || ftruncate(fd_, sizeof(shmStruct)) != 0) {
      std::printf("Failed to create shared memory slab\n");
      std::exit(EXIT_FAILURE);
    }
    void* addr = mmap(0, sizeof(shmStruct), PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
    if (addr == MAP_FAILED) {
      std::printf("Failed to create shared memory slab\n");
      std::exit(EXIT_FAILURE);
    }
    shm_ = (volatile shmStruct*)addr;
    std::memset((void*)shm_, 0, sizeof(shmStruct));
    ids_.resize(device_count);
    float_features_.resize(device_count);
    labels_.resize(device_count);
    agg_src_.resize(device_count);
    agg_dst_.resize(device_count);
    node_counter_.resize(device_count);
    edge_counter_.resize(device_count);
    feature_rows_.assign(device_count, 0);
    semr_.resize(device_count);
    semw_.resize(device_count);
    OpenExt();
  }

  // the optional side channel (include/legion_b200_ext.h); LEGION_EXT_SHM=0 = the reference wire and nothing else
  void OpenExt() {
    const char* e = std::getenv("LEGION_EXT_SHM");
    if (e && std::atoi(e) == 0) return;
    shm_unlink(LG_EXT_SHM_NAME);  // never inherit a crashed run's sequence numbers
    ext_fd_ = shm_open(LG_EXT_SHM_NAME, O_RDWR | O_CREAT, 0777);
    if (ext_fd_ < 0 || ftruncate(ext_fd_, sizeof(lg_ext_shm)) != 0) return;
    void* addr = mmap(0, sizeof(lg_ext_shm), PROT_READ | PROT_WRITE, MAP_SHARED, ext_fd_, 0);
    if (addr == MAP_FAILED) return;
    std::memset(addr, 0, sizeof(lg_ext_shm));
    if (lg_host_register(addr, sizeof(lg_ext_shm), nullptr) != 0) {  // pageable would make the copies synchronous: do without
      munmap(addr, sizeof(lg_ext_shm));
      return;
    }
    ext_ = (lg_ext_shm*)addr;
    ext_->n_gpus = device_count_;
    ext_->version = LG_EXT_VERSION;
    ext_->magic = LG_EXT_MAGIC;
  }

  void Coordinate(BuildInfo* info) override {
    int32_t P = info->partition_count;
    epoch_ = info->epoch;
    raw_batch_size_ = info->raw_batch_size;
    int32_t min_train = 1000000000, max_valid = 0, max_test = 0;
    for (int32_t i = 0; i < P; i++) {
      if (info->training_set_num[i] < min_train) min_train = info->training_set_num[i];
      if (info->validation_set_num[i] > max_valid) max_valid = info->validation_set_num[i];
      if (info->testing_set_num[i] > max_test) max_test = info->testing_set_num[i];
    }
    train_step_ = (min_train - 1) / raw_batch_size_;
    valid_step_ = (max_valid - 1) / 512 + 1;
    test_step_ = (max_test - 1) / 512 + 1;
    for (int32_t i = 0; i < P; i++) {
      train_batch_size_.push_back(raw_batch_size_);
      valid_batch_size_.push_back((info->validation_set_num[i] - 1) / valid_step_ + 1);
      test_batch_size_.push_back((info->testing_set_num[i] - 1) / test_step_ + 1);
    }
    std::cout << "Train Steps: " << train_step_ << "\n";
    std::cout << "Valid Steps: " << valid_step_ << "\n";
    std::cout << "Test Steps: " << test_step_ << "\n";
    shm_->steps[0] = train_step_;
    shm_->steps[1] = valid_step_;
    shm_->steps[2] = test_step_;
  }

  int32_t GetMaxStep() override { return (train_step_ + valid_step_) * epoch_ + test_step_; }

  void* Alloc(int64_t bytes, int32_t dev, int32_t pipe, int32_t slot) {
    void* p = nullptr;
    LGCHECK(lg_device_alloc(&p, bytes));
    LGCHECK(lg_memset_async(p, 0, bytes, nullptr));
    LGCHECK(lg_ipc_export(p, (unsigned char*)shm_->memHandle[dev][pipe][slot].b));
    return p;
  }

  void InitializeSamplesBuffer(int32_t batch_size, int32_t num_ids, int32_t, int32_t dev, int32_t depth) override {
    LGCHECK(lg_set_device(dev));
    semr_[dev].resize(depth);
    semw_[dev].resize(depth);
    for (int32_t i = 0; i < depth; i++) {
      ids_[dev].push_back(Alloc((int64_t)num_ids * 4, dev, i, 0));
      labels_[dev].push_back(Alloc((int64_t)batch_size * 4, dev, i, 2));
      agg_src_[dev].push_back(Alloc((int64_t)num_ids * 4, dev, i, 3));
      agg_dst_[dev].push_back(Alloc((int64_t)num_ids * 4, dev, i, 4));
      node_counter_[dev].push_back(Alloc(16 * 4, dev, i, 5));
      edge_counter_[dev].push_back(Alloc(16 * 4, dev, i, 6));
      std::string r = "sem_r_" + std::to_string(dev) + "_" + std::to_string(i);
      std::string w = "sem_w_" + std::to_string(dev) + "_" + std::to_string(i);
      sem_unlink(r.c_str());  // drop counts left behind by a crashed run
      sem_unlink(w.c_str());
      semr_[dev][i] = sem_open(r.c_str(), O_CREAT | O_RDWR, 0666, 0);
      semw_[dev][i] = sem_open(w.c_str(), O_CREAT | O_RDWR, 0666, 0);
      if (semr_[dev][i] == SEM_FAILED || semw_[dev][i] == SEM_FAILED) {
        std::printf("errno = %d\n", errno);
        std::exit(EXIT_FAILURE);
      }
    }
    pipeline_depth_ = depth;
  }

  void InitializeFeaturesBuffer(int32_t, int32_t num_ids, int32_t feature_dim, int32_t dev, int32_t depth) override {
    LGCHECK(lg_set_device(dev));
    feature_rows_[dev] = num_ids;
    for (int32_t i = 0; i < depth; i++)
      float_features_[dev].push_back(Alloc((int64_t)num_ids * feature_dim * 4, dev, i, 1));
    LGCHECK(lg_stream_synchronize(nullptr));
  }

  int32_t GetRawBatchsize() override { return raw_batch_size_; }

  int32_t GetLocalBatchId(int32_t g) override {
    int32_t tv = train_step_ + valid_step_;
    if (g < tv * epoch_) {
      int32_t e = g % tv;
      return e < train_step_ ? e : e - train_step_;
    }
    return (g - tv * epoch_) % test_step_;
  }
  int32_t GetCurrentMode(int32_t g) override {
    int32_t tv = train_step_ + valid_step_;
    if (g < tv * epoch_) return (g % tv) < train_step_ ? TRAINMODE : VALIDMODE;
    return TESTMODE;
  }
  int32_t GetCurrentBatchsize(int32_t dev, int32_t mode) override {
    if (mode == TRAINMODE) return train_batch_size_[dev];
    if (mode == VALIDMODE) return valid_batch_size_[dev];
    return test_batch_size_[dev];
  }

  int32_t* GetIds(int32_t d, int32_t p) override { return (int32_t*)ids_[d][p % pipeline_depth_]; }
  float* GetFloatFeatures(int32_t d, int32_t p) override { return (float*)float_features_[d][p % pipeline_depth_]; }
  int32_t* GetLabels(int32_t d, int32_t p) override { return (int32_t*)labels_[d][p % pipeline_depth_]; }
  int32_t* GetAggSrc(int32_t d, int32_t p) override { return (int32_t*)agg_src_[d][p % pipeline_depth_]; }
  int32_t* GetAggDst(int32_t d, int32_t p) override { return (int32_t*)agg_dst_[d][p % pipeline_depth_]; }
  int32_t* GetNodeCounter(int32_t d, int32_t p) override { return (int32_t*)node_counter_[d][p % pipeline_depth_]; }
  int32_t* GetEdgeCounter(int32_t d, int32_t p) override { return (int32_t*)edge_counter_[d][p % pipeline_depth_]; }
  int64_t GetFeatureRows(int32_t d) override { return feature_rows_[d]; }

  int32_t* GetHostCounters(int32_t d, int32_t p) override {
    return (ext_ && d < LG_EXT_MAX_DEVICE && p < LG_EXT_SLOTS) ? (int32_t*)ext_->counters[d][p] : nullptr;
  }

  bool InitializeCscBuffers(int32_t dev, int32_t depth, int32_t hops, const int64_t* max_edges, const int64_t* max_dst) override {
    const char* e = std::getenv("LEGION_EMIT_CSC");
    if (!ext_ || !e || std::atoi(e) == 0 || hops > LG_EXT_MAX_HOPS || dev >= LG_EXT_MAX_DEVICE || depth > LG_EXT_SLOTS) return false;
    LGCHECK(lg_set_device(dev));
    if ((int32_t)csc_.size() < device_count_) csc_.resize(device_count_);
    csc_[dev].assign((size_t)depth * hops * 3, nullptr);
    for (int32_t p = 0; p < depth; p++)
      for (int32_t h = 0; h < hops; h++)
        for (int32_t k = 0; k < 3; k++) {
          const int64_t words = k == 0 ? max_dst[h] + 1 : max_edges[h];
          void* ptr = nullptr;
          LGCHECK(lg_device_alloc(&ptr, words * 4));
          LGCHECK(lg_memset_async(ptr, 0, words * 4, nullptr));
          LGCHECK(lg_ipc_export(ptr, ext_->csc_handle[dev][p][h][k]));
          csc_[dev][((size_t)p * hops + h) * 3 + k] = ptr;
        }
    LGCHECK(lg_stream_synchronize(nullptr));
    csc_hops_ = hops;
    ext_->csc_hops = hops;
    return true;
  }
  int32_t* GetCsc(int32_t dev, int32_t p, int32_t hop, int32_t which) override {
    if (csc_hops_ == 0 || dev >= (int32_t)csc_.size() || csc_[dev].empty()) return nullptr;
    return (int32_t*)csc_[dev][((size_t)(p % pipeline_depth_) * csc_hops_ + (hop - 1)) * 3 + which];
  }

  void IPCPost(int32_t d, int32_t p) override {
    if (ext_ && d < LG_EXT_MAX_DEVICE && p < LG_EXT_SLOTS) {
      ext_->seq[d][p] = ext_->seq[d][p] + 1;  // the counters of this batch are in place (the runner joined its streams)
      __sync_synchronize();
    }
    sem_post(semw_[d][p]);
  }
  void IPCWait(int32_t d, int32_t p) override { sem_wait(semr_[d][p]); }

  void Finalize() override {
    for (int32_t i = 0; i < device_count_; i++) {
      LGCHECK(lg_set_device(i));
      for (int32_t j = 0; j < pipeline_depth_; j++) {
        lg_device_free(ids_[i][j]);
        if (j < (int32_t)float_features_[i].size()) lg_device_free(float_features_[i][j]);
        lg_device_free(labels_[i][j]);
        lg_device_free(agg_src_[i][j]);
        lg_device_free(agg_dst_[i][j]);
        lg_device_free(node_counter_[i][j]);
        lg_device_free(edge_counter_[i][j]);
        sem_close(semw_[i][j]);
        sem_close(semr_[i][j]);
        sem_unlink(("sem_r_" + std::to_string(i) + "_" + std::to_string(j)).c_str());
        sem_unlink(("sem_w_" + std::to_string(i) + "_" + std::to_string(j)).c_str());
      }
    }
    for (size_t i = 0; i < csc_.size(); i++) {
      if (csc_[i].empty()) continue;
      LGCHECK(lg_set_device((int32_t)i));
      for (void* ptr : csc_[i]) lg_device_free(ptr);
    }
    munmap((void*)shm_, sizeof(shmStruct));
    close(fd_);
    shm_unlink(kShmName);
    if (ext_) {
      lg_host_unregister((void*)ext_);
      munmap((void*)ext_, sizeof(lg_ext_shm));
      close(ext_fd_);
      shm_unlink(LG_EXT_SHM_NAME);
      ext_ = nullptr;
    }
  }

  int32_t GetTrainStep() override { return train_step_; }

 private:
  volatile shmStruct* shm_ = nullptr;
  int fd_ = -1;
  lg_ext_shm* ext_ = nullptr;
  int ext_fd_ = -1;
  std::vector<std::vector<void*>> csc_;  // [gpu][(slot * hops + h) * 3 + k]
  int32_t csc_hops_ = 0;
  std::vector<std::vector<void*>> ids_, float_features_, labels_, agg_src_, agg_dst_, node_counter_, edge_counter_;
  std::vector<int64_t> feature_rows_;
  std::vector<std::vector<sem_t*>> semr_, semw_;
  int32_t raw_batch_size_ = 0;
  std::vector<int32_t> train_batch_size_, valid_batch_size_, test_batch_size_;
  int32_t device_count_ = 0, train_step_ = 0, valid_step_ = 0, test_step_ = 0, epoch_ = 0, pipeline_depth_ = 0;
};
}  // namespace

IPCEnv* NewIPCEnv(int32_t device_count) { return new CUDAIPCEnv(device_count); }

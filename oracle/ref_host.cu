// ref_host.cu — HOST code of the reference's IPC environment, compiled from the sources where they lie under
// /root/reference (nothing is copied): CUDAIPCEnv::Coordinate / GetMaxStep / GetCurrentMode / GetLocalBatchId /
// GetCurrentBatchsize (engine/ipc_service.cu:60-132,213-253).  Output: oracle/_ref/libref_host.so.  No kernels: it loads
// and runs on a box without a GPU, so the schedule oracle (lgo_coordinate, lgo_mode_of) is pinned in the CPU test suite.
// TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#define private public
#include "engine/ipc_service.cu"
#undef private
#include "engine/helper_multiprocess.cu"

extern "C" {

// Builds the environment (creates + zeroes the `simpleIPCshm` segment like the server does), runs Coordinate.
void* ref_env_coordinate(const int32_t* train_num, const int32_t* valid_num, const int32_t* test_num, int32_t parts,
                         int32_t raw_batch, int32_t epoch, int32_t* steps3, int32_t* train_bs, int32_t* valid_bs,
                         int32_t* test_bs, int32_t* max_step, int32_t* shm_steps3) {
  CUDAIPCEnv* env = new CUDAIPCEnv(parts);
  BuildInfo info;
  info.partition_count = parts;
  info.epoch = epoch;
  info.raw_batch_size = raw_batch;
  for (int i = 0; i < parts; i++) {
    info.training_set_num.push_back(train_num[i]);
    info.validation_set_num.push_back(valid_num[i]);
    info.testing_set_num.push_back(test_num[i]);
  }
  env->Coordinate(&info);
  steps3[0] = env->train_step_;
  steps3[1] = env->valid_step_;
  steps3[2] = env->test_step_;
  for (int i = 0; i < parts; i++) {
    train_bs[i] = env->GetCurrentBatchsize(i, TRAINMODE);
    valid_bs[i] = env->GetCurrentBatchsize(i, VALIDMODE);
    test_bs[i] = env->GetCurrentBatchsize(i, TESTMODE);
  }
  *max_step = env->GetMaxStep();
  for (int k = 0; k < 3; k++) shm_steps3[k] = env->shm_->steps[k];  // what the trainer reads (ipc_cuda_kernel.cu:44-52)
  return env;
}

void ref_env_schedule(void* env_, int32_t global_batch_id, int32_t* mode, int32_t* local_batch_id) {
  CUDAIPCEnv* env = (CUDAIPCEnv*)env_;
  *mode = env->GetCurrentMode(global_batch_id);
  *local_batch_id = env->GetLocalBatchId(global_batch_id);
}

int32_t ref_shm_struct_size(void) { return (int32_t)sizeof(shmStruct); }

void ref_env_free(void* env_) {  // not env->Finalize(): that frees device buffers this test never allocated
  CUDAIPCEnv* env = (CUDAIPCEnv*)env_;
  sharedMemoryClose(&env->info_);
  shm_unlink("simpleIPCshm");
  delete env;
}

}  // extern "C"

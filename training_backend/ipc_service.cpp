// ipc_service.cpp — trainer side of the hand-off: torch extension `ipc_service` with the six
// functions legion_graphsage/gcn/gat.py call (reference: training_backend/ipc_service.cpp:93-100,
// ipc_cuda_kernel.cu:35-235).  Same wire contract: shm "simpleIPCshm" {int32 steps[3];
// cudaIpcMemHandle_t memHandle[8][2][7]}, semaphores sem_{r,w}_<gpu>_<slot>, zero-copy
// torch::from_blob views of the server's device buffers.  Host C++ only (CUDA runtime calls, no kernels).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <semaphore.h>
#include <sys/mman.h>
#include <torch/extension.h>
#include <unistd.h>

#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

#define INTRABATCH_CON 3
#define INTERBATCH_CON 2
#define MAX_DEVICE 8
#define MEMORY_USAGE 7

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      printf("Cuda failure %s:%d: '%s'\n", __FILE__, __LINE__, cudaGetErrorString(e_));                 \
      exit(EXIT_FAILURE);                                                                               \
    }                                                                                                   \
  } while (0)

typedef struct shmStruct_st {
  int32_t steps[3];
  cudaIpcMemHandle_t memHandle[MAX_DEVICE][INTERBATCH_CON][MEMORY_USAGE];
} shmStruct;
static_assert(sizeof(shmStruct) == 7180, "simpleIPCshm layout");

namespace {
struct Env {
  int device = -1;
  int32_t steps[3] = {0, 0, 0};
  void* buf[INTERBATCH_CON][MEMORY_USAGE];
  sem_t* semr[INTERBATCH_CON];
  sem_t* semw[INTERBATCH_CON];
  int current_pipe = 0;
  int32_t* h_counters = nullptr;  // pinned: node_counter[16] | edge_counter[16]
} env;
}  // namespace

void InitializeIPC() {
  CUDA_OK(cudaGetDevice(&env.device));
  int fd = shm_open("simpleIPCshm", O_RDWR | O_CREAT, 0777);  // sharedMemoryCreate semantics
  if (fd < 0 || ftruncate(fd, sizeof(shmStruct)) != 0) {
    printf("Failed to create shared memory slab\n");
    exit(EXIT_FAILURE);
  }
  auto* shm = (volatile shmStruct*)mmap(0, sizeof(shmStruct), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if ((void*)shm == MAP_FAILED) {
    printf("Failed to create shared memory slab\n");
    exit(EXIT_FAILURE);
  }
  for (int i = 0; i < 3; i++) env.steps[i] = shm->steps[i];
  for (int i = 0; i < INTERBATCH_CON; i++)
    for (int k = 0; k < MEMORY_USAGE; k++)
      CUDA_OK(cudaIpcOpenMemHandle(&env.buf[i][k], *(cudaIpcMemHandle_t*)&shm->memHandle[env.device][i][k],
                                   cudaIpcMemLazyEnablePeerAccess));
  std::cout << "CUDA: " << env.device << " IPC shared memory opened\n";
  for (int i = 0; i < INTERBATCH_CON; i++) {
    std::string r = "sem_r_" + std::to_string(env.device) + "_" + std::to_string(i);
    std::string w = "sem_w_" + std::to_string(env.device) + "_" + std::to_string(i);
    env.semr[i] = sem_open(r.c_str(), O_CREAT | O_RDWR, 0666, 0);
    env.semw[i] = sem_open(w.c_str(), O_CREAT | O_RDWR, 0666, 0);
    if (env.semr[i] == SEM_FAILED || env.semw[i] == SEM_FAILED) {
      printf("errno = %d\n", errno);
      exit(EXIT_FAILURE);
    }
    sem_post(env.semr[i]);  // both slots start free
  }
  CUDA_OK(cudaMallocHost(&env.h_counters, 32 * sizeof(int32_t)));
  env.current_pipe = 0;
  munmap((void*)shm, sizeof(shmStruct));
  close(fd);
}

void FinalizeIPC() {
  for (int i = 0; i < INTERBATCH_CON; i++) {
    for (int k = 0; k < MEMORY_USAGE; k++) cudaIpcCloseMemHandle(env.buf[i][k]);
    sem_close(env.semw[i]);
    sem_close(env.semr[i]);
  }
}

std::vector<torch::Tensor> get_next(int feature_dim) {
  sem_wait(env.semw[env.current_pipe]);
  void** b = env.buf[env.current_pipe];
  int32_t* nc = env.h_counters;
  int32_t* ec = env.h_counters + 16;
  CUDA_OK(cudaMemcpy(nc, b[5], 16 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(ec, b[6], 16 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  const int hop_num = nc[INTRABATCH_CON * 3 - 1];
  auto dev = torch::Device(torch::kCUDA, env.device);
  auto i32 = torch::TensorOptions().dtype(torch::kI32).device(dev);
  auto f32 = torch::TensorOptions().dtype(torch::kF32).device(dev);
  const long long n = nc[INTRABATCH_CON * 3 + hop_num];
  std::vector<torch::Tensor> ret;
  ret.push_back(torch::from_blob(b[0], {n}, i32));
  ret.push_back(torch::from_blob(b[1], {n, (long long)feature_dim}, f32));
  ret.push_back(torch::from_blob(b[2], {(long long)nc[INTRABATCH_CON * 3]}, i32));
  for (int i = hop_num; i > 0; i--) {  // block i holds the edges of hops 1..i (cumulative views)
    const long long e = ec[INTRABATCH_CON * 3 + i];
    ret.push_back(torch::from_blob(b[3], {e}, i32));
    ret.push_back(torch::from_blob(b[4], {e}, i32));
  }
  return ret;
}

std::vector<int> get_block_size() {
  std::vector<int> ret;
  const int32_t* nc = env.h_counters;
  const int hop_num = nc[INTRABATCH_CON * 3 - 1];
  for (int i = hop_num; i > 0; i--) {
    ret.push_back(nc[INTRABATCH_CON * 3 + i]);
    ret.push_back(nc[INTRABATCH_CON * 3 + i - 1]);
  }
  return ret;
}

std::vector<int32_t> get_steps() { return {env.steps[0], env.steps[1], env.steps[2]}; }

void Synchronize() {
  sem_post(env.semr[env.current_pipe]);
  env.current_pipe = (env.current_pipe + 1) % INTERBATCH_CON;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("get_next", &get_next, "dataset get next (CUDA)");
  m.def("get_block_size", &get_block_size, "get dgl block size(CUDA)");
  m.def("get_steps", &get_steps, "get steps(CUDA)");
  m.def("initialize", &InitializeIPC, "InitializeIPC (CUDA)");
  m.def("finalize", &FinalizeIPC, "FinalizeIPC (CUDA)");
  m.def("synchronize", &Synchronize, "synchronize (CUDA)");
}

// ipc_service.cpp — trainer side of the hand-off: torch extension `ipc_service` with the six
// functions legion_graphsage/gcn/gat.py call (reference: training_backend/ipc_service.cpp:93-100,
// ipc_cuda_kernel.cu:35-235).  Same wire contract: shm "simpleIPCshm" {int32 steps[3];
// cudaIpcMemHandle_t memHandle[8][2][7]}, semaphores sem_{r,w}_<gpu>_<slot>, zero-copy
// torch::from_blob views of the server's device buffers.  Host C++ only (CUDA runtime calls, no kernels of its own).
//
// Two additions next to the reference's six functions (the unchanged trainers never need them):
//   * get_next reads the batch's counters from the server's side channel (include/legion_b200_ext.h) when the server
//     provides it — no device copy, no synchronisation; against a reference server it falls back to the reference's two
//     blocking cudaMemcpy (training_backend/ipc_cuda_kernel.cu:194-195);
//   * get_next_csc(feature_dim) = get_next + the CSC of every block (indptr over destinations, sources, COO positions),
//     built on the caller's current CUDA stream by lg_block_csc (liblegion_b200.so, csrc/blocks.cu) straight from the
//     CUDA-IPC buffers: what dgl.create_block(('csc', (indptr, indices, eids)), ...) takes, instead of DGL's own
//     COO -> CSC conversion behind create_unitgraph_from_coo (training_backend/legion_graphsage.py:66-79).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <semaphore.h>
#include <sys/mman.h>
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../include/legion_b200.h"
#include "../include/legion_b200_ext.h"

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
  lg_ext_shm* ext = nullptr;      // the server's side channel, when there is one
  uint32_t expect_seq[INTERBATCH_CON] = {1, 1};
  bool ext_used = false;
  torch::Tensor csc_workspace;
  int64_t csc_edges = 0;
  int csc_hops = 0;  // > 0: the server builds the blocks; three more IPC buffers per (slot, block)
  void* csc_buf[INTERBATCH_CON][LG_EXT_MAX_HOPS][3];
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
  env.ext = nullptr;
  env.ext_used = false;
  for (int i = 0; i < INTERBATCH_CON; i++) env.expect_seq[i] = 1;
  {  // side channel: only if a server created it (no O_CREAT), it is the right version and it serves this GPU
    const char* e = getenv("LEGION_EXT_SHM");
    int efd = (e && atoi(e) == 0) ? -1 : shm_open(LG_EXT_SHM_NAME, O_RDWR, 0);
    if (efd >= 0) {
      void* addr = mmap(0, sizeof(lg_ext_shm), PROT_READ | PROT_WRITE, MAP_SHARED, efd, 0);
      close(efd);
      if (addr != MAP_FAILED) {
        auto* x = (lg_ext_shm*)addr;
        if (x->magic == LG_EXT_MAGIC && x->version == LG_EXT_VERSION && env.device < x->n_gpus && env.device < LG_EXT_MAX_DEVICE)
          env.ext = x;
        else
          munmap(addr, sizeof(lg_ext_shm));
        env.csc_hops = 0;
        if (env.ext && env.ext->csc_hops > 0 && env.ext->csc_hops <= LG_EXT_MAX_HOPS) {
          env.csc_hops = env.ext->csc_hops;
          for (int i = 0; i < INTERBATCH_CON; i++)
            for (int h = 0; h < env.csc_hops; h++)
              for (int k = 0; k < 3; k++)
                CUDA_OK(cudaIpcOpenMemHandle(&env.csc_buf[i][h][k], *(cudaIpcMemHandle_t*)env.ext->csc_handle[env.device][i][h][k],
                                             cudaIpcMemLazyEnablePeerAccess));
        }
      }
    }
  }
  munmap((void*)shm, sizeof(shmStruct));
  close(fd);
}

void FinalizeIPC() {
  for (int i = 0; i < INTERBATCH_CON && env.csc_hops > 0; i++)
    for (int h = 0; h < env.csc_hops; h++)
      for (int k = 0; k < 3; k++) cudaIpcCloseMemHandle(env.csc_buf[i][h][k]);
  env.csc_hops = 0;
  if (env.ext) munmap((void*)env.ext, sizeof(lg_ext_shm));
  env.ext = nullptr;
  env.csc_workspace = torch::Tensor();
  for (int i = 0; i < INTERBATCH_CON; i++) {
    for (int k = 0; k < MEMORY_USAGE; k++) cudaIpcCloseMemHandle(env.buf[i][k]);
    sem_close(env.semw[i]);
    sem_close(env.semr[i]);
  }
}

// counters of the batch in the current slot -> env.h_counters
static void fetch_counters() {
  const int p = env.current_pipe;
  void** b = env.buf[p];
  int32_t* nc = env.h_counters;
  int32_t* ec = env.h_counters + 16;
  if (env.ext) {
    // the server bumps seq[gpu][slot] once per batch after the counters landed and before sem_post: anything else means
    // the segment belongs to another (dead) server -> the reference's path from here on
    if (env.ext->seq[env.device][p] == env.expect_seq[p]) {
      __sync_synchronize();
      for (int i = 0; i < 32; i++) env.h_counters[i] = env.ext->counters[env.device][p][i];
      env.expect_seq[p]++;
      env.ext_used = true;
      return;
    }
    munmap((void*)env.ext, sizeof(lg_ext_shm));
    env.ext = nullptr;
  }
  CUDA_OK(cudaMemcpy(nc, b[5], 16 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(ec, b[6], 16 * sizeof(int32_t), cudaMemcpyDeviceToHost));
}

std::vector<torch::Tensor> get_next(int feature_dim) {
  sem_wait(env.semw[env.current_pipe]);
  void** b = env.buf[env.current_pipe];
  int32_t* nc = env.h_counters;
  int32_t* ec = env.h_counters + 16;
  fetch_counters();
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

// get_next + CSC of every block: [ids, features, labels, (indptr, indices, eids) for block H..1]
std::vector<torch::Tensor> get_next_csc(int feature_dim) {
  std::vector<torch::Tensor> coo = get_next(feature_dim);
  const int32_t* nc = env.h_counters;
  const int32_t* ec = env.h_counters + 16;
  const int hop_num = nc[INTRABATCH_CON * 3 - 1];
  auto dev = torch::Device(torch::kCUDA, env.device);
  auto i32 = torch::TensorOptions().dtype(torch::kI32).device(dev);
  if (env.csc_hops >= hop_num && env.ext_used && env.ext) {  // built by the server next to the gather: zero-copy views
    std::vector<torch::Tensor> ret = {coo[0], coo[1], coo[2]};
    for (int i = hop_num; i > 0; i--) {
      void** c = env.csc_buf[env.current_pipe][i - 1];
      const long long e = ec[INTRABATCH_CON * 3 + i], num_dst = nc[INTRABATCH_CON * 3 + i - 1];
      ret.push_back(torch::from_blob(c[0], {num_dst + 1}, i32));
      ret.push_back(torch::from_blob(c[1], {e}, i32));
      ret.push_back(torch::from_blob(c[2], {e}, i32));
    }
    return ret;
  }
  // built here, on the caller's current stream: all blocks in one set of launches (lg_block_csc_batch)
  int64_t max_edges[LG_MAX_HOPS];
  int32_t max_dst[LG_MAX_HOPS];
  int32_t *indptr[LG_MAX_HOPS], *indices[LG_MAX_HOPS], *eids[LG_MAX_HOPS];
  std::vector<torch::Tensor> t_indptr(hop_num), t_indices(hop_num), t_eids(hop_num);
  for (int h = 1; h <= hop_num; h++) {
    max_edges[h - 1] = ec[INTRABATCH_CON * 3 + h] > 0 ? ec[INTRABATCH_CON * 3 + h] : 1;
    max_dst[h - 1] = nc[INTRABATCH_CON * 3 + h - 1] > 0 ? nc[INTRABATCH_CON * 3 + h - 1] : 1;
    t_indptr[h - 1] = torch::empty({(long long)nc[INTRABATCH_CON * 3 + h - 1] + 1}, i32);
    t_indices[h - 1] = torch::empty({(long long)ec[INTRABATCH_CON * 3 + h]}, i32);
    t_eids[h - 1] = torch::empty({(long long)ec[INTRABATCH_CON * 3 + h]}, i32);
    indptr[h - 1] = t_indptr[h - 1].data_ptr<int32_t>();
    indices[h - 1] = t_indices[h - 1].data_ptr<int32_t>();
    eids[h - 1] = t_eids[h - 1].data_ptr<int32_t>();
  }
  int64_t nb = 0;
  if (lg_block_csc_batch_workspace(hop_num, max_edges, &nb) != 0) throw std::runtime_error(lg_last_error());
  if (!env.csc_workspace.defined() || nb > env.csc_workspace.numel())
    env.csc_workspace = torch::empty({nb + nb / 4}, torch::TensorOptions().dtype(torch::kU8).device(dev));
  void** b = env.buf[env.current_pipe];
  lg_batch batch;
  memset(&batch, 0, sizeof(batch));
  batch.agg_src = (int32_t*)b[3];
  batch.agg_dst = (int32_t*)b[4];
  batch.node_counter = (int32_t*)b[5];
  batch.edge_counter = (int32_t*)b[6];
  void* st = (void*)c10::cuda::getCurrentCUDAStream(env.device).stream();
  if (lg_block_csc_batch(st, &batch, hop_num, max_edges, max_dst, indptr, indices, eids, env.csc_workspace.data_ptr(),
                         env.csc_workspace.numel()) != 0)
    throw std::runtime_error(lg_last_error());
  std::vector<torch::Tensor> ret = {coo[0], coo[1], coo[2]};
  for (int i = hop_num; i > 0; i--) {
    ret.push_back(t_indptr[i - 1]);
    ret.push_back(t_indices[i - 1]);
    ret.push_back(t_eids[i - 1]);
  }
  return ret;
}

// true when the last get_next read its counters from the server's side channel (no device copy)
bool counters_from_host() { return env.ext_used && env.ext != nullptr; }
// true when get_next_csc hands out the server's own CSC buffers (LEGION_EMIT_CSC=1 on the server)
bool csc_from_server() { return env.csc_hops > 0 && env.ext != nullptr; }

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
  m.def("get_next_csc", &get_next_csc, "get_next + CSC (indptr, indices, eids) of every block, built on the current stream");
  m.def("csc_from_server", &csc_from_server, "get_next_csc returns views of blocks the server built (no trainer-side kernels)");
  m.def("counters_from_host", &counters_from_host, "the last get_next needed no device copy (server side channel)");
}

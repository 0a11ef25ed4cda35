// ipc_service.h — server side of the hand-off (reference: sampling_server/src/engine/ipc_service.h:6-35).
#pragma once
#include <cstdint>

#include "buildinfo.h"
#include "system_config.h"

// shm wire layout shared with training_backend (reference engine/ipc_service.cu:28-31):
// 12 + 8*2*7*64 = 7180 bytes, cudaIpcMemHandle_t is 64 opaque bytes.
struct IpcHandleBytes {
  unsigned char b[64];
};
typedef struct shmStruct_st {
  int32_t steps[3];
  IpcHandleBytes memHandle[MAX_DEVICE][INTERBATCH_CON][MEMORY_USAGE];
} shmStruct;
static_assert(sizeof(shmStruct) == 7180, "simpleIPCshm layout");

class IPCEnv {
 public:
  virtual ~IPCEnv() {}
  virtual void Coordinate(BuildInfo* info) = 0;
  virtual int32_t GetMaxStep() = 0;
  virtual void InitializeSamplesBuffer(int32_t batch_size, int32_t num_ids, int32_t feature_dim, int32_t device_id,
                                       int32_t pipeline_depth) = 0;
  virtual void InitializeFeaturesBuffer(int32_t batch_size, int32_t num_ids, int32_t feature_dim, int32_t device_id,
                                        int32_t pipeline_depth) = 0;
  virtual int32_t GetRawBatchsize() = 0;
  virtual int32_t GetLocalBatchId(int32_t global_batch_id) = 0;
  virtual int32_t GetCurrentBatchsize(int32_t dev_id, int32_t current_mode) = 0;
  virtual int32_t GetCurrentMode(int32_t global_batch_id) = 0;
  virtual int32_t* GetIds(int32_t dev_id, int32_t current_pipe) = 0;
  virtual float* GetFloatFeatures(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int32_t* GetLabels(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int32_t* GetAggSrc(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int32_t* GetAggDst(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int32_t* GetNodeCounter(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int32_t* GetEdgeCounter(int32_t dev_id, int32_t current_pipe) = 0;
  virtual int64_t GetFeatureRows(int32_t dev_id) = 0;
  // host words (node_counter[16] | edge_counter[16]) of the side channel include/legion_b200_ext.h, or nullptr when it
  // is disabled: the runner copies the batch's counters there behind its last kernel, IPCPost publishes them
  virtual int32_t* GetHostCounters(int32_t dev_id, int32_t current_pipe) = 0;
  // LEGION_EMIT_CSC=1 and the side channel is up: three more CUDA-IPC buffers per (gpu, slot, block) for the blocks as
  // CSC (include/legion_b200_ext.h).  max_edges[h-1] / max_dst[h-1] bound block h.  Returns false when disabled.
  virtual bool InitializeCscBuffers(int32_t dev_id, int32_t pipeline_depth, int32_t hops, const int64_t* max_edges,
                                    const int64_t* max_dst) = 0;
  virtual int32_t* GetCsc(int32_t dev_id, int32_t current_pipe, int32_t hop, int32_t which) = 0;  // which: 0 indptr, 1 indices, 2 eids
  virtual void IPCPost(int32_t dev_id, int32_t current_pipe) = 0;
  virtual void IPCWait(int32_t dev_id, int32_t current_pipe) = 0;
  virtual void Finalize() = 0;
  virtual int32_t GetTrainStep() = 0;
};
IPCEnv* NewIPCEnv(int32_t device_count);

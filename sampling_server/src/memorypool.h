// memorypool.h — per-GPU op context (reference: engine/memorypool.cuh:20-221).  The reference's
// pool is a bag of raw scratch pointers; here the scratch lives inside the lg_sampler handle
// and the pool only carries the per-slot IPC buffers + the iteration state.
#pragma once
#include <vector>

#include "system_config.h"

class MemoryPool {
 public:
  explicit MemoryPool(int32_t pipeline_depth) : batch_(pipeline_depth) {}
  int32_t GetIter() const { return iter_; }
  void SetIter(int32_t i) { iter_ = i; }
  int32_t GetCurrentMode() const { return mode_; }
  void SetCurrentMode(int32_t m) { mode_ = m; }
  void SetCurrentPipe(int32_t p) { current_pipe_ = p; }
  int32_t GetCurrentPipe() const { return current_pipe_; }
  void SetGlobalBatchId(int32_t g) { global_batch_id_ = g; }
  int32_t GetGlobalBatchId() const { return global_batch_id_; }
  // the buffers the operators of the batch in flight work on: the IPC slot, or (staged sampling, server.cc) a
  // server-private staging set while the trainer still owns the slot
  lg_batch* Batch() { return view_ ? view_ : &batch_[current_pipe_]; }
  lg_batch* Batch(int pipe) { return &batch_[pipe]; }
  void SetBatchView(lg_batch* v) { view_ = v; }
  void SetCurrentSampler(int32_t i) { current_sampler_ = i; }
  // one sampler handle (position map, scan state, frontier scratch) per pipeline slot: the sampling of batch i+1
  // does not queue behind batch i's on a shared handle (the reference has one scratch set per GPU, engine/server.cu:221-234)
  lg_sampler* Sampler() {
    if (current_sampler_ >= 0 && current_sampler_ < (int32_t)samplers.size()) return samplers[current_sampler_];
    return samplers[samplers.size() > 1 ? current_pipe_ : 0];
  }
  lg_sampler* Sampler(int pipe) { return samplers[samplers.size() > 1 ? pipe : 0]; }
  std::vector<lg_sampler*> samplers;
  int32_t rng_kind = LG_RNG_PHILOX;
  uint64_t rng_seed = 0x1E910;

 private:
  std::vector<lg_batch> batch_;
  int32_t iter_ = 0, mode_ = 0, current_pipe_ = 0, global_batch_id_ = 0, current_sampler_ = -1;
  lg_batch* view_ = nullptr;
};

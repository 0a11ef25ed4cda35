// server.h — in-process API of the sampling server, identical in shape to the reference's
// (sampling_server/src/engine/server.h:5-33) so main() and the pybind wrapper read the same.
#pragma once
#include <vector>

struct RunnerParams {
  int device_id;
  std::vector<int> fanout;
  void* cache;
  void* graph;
  void* feature;
  void* env;
  int global_batch_id;
  bool in_memory;
};

class Server {
 public:
  virtual ~Server() {}
  virtual void Initialize(int global_shard_count, std::vector<int> fanout, int in_memory_mode) = 0;
  virtual void PreSc(int cache_agg_mode) = 0;
  virtual void Run() = 0;
  virtual void Finalize() = 0;
};
Server* NewGPUServer();

class Runner {
 public:
  virtual ~Runner() {}
  virtual void Initialize(RunnerParams* params) = 0;
  virtual void InitializeFeaturesBuffer(RunnerParams* params) = 0;
  virtual void RunPreSc(RunnerParams* params) = 0;
  virtual void RunOnce(RunnerParams* params) = 0;
  virtual void Finalize(RunnerParams* params) = 0;
};
Runner* NewGPURunner();

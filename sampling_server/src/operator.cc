// operator.cc — the five operators (reference: engine/operator.cu:10-129).  Each run() unpacks
// OpParams exactly like the reference, calls ONE C-ABI entry point on params->stream and records
// params->event.
#include "operator.h"

#include <cstdlib>

#include "cache.h"
#include "ipc_service.h"
#include "memorypool.h"
#include "storage.h"

namespace {
class BatchGenerateOP : public Operator {
 public:
  explicit BatchGenerateOP(int op_id) : op_id_(op_id) {}
  void run(OpParams* params) override {
    auto* feature = (FeatureStorage*)params->feature;
    auto* pool = (MemoryPool*)params->memorypool;
    auto* env = (IPCEnv*)params->env;
    int32_t dev = params->device_id, mode = pool->GetCurrentMode(), iter = pool->GetIter();
    int32_t batch_size = env->GetCurrentBatchsize(dev, mode);
    const int32_t *ids, *labels;
    int32_t cap;
    if (mode == TRAINMODE) {
      ids = feature->GetTrainingSetIds(dev); labels = feature->GetTrainingLabels(dev); cap = feature->TrainingSetSize(dev);
    } else if (mode == VALIDMODE) {
      ids = feature->GetValidationSetIds(dev); labels = feature->GetValidationLabels(dev); cap = feature->ValidationSetSize(dev);
    } else {
      ids = feature->GetTestingSetIds(dev); labels = feature->GetTestingLabels(dev); cap = feature->TestingSetSize(dev);
    }
    LGCHECK(lg_batch_generate(pool->Sampler(), params->stream, ids, labels, cap, batch_size, iter, pool->Batch()));
    if (params->event) LGCHECK(lg_event_record(params->event, params->stream));
  }
 private:
  int op_id_;
};

class RandomSampleOP : public Operator {
 public:
  explicit RandomSampleOP(int op_id) : op_id_(op_id) {}
  void run(OpParams* params) override {
    auto* pool = (MemoryPool*)params->memorypool;
    auto* graph = (GraphStorage*)params->graph;
    auto* cache = (UnifiedCache*)params->cache;
    int32_t dev = params->device_id;
    // presampling reads the host CSR directly and counts edge accesses (pre_sample, operator_impl.cu:301-397)
    unsigned long long* edge_hot = params->is_presc ? cache->GetEdgeAccessedMap(dev) : nullptr;
    LGCHECK(lg_random_sample(pool->Sampler(), params->stream, graph->Topology(dev), op_id_ / INTRABATCH_CON, pool->rng_kind,
                             pool->rng_seed, (uint32_t)pool->GetGlobalBatchId(), (uint32_t)dev, pool->Batch(), edge_hot));
    if (params->event) LGCHECK(lg_event_record(params->event, params->stream));
  }
 private:
  int op_id_;
};

// The reference gathers after every sampling op (seeds, hop 1, ..., hop H: H+1 launches per batch).  By default the
// lookup ops before the last one only record their event and the last one gathers the rows of all hops in one launch
// (lg_feature_cache_lookup_range): the trainer sees the same buffers and the same final counters, and every launch
// less matters next to the streaming gather.  LEGION_FUSE_GATHERS=0 restores one gather per op.
class CacheLookupOP : public Operator {
 public:
  explicit CacheLookupOP(int op_id) : op_id_(op_id) {}
  void run(OpParams* params) override {
    static const bool fuse = [] {
      const char* e = std::getenv("LEGION_FUSE_GATHERS");
      return !e || std::atoi(e) != 0;
    }();
    auto* pool = (MemoryPool*)params->memorypool;
    auto* cache = (UnifiedCache*)params->cache;
    int32_t dev = params->device_id;
    const bool last = op_id_ / INTRABATCH_CON == params->hop_num;
    if (!fuse)
      LGCHECK(lg_feature_cache_lookup(pool->Sampler(), params->stream, cache->FeatureCache(dev), op_id_, cache->LocalPart(dev),
                                      pool->Batch(), cache->TierRows(dev)));
    else if (last)
      LGCHECK(lg_feature_cache_lookup_range(pool->Sampler(), params->stream, cache->FeatureCache(dev), op_id_, 0,
                                            cache->LocalPart(dev), pool->Batch(), cache->TierRows(dev)));
    if (params->event) LGCHECK(lg_event_record(params->event, params->stream));
  }
 private:
  int op_id_;
};

class SSDIOSubmitOP : public Operator {
 public:
  explicit SSDIOSubmitOP(int op_id) : op_id_(op_id) {}
  void run(OpParams* params) override {
    auto* pool = (MemoryPool*)params->memorypool;
    LGCHECK(lg_io_submit(pool->Sampler(), params->stream, op_id_, pool->Batch()));
    if (params->event) LGCHECK(lg_event_record(params->event, params->stream));
  }
 private:
  int op_id_;
};

class SSDIOCompleteOP : public Operator {
 public:
  explicit SSDIOCompleteOP(int op_id) : op_id_(op_id) {}
  void run(OpParams* params) override {
    auto* pool = (MemoryPool*)params->memorypool;
    auto* cache = (UnifiedCache*)params->cache;
    int32_t dev = params->device_id;
    bool presc = params->is_presc;
    LGCHECK(lg_io_complete(pool->Sampler(), params->stream, pool->GetCurrentMode(), pool->Batch(),
                           presc ? cache->GetNodeAccessedMap(dev) : nullptr, presc ? cache->MaxIdsDevice(dev) : nullptr));
    if (params->event) LGCHECK(lg_event_record(params->event, params->stream));
  }
 private:
  int op_id_;
};
}  // namespace

Operator* NewBatchGenerateOP(int op_id) { return new BatchGenerateOP(op_id); }
Operator* NewRandomSampleOP(int op_id) { return new RandomSampleOP(op_id); }
Operator* NewCacheLookupOP(int op_id) { return new CacheLookupOP(op_id); }
Operator* NewSSDIOSubmitOP(int op_id) { return new SSDIOSubmitOP(op_id); }
Operator* NewSSDIOCompleteOP(int op_id) { return new SSDIOCompleteOP(op_id); }

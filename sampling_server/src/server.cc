// server.cc — GPUServer / GPURunner (reference: engine/server.cu:44-370).
// One host thread per GPU; per GPU three streams and an op DAG of (hops+1)*3+1 operators.  The
// gather of hop h (stream 1) overlaps the sampling of hop h+1 (stream 0) exactly as in the
// reference; unlike the reference (which polls only the last op's event, server.cu:319-324, and
// can post a batch whose gather is still running) every stream is joined before IPCPost.
#include "server.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <thread>

#include "cache.h"
#include "ipc_service.h"
#include "memorypool.h"
#include "operator.h"
#include "storage.h"

namespace {
void PreSCLoop(int train_step, Runner* runner, RunnerParams* params) {
  for (int i = 0; i < train_step; i++) {
    params->global_batch_id = i;
    runner->RunPreSc(params);
  }
  runner->InitializeFeaturesBuffer(params);
}
void RunnerLoop(int max_step, Runner* runner, RunnerParams* params) {
  for (int i = 0; i < max_step; i++) {
    params->global_batch_id = i;
    runner->RunOnce(params);
  }
}

class GPUServer : public Server {
 public:
  void Initialize(int global_shard_count, std::vector<int> fanout, int in_memory_mode) override {
    shard_count_ = global_shard_count;
    std::cout << (in_memory_mode ? "In Memory Mode\n" : "In Disk Mode\n");
    storage_ = new StorageManagement();
    storage_->Initialze(shard_count_, in_memory_mode, fanout);
    if (!storage_->GetInfo()->fanout.empty()) fanout = storage_->GetInfo()->fanout;  // meta_config extension
    graph_ = storage_->GetGraph();
    feature_ = storage_->GetFeature();
    cache_ = storage_->GetCache();
    ipc_env_ = storage_->GetIPCEnv();
    train_step_ = ipc_env_->GetTrainStep();
    max_step_ = ipc_env_->GetMaxStep();
    runners_.resize(shard_count_);
    params_.resize(shard_count_);
    for (int i = 0; i < shard_count_; i++) {
      LGCHECK(lg_set_device(i));
      auto* p = new RunnerParams();
      p->device_id = i;
      p->fanout = fanout;
      p->cache = cache_;
      p->graph = graph_;
      p->feature = feature_;
      p->env = ipc_env_;
      p->global_batch_id = 0;
      p->in_memory = true;
      params_[i] = p;
      runners_[i] = NewGPURunner();
      runners_[i]->Initialize(p);
    }
  }

  void PreSc(int cache_agg_mode) override {
    auto t1 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int i = 0; i < shard_count_; i++) pool.emplace_back(&PreSCLoop, train_step_, runners_[i], params_[i]);
    for (auto& th : pool) th.join();
    std::vector<uint64_t> counters(2, 0);  // PCM PCIe counters: disabled in the reference (server.cu:106)
    double t = std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::steady_clock::now() - t1).count();
    cache_->CandidateSelection(cache_agg_mode, feature_, graph_);
    cache_->CostModel(cache_agg_mode, feature_, graph_, counters, train_step_);
    cache_->FillUp(cache_agg_mode, feature_, graph_);
    std::cout << "Preprocessing cost: " << t << " s\n";
    std::cout << "System is ready for serving\n" << std::flush;
  }

  void Run() override {
    auto t1 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int i = 0; i < shard_count_; i++) pool.emplace_back(&RunnerLoop, max_step_, runners_[i], params_[i]);
    for (auto& th : pool) th.join();
    serve_seconds_ = std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::steady_clock::now() - t1).count();
  }

  // One JSON line per GPU: rows gathered per tier over the whole run (local HBM shard / peer shard over NVLink / host
  // backing matrix over PCIe), and the bytes those rows moved per second of serving (metrics row: the
  // reference only has commented-out hit-rate printing, cache/cache.cu:197-214).
  void Telemetry() {
    const int dim = feature_->GetFloatFeatureLen();
    for (int i = 0; i < shard_count_; i++) {
      unsigned long long rows[3] = {0, 0, 0};
      cache_->ReadTierRows(i, rows);
      const double tot = (double)(rows[0] + rows[1] + rows[2]), s = serve_seconds_ > 0 ? serve_seconds_ : 1.0;
      std::printf("{\"legion_b200_telemetry\": {\"gpu\": %d, \"batches\": %d, \"serve_seconds\": %.4f, \"rows\": {\"local\": %llu, "
                  "\"peer\": %llu, \"host\": %llu}, \"hit_mix\": {\"local\": %.4f, \"peer\": %.4f, \"host\": %.4f}, "
                  "\"GBps\": {\"local_hbm\": %.2f, \"peer_nvlink\": %.2f, \"host_pcie\": %.2f}, \"seeds_per_s\": %.1f}}\n",
                  i, max_step_, serve_seconds_, rows[0], rows[1], rows[2], tot > 0 ? rows[0] / tot : 0.0,
                  tot > 0 ? rows[1] / tot : 0.0, tot > 0 ? rows[2] / tot : 0.0, rows[0] * 4.0 * dim / s / 1e9,
                  rows[1] * 4.0 * dim / s / 1e9, rows[2] * 4.0 * dim / s / 1e9, (double)seeds_served_[i] / s);
    }
    std::fflush(stdout);
  }

  void Finalize() override {
    for (int i = 0; i < shard_count_; i++) runners_[i]->Finalize(params_[i]);
    seeds_served_.assign(shard_count_, 0);
    for (int i = 0; i < shard_count_; i++)
      for (int g = 0; g < max_step_; g++) seeds_served_[i] += ipc_env_->GetCurrentBatchsize(i, ipc_env_->GetCurrentMode(g));
    Telemetry();
    graph_->Finalize();
    feature_->Finalize();
    ipc_env_->Finalize();
    std::cout << "Server Stopped\n" << std::flush;
  }

 private:
  StorageManagement* storage_ = nullptr;
  GraphStorage* graph_ = nullptr;
  FeatureStorage* feature_ = nullptr;
  UnifiedCache* cache_ = nullptr;
  IPCEnv* ipc_env_ = nullptr;
  int shard_count_ = 0, train_step_ = 0, max_step_ = 0;
  double serve_seconds_ = 0.0;
  std::vector<long long> seeds_served_;
  std::vector<Runner*> runners_;
  std::vector<RunnerParams*> params_;
};

class GPURunner : public Runner {
 public:
  void Initialize(RunnerParams* params) override {
    LGCHECK(lg_set_device(params->device_id));
    local_dev_id_ = params->device_id;
    auto* cache = (UnifiedCache*)params->cache;
    auto* feature = (FeatureStorage*)params->feature;
    auto* env = (IPCEnv*)params->env;
    interbatch_concurrency_ = INTERBATCH_CON;
    {  // LEGION_SLOT_SAMPLERS=0: one sampler handle and one stream triple per GPU, as in the reference
      const char* e = std::getenv("LEGION_SLOT_SAMPLERS");
      slot_samplers_ = !(e && std::atoi(e) == 0);
    }
    // three streams per pipeline slot (the reference has three per GPU, engine/server.cu:181-184): with a sampler handle per
    // slot the chains of consecutive batches are independent, so they must not share stream 0 either
    streams_.resize(slot_samplers_ ? interbatch_concurrency_ : 1);
    for (auto& ss : streams_) {
      ss.resize(INTRABATCH_CON);
      for (auto& s : ss) LGCHECK(lg_stream_create(&s));
    }
    int batch_size = env->GetRawBatchsize();
    int hop_num = (int)params->fanout.size();
    num_ids_ = (int32_t)lg_num_ids(batch_size, params->fanout.data(), hop_num);  // server.cu:187-199
    op_num_ = (hop_num + 1) * INTRABATCH_CON + 1;
    ops_.resize(op_num_);
    ops_[0] = NewBatchGenerateOP(0);
    ops_[1] = NewCacheLookupOP(1);
    ops_[2] = NewSSDIOSubmitOP(2);
    for (int i = 0; i < hop_num; i++) {
      ops_[INTRABATCH_CON * i + 3] = NewRandomSampleOP(INTRABATCH_CON * i + 3);
      ops_[INTRABATCH_CON * i + 4] = NewCacheLookupOP(INTRABATCH_CON * i + 4);
      ops_[INTRABATCH_CON * i + 5] = NewSSDIOSubmitOP(INTRABATCH_CON * i + 5);
    }
    ops_[op_num_ - 1] = NewSSDIOCompleteOP(op_num_ - 1);
    cache->InitializeCacheController(local_dev_id_, feature->TotalNodeNum());
    memorypool_ = new MemoryPool(interbatch_concurrency_);
    // valid/test batches can be larger than the raw batch only if their sets are (512-sized steps): take the max
    int max_batch = batch_size;
    for (int m : {VALIDMODE, TESTMODE})
      if (env->GetCurrentBatchsize(local_dev_id_, m) > max_batch) max_batch = env->GetCurrentBatchsize(local_dev_id_, m);
    {  // staged sampling (RunOnceStaged): LEGION_STAGING=k sampling stages (default 3; 0 = sample straight into the IPC slot)
      const char* e = std::getenv("LEGION_STAGING");
      n_stage_ = e ? std::atoi(e) : 3;
      if (!slot_samplers_ || n_stage_ < 2) n_stage_ = 0;
      if (n_stage_ > 4) n_stage_ = 4;
    }
    memorypool_->samplers.assign(n_stage_ ? n_stage_ : (slot_samplers_ ? interbatch_concurrency_ : 1), nullptr);
    if (max_batch > batch_size) num_ids_ = (int32_t)lg_num_ids(max_batch, params->fanout.data(), hop_num);
    if (const char* e = std::getenv("LEGION_RNG")) memorypool_->rng_kind = std::strcmp(e, "minstd") == 0 ? LG_RNG_MINSTD : LG_RNG_PHILOX;
    if (const char* e = std::getenv("LEGION_SEED")) memorypool_->rng_seed = std::strtoull(e, nullptr, 0);
    for (auto& smp : memorypool_->samplers) {
      LGCHECK(lg_sampler_create(local_dev_id_, max_batch, params->fanout.data(), hop_num, feature->TotalNodeNum(), &smp));
      if (const char* e = std::getenv("LEGION_GATHER")) LGCHECK(lg_sampler_set_gather_variant(smp, std::atoi(e)));
      if (const char* e = std::getenv("LEGION_TAIL"))  // "reference": the clipped tail batch strides like the reference's
        LGCHECK(lg_sampler_set_tail_mode(smp, std::strcmp(e, "reference") == 0 ? LG_TAIL_REFERENCE : LG_TAIL_EXACT));
      // the trainer only sees a batch after its last op: let hop h+1 finish hop h's construct_graph and the last
      // sampling op release the position map (two launches less per batch); LEGION_LAZY_RELABEL=0 = every op complete
      const char* e = std::getenv("LEGION_LAZY_RELABEL");
      LGCHECK(lg_sampler_set_lazy_relabel(smp, (e && std::atoi(e) == 0) ? 0 : 1));
    }
    float_feature_len_ = feature->GetFloatFeatureLen();
    max_batch_ = max_batch;
    env->InitializeSamplesBuffer(max_batch, num_ids_, float_feature_len_, local_dev_id_, interbatch_concurrency_);
    {  // blocks as CSC, built by the server on the third stream next to the gather (LEGION_EMIT_CSC=1): bounds per block
      csc_max_edges_.assign(hop_num, 0);
      csc_max_dst_.assign(hop_num, 0);
      int64_t per = max_batch, nodes = max_batch, edges = 0;
      for (int h = 0; h < hop_num; h++) {
        csc_max_dst_[h] = nodes;  // destinations of block h+1 = every vertex of hops 0..h
        per *= params->fanout[h];
        edges += per;
        nodes += per;
        csc_max_edges_[h] = edges;  // cumulative edges of hops 1..h+1
      }
      emit_csc_ = env->InitializeCscBuffers(local_dev_id_, interbatch_concurrency_, hop_num, csc_max_edges_.data(), csc_max_dst_.data());
      if (emit_csc_) {
        int64_t nb = 0;
        LGCHECK(lg_block_csc_batch_workspace(hop_num, csc_max_edges_.data(), &nb));
        csc_max_dst32_.assign(csc_max_dst_.begin(), csc_max_dst_.end());
        csc_ptrs_.assign((size_t)interbatch_concurrency_ * 3 * hop_num, nullptr);
        for (int p = 0; p < interbatch_concurrency_; p++)
          for (int k = 0; k < 3; k++)
            for (int h = 1; h <= hop_num; h++) csc_ptrs_[((size_t)p * 3 + k) * hop_num + (h - 1)] = env->GetCsc(local_dev_id_, p, h, k);
        csc_workspace_bytes_ = nb;
        csc_workspace_.assign(interbatch_concurrency_, nullptr);
        csc_ev_.resize(interbatch_concurrency_);
        for (int i = 0; i < interbatch_concurrency_; i++) {
          LGCHECK(lg_device_alloc(&csc_workspace_[i], nb));
          LGCHECK(lg_event_create(&csc_ev_[i]));
        }
        if (local_dev_id_ == 0) std::cout << "Blocks are emitted as CSC (LEGION_EMIT_CSC)\n";
      }
    }
    current_pipe_ = 0;
    stage_b_.resize(n_stage_);
    stage_stream_.resize(n_stage_);
    sampled_ev_.resize(n_stage_);
    copied_ev_.resize(n_stage_);
    for (int i = 0; i < n_stage_; i++) {  // server-private twins of the slot's small buffers (everything but the features)
      lg_batch& b = stage_b_[i];
      std::memset(&b, 0, sizeof(b));
      auto alloc = [&](int64_t words) {
        void* p = nullptr;
        LGCHECK(lg_device_alloc(&p, words * 4));
        LGCHECK(lg_memset_async(p, 0, words * 4, nullptr));
        return (int32_t*)p;
      };
      b.ids = alloc(num_ids_);
      b.labels = alloc(max_batch);
      b.agg_src = alloc(num_ids_);
      b.agg_dst = alloc(num_ids_);
      b.node_counter = alloc(LG_COUNTER_SLOTS);
      b.edge_counter = alloc(LG_COUNTER_SLOTS);
      b.num_ids = num_ids_;
      b.feature_rows = 0;
      LGCHECK(lg_stream_create(&stage_stream_[i]));
      LGCHECK(lg_event_create(&sampled_ev_[i]));
      LGCHECK(lg_event_create(&copied_ev_[i]));
    }

    for (int i = 0; i < interbatch_concurrency_; i++) {
      lg_batch* b = memorypool_->Batch(i);
      std::memset(b, 0, sizeof(*b));
      b->ids = env->GetIds(local_dev_id_, i);
      b->labels = env->GetLabels(local_dev_id_, i);
      b->agg_src = env->GetAggSrc(local_dev_id_, i);
      b->agg_dst = env->GetAggDst(local_dev_id_, i);
      b->node_counter = env->GetNodeCounter(local_dev_id_, i);
      b->edge_counter = env->GetEdgeCounter(local_dev_id_, i);
      b->num_ids = num_ids_;
      b->feature_rows = 0;
    }
    status_host_.assign(interbatch_concurrency_, nullptr);
    status_ev_.resize(interbatch_concurrency_);
    for (int i = 0; i < interbatch_concurrency_; i++) {
      void* h = nullptr;
      LGCHECK(lg_host_alloc_mapped(&h, nullptr, sizeof(int32_t)));
      status_host_[i] = (int32_t*)h;
      *status_host_[i] = 0;
      LGCHECK(lg_event_create(&status_ev_[i]));
    }
    events_.resize(interbatch_concurrency_);
    for (auto& ev : events_) {
      ev.resize(op_num_);
      for (auto& e : ev) LGCHECK(lg_event_create(&e));
    }
    op_params_.resize(op_num_);
    for (int i = 0; i < op_num_; i++) {
      auto* op = new OpParams();
      op->device_id = local_dev_id_;
      op->stream = streams_[0][i % INTRABATCH_CON];
      op->event = events_[0][i];
      op->memorypool = memorypool_;
      op->cache = cache;
      op->graph = params->graph;
      op->feature = feature;
      op->env = env;
      op->in_memory = params->in_memory;
      op->hop_num = hop_num;
      op->neighbor_count = 0;
      op->is_presc = false;
      op_params_[i] = op;
    }
    for (int i = 0; i < hop_num; i++) op_params_[INTRABATCH_CON * i + INTRABATCH_CON]->neighbor_count = params->fanout[i];
  }

  void InitializeFeaturesBuffer(RunnerParams* params) override {
    auto* cache = (UnifiedCache*)params->cache;
    auto* env = (IPCEnv*)params->env;
    LGCHECK(lg_set_device(local_dev_id_));
    int64_t rows = (int64_t)(cache->MaxIdNum(local_dev_id_) * 1.2);  // server.cu:277
    if (max_batch_ > env->GetRawBatchsize())  // eval batches may be larger than anything presampling saw
      rows = rows * max_batch_ / env->GetRawBatchsize() + 1;
    if (rows > num_ids_) rows = num_ids_;
    if (rows < 1) rows = 1;
    env->InitializeFeaturesBuffer(0, (int32_t)rows, float_feature_len_, local_dev_id_, interbatch_concurrency_);
    for (int i = 0; i < interbatch_concurrency_; i++) {
      memorypool_->Batch(i)->features = env->GetFloatFeatures(local_dev_id_, i);
      memorypool_->Batch(i)->feature_rows = rows;
    }
  }

  void RunPreSc(RunnerParams* params) override {
    LGCHECK(lg_set_device(local_dev_id_));
    memorypool_->SetCurrentMode(TRAINMODE);
    memorypool_->SetIter(params->global_batch_id);
    memorypool_->SetGlobalBatchId(params->global_batch_id);
    memorypool_->SetCurrentPipe(0);
    for (int i = 0; i < op_num_; i += INTRABATCH_CON) {  // ops 0,3,6,..: all on stream 0
      op_params_[i]->is_presc = true;
      op_params_[i]->stream = streams_[0][i % INTRABATCH_CON];
      op_params_[i]->event = events_[0][i];
      ops_[i]->run(op_params_[i]);
    }
    LGCHECK(lg_event_synchronize(events_[0][op_num_ - 1]));
  }

  // RunOnce(i) launches batch i and then completes batch i-1 (host-level software pipeline over the
  // INTERBATCH_CON slots): the gather tail of batch i-1 (stream 1) overlaps the sampling of batch i
  // (stream 0).  The trainer still sees batches strictly in order, one semaphore post per batch.
  // ---- staged sampling: three batches in flight over the wire's two slots ----
  // The trainer owns a slot until it calls synchronize(); the reference (and RunOnce below) can therefore only have two
  // batches in flight.  But only the GATHER needs the slot's feature buffer: the sampling chain of batch g runs into a
  // server-private staging set (ids, labels, COO, counters: 27 MB) without waiting for anybody; when the slot of batch g-1
  // is free, one copy kernel (lg_batch_publish) publishes its staging set into the slot and the gather, the CSC builder and the
  // counter / status copies follow on the slot's streams; batch g-2 is handed over.  Per iteration:
  //   A(g)   sample batch g into stage g % S                       (stream of the stage, sampler of the stage)
  //   B(g-1) IPCWait(slot); stage -> slot copies; lookup ops, ...  (streams of the slot)
  //   C(g-2) join, IPCPost
  // The trainer sees the same buffers, the same counters and the same order.  LEGION_STAGING=0 disables it.
  void StageA(IPCEnv* env, int32_t g) {
    const int s = g % n_stage_;
    memorypool_->SetCurrentMode(env->GetCurrentMode(g));
    memorypool_->SetIter(env->GetLocalBatchId(g));
    memorypool_->SetGlobalBatchId(g);
    memorypool_->SetCurrentSampler(s);
    memorypool_->SetBatchView(&stage_b_[s]);
    lg_stream_t st = stage_stream_[s];
    LGCHECK(lg_stream_wait_event(st, copied_ev_[s]));  // the stage's previous batch has been published (no-op the first time)
    for (int i = 0; i < op_num_; i += INTRABATCH_CON) {  // ops 0, 3, 6, ..: batch_generate, the sampling ops, io_complete
      op_params_[i]->is_presc = false;
      op_params_[i]->stream = st;
      op_params_[i]->event = nullptr;  // nothing waits for a single op of the chain: no event between its kernels, so
                                       // they stay chained by programmatic dependent launch; the stage's event follows
      ops_[i]->run(op_params_[i]);
    }
    LGCHECK(lg_event_record(sampled_ev_[s], st));
    memorypool_->SetBatchView(nullptr);
  }
  void StageB(IPCEnv* env, int32_t g) {
    const int s = g % n_stage_, pipe = g % interbatch_concurrency_;
    env->IPCWait(local_dev_id_, pipe);
    memorypool_->SetCurrentPipe(pipe);
    memorypool_->SetCurrentSampler(s);
    memorypool_->SetBatchView(nullptr);
    auto& st = streams_[pipe];
    auto& ev = events_[pipe];
    const lg_batch& from = stage_b_[s];
    lg_batch* to = memorypool_->Batch(pipe);
    LGCHECK(lg_stream_wait_event(st[0], sampled_ev_[s]));
    LGCHECK(lg_batch_publish(st[0], &from, to));  // the used part of ids / labels / COO / counters, one launch
    LGCHECK(lg_event_record(copied_ev_[s], st[0]));
    for (int i = 0; i < op_num_; i += INTRABATCH_CON) LGCHECK(lg_event_record(ev[i], st[0]));  // "the sampling op of the hop is done"
    for (int i = 0; i < op_num_; i++) {
      if (i % INTRABATCH_CON == 0) continue;  // ran in stage A
      LGCHECK(lg_stream_wait_event(st[i % INTRABATCH_CON], ev[i / INTRABATCH_CON * INTRABATCH_CON]));
      op_params_[i]->is_presc = false;
      op_params_[i]->stream = st[i % INTRABATCH_CON];
      op_params_[i]->event = ev[i];
      ops_[i]->run(op_params_[i]);
    }
    AfterOps(env, pipe);
  }
  void RunOnceStaged(RunnerParams* params) {
    auto* env = (IPCEnv*)params->env;
    const int32_t g = params->global_batch_id, last = env->GetMaxStep() - 1;
    StageA(env, g);
    if (g >= 1) StageB(env, g - 1);
    if (g >= 2) Complete(env, (g - 2) % interbatch_concurrency_, g - 2);
    if (g == last) {  // drain
      StageB(env, g);
      if (g >= 1) Complete(env, (g - 1) % interbatch_concurrency_, g - 1);
      Complete(env, g % interbatch_concurrency_, g);
    }
    current_pipe_ = (g + 1) % interbatch_concurrency_;
    memorypool_->SetCurrentSampler(-1);
  }

  // what follows a batch's ops on the slot's streams: blocks as CSC, counters to the host side channel, status
  void AfterOps(IPCEnv* env, int pipe) {
    auto& ev = events_[pipe];
    auto& st = streams_[slot_samplers_ ? pipe : 0];
    if (emit_csc_) {  // the third stream is idle in the reference (SSDIOSubmit is a no-op): the blocks are built there, behind
                      // the last sampling op, while the gather streams on the second
      LGCHECK(lg_stream_wait_event(st[2], ev[op_num_ - 1 - INTRABATCH_CON]));
      const int hops = (int)csc_max_edges_.size();
      int32_t** ptrs = &csc_ptrs_[(size_t)pipe * 3 * hops];  // [indptr | indices | eids][block]
      LGCHECK(lg_block_csc_batch(st[2], memorypool_->Batch(pipe), hops, csc_max_edges_.data(), csc_max_dst32_.data(),
                                 ptrs, ptrs + hops, ptrs + 2 * hops, csc_workspace_[pipe], csc_workspace_bytes_));
      LGCHECK(lg_event_record(csc_ev_[pipe], st[2]));
    }
    // the side channel (include/legion_b200_ext.h): the batch's counters go to host memory behind its last lookup (which
    // waited for the last sampling op), so that the trainer's get_next needs no device copy of its own
    if (int32_t* hc = env->GetHostCounters(local_dev_id_, pipe)) {
      const lg_batch* b = memorypool_->Batch(pipe);
      LGCHECK(lg_memcpy_d2h(hc, b->node_counter, LG_COUNTER_SLOTS * sizeof(int32_t), st[1]));
      LGCHECK(lg_memcpy_d2h(hc + LG_COUNTER_SLOTS, b->edge_counter, LG_COUNTER_SLOTS * sizeof(int32_t), st[1]));
    }
    // the sticky overflow flag as of this batch's last lookup: an asynchronous copy behind it (a synchronous read on any
    // of the three streams would wait for the batch just launched, not for the one being handed off)
    LGCHECK(lg_sampler_status_async(memorypool_->Sampler(), st[1], status_host_[pipe]));
    LGCHECK(lg_event_record(status_ev_[pipe], st[1]));
  }

  void RunOnce(RunnerParams* params) override {
    LGCHECK(lg_set_device(local_dev_id_));
    if (n_stage_ > 0) return RunOnceStaged(params);
    auto* env = (IPCEnv*)params->env;
    int32_t batch_id = params->global_batch_id;
    mode_ = env->GetCurrentMode(batch_id);
    memorypool_->SetCurrentMode(mode_);
    memorypool_->SetIter(env->GetLocalBatchId(batch_id));
    memorypool_->SetGlobalBatchId(batch_id);
    memorypool_->SetCurrentPipe(current_pipe_);
    env->IPCWait(local_dev_id_, current_pipe_);
    auto& ev = events_[current_pipe_];
    auto& st = streams_[slot_samplers_ ? current_pipe_ : 0];
    for (int i = 0; i < op_num_; i++) {
      if (i % INTRABATCH_CON >= 1)  // lookup / io ops wait for the sampling op of their hop (server.cu:312-314)
        LGCHECK(lg_stream_wait_event(st[i % INTRABATCH_CON], ev[i / INTRABATCH_CON * INTRABATCH_CON]));
      op_params_[i]->is_presc = false;
      op_params_[i]->stream = st[i % INTRABATCH_CON];
      op_params_[i]->event = ev[i];
      ops_[i]->run(op_params_[i]);
    }
    AfterOps(env, current_pipe_);
    if (pending_pipe_ >= 0) Complete(env, pending_pipe_, pending_batch_);
    pending_pipe_ = current_pipe_;
    pending_batch_ = batch_id;
    current_pipe_ = (current_pipe_ + 1) % interbatch_concurrency_;
    if (batch_id == env->GetMaxStep() - 1) {  // last batch: nothing left to overlap with, hand it off from this thread
      Complete(env, pending_pipe_, pending_batch_);
      pending_pipe_ = -1;
    }
  }

  void Complete(IPCEnv* env, int pipe, int32_t batch_id) {
    auto& ev = events_[pipe];
    for (int i = op_num_ - INTRABATCH_CON; i < op_num_; i++)  // join all three streams before the hand-off
      LGCHECK(lg_event_synchronize(ev[i]));
    LGCHECK(lg_event_synchronize(status_ev_[pipe]));
    if (emit_csc_) LGCHECK(lg_event_synchronize(csc_ev_[pipe]));
    const int32_t st = *status_host_[pipe];
    if (st != 0) {
      std::fprintf(stderr, "batch %d on GPU %d overflowed its buffers (status %d)\n", batch_id, local_dev_id_, st);
      std::exit(EXIT_FAILURE);
    }
    env->IPCPost(local_dev_id_, pipe);
    if (batch_id % 1000 == 0 && local_dev_id_ == 0) std::cout << "batch id: " << batch_id << "\n" << std::flush;
  }

  void Finalize(RunnerParams* params) override {
    auto* env = (IPCEnv*)params->env;
    LGCHECK(lg_set_device(local_dev_id_));
    if (pending_pipe_ >= 0) Complete(env, pending_pipe_, pending_batch_);
    pending_pipe_ = -1;
    env->IPCWait(local_dev_id_, (current_pipe_ + 1) % interbatch_concurrency_);  // server.cu:336
    LGCHECK(lg_set_device(local_dev_id_));
    for (auto* smp : memorypool_->samplers) lg_sampler_destroy(smp);
    for (auto& ev : events_)
      for (auto& e : ev) lg_event_destroy(e);
    for (auto& e : status_ev_) lg_event_destroy(e);
    for (auto* h : status_host_) lg_host_free(h);
    for (auto& ss : streams_)
      for (auto& s : ss) lg_stream_destroy(s);
  }

 private:
  int32_t num_ids_ = 0, float_feature_len_ = 0, max_batch_ = 0;
  MemoryPool* memorypool_ = nullptr;
  int current_pipe_ = 0, interbatch_concurrency_ = INTERBATCH_CON, local_dev_id_ = 0, mode_ = 0, op_num_ = 0;
  bool slot_samplers_ = true;
  bool emit_csc_ = false;
  int n_stage_ = 0;  // staged sampling: number of staging sets (0 = off)
  std::vector<lg_batch> stage_b_;
  std::vector<lg_stream_t> stage_stream_;
  std::vector<lg_event_t> sampled_ev_, copied_ev_;
  std::vector<int64_t> csc_max_edges_, csc_max_dst_;
  std::vector<int32_t> csc_max_dst32_;
  std::vector<int32_t*> csc_ptrs_;  // [slot][indptr | indices | eids][block]
  std::vector<void*> csc_workspace_;
  int64_t csc_workspace_bytes_ = 0;
  std::vector<lg_event_t> csc_ev_;
  std::vector<std::vector<lg_stream_t>> streams_;  // [slot][INTRABATCH_CON]
  std::vector<std::vector<lg_event_t>> events_;  // [slot][op]
  int pending_pipe_ = -1;
  int32_t pending_batch_ = 0;
  std::vector<int32_t*> status_host_;  // [slot] pinned copy of the sampler's sticky status
  std::vector<lg_event_t> status_ev_;
  std::vector<Operator*> ops_;
  std::vector<OpParams*> op_params_;
};
}  // namespace

Server* NewGPUServer() { return new GPUServer(); }
Runner* NewGPURunner() { return new GPURunner(); }

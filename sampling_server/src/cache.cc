// cache.cc — presampling statistics -> ranking -> cost model -> interleaved fill, through the C ABI.
// Flow and arithmetic follow reference cache/cache.cu:295-611; what differs is mechanical: dense
// directories instead of bcht maps, one descriptor struct per GPU instead of device pointer tables.
#include "cache.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

static void* DevAlloc(int64_t bytes, bool zero) {
  void* p = nullptr;
  LGCHECK(lg_device_alloc(&p, bytes));
  if (zero) LGCHECK(lg_memset_async(p, 0, bytes, nullptr));
  return p;
}

void UnifiedCache::Initialize(int64_t cache_memory, int32_t float_feature_len, int32_t train_step, int32_t device_count) {
  cache_memory_ = cache_memory;
  float_feature_len_ = float_feature_len;
  train_step_ = train_step;
  device_count_ = device_count;
  node_access_.assign(device_count, nullptr);
  edge_access_.assign(device_count, nullptr);
  max_ids_dev_.assign(device_count, nullptr);
  tier_rows_.assign(device_count, nullptr);
  fcache_.resize(device_count);
  for (auto& c : fcache_) std::memset(&c, 0, sizeof(c));
  is_presc_ = true;
}

void UnifiedCache::InitializeCacheController(int32_t dev, int32_t total_num_nodes) {
  LGCHECK(lg_set_device(dev));
  total_num_nodes_ = total_num_nodes;
  node_access_[dev] = (unsigned long long*)DevAlloc((int64_t)total_num_nodes * 8, true);  // cache.cu:21-25
  edge_access_[dev] = (unsigned long long*)DevAlloc((int64_t)total_num_nodes * 8, true);
  max_ids_dev_[dev] = (int32_t*)DevAlloc(4, true);
  tier_rows_[dev] = (unsigned long long*)DevAlloc(3 * 8, true);
  LGCHECK(lg_stream_synchronize(nullptr));
}

int32_t UnifiedCache::MaxIdNum(int32_t dev) {
  LGCHECK(lg_set_device(dev));
  int32_t v = 0;
  LGCHECK(lg_memcpy_d2h(&v, max_ids_dev_[dev], 4, nullptr));
  LGCHECK(lg_stream_synchronize(nullptr));
  return v;
}

void UnifiedCache::ReadTierRows(int32_t dev, unsigned long long out[3]) {
  LGCHECK(lg_set_device(dev));
  LGCHECK(lg_memcpy_d2h(out, tier_rows_[dev], 3 * 8, nullptr));
  LGCHECK(lg_stream_synchronize(nullptr));
}

void UnifiedCache::CandidateSelection(int cache_agg_mode, FeatureStorage* feature, GraphStorage*) {
  Kg_ = 1 << cache_agg_mode;  // cache.cu:375-389
  if (Kg_ > device_count_) Kg_ = device_count_;
  Kc_ = device_count_ / Kg_;
  const int64_t n = feature->TotalNodeNum();
  for (int32_t i = 0; i < Kc_; i++) {
    LGCHECK(lg_set_device(i * Kg_));
    int64_t tmp_bytes = 0;
    LGCHECK(lg_hotness_rank(nullptr, nullptr, n, nullptr, nullptr, nullptr, &tmp_bytes));
    void* tmp = DevAlloc(tmp_bytes, false);
    for (int pass = 0; pass < 2; pass++) {  // features ranking (QF), then topology ranking (QT)
      auto* agg = (unsigned long long*)DevAlloc(n * 8, true);
      for (int32_t j = 0; j < Kg_; j++)  // peer reads of the other GPUs' counters (:408-411,428-431)
        LGCHECK(lg_hotness_accumulate(nullptr, agg, pass == 0 ? node_access_[i * Kg_ + j] : edge_access_[i * Kg_ + j], n));
      auto* order = (int32_t*)DevAlloc(n * 4, false);
      auto* sorted = (unsigned long long*)DevAlloc(n * 8, false);
      LGCHECK(lg_hotness_rank(nullptr, agg, n, order, sorted, tmp, &tmp_bytes));
      LGCHECK(lg_stream_synchronize(nullptr));
      lg_device_free(agg);
      (pass == 0 ? QF_ : QT_).push_back(order);
      (pass == 0 ? AF_ : AT_).push_back(sorted);
    }
    lg_device_free(tmp);
  }
  is_presc_ = false;
}

void UnifiedCache::CostModel(int, FeatureStorage* feature, GraphStorage* graph, std::vector<uint64_t>& counters,
                             int32_t train_step) {
  const int64_t n = feature->TotalNodeNum();
  for (int32_t i = 0; i < Kc_; i++) {
    LGCHECK(lg_set_device(i * Kg_));
    uint64_t topo_trans = counters[0] + counters[1];  // PCM counters: {0,0} in the reference (server.cu:106)
    uint64_t feat_trans = 0;
    for (int j = 0; j < Kg_; j++)  // cache.cu:461-463 (the reference indexes controllers 0..Kg-1 for every clique)
      feat_trans += (uint64_t)((int64_t)MaxIdNum(j) * train_step * float_feature_len_ * 4 / 64);
    LGCHECK(lg_set_device(i * Kg_));
    std::vector<unsigned long long> af(n), at(n);
    std::vector<int32_t> qt(n);
    LGCHECK(lg_memcpy_d2h(af.data(), AF_[i], n * 8, nullptr));
    LGCHECK(lg_memcpy_d2h(at.data(), AT_[i], n * 8, nullptr));
    LGCHECK(lg_memcpy_d2h(qt.data(), QT_[i], n * 4, nullptr));
    LGCHECK(lg_stream_synchronize(nullptr));
    // The reference feeds Intel-PCM PCIe counters here and they are disabled ({0,0}, server.cu:106), which removes the
    // topology term altogether.  Their stand-in: the presampling epoch issued one PCIe read transaction per sampled
    // neighbour (a 4-byte read still costs a whole transaction) = the sum of the edge hotness counters.
    if (topo_trans == 0)
      for (int64_t v = 0; v < n; v++) topo_trans += at[v];
    int32_t ncap = 0, ecap = 0;
    double alpha = 0;
    const char* cm = std::getenv("LEGION_COSTMODEL");
    if (cm && std::strcmp(cm, "reference") == 0)  // the reference's rule verbatim: capacity 0 once everything fits
      LGCHECK(lg_cost_model(af.data(), at.data(), qt.data(), graph->HostIndptr(), n, float_feature_len_, cache_memory_, Kg_,
                            counters[0] + counters[1], feat_trans, &ncap, &ecap, &alpha));
    else
      LGCHECK(lg_cost_model_saturating(af.data(), at.data(), qt.data(), graph->HostIndptr(), n, float_feature_len_,
                                       cache_memory_, Kg_, topo_trans, feat_trans, &ncap, &ecap, &alpha));
    std::cout << "Alpha: " << alpha << " on Clique: " << i << std::endl;
    std::cout << "Feat capacity: " << ncap - 1 << " Topo capacity: " << ecap - 1 << " on Clique: " << i << std::endl;
    node_capacity_.push_back(ncap);
    edge_capacity_.push_back(ecap);
  }
}

// LEGION_REPLICATE_RATIO=r (default 0 = the reference placement): the first r * node_capacity rows of every feature
// shard hold the hottest ranks on EVERY GPU of the clique (hybrid placement, lg_place_features_hybrid); only the ranks
// after them are interleaved.  Same shard size, fewer distinct rows cached, local reads for the head of the order.
static int32_t ReplicatedRows(int32_t ncap, int32_t kg) {
  const char* e = std::getenv("LEGION_REPLICATE_RATIO");
  if (!e || kg <= 1) return 0;
  double r = std::atof(e);
  if (r <= 0) return 0;
  if (r > 1) r = 1;
  return (int32_t)(r * ncap);
}

// A cache that holds EVERYTHING on one GPU (Kg = 1 and the capacity chosen by the cost model covers every vertex — the
// normal case with 180 GB of HBM) needs no lookup structure at all: feature rows are stored at row index = vertex id
// (LG_CACHE_IDENTITY) and the CSR is simply resident in HBM (slot P of the pointer table, no topology directory).  The
// reference still builds and probes its three hash maps in that case (cache/cache.cu:80-136,180-225).
// LEGION_IDENTITY=0 keeps the rank-ordered shards + directories.
static bool IdentityEnabled() {
  const char* e = std::getenv("LEGION_IDENTITY");
  return !(e && std::atoi(e) == 0);
}

void UnifiedCache::FillUp(int, FeatureStorage* feature, GraphStorage* graph) {
  const int64_t n = feature->TotalNodeNum();
  const int32_t dim = feature->GetFloatFeatureLen();
  for (int32_t i = 0; i < Kc_; i++) {
    const int32_t ncap = node_capacity_[i], ecap = edge_capacity_[i];
    const int32_t rep = ReplicatedRows(ncap, Kg_);
    if (Kg_ == 1 && IdentityEnabled() && ncap >= n && ecap >= n) {
      const int32_t dev = i;
      LGCHECK(lg_set_device(dev));
      const int64_t n_edges = graph->HostIndptr()[n];
      auto* rows = (float*)DevAlloc(n * dim * 4, false);
      LGCHECK(lg_memcpy_d2d(rows, feature->GetAllFloatFeature(), n * dim * 4, nullptr));
      auto* ip = (int64_t*)DevAlloc((n + 1) * 8, false);
      auto* ix = (int32_t*)DevAlloc((n_edges > 0 ? n_edges : 1) * 4, false);
      LGCHECK(lg_memcpy_d2d(ip, graph->GetCSRNodeIndexCPU(), (n + 1) * 8, nullptr));
      LGCHECK(lg_memcpy_d2d(ix, graph->GetCSRNodeMatrixCPU(), n_edges * 4, nullptr));
      LGCHECK(lg_stream_synchronize(nullptr));
      lg_feature_cache& c = fcache_[dev];
      c.n_parts = 1;
      c.shard_rows = (int32_t)n;
      c.dim = dim;
      c.flags = LG_CACHE_IDENTITY;
      c.num_nodes = n;
      c.shard[0] = rows;
      c.backing = feature->GetAllFloatFeature();
      c.directory = nullptr;
      lg_topology* t = graph->Topology(dev);
      t->n_parts = 0;
      t->shard_rows = 0;
      t->indptr[0] = ip;
      t->indices[0] = ix;
      t->directory = nullptr;
      std::cout << "Everything fits GPU " << dev << ": rows and CSR resident in HBM, no lookup directories" << std::endl;
      continue;
    }
    if (rep > 0) std::cout << "Replicated feature rows: " << rep << " of " << ncap << " per GPU on Clique: " << i << std::endl;
    std::vector<float*> shard(Kg_);
    std::vector<int64_t*> sip(Kg_);
    std::vector<int32_t*> six(Kg_);
    for (int32_t j = 0; j < Kg_; j++) {  // shards: FeatFillUp + GraphCache (cache.cu:584-608)
      const int32_t dev = i * Kg_ + j;
      LGCHECK(lg_set_device(dev));
      shard[j] = (float*)DevAlloc((int64_t)ncap * dim * 4, false);
      LGCHECK(lg_fill_feature_shard_hybrid(nullptr, QF_[i], ncap, Kg_, rep, j, dim, n, feature->GetAllFloatFeature(), shard[j]));
      sip[j] = (int64_t*)DevAlloc((int64_t)(ecap + 1) * 8, false);
      LGCHECK(lg_topo_shard_indptr(nullptr, QT_[i], ecap, Kg_, j, n, graph->GetCSRNodeIndexCPU(), sip[j]));
      int64_t total = 0;
      LGCHECK(lg_memcpy_d2h(&total, sip[j] + ecap, 8, nullptr));
      LGCHECK(lg_stream_synchronize(nullptr));
      six[j] = (int32_t*)DevAlloc((total > 0 ? total : 1) * 4, false);
      LGCHECK(lg_topo_shard_fill(nullptr, QT_[i], ecap, Kg_, j, n, graph->GetCSRNodeIndexCPU(),
                                 graph->GetCSRNodeMatrixCPU(), sip[j], six[j]));
      LGCHECK(lg_stream_synchronize(nullptr));
    }
    for (int32_t j = 0; j < Kg_; j++) {  // directories + descriptors, one per GPU (cache.cu:565-569,572-602)
      const int32_t dev = i * Kg_ + j;
      LGCHECK(lg_set_device(dev));
      auto* fdir = (int32_t*)DevAlloc(n * 4, false);
      LGCHECK(lg_fill_i32(nullptr, fdir, CACHEMISS_FLAG, n));
      LGCHECK(lg_place_features_hybrid(nullptr, QF_[i], ncap, Kg_, rep, j, n, fdir));
      auto* tdir = (int32_t*)DevAlloc(n * 4, false);
      LGCHECK(lg_fill_i32(nullptr, tdir, CACHEMISS_FLAG, n));
      LGCHECK(lg_place_topology(nullptr, QT_[i], ecap, Kg_, 0, n, tdir));  // parts are clique-local slots
      LGCHECK(lg_stream_synchronize(nullptr));
      lg_feature_cache& c = fcache_[dev];
      c.n_parts = Kg_;
      c.shard_rows = ncap;
      c.dim = dim;
      c.num_nodes = n;
      for (int32_t k = 0; k < Kg_; k++) c.shard[k] = shard[k];
      c.backing = feature->GetAllFloatFeature();
      c.directory = fdir;
      lg_topology* t = graph->Topology(dev);
      const int64_t* full_ip = t->indptr[0];
      const int32_t* full_ix = t->indices[0];
      t->n_parts = Kg_;
      t->shard_rows = ecap;
      for (int32_t k = 0; k < Kg_; k++) {
        t->indptr[k] = sip[k];
        t->indices[k] = six[k];
      }
      t->indptr[Kg_] = full_ip;
      t->indices[Kg_] = full_ix;
      t->directory = tdir;
    }
  }
}

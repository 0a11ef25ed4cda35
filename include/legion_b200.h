/*
 * legion_b200.h — C ABI of liblegion_b200.so: Legion's mini-batch data path
 * (k-hop neighbour sampling + unified-cache lookup + feature gather) for B200.
 *
 * This is the drop-in boundary.  Every entry point below replaces one piece of the
 * reference's operator FFI (the five `extern "C"` op bodies declared in
 * sampling_server/src/engine/operator_impl.cuh:11-63) or of the cache / storage code
 * those bodies call.  Where the reference passes C++ object pointers (MemoryPool*,
 * UnifiedCache*, GraphStorage*) this ABI passes flat descriptors of raw device
 * pointers and sizes, so it can be bound from C++, ctypes, cgo, JNI …
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from lg_last_error() (thread local).  The reference exits the process on
 *     any CUDA error (engine/operator_impl.cu:16-24); the host layer above this ABI keeps
 *     that behaviour, the ABI itself never calls exit().
 *   - `stream` is a cudaStream_t passed as void*.  Everything is enqueued on it; no call
 *     synchronises the device unless its comment says so (the reference performs >= 5
 *     blocking cudaMemcpy per hop: engine/operator_impl.cu:439-445, cache/cache.cu:187-188).
 *   - the caller owns every buffer it passes in; handles own only their private scratch.
 *   - one host thread per GPU drives one lg_sampler (engine/server.cu:122-130).
 */
#ifndef LEGION_B200_H_
#define LEGION_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants shared with the reference wire format (include/system_config.cuh:47-57) ---- */
#define LG_INTERBATCH_CON 2   /* pipeline slots per GPU                     */
#define LG_INTRABATCH_CON 3   /* ops per hop (sample, lookup, io)           */
#define LG_MAX_DEVICE 8
#define LG_MEMORY_USAGE 7     /* IPC buffers per (gpu, slot)                */
#define LG_COUNTER_SLOTS 16   /* int32 node_counter[16], edge_counter[16]   */
#define LG_CACHEMISS_FLAG (-2)
#define LG_MAX_HOPS 6         /* nc[9+h] must stay inside 16 slots          */
#define LG_TRAINMODE 0
#define LG_VALIDMODE 1
#define LG_TESTMODE 2

/* random streams for the neighbour pick */
#define LG_RNG_MINSTD 0 /* thrust::minstd_rand seed 1, discard(slot) — the reference stream
                           (engine/operator_impl.cu:235-238)                              */
#define LG_RNG_PHILOX 1 /* Philox4x32-10, key=(seed lo,hi), ctr=(slot, hop, batch, stream) */

/* data movers of the gather */
#define LG_GATHER_AUTO 0
#define LG_GATHER_LDG 1 /* warp-per-row 16-byte ld.global.nc / st.global */
#define LG_GATHER_TMA 2 /* cp.async.bulk row loads into smem + bulk tile store */

typedef void* lg_stream_t;
typedef struct lg_sampler lg_sampler;

/* ---- descriptors (plain structs of device pointers; passed by const pointer, copied) ---- */

/* CSR pointer tables: replaces GraphStorage's per-GPU table of P+1 pointer pairs
 * (storage/graph_storage.cu:30-33,60-62,162-167).  Slots 0..n_parts-1 are the cached CSR
 * shards of the NVLink clique (local or peer HBM); slot n_parts is the full CSR (host
 * pinned via UVA, or HBM when the whole graph is resident).  `directory` replaces the two
 * bcht maps queried by UnifiedCache::FindTopo (cache/cache.cu:217-225): entry v is
 * part*shard_rows + row, or LG_CACHEMISS_FLAG.  NULL directory = every vertex misses. */
typedef struct lg_topology {
  int32_t n_parts;
  int32_t shard_rows; /* edge_capacity_ (rows per shard) */
  int64_t num_nodes;
  const int64_t* indptr[LG_MAX_DEVICE + 1];
  const int32_t* indices[LG_MAX_DEVICE + 1];
  const int32_t* directory;
} lg_topology;

/* Feature cache: replaces float** gpu_float_feature + cpu_float_features + node_map_
 * (cache/cache.cu:572-602, cache/cache_impl.cuh:239-272).  directory[v] = gidx =
 * part*shard_rows + row (cache_impl.cuh:104-109) or LG_CACHEMISS_FLAG. */
/* lg_feature_cache.flags.  LG_CACHE_IDENTITY: shard[local_part] holds EVERY vertex's row at row index = vertex id (a cache
 * at least as large as the dataset on one GPU — the normal case with 180 GB of HBM); the lookup needs no directory (the
 * reference's FindFeat probe, cache/cache.cu:180-215, and the 64-byte DRAM granule it costs per row disappear) and every
 * row counts as a local hit.  `directory` must be NULL.  Gathered features are the same bits either way. */
#define LG_CACHE_IDENTITY 1

typedef struct lg_feature_cache {
  int32_t n_parts;
  int32_t shard_rows; /* node_capacity_ (rows per shard) */
  int32_t dim;        /* float_feature_len */
  int32_t flags;      /* LG_CACHE_* */
  int64_t num_nodes;
  const float* shard[LG_MAX_DEVICE];
  const float* backing;     /* full [num_nodes x dim] matrix: host UVA pointer or HBM; may be NULL when the
                               directory covers every vertex (a miss then sets status 3) */
  const int32_t* directory; /* NULL = every row misses */
} lg_feature_cache;

/* The seven per-(gpu, slot) buffers the trainer opens over CUDA IPC
 * (engine/ipc_service.cu:134-206; training_backend/ipc_cuda_kernel.cu:62-68). */
typedef struct lg_batch {
  int32_t* ids;          /* [num_ids]   global ids, seeds first                  */
  float* features;       /* [feature_rows x dim]                                 */
  int32_t* labels;       /* [batch]                                              */
  int32_t* agg_src;      /* [num_ids]   batch-local index of sampled neighbour   */
  int32_t* agg_dst;      /* [num_ids]   batch-local index of the frontier vertex */
  int32_t* node_counter; /* [16]        protocol: SURVEY 3.3 / operator_impl.cu:57-89 */
  int32_t* edge_counter; /* [16]                                                 */
  int64_t feature_rows;  /* capacity of `features` in rows (reference: 1.2 x presampled max, unchecked) */
  int32_t num_ids;       /* capacity of ids / agg_src / agg_dst                  */
  int32_t reserved;
} lg_batch;

/* ---- library ---- */
const char* lg_last_error(void);
int lg_version(void);
/* bytes / counts a host needs to size things without guessing */
int64_t lg_num_ids(int32_t batch_size, const int32_t* fanout, int32_t n_hops); /* engine/server.cu:187-199 */

/* ---- sampler handle: private scratch (position map, scan state, global-id frontier) ----
 * num_nodes sizes the position map: one 32-bit word per vertex, the reference's position_map
 * (engine/server.cu:224; 4*num_nodes bytes per handle).  Words are claimed by lg_batch_generate /
 * lg_random_sample and released by lg_io_complete (ClearPosMap, engine/operator_impl.cu:542-548); a batch
 * that never reaches lg_io_complete is released by the next lg_batch_generate on the same handle. */
int lg_sampler_create(int32_t device, int32_t max_batch, const int32_t* fanout, int32_t n_hops,
                      int64_t num_nodes, lg_sampler** out);
int lg_sampler_destroy(lg_sampler* s);
/* full reset of the position map and the sticky status (after an overflow status or an aborted batch) */
int lg_sampler_reset(lg_sampler* s, lg_stream_t stream);
int64_t lg_sampler_scratch_bytes(const lg_sampler* s);
/* layout of the position map chosen for this handle: 0 = dense (one 32-bit word per vertex, the reference's
 * position_map), 1 = hashed (O(batch) L2-resident table; picked when 4*num_nodes bytes would not stay in L2).
 * Environment: LG_DEDUP=dense|hash forces one, LG_DENSE_MAX_MB moves the threshold (default 48). */
int32_t lg_sampler_dedup_layout(const lg_sampler* s);
/* data mover used by lg_feature_cache_lookup: LG_GATHER_AUTO / LG_GATHER_LDG / LG_GATHER_TMA */
int lg_sampler_set_gather_variant(lg_sampler* s, int32_t variant);
/* how many gather launches lg_run_batch issues: 0 = one per lookup op (the reference's schedule);
 * 1 (default) = the seeds' rows are moved together with hop 1's; 2 = one gather of all rows after the
 * last hop.  Results (features, all counter slots) are identical in every mode. */
int lg_sampler_set_gather_fusion(lg_sampler* s, int32_t mode);
/* how lg_run_batch schedules the gathers: 0 = everything on the caller's stream; 1 (default) = gathers
 * on an internal side stream, overlapping the sampling of the next hop, joined before returning
 * (the reference's stream split, engine/server.cu:311-317); 2 = pipelined across batches like the
 * reference's two INTERBATCH_CON slots: the last gather of batch k overlaps the sampling of batch
 * k+1; a consumer (or anyone reusing the buffers on another stream) calls lg_batch_wait first. */
int lg_sampler_set_overlap(lg_sampler* s, int32_t mode);
/* Lazy construct_graph for the op-by-op calls (default 0 = every lg_random_sample leaves its hop complete, like the
 * reference's RandomSample op, engine/operator_impl.cu:400-499).  1 = a host that only hands the batch on after the
 * last op (the server: IPCPost follows the last op) lets hop h's agg_src be finished by hop h+1's sample kernel, and
 * the last hop's relabel kernel also releases the position map (lg_io_complete then has nothing left to launch):
 * two kernel launches less per batch, same final buffers and counters. */
int lg_sampler_set_lazy_relabel(lg_sampler* s, int32_t mode);
/* Seed offset of the clipped tail batch of a set (batch_size * (counter+1) > total_cap).  LG_TAIL_EXACT (default): the
 * batch starts at batch_size * counter — the true tail of the set.  LG_TAIL_REFERENCE: it starts at clipped_size * counter,
 * what the reference computes (engine/operator_impl.cu:159-162 pass the clipped size as the kernel's batch_size, :40,:44
 * multiply it by the counter): with the server's valid/test batch size ceil(n/steps) the last eval batch re-reads seeds
 * from the middle of the set and the real tail is never served.  Full batches are identical in both modes. */
#define LG_TAIL_EXACT 0
#define LG_TAIL_REFERENCE 1
int lg_sampler_set_tail_mode(lg_sampler* s, int32_t mode);
/* make `stream` wait until the batch last produced into `batch` is complete (mode 2; no-op otherwise) */
int lg_batch_wait(lg_sampler* s, lg_stream_t stream, const lg_batch* batch);
/* sticky status (0 ok, 1 ids overflow, 2 features buffer too small — the reference sizes it 1.2 x presampled max without a
 * bound check, engine/server.cu:277 —, 3 cache miss without a backing matrix, 4 edge_dst holds a vertex id outside
 * [0, num_nodes): the reference drops the edge when the id is negative, engine/operator_impl.cu:244, and reads out of
 * bounds otherwise; here the batch is flagged invalid and stays memory-safe).  Synchronises. */
int lg_sampler_status(lg_sampler* s, lg_stream_t stream, int32_t* host_status);
/* the same copy without the synchronisation: `pinned_host_status` (page-locked) holds the flag once the work enqueued on
 * `stream` so far has completed — a pipelined host reads it after the event it waits for anyway */
int lg_sampler_status_async(lg_sampler* s, lg_stream_t stream, int32_t* pinned_host_status);

/* diagnostics (kernel-launch counter, phase traces, spin kernel) live in legion_b200_debug.h, not in this ABI */

/* ---- the five operator bodies (engine/operator_impl.cuh:11-63) ---- */

/* BatchGenerate (engine/operator_impl.cu:27-55,92-172): seeds of batch `counter` =
 * all_ids[batch_size*counter ...], -1 past total_cap; labels; counters reset + op-0 update;
 * seeds enter the position map with local index = position. */
int lg_batch_generate(lg_sampler* s, lg_stream_t stream, const int32_t* all_ids,
                      const int32_t* all_labels, int32_t total_cap, int32_t batch_size,
                      int32_t counter, const lg_batch* batch);

/* RandomSample (engine/operator_impl.cu:175-296,400-499): one hop.  hop is 1-based
 * (op_id = 3*hop).  Slot rule, with-replacement pick, first-seen dedup, local re-indexing
 * (construct_graph) and the op's counter_update are all done on the device; output order is
 * ascending slot index.  edge_hotness (u64[num_nodes], may be NULL) receives the
 * presampling count of engine/operator_impl.cu:358. */
int lg_random_sample(lg_sampler* s, lg_stream_t stream, const lg_topology* topo, int32_t hop,
                     int32_t rng_kind, uint64_t rng_seed, uint32_t batch_id, uint32_t stream_id,
                     const lg_batch* batch, unsigned long long* edge_hotness);

/* FeatureCacheLookup (engine/operator_impl.cu:501-519 -> cache/cache.cu:180-215,726-748 ->
 * cache/cache_impl.cuh:239-272): counter snapshot for op_id, directory lookup (FindFeat) and
 * gather of rows [nc[2], nc[2]+nc[3]) of `ids` into batch->features.  tier_rows (u64[3],
 * may be NULL) accumulates local / peer / miss row counts (hit-mix telemetry). */
int lg_feature_cache_lookup(lg_sampler* s, lg_stream_t stream, const lg_feature_cache* cache,
                            int32_t op_id, int32_t local_part, const lg_batch* batch,
                            unsigned long long* tier_rows);

/* same op, but one launch moves the rows of hops [first_hop, op_id/3] (fused gathers of lg_run_batch);
 * the counter snapshot written is still the one of op_id */
int lg_feature_cache_lookup_range(lg_sampler* s, lg_stream_t stream, const lg_feature_cache* cache,
                                  int32_t op_id, int32_t first_hop, int32_t local_part,
                                  const lg_batch* batch, unsigned long long* tier_rows);

/* IOSubmit is a no-op in the reference (engine/operator_impl.cu:521-539); kept for API parity. */
int lg_io_submit(lg_sampler* s, lg_stream_t stream, int32_t op_id, const lg_batch* batch);

/* Copies the used part of one set of batch buffers into another: ids[0, nc[9+H]), labels[0, nc[9]), agg_src / agg_dst
 * [0, ec[9+H]) and both counter arrays; the sizes are read from `from`'s counters on the device (one launch, no host
 * synchronisation).  For hosts that sample into private staging buffers while the consumer still owns the hand-off
 * buffers (the reference samples straight into them and therefore waits, engine/server.cu:309). */
int lg_batch_publish(lg_stream_t stream, const lg_batch* from, const lg_batch* to);

/* IOComplete (engine/operator_impl.cu:542-580): end-of-batch cleanup (ClearPosMap: the position-map words
 * of the batch's vertices are released, in every mode) and, in train mode when node_hotness != NULL, HotnessMeasure
 * (cache/cache_impl.cuh:190-198) plus the running max of unique ids (cache/cache.cu:59-61,
 * written to max_ids, a device int32, may be NULL). */
int lg_io_complete(lg_sampler* s, lg_stream_t stream, int32_t mode, const lg_batch* batch,
                   unsigned long long* node_hotness, int32_t* max_ids);

/* ---- fused batch: ops 0..3*hops+3 back to back on one stream (GPURunner::RunOnce body,
 *      engine/server.cu:311-317) ---- */
typedef struct lg_batch_params {
  const int32_t* all_ids;    /* device */
  const int32_t* all_labels; /* device */
  int32_t total_cap;
  int32_t batch_size;
  int32_t counter; /* local batch id */
  int32_t mode;
  int32_t rng_kind;
  uint32_t batch_id; /* global batch id (Philox ctr word 2) */
  uint32_t stream_id; /* GPU / rank (Philox ctr word 3) */
  int32_t local_part; /* this GPU's slot in the clique */
  uint64_t rng_seed;
} lg_batch_params;

int lg_run_batch(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                 const lg_feature_cache* cache, const lg_batch_params* p, const lg_batch* batch,
                 unsigned long long* tier_rows);

/* End-to-end form with HOST buffers (pinned or pageable): H2D of the seed ids + labels of
 * this batch, lg_run_batch, D2H of both counter arrays (what get_next reads:
 * training_backend/ipc_cuda_kernel.cu:194-195).  Synchronises `stream` before returning. */
int lg_run_batch_host(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                      const lg_feature_cache* cache, const lg_batch_params* p,
                      const int32_t* host_seed_ids, const int32_t* host_seed_labels,
                      const lg_batch* batch, int32_t* host_node_counter,
                      int32_t* host_edge_counter);

/* same without the final synchronisation: the host buffers (pinned) are valid once `stream` has drained past
 * this call — several host-fed batches can be in flight, like the trainer's two pipeline slots */
int lg_run_batch_host_async(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                            const lg_feature_cache* cache, const lg_batch_params* p,
                            const int32_t* host_seed_ids, const int32_t* host_seed_labels,
                            const lg_batch* batch, int32_t* host_node_counter,
                            int32_t* host_edge_counter);

/* ---- standalone gather (the roofline kernel) : dst[r,:] = row of ids[r] for r in [0,n) ---- */
int lg_gather_rows(lg_stream_t stream, const lg_feature_cache* cache, const int32_t* ids,
                   int64_t n, float* dst, int32_t local_part, int32_t variant,
                   unsigned long long* tier_rows);

/* ---- trainer-side block construction ----
 * The reference trainer builds a DGL block from the COO of every layer on every step
 * (training_backend/legion_graphsage.py:66-79, create_unitgraph_from_coo) and DGL then converts it to CSC for
 * the SpMM.  lg_block_csc builds that CSC once from the CUDA-IPC buffers, on the trainer's stream:
 * block h = edges [0, ec[9+h]) of agg_src/agg_dst, num_dst = nc[9+h-1], num_src = nc[9+h]
 * (training_backend/ipc_cuda_kernel.cu:200-232).  indptr[num_dst+1]; indices[n_edges] = batch-local source of
 * each in-edge; eids[n_edges] (may be NULL) = position of the edge in the COO.  In-edges of one destination keep
 * their COO order (stable), so the result is a pure function of the COO.  Edges whose dst >= num_dst are invalid
 * input.  workspace: lg_block_csc_workspace(max_edges) bytes of device memory. */
int lg_block_csc_workspace(int64_t max_edges, int64_t* bytes);
int lg_block_csc(lg_stream_t stream, const int32_t* agg_src, const int32_t* agg_dst, int64_t n_edges,
                 int32_t num_dst, int32_t* indptr, int32_t* indices, int32_t* eids, void* workspace,
                 int64_t workspace_bytes);

/* Every block of a batch in ONE set of launches, sizes read on the device: block h (1..n_hops) = edges [0, ec[9+h]),
 * destinations [0, nc[9+h-1]) of batch->agg_src / agg_dst (training_backend/ipc_cuda_kernel.cu:218-231).  Nothing is read
 * on the host, so the call can be enqueued right behind the batch's sampling ops (the server builds the blocks next to
 * the gather).  max_edges[h-1] / max_dst[h-1] bound block h; indptr / indices / eids are arrays of n_hops device pointers
 * (eids may be NULL); workspace: lg_block_csc_batch_workspace(n_hops, max_edges) bytes. */
int lg_block_csc_batch_workspace(int32_t n_hops, const int64_t* max_edges, int64_t* bytes);
int lg_block_csc_batch(lg_stream_t stream, const lg_batch* batch, int32_t n_hops, const int64_t* max_edges,
                       const int32_t* max_dst, int32_t* const* indptr, int32_t* const* indices, int32_t* const* eids,
                       void* workspace, int64_t workspace_bytes);

/* ---- unified cache construction (cache/cache.cu:360-443,71-136,553-611) ---- */

/* aggregate_access (cache/cache_impl.cuh:72-76): agg[i] += part[i]; `part` may be peer memory */
int lg_hotness_accumulate(lg_stream_t stream, unsigned long long* agg,
                          const unsigned long long* part, int64_t n);
/* CandidateSelection ranking (cache/cache.cu:414-415,434-435): order = ids sorted by hotness
 * descending, ties by ascending id (the reference's thrust sort leaves ties unspecified).
 * sorted_hotness may be NULL.  tmp/tmp_bytes: pass tmp=NULL to query the size. */
int lg_hotness_rank(lg_stream_t stream, const unsigned long long* hotness, int64_t n,
                    int32_t* order, unsigned long long* sorted_hotness, void* tmp,
                    int64_t* tmp_bytes);
/* InitPair + insert (cache/cache_impl.cuh:104-109; cache/cache.cu:94-102): for rank r in
 * [0, cap*kg): directory[order[r]] = (r % kg)*cap + r/kg.  directory must be pre-filled with
 * LG_CACHEMISS_FLAG (lg_fill_i32). */
int lg_place_features(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                      int64_t num_nodes, int32_t* directory);
/* InitIndexPair/InitOffsetPair + insert (cache/cache_impl.cuh:89-101): part = r%kg + ki*kg,
 * row = r/kg, stored packed as part*cap + row. */
int lg_place_topology(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                      int32_t ki, int64_t num_nodes, int32_t* directory);
/* FeatFillUp (cache/cache_impl.cuh:183-188): shard[r,:] = backing[order[r*kg + j],:] */
int lg_fill_feature_shard(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                          int32_t j, int32_t dim, int64_t num_nodes, const float* backing,
                          float* shard);
/* Hybrid placement (no reference counterpart: a B200 has room to spare, Legion's GPUs did not): the `rep` hottest ranks
 * are stored on EVERY GPU of the clique (rows 0..rep-1 of each shard) and resolve to the reader's own part `j` — local
 * HBM reads for the head of the distribution instead of (Kg-1)/Kg peer reads — and the ranks after them are interleaved
 * over the parts below row `rep` exactly like lg_place_features / lg_fill_feature_shard do from row 0 (rep = 0 is the
 * reference placement).  Every GPU builds its own directory.  Rows cached per clique: rep + (cap-rep)*Kg. */
int lg_place_features_hybrid(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t rep, int32_t j,
                             int64_t num_nodes, int32_t* directory);
int lg_fill_feature_shard_hybrid(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg, int32_t rep, int32_t j,
                                 int32_t dim, int64_t num_nodes, const float* backing, float* shard);
/* GraphCache (storage/graph_storage.cu:76-111; graph_storage_impl.cuh:33-53), two calls:
 * counts -> shard_indptr[cap+1] (exclusive scan, shard_indptr[cap] = total edges), then fill. */
int lg_topo_shard_indptr(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                         int32_t j, int64_t num_nodes, const int64_t* indptr,
                         int64_t* shard_indptr);
int lg_topo_shard_fill(lg_stream_t stream, const int32_t* order, int32_t cap, int32_t kg,
                       int32_t j, int64_t num_nodes, const int64_t* indptr,
                       const int32_t* indices, const int64_t* shard_indptr,
                       int32_t* shard_indices);
int lg_fill_i32(lg_stream_t stream, int32_t* dst, int32_t value, int64_t n);

/* CostModel (cache/cache.cu:445-551), host arithmetic on host arrays: picks the feature /
 * topology split.  sorted_*_hotness are the descending arrays from lg_hotness_rank copied to
 * the host; topo_order is the topology ranking; indptr the full CSR offsets (host).
 * topo_trans / feat_trans are the measured transaction totals (reference: PCM counters,
 * always 0 — engine/server.cu:106).  Outputs capacities per GPU (rows). */
int lg_cost_model(const unsigned long long* sorted_node_hotness,
                  const unsigned long long* sorted_edge_hotness, const int32_t* topo_order,
                  const int64_t* indptr, int64_t num_nodes, int32_t dim, int64_t cache_bytes,
                  int32_t kg, uint64_t topo_trans, uint64_t feat_trans, int32_t* node_capacity,
                  int32_t* edge_capacity, double* alpha);
/* Same sweep with the saturation the reference lacks: a split whose feature (or topology) share already holds EVERY
 * vertex still counts its transactions and capacity (the reference skips the update when node_num == total,
 * cache/cache.cu:528-535, so a cache larger than the dataset — the normal case with 180 GB of HBM — ends with both
 * capacities 0), and node_num_feat is clamped to num_nodes.  Used by sampling_server (DESIGN.md deviation 8). */
int lg_cost_model_saturating(const unsigned long long* sorted_node_hotness,
                  const unsigned long long* sorted_edge_hotness, const int32_t* topo_order,
                  const int64_t* indptr, int64_t num_nodes, int32_t dim, int64_t cache_bytes,
                  int32_t kg, uint64_t topo_trans, uint64_t feat_trans, int32_t* node_capacity,
                  int32_t* edge_capacity, double* alpha);

/* ---- device plumbing a non-CUDA host needs (storage/storage_management.cu:5-23,100-115;
 *      engine/ipc_service.cu:163-169; training_backend/ipc_cuda_kernel.cu:62-68) ---- */
int lg_device_count(int32_t* n);
int lg_set_device(int32_t device);
int lg_enable_peer_access(int32_t n_devices);
int lg_device_alloc(void** ptr, int64_t bytes); /* plain cudaMalloc: legacy-IPC exportable */
int lg_device_free(void* ptr);
/* Device memory that OTHER PROCESSES map at full page size (cache shards of the one-process-per-GPU deployment):
 * cuMemCreate + POSIX-fd export / cuMemImportFromShareableHandle + cuMemMap.  A legacy cudaIpcOpenMemHandle mapping
 * is read through small pages by the importer and random row reads out of a multi-GB peer shard collapse
 * (profiles/r01b_peer_mapping.md); the trainer-facing batch buffers stay legacy-IPC (engine/ipc_service.cu:163-169).
 * lg_vmm_alloc allocates on the current device and returns a file descriptor to pass to the peers (SCM_RIGHTS);
 * lg_vmm_import maps a received descriptor for the current device; both sizes are rounded up to 2 MB. */
int64_t lg_vmm_round_up(int64_t bytes);
int lg_vmm_alloc(int64_t bytes, void** ptr, int32_t* shareable_fd);
int lg_vmm_import(int32_t shareable_fd, int64_t bytes, void** ptr);
int lg_vmm_free(void* ptr);
int lg_host_alloc_mapped(void** host_ptr, void** device_ptr, int64_t bytes); /* cudaHostAllocMapped */
int lg_host_free(void* host_ptr);
/* page-lock and map memory the host already owns (a POSIX shm mapping shared by the processes of a box): asynchronous
 * copies into it stay asynchronous, and kernels can read it through *device_ptr (UVA; may be NULL when not needed) */
int lg_host_register(void* host_ptr, int64_t bytes, void** device_ptr);
int lg_host_unregister(void* host_ptr);
int lg_ipc_export(const void* device_ptr, unsigned char handle[64]);
int lg_ipc_open(const unsigned char handle[64], void** device_ptr);
int lg_ipc_close(void* device_ptr);
int lg_stream_create(lg_stream_t* stream);
int lg_stream_destroy(lg_stream_t stream);
int lg_stream_synchronize(lg_stream_t stream);
int lg_memcpy_h2d(void* dst, const void* src, int64_t bytes, lg_stream_t stream);
int lg_memcpy_d2h(void* dst, const void* src, int64_t bytes, lg_stream_t stream);
int lg_memcpy_d2d(void* dst, const void* src, int64_t bytes, lg_stream_t stream);
int lg_memset_async(void* dst, int32_t byte_value, int64_t bytes, lg_stream_t stream);
/* events: the op DAG of GPURunner (engine/server.cu:250-260,311-324) */
typedef void* lg_event_t;
int lg_event_create(lg_event_t* ev);
int lg_event_destroy(lg_event_t ev);
int lg_event_record(lg_event_t ev, lg_stream_t stream);
int lg_event_query(lg_event_t ev, int32_t* ready); /* ready = 1 when complete */
int lg_event_synchronize(lg_event_t ev);
int lg_stream_wait_event(lg_stream_t stream, lg_event_t ev);
int lg_device_mem_info(int64_t* free_bytes, int64_t* total_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LEGION_B200_H_ */

// sampler_state.cuh — private state of an lg_sampler handle (shared by sampler.cu and gather.cu).
#pragma once
#include "common.cuh"

struct HopState {  // zeroed once per batch (batch_generate_body zeroes the whole small region)
  int32_t sample_ticket;
  int32_t rank_ticket;
  int32_t rank_done;
  int32_t new_nodes;  // C_h, written by the last rank tile
};

struct lg_sampler {
  int32_t device;
  int32_t max_batch;
  int32_t n_hops;
  int32_t gather_variant;
  int32_t fanout[LG_MAX_HOPS];
  int64_t slots_per_hop[LG_MAX_HOPS + 1];  // S_0 = batch, S_h = S_{h-1} * fanout_h
  int64_t num_ids;
  int64_t num_nodes;
  uint32_t* pm;          // position map [num_nodes]: 0xFFFFFFFF absent, kNewBit|edge position while a hop is open,
                         // else the batch-local id (engine/server.cu:224 position_map)
  u64* table;            // HASHED layout of the same map: (vertex << 32 | value) words, open addressing
  uint32_t table_mask;
  int32_t* seed_local;   // HASHED: batch-local ids of the seeds [max_batch]
  int32_t hashed;        // 0 dense pm, 1 hashed table (chosen at create time by the size of the graph)
  int32_t pm_dirty;      // a batch was generated and its words not yet released (lg_io_complete)
  lg_batch dirty_batch;
  cudaEvent_t ev_clear;        // recorded after the last release of the position map (lg_io_complete may run on
  cudaStream_t clear_stream;   // another stream than the next lg_batch_generate: engine/server.cu puts it on stream 2)
  int32_t clear_recorded;
  int32_t table_clean;         // HASHED: the table was re-initialised by the previous batch's last kernel
  unsigned* chain_bar;         // DENSE: grid-barrier words of chain_kernel ([0] arrivals, [1] generation)
  int32_t pm_fill_mb;          // DENSE maps up to this size are released by a streaming fill (LG_PM_FILL_MB at create time)
  int32_t chain;               // LG_CHAIN=1 at create time: lg_run_batch uses chain_kernel (opt-in: measured slower)
  int32_t* gid[2];       // double-buffered hop-relative global ids (next frontier)
  uint8_t* small;        // memset-per-batch region: HopState[hops] + chained-scan tile states
  int64_t small_bytes;
  HopState* hs;
  u64* sample_state[LG_MAX_HOPS];
  u64* rank_state[LG_MAX_HOPS];
  u64* sample_anchor[LG_MAX_HOPS];
  u64* rank_anchor[LG_MAX_HOPS];
  int32_t rank_items[LG_MAX_HOPS];  // edges per thread of the hop's rank kernel
  int32_t sample_tiles[LG_MAX_HOPS];
  int32_t sample_tile_f[LG_MAX_HOPS];
  int32_t rank_tiles[LG_MAX_HOPS];
  int32_t* status;       // device int32: 1 = ids overflow, 2 = features buffer overflow, 3 = miss without a backing matrix,
                         // 4 = edge_dst holds an id outside [0, num_nodes)
  int32_t* gather_ticket;  // [2] dynamic tile claims of the gather (re-armed by the kernel itself)
  int32_t gather_chunk;    // tiles per claim
  int32_t gather_static_pct;  // share of the launch that keeps the static tile order before the counter takes over
  int32_t* pinned_seeds;
  // gather/sampling overlap inside lg_run_batch: the gather of hop h runs on `side` while hop h+1
  // is sampled on the caller's stream (the reference's stream-1 / stream-0 split, server.cu:311-317)
  cudaStream_t side;
  cudaEvent_t ev_fork[LG_MAX_HOPS + 1];
  cudaEvent_t ev_join;
  int32_t fuse_gathers;  // lg_run_batch: 0 = one gather per op (reference schedule); 1 = seeds' rows ride with hop 1;
                         // 2 = a single gather of all rows after the last hop
  int32_t lazy_relabel;  // op-by-op calls: hop h's construct_graph is finished by hop h+1 (lg_sampler_set_lazy_relabel)
  int32_t tail_reference;  // lg_sampler_set_tail_mode: 1 = the tail batch strides by its clipped size like the reference
  int32_t overlap;  // 0 one stream, 1 fork/join inside a batch, 2 pipelined across batches (see lg_batch_wait)
  // pipelined mode: completion event of the last batch that used a given set of buffers
  struct Done {
    const void* key;  // batch->ids
    cudaEvent_t ev;
  } done[4];
  int32_t n_done;
  u64* trace;  // diagnostics: per-tile phase timestamps (lg_debug_set_trace), nullptr in production
};

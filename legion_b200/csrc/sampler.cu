// sampler.cu — k-hop neighbour sampler of the Legion data path for sm_100a.
//
// Replaces, per hop, the reference's chain  FindTopo (2 bcht probes) -> random_sample ->
// construct_graph -> counter_update (+ 2 blocking D2H copies)   [engine/operator_impl.cu:175-296,
// 400-499; cache/cache.cu:217-225]  with TWO launches and no host synchronisation:
//
//   sample_hop_kernel   tile of frontier entries -> row lookup (directory -> HBM shard / peer
//                       shard / host UVA), per-entry edge count min(deg, fanout), chained scan
//                       (decoupled look-back) for the canonical edge offsets, with-replacement
//                       pick (Philox4x32-10 or the reference's minstd stream), edge emission in
//                       ascending slot order, insert-min of (vertex -> first edge position) into
//                       the batch dedup table.
//   rank_relabel_kernel first-occurrence flags -> chained scan -> batch-local ids in first-seen
//                       order, `ids` append, COO source relabel (construct_graph), the op's
//                       counter_update done by the last CTA.
//
// State that the reference keeps O(N) per GPU (accessed bitmap + position_map, memset / cleared
// every batch: engine/operator_impl.cu:151,542-548) is an O(batch) open-addressing table of
// (vertex, local id) words that stays L2-resident.
#include "common.cuh"
#include "sampler_state.cuh"

using namespace lg;

namespace {

constexpr u64 kEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kNewBit = 0x80000000u;  // value = kNewBit | first edge position while a hop is open
constexpr int kBlock = 256;
constexpr int kRankItems = 4;                      // edges per thread in rank_relabel_kernel
constexpr int kRankTile = kBlock * kRankItems;     // 1024 edges per tile
static_assert(kRankItems * (kBlock / 32) == 32, "rank tile partial counts must fill one warp");
constexpr int kSlotUnroll = 4;                     // neighbour reads in flight per thread in sample_hop_kernel

// optional per-tile phase timestamps (diagnostics only: lg_debug_set_trace; nullptr in production)
constexpr int kTraceTiles = 2048, kTracePhases = 8;
__device__ __forceinline__ void trace_mark(u64* trace, int kernel_slot, int tile, int phase) {
  if (!trace || tile >= kTraceTiles) return;
  u64 t, c;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  c = (u64)clock64();
  u64* p = trace + ((size_t)kernel_slot * kTraceTiles + tile) * kTracePhases * 2 + phase * 2;
  p[0] = t;
  p[1] = c;
}

// Insert-min of (key -> val).  The pre-check reads the slot through L1 (ld.ca): popular vertices are sampled
// thousands of times per batch, and sending every one of those reads to the single L2 slice that owns the
// slot serialises the whole kernel.  A stale line is harmless: keys never change once written, values only
// decrease, so a stale value can only make us issue an atomic that turns out to be a no-op — never skip
// one that was needed (stale >= actual, and we skip only when stale <= val).
__device__ __forceinline__ void table_insert_min(u64* table, uint32_t mask, int32_t key, uint32_t val) {
  uint32_t slot = hash32((uint32_t)key) & mask;
  const u64 packed = ((u64)(uint32_t)key << 32) | val;
  while (true) {
    u64 cur = __ldca(table + slot);
    if (cur == kEmpty) {
      cur = atomicCAS(table + slot, kEmpty, packed);
      if (cur == kEmpty) return;
    }
    if ((uint32_t)(cur >> 32) == (uint32_t)key) {
      if ((uint32_t)cur > val) atomicMin(table + slot, packed);
      return;
    }
    slot = (slot + 1) & mask;
  }
}
// Batched form: the caller has already loaded `cur` = table[slot] for several keys at once (memory-level
// parallelism across a thread's slots) and, for empty slots, already issued the CAS (`cur` = its return value,
// `claimed` = the CAS found the slot empty).  Finishes the insert; falls back to linear probing on a collision.
__device__ __forceinline__ void table_insert_finish(u64* table, uint32_t mask, int32_t key, uint32_t val, uint32_t slot,
                                                    u64 cur, bool claimed) {
  if (claimed) return;
  if ((uint32_t)(cur >> 32) == (uint32_t)key) {
    if ((uint32_t)cur > val) atomicMin(table + slot, ((u64)(uint32_t)key << 32) | val);
    return;
  }
  // another key owns this slot: continue with the generic probe sequence from the next slot
  const u64 packed = ((u64)(uint32_t)key << 32) | val;
  slot = (slot + 1) & mask;
  while (true) {
    cur = __ldca(table + slot);
    if (cur == kEmpty) {
      cur = atomicCAS(table + slot, kEmpty, packed);
      if (cur == kEmpty) return;
    }
    if ((uint32_t)(cur >> 32) == (uint32_t)key) {
      if ((uint32_t)cur > val) atomicMin(table + slot, packed);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// key must be present.  `cached`: probe through L1 — valid when the caller only needs the key/slot or a value
// that cannot have changed since the kernel started (keys are immutable within a batch).
template <bool CACHED>
__device__ __forceinline__ uint32_t table_find(const u64* table, uint32_t mask, int32_t key, u64* word) {
  uint32_t slot = hash32((uint32_t)key) & mask;
  while (true) {
    u64 cur = CACHED ? __ldca(table + slot) : ld_relaxed(table + slot);
    if ((uint32_t)(cur >> 32) == (uint32_t)key || cur == kEmpty) {
      *word = cur;
      return slot;
    }
    slot = (slot + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------------
// batch_generate + op-0 counter_update (engine/operator_impl.cu:27-89,159-165)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) batch_generate_kernel(
    const int32_t* __restrict__ all_ids, const int32_t* __restrict__ all_labels, int32_t total_cap,
    int32_t size, int32_t counter, int32_t hop_num, int32_t* __restrict__ ids, int32_t* __restrict__ labels,
    int32_t* __restrict__ nc, int32_t* __restrict__ ec, u64* table, uint32_t mask) {
  int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < LG_COUNTER_SLOTS) {
    int t = threadIdx.x;
    int32_t v = 0;
    if (t == 1 || t == LG_INTRABATCH_CON * 3) v = size;  // nc[1], nc[9]
    if (t == LG_INTRABATCH_CON * 3 - 1) v = hop_num;     // nc[8]
    nc[t] = v;
    ec[t] = 0;
  }
  if (idx >= size) return;
  long long pos = (long long)size * counter + idx;  // the reference strides by the clipped size (:40,:44,:162)
  if (pos >= total_cap) {
    ids[idx] = -1;
    labels[idx] = -1;
    return;
  }
  int32_t v = all_ids[pos % total_cap];
  ids[idx] = v;
  labels[idx] = all_labels[pos % total_cap];
  if (v >= 0) table_insert_min(table, mask, v, (uint32_t)idx);  // local index = first position
}

// ------------------------------------------------------------------------------------------
// sample_hop_kernel
// ------------------------------------------------------------------------------------------
struct SampleArgs {
  lg_topology topo;
  const int32_t* frontier_prev;  // hop > 1: global ids of the previous hop's sampled sources
  int32_t* gid_out;              // this hop's sampled sources (global ids), hop-relative positions
  int32_t* ids;
  int32_t* agg_src;
  int32_t* agg_dst;
  int32_t* nc;
  int32_t* ec;
  u64* table;
  u64* tile_state;
  HopState* hs;
  u64* edge_hot;
  uint32_t mask;
  int32_t hop;
  int32_t fanout;
  uint32_t batch_id, stream_id, k0, k1;
  u64* trace;
};

template <int TILE_F, int RNG, int INS>
__global__ void __launch_bounds__(kBlock) sample_hop_kernel(const SampleArgs a) {
  __shared__ long long s_start[TILE_F];
  __shared__ int32_t s_deg[TILE_F];
  __shared__ int32_t s_cnt[TILE_F];
  __shared__ int32_t s_off[TILE_F];
  __shared__ int32_t s_flocal[TILE_F];
  __shared__ const int32_t* s_indices[TILE_F];
  __shared__ int32_t s_warp[kBlock / 32];
  __shared__ int32_t s_tile, s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&a.hs->sample_ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int tslot = (a.hop - 1) * 2;
  if (tid == 0) trace_mark(a.trace, tslot, tile, 0);

  const bool first_hop = (a.hop == 1);
  const int32_t F = first_hop ? a.nc[1] : a.ec[1];          // :201-206
  const int32_t prev_edge_off = a.ec[0];
  const int32_t edge_base = a.ec[0] + a.ec[1];              // :275
  const int n_tiles = (F + TILE_F - 1) / TILE_F;
  if (tile >= n_tiles) {
    if (n_tiles == 0 && tile == 0 && tid == 0) a.ec[2] = 0;
    return;
  }
  const int32_t c = a.fanout;
  const int32_t i0 = tile * TILE_F;

  // 1. row lookup for the tile's frontier entries
  if (tid < TILE_F) {
    int32_t i = i0 + tid;
    int32_t cnt = 0, deg = 0, fl = 0;
    long long start = 0;
    const int32_t* ind = a.topo.indices[a.topo.n_parts];
    if (i < F) {
      int32_t v = first_hop ? a.ids[i] : a.frontier_prev[i];
      if (v >= 0) {
        if (first_hop) {
          u64 w;
          table_find<true>(a.table, a.mask, v, &w);
          fl = (int32_t)(uint32_t)w;
        } else {
          fl = a.agg_src[prev_edge_off + i];
        }
        int part = a.topo.n_parts;
        long long row = v;
        if (a.topo.directory) {
          int32_t loc = a.topo.directory[v];
          if (loc >= 0) {
            part = loc / a.topo.shard_rows;
            row = loc - part * a.topo.shard_rows;
          }
        }
        const int64_t* ip = a.topo.indptr[part];
        start = ip[row];
        deg = (int32_t)(ip[row + 1] - start);  // :226 (int32 col_size)
        ind = a.topo.indices[part];
        cnt = deg < c ? deg : c;
        if (cnt < 0) cnt = 0;
        if (a.edge_hot && cnt > 0) atomicAdd(a.edge_hot + v, (u64)cnt);  // pre_sample :358, summed per entry
      }
    }
    s_start[tid] = start;
    s_deg[tid] = deg;
    s_cnt[tid] = cnt;
    s_flocal[tid] = fl;
    s_indices[tid] = ind;
  }
  __syncthreads();
  if (tid == 0) trace_mark(a.trace, tslot, tile, 1);

  // 2. exclusive scan of the per-entry edge counts inside the tile
  {
    int32_t v = (tid < TILE_F) ? s_cnt[tid] : 0;
    int32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = (lane < kBlock / 32) ? s_warp[lane] : 0;
      int32_t winc = warp_incl_scan(w, lane);
      if (lane < kBlock / 32) s_warp[lane] = winc - w;
      int32_t total = __shfl_sync(0xffffffffu, winc, kBlock / 32 - 1);
      // 3. chained scan across tiles
      int32_t excl = lookback_exclusive(a.tile_state, tile, total, lane);
      if (lane == 0) {
        s_base = excl;
        if (tile == n_tiles - 1) a.ec[2] = excl + total;  // E_h (:264)
      }
    }
    __syncthreads();
    if (tid < TILE_F) s_off[tid] = s_warp[warp] + inc - v;
  }
  __syncthreads();
  const int32_t base = s_base;
  if (tid == 0) trace_mark(a.trace, tslot, tile, 2);

  // 4. one thread per slot: pick, emit, insert-min.  kSlotUnroll independent neighbour reads are issued
  //    back to back before any of them is consumed (the read is a random 4-byte HBM/NVLink/PCIe access).
  const int n_slots = TILE_F * c;
  for (int k0 = tid; k0 < n_slots; k0 += kBlock * kSlotUnroll) {
    int32_t w[kSlotUnroll], p[kSlotUnroll], fl[kSlotUnroll];
#pragma unroll
    for (int u = 0; u < kSlotUnroll; u++) {
      const int k = k0 + u * kBlock;
      p[u] = -1;
      if (k < n_slots) {
        const int t = k / c, j = k - t * c;
        if (j < s_cnt[t]) {  // :232  neighbor_offset >= col_size -> none
          const uint32_t slot = (uint32_t)(i0 + t) * (uint32_t)c + (uint32_t)j;
          const int32_t pick =
              pick_neighbor<RNG>(slot, s_deg[t], (uint32_t)a.hop, a.batch_id, a.stream_id, a.k0, a.k1);
          w[u] = __ldg(s_indices[t] + s_start[t] + pick);  // :240-242
          p[u] = base + s_off[t] + j;
          fl[u] = s_flocal[t];
        }
      }
    }
    // dedup-table insert-min.  INS 0: one slot after the other; 1: all first probes of the group in flight
    // together, claims one by one; 2: probes and claims both batched
    if (INS == 0) {
#pragma unroll
      for (int u = 0; u < kSlotUnroll; u++) {
        if (p[u] >= 0) {
          a.gid_out[p[u]] = w[u];
          a.agg_dst[edge_base + p[u]] = fl[u];  // construct_graph :292,294
          table_insert_min(a.table, a.mask, w[u], kNewBit | (uint32_t)p[u]);
        }
      }
    } else {
      uint32_t sl[kSlotUnroll];
      u64 cur[kSlotUnroll];
#pragma unroll
      for (int u = 0; u < kSlotUnroll; u++) {
        if (p[u] >= 0) {
          a.gid_out[p[u]] = w[u];
          a.agg_dst[edge_base + p[u]] = fl[u];
          sl[u] = hash32((uint32_t)w[u]) & a.mask;
          cur[u] = __ldca(a.table + sl[u]);
        }
      }
      bool claimed[kSlotUnroll];
      if (INS == 2) {
#pragma unroll
        for (int u = 0; u < kSlotUnroll; u++) {
          claimed[u] = false;
          if (p[u] >= 0 && cur[u] == kEmpty) {
            cur[u] = atomicCAS(a.table + sl[u], kEmpty, ((u64)(uint32_t)w[u] << 32) | (kNewBit | (uint32_t)p[u]));
            claimed[u] = (cur[u] == kEmpty);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kSlotUnroll; u++) {
        if (p[u] < 0) continue;
        if (INS == 1) {
          claimed[u] = false;
          if (cur[u] == kEmpty) {
            cur[u] = atomicCAS(a.table + sl[u], kEmpty, ((u64)(uint32_t)w[u] << 32) | (kNewBit | (uint32_t)p[u]));
            claimed[u] = (cur[u] == kEmpty);
          }
        }
        table_insert_finish(a.table, a.mask, w[u], kNewBit | (uint32_t)p[u], sl[u], cur[u], claimed[u]);
      }
    }
  }
  if (a.trace) {
    if (tid == 0) trace_mark(a.trace, tslot, tile, 3);  // thread 0 done
    __syncthreads();
    if (tid == 0) trace_mark(a.trace, tslot, tile, 4);  // whole CTA done
  }
}

// ------------------------------------------------------------------------------------------
// rank_relabel_kernel
// ------------------------------------------------------------------------------------------
struct RankArgs {
  const int32_t* gid;  // this hop's sampled sources
  int32_t* ids;
  int32_t* agg_src;
  int32_t* nc;
  int32_t* ec;
  u64* table;
  u64* tile_state;
  HopState* hs;
  uint32_t mask;
  int32_t hop;
  int32_t ids_cap;
  int32_t* status;
  u64* trace;
};

__global__ void __launch_bounds__(kBlock) rank_relabel_kernel(const RankArgs a) {
  __shared__ int32_t s_cnt[kRankItems * (kBlock / 32)];
  __shared__ int32_t s_tile, s_base, s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&a.hs->rank_ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int tslot = (a.hop - 1) * 2 + 1;
  if (tid == 0) trace_mark(a.trace, tslot, tile, 0);
  const int32_t E = a.ec[2];
  const int32_t node_base = a.nc[0] + a.nc[1];  // :268
  const int32_t edge_base = a.ec[0] + a.ec[1];
  const int n_tiles = (E + kRankTile - 1) / kRankTile;

  if (tile < n_tiles) {
    const int32_t p0 = tile * kRankTile;
    int32_t w[kRankItems];
    uint32_t slot[kRankItems];
    bool first[kRankItems];
    unsigned bal[kRankItems];
    u64 word[kRankItems];
#pragma unroll
    for (int k = 0; k < kRankItems; k++) {  // all first probes of the thread's edges in flight together
      const int32_t p = p0 + k * kBlock + tid;
      w[k] = (p < E) ? a.gid[p] : -1;
    }
#pragma unroll
    for (int k = 0; k < kRankItems; k++) {
      slot[k] = hash32((uint32_t)w[k]) & a.mask;
      // L1-cached probe: a line fetched before the owner publishes still holds kNewBit|p_first, one fetched
      // after holds the final id — neither can equal kNewBit|p for a non-owner, and the owner's own word is
      // only ever rewritten by the owner itself
      word[k] = (w[k] >= 0) ? __ldca(a.table + slot[k]) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kRankItems; k++) {
      const int32_t p = p0 + k * kBlock + tid;
      first[k] = false;
      if (w[k] >= 0) {
        while ((uint32_t)(word[k] >> 32) != (uint32_t)w[k] && word[k] != kEmpty) {  // collision: next slot
          slot[k] = (slot[k] + 1) & a.mask;
          word[k] = __ldca(a.table + slot[k]);
        }
        first[k] = ((uint32_t)word[k] == (kNewBit | (uint32_t)p));
      }
      bal[k] = __ballot_sync(0xffffffffu, first[k]);
      if (lane == 0) s_cnt[k * (kBlock / 32) + warp] = __popc(bal[k]);
    }
    if (tid == 0) trace_mark(a.trace, tslot, tile, 1);  // warp 0 probes done
    __syncthreads();
    if (tid == 0) trace_mark(a.trace, tslot, tile, 2);  // all probes done
    if (warp == 0) {  // kRankItems * 8 == 32 partial counts, in edge order
      int32_t v = s_cnt[lane];
      int32_t inc = warp_incl_scan(v, lane);
      s_cnt[lane] = inc - v;
      int32_t total = __shfl_sync(0xffffffffu, inc, 31);
      int32_t excl = lookback_exclusive(a.tile_state, tile, total, lane);
      if (lane == 0) {
        s_base = excl;
        if (tile == n_tiles - 1) a.hs->new_nodes = excl + total;  // C_h (:263)
      }
    }
    __syncthreads();
    const int32_t base = node_base + s_base;
    if (tid == 0) trace_mark(a.trace, tslot, tile, 3);  // look-back done
    const unsigned lt = (1u << lane) - 1u;
    // first occurrences: assign the local id, append to ids, publish in the table
#pragma unroll
    for (int k = 0; k < kRankItems; k++) {
      if (first[k]) {
        int32_t local = base + s_cnt[k * (kBlock / 32) + warp] + __popc(bal[k] & lt);
        if (local < a.ids_cap) a.ids[local] = w[k];  // :270
        else *a.status = 1;
        st_relaxed(a.table + slot[k], ((u64)(uint32_t)w[k] << 32) | (uint32_t)local);  // position_map :271
        a.agg_src[edge_base + p0 + k * kBlock + tid] = local;
      }
    }
    if (tid == 0) trace_mark(a.trace, tslot, tile, 4);  // published
    // repeats: wait for the owner (an earlier edge, in this or an earlier tile) to publish
#pragma unroll
    for (int k = 0; k < kRankItems; k++) {
      int32_t p = p0 + k * kBlock + tid;
      if (p < E && !first[k]) {
        u64 word = ld_relaxed(a.table + slot[k]);
        while ((uint32_t)word & kNewBit) {  // back off: thousands of repeats of a hub vertex poll the same word
          __nanosleep(64);
          word = ld_relaxed(a.table + slot[k]);
        }
        a.agg_src[edge_base + p] = (int32_t)(uint32_t)word;  // construct_graph :291,293
      }
    }
  }

  // the op's counter_update (:69-82), by the last CTA to finish
  if (tid == 0) trace_mark(a.trace, tslot, tile, 5);  // thread 0 spins done
  __syncthreads();
  if (tid == 0) {
    trace_mark(a.trace, tslot, tile, 6);  // all spins done
    __threadfence();
    int32_t done = atomicAdd(&a.hs->rank_done, 1);
    s_last = (done == (int32_t)gridDim.x - 1);
  }
  __syncthreads();
  if (tid == 0) trace_mark(a.trace, tslot, tile, 7);
  if (s_last && tid == 0) {
    __threadfence();
    volatile int32_t* nc = a.nc;
    volatile int32_t* ec = a.ec;
    int32_t C = (n_tiles > 0) ? *((volatile int32_t*)&a.hs->new_nodes) : 0;
    int32_t nc0 = nc[0] + nc[1];
    nc[0] = nc0;
    nc[1] = C;
    nc[LG_INTRABATCH_CON * 2] = 0;
    nc[LG_INTRABATCH_CON * 2 + 1] = nc0 + C;
    int32_t ec0 = ec[0] + ec[1];
    ec[0] = ec0;
    ec[1] = E;
    ec[2] = 0;
    nc[LG_INTRABATCH_CON * 3 + a.hop] = nc0 + C;
    ec[LG_INTRABATCH_CON * 3 + a.hop] = ec0 + E;
  }
}

// HotnessMeasure (cache/cache_impl.cuh:190-198) + max_ids_ (cache/cache.cu:59-61)
__global__ void __launch_bounds__(kBlock) hotness_measure_kernel(const int32_t* __restrict__ ids,
                                                                 const int32_t* __restrict__ nc, u64* node_hot,
                                                                 int32_t* max_ids) {
  const int32_t n = nc[LG_INTRABATCH_CON * 2 + 1];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int32_t cid = ids[i];
    if (cid >= 0) atomicAdd(node_hot + cid, 1ull);
  }
  if (max_ids && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(max_ids, n);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
extern "C" int64_t lg_num_ids(int32_t batch_size, const int32_t* fanout, int32_t n_hops) {
  int64_t tot = batch_size, per = batch_size;  // engine/server.cu:187-199
  for (int i = 0; i < n_hops; i++) {
    per *= fanout[i];
    tot += per;
  }
  return tot;
}

static int64_t next_pow2(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

static int pick_tile_f(int64_t frontier_max) {
  // keep >= ~2 waves of CTAs on 148 SMs when the frontier allows it
  if (frontier_max >= 256ll * kSMs * 2) return 256;
  if (frontier_max >= 128ll * kSMs * 2) return 128;
  if (frontier_max >= 64ll * kSMs) return 64;
  return 32;
}

static int sampler_alloc_table(lg_sampler* s, int64_t slots) {
  if (s->table) cudaFree(s->table);
  s->table = nullptr;
  s->table_slots = slots;
  LG_CUDA(cudaMalloc(&s->table, (size_t)slots * sizeof(u64)));
  LG_CUDA(cudaMemset(s->table, 0xFF, (size_t)slots * sizeof(u64)));
  return 0;
}

extern "C" int lg_sampler_create(int32_t device, int32_t max_batch, const int32_t* fanout, int32_t n_hops,
                                 lg_sampler** out) {
  LG_REQUIRE(out && fanout, "lg_sampler_create: null argument");
  LG_REQUIRE(n_hops >= 1 && n_hops <= LG_MAX_HOPS, "lg_sampler_create: n_hops %d outside [1,%d]", n_hops, LG_MAX_HOPS);
  LG_REQUIRE(max_batch >= 1, "lg_sampler_create: max_batch %d", max_batch);
  LG_CUDA(cudaSetDevice(device));
  lg_sampler* s = new lg_sampler();
  memset(s, 0, sizeof(*s));
  s->device = device;
  s->max_batch = max_batch;
  s->n_hops = n_hops;
  s->gather_variant = LG_GATHER_AUTO;
  s->slots_per_hop[0] = max_batch;
  for (int h = 0; h < n_hops; h++) {
    LG_REQUIRE(fanout[h] >= 1, "lg_sampler_create: fanout[%d]=%d", h, fanout[h]);
    s->fanout[h] = fanout[h];
    s->slots_per_hop[h + 1] = s->slots_per_hop[h] * fanout[h];
  }
  s->num_ids = lg_num_ids(max_batch, fanout, n_hops);
  LG_REQUIRE(s->num_ids < (1ll << 31), "lg_sampler_create: num_ids %lld does not fit int32", (long long)s->num_ids);
  int64_t smax = s->slots_per_hop[n_hops];
  if (smax < 2ll * max_batch) smax = 2ll * max_batch;  // head of gid[1] doubles as the e2e seed staging area
  for (int b = 0; b < 2; b++) LG_CUDA(cudaMalloc(&s->gid[b], (size_t)smax * sizeof(int32_t)));
  // small per-batch state
  int64_t bytes = sizeof(HopState) * LG_MAX_HOPS;
  bytes = (bytes + 255) & ~255ll;
  int64_t off_state[2][LG_MAX_HOPS];
  for (int h = 0; h < n_hops; h++) {
    int tf = pick_tile_f(s->slots_per_hop[h]);
    s->sample_tile_f[h] = tf;
    s->sample_tiles[h] = (int32_t)((s->slots_per_hop[h] + tf - 1) / tf);
    s->rank_tiles[h] = (int32_t)((s->slots_per_hop[h + 1] + kRankTile - 1) / kRankTile);
    off_state[0][h] = bytes;
    bytes += (int64_t)s->sample_tiles[h] * 8;
    off_state[1][h] = bytes;
    bytes += (int64_t)s->rank_tiles[h] * 8;
  }
  s->small_bytes = bytes;
  LG_CUDA(cudaMalloc(&s->small, (size_t)bytes));
  LG_CUDA(cudaMemset(s->small, 0, (size_t)bytes));
  s->hs = (HopState*)s->small;
  for (int h = 0; h < n_hops; h++) {
    s->sample_state[h] = (u64*)(s->small + off_state[0][h]);
    s->rank_state[h] = (u64*)(s->small + off_state[1][h]);
  }
  LG_CUDA(cudaMalloc(&s->status, sizeof(int32_t)));
  LG_CUDA(cudaMemset(s->status, 0, sizeof(int32_t)));
  LG_CUDA(cudaMallocHost(&s->pinned_seeds, (size_t)max_batch * 2 * sizeof(int32_t)));
  LG_CUDA(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  for (int h = 0; h <= LG_MAX_HOPS; h++) LG_CUDA(cudaEventCreateWithFlags(&s->ev_fork[h], cudaEventDisableTiming));
  LG_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  s->overlap = 1;
  s->fuse_gathers = 1;
  int rc = sampler_alloc_table(s, next_pow2(s->num_ids + s->num_ids / 2));
  if (rc) return rc;
  *out = s;
  return 0;
}

extern "C" int lg_sampler_destroy(lg_sampler* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->table);
  cudaFree(s->gid[0]);
  cudaFree(s->gid[1]);
  cudaFree(s->small);
  cudaFree(s->status);
  cudaFreeHost(s->pinned_seeds);
  cudaStreamDestroy(s->side);
  for (int h = 0; h <= LG_MAX_HOPS; h++) cudaEventDestroy(s->ev_fork[h]);
  cudaEventDestroy(s->ev_join);
  for (int i = 0; i < s->n_done; i++) cudaEventDestroy(s->done[i].ev);
  delete s;
  return 0;
}

extern "C" int lg_sampler_set_table_slots(lg_sampler* s, int64_t slots) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(slots > s->num_ids && (slots & (slots - 1)) == 0 && slots <= (1ll << 31),
             "table slots %lld must be a power of two > num_ids %lld", (long long)slots, (long long)s->num_ids);
  LG_CUDA(cudaSetDevice(s->device));
  LG_CUDA(cudaDeviceSynchronize());
  return sampler_alloc_table(s, slots);
}

extern "C" int lg_debug_set_trace(lg_sampler* s, unsigned long long* device_buf) {
  LG_REQUIRE(s, "null sampler");
  s->trace = (u64*)device_buf;
  return 0;
}
extern "C" int64_t lg_debug_trace_words(void) { return (int64_t)LG_MAX_HOPS * 2 * kTraceTiles * kTracePhases * 2; }

extern "C" int lg_sampler_set_gather_variant(lg_sampler* s, int32_t variant) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(variant >= LG_GATHER_AUTO && variant <= LG_GATHER_TMA, "gather variant %d", variant);
  s->gather_variant = variant;
  return 0;
}

extern "C" int lg_sampler_set_gather_fusion(lg_sampler* s, int32_t mode) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(mode >= 0 && mode <= 2, "gather fusion mode %d", mode);
  s->fuse_gathers = mode;
  return 0;
}

extern "C" int lg_sampler_set_overlap(lg_sampler* s, int32_t mode) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(mode >= 0 && mode <= 2, "overlap mode %d", mode);
  s->overlap = mode;
  return 0;
}

static int done_event(lg_sampler* s, const lg_batch* b, bool create, cudaEvent_t* ev) {
  *ev = nullptr;
  for (int i = 0; i < s->n_done; i++)
    if (s->done[i].key == b->ids) {
      *ev = s->done[i].ev;
      return 0;
    }
  if (!create) return 0;
  LG_REQUIRE(s->n_done < 4, "pipelined mode supports at most 4 distinct batch buffer sets");
  LG_CUDA(cudaEventCreateWithFlags(&s->done[s->n_done].ev, cudaEventDisableTiming));
  s->done[s->n_done].key = b->ids;
  *ev = s->done[s->n_done].ev;
  s->n_done++;
  return 0;
}

extern "C" int lg_batch_wait(lg_sampler* s, lg_stream_t stream, const lg_batch* b) {
  LG_REQUIRE(s && b, "lg_batch_wait: null argument");
  cudaEvent_t ev;
  int rc = done_event(s, b, false, &ev);
  if (rc) return rc;
  if (ev) LG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ev, 0));
  return 0;
}

extern "C" int lg_sampler_status(lg_sampler* s, lg_stream_t stream, int32_t* host_status) {
  LG_REQUIRE(s && host_status, "null argument");
  LG_CUDA(cudaMemcpyAsync(host_status, s->status, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  LG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

extern "C" int64_t lg_sampler_scratch_bytes(const lg_sampler* s) {
  if (!s) return 0;
  return s->table_slots * 8 + 2 * s->slots_per_hop[s->n_hops] * 4 + s->small_bytes + 4;
}

extern "C" int lg_batch_generate(lg_sampler* s, lg_stream_t stream_, const int32_t* all_ids,
                                 const int32_t* all_labels, int32_t total_cap, int32_t batch_size, int32_t counter,
                                 const lg_batch* b) {
  LG_REQUIRE(s && b && all_ids && all_labels, "lg_batch_generate: null argument");
  LG_REQUIRE(batch_size <= s->max_batch, "lg_batch_generate: batch %d > max_batch %d", batch_size, s->max_batch);
  LG_REQUIRE(b->num_ids >= s->num_ids, "lg_batch_generate: batch buffers hold %d ids, need %lld", b->num_ids,
             (long long)s->num_ids);
  cudaStream_t st = (cudaStream_t)stream_;
  // reset of the per-batch state: dedup table (replaces the O(N) bitmap memset, :151) + scan state
  LG_CUDA(cudaMemsetAsync(s->table, 0xFF, (size_t)s->table_slots * sizeof(u64), st));
  LG_CUDA(cudaMemsetAsync(s->small, 0, (size_t)s->small_bytes, st));
  long long done = (long long)batch_size * ((long long)counter + 1);
  int32_t size = (done >= total_cap) ? (int32_t)(total_cap - (long long)batch_size * counter) : batch_size;  // :159
  if (size < 0) size = 0;
  int grid = size > 0 ? (size + kBlock - 1) / kBlock : 1;
  batch_generate_kernel<<<grid, kBlock, 0, st>>>(all_ids, all_labels, total_cap, size, counter, s->n_hops, b->ids,
                                                 b->labels, b->node_counter, b->edge_counter, s->table,
                                                 (uint32_t)(s->table_slots - 1));
  LG_LAUNCH_OK();
  return 0;
}

template <int RNG, int INS>
static void launch_sample_ins(int tile_f, int grid, cudaStream_t st, const SampleArgs& a) {
  switch (tile_f) {
    case 256: sample_hop_kernel<256, RNG, INS><<<grid, kBlock, 0, st>>>(a); break;
    case 128: sample_hop_kernel<128, RNG, INS><<<grid, kBlock, 0, st>>>(a); break;
    case 64: sample_hop_kernel<64, RNG, INS><<<grid, kBlock, 0, st>>>(a); break;
    default: sample_hop_kernel<32, RNG, INS><<<grid, kBlock, 0, st>>>(a); break;
  }
}
template <int RNG>
static void launch_sample(int tile_f, int grid, cudaStream_t st, const SampleArgs& a) {
  static const int ins = [] { const char* e = getenv("LG_SAMPLE_INS"); return e ? atoi(e) : 0; }();
  if (ins == 2) launch_sample_ins<RNG, 2>(tile_f, grid, st, a);
  else if (ins == 1) launch_sample_ins<RNG, 1>(tile_f, grid, st, a);
  else launch_sample_ins<RNG, 0>(tile_f, grid, st, a);
}

extern "C" int lg_random_sample(lg_sampler* s, lg_stream_t stream_, const lg_topology* topo, int32_t hop,
                                int32_t rng_kind, uint64_t rng_seed, uint32_t batch_id, uint32_t stream_id,
                                const lg_batch* b, unsigned long long* edge_hotness) {
  LG_REQUIRE(s && topo && b, "lg_random_sample: null argument");
  LG_REQUIRE(hop >= 1 && hop <= s->n_hops, "lg_random_sample: hop %d outside [1,%d]", hop, s->n_hops);
  LG_REQUIRE(rng_kind == LG_RNG_MINSTD || rng_kind == LG_RNG_PHILOX, "lg_random_sample: rng_kind %d", rng_kind);
  LG_REQUIRE(topo->n_parts >= 0 && topo->n_parts <= LG_MAX_DEVICE, "lg_random_sample: n_parts %d", topo->n_parts);
  LG_REQUIRE(topo->indptr[topo->n_parts] && topo->indices[topo->n_parts], "lg_random_sample: full CSR slot is null");
  LG_REQUIRE(!topo->directory || topo->shard_rows > 0, "lg_random_sample: directory without shard_rows");
  cudaStream_t st = (cudaStream_t)stream_;
  const int h = hop - 1;
  SampleArgs a;
  a.topo = *topo;
  a.frontier_prev = s->gid[(h + 1) & 1];
  a.gid_out = s->gid[h & 1];
  a.ids = b->ids;
  a.agg_src = b->agg_src;
  a.agg_dst = b->agg_dst;
  a.nc = b->node_counter;
  a.ec = b->edge_counter;
  a.table = s->table;
  a.tile_state = s->sample_state[h];
  a.hs = s->hs + h;
  a.edge_hot = (u64*)edge_hotness;
  a.mask = (uint32_t)(s->table_slots - 1);
  a.hop = hop;
  a.fanout = s->fanout[h];
  a.batch_id = batch_id;
  a.stream_id = stream_id;
  a.k0 = (uint32_t)rng_seed;
  a.k1 = (uint32_t)(rng_seed >> 32);
  a.trace = s->trace;

  if (rng_kind == LG_RNG_MINSTD)
    launch_sample<LG_RNG_MINSTD>(s->sample_tile_f[h], s->sample_tiles[h], st, a);
  else
    launch_sample<LG_RNG_PHILOX>(s->sample_tile_f[h], s->sample_tiles[h], st, a);
  LG_LAUNCH_OK();
  RankArgs r;
  r.gid = s->gid[h & 1];
  r.ids = b->ids;
  r.agg_src = b->agg_src;
  r.nc = b->node_counter;
  r.ec = b->edge_counter;
  r.table = s->table;
  r.tile_state = s->rank_state[h];
  r.hs = s->hs + h;
  r.mask = a.mask;
  r.hop = hop;
  r.ids_cap = b->num_ids;
  r.status = s->status;
  r.trace = s->trace;
  rank_relabel_kernel<<<s->rank_tiles[h], kBlock, 0, st>>>(r);
  LG_LAUNCH_OK();
  return 0;
}

extern "C" int lg_io_submit(lg_sampler*, lg_stream_t, int32_t, const lg_batch*) { return 0; }

extern "C" int lg_io_complete(lg_sampler* s, lg_stream_t stream_, int32_t mode, const lg_batch* b,
                              unsigned long long* node_hotness, int32_t* max_ids) {
  LG_REQUIRE(s && b, "lg_io_complete: null argument");
  if (mode != LG_TRAINMODE) return 0;  // :558
  if (node_hotness) {
    hotness_measure_kernel<<<kSMs * 2, kBlock, 0, (cudaStream_t)stream_>>>(b->ids, b->node_counter,
                                                                          (u64*)node_hotness, max_ids);
    LG_LAUNCH_OK();
  }
  return 0;
}

extern "C" int lg_run_batch(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                            const lg_feature_cache* cache, const lg_batch_params* p, const lg_batch* b,
                            unsigned long long* tier_rows) {
  LG_REQUIRE(s && topo && p && b, "lg_run_batch: null argument");
  cudaStream_t main_st = (cudaStream_t)stream;
  const bool fork = cache && s->overlap;
  const bool pipelined = cache && s->overlap == 2;
  cudaEvent_t ev_done = nullptr;
  if (pipelined) {  // these buffers may still be written by the gathers of the batch that used them last
    int rc0 = done_event(s, b, true, &ev_done);
    if (rc0) return rc0;
    LG_CUDA(cudaStreamWaitEvent(main_st, ev_done, 0));
  }
  // gathers go to the side stream (HBM-bound) while the next hop is sampled on the caller's stream
  // (latency-bound): both kinds of kernels are resident on the SMs at once.
  lg_stream_t gst = fork ? (lg_stream_t)s->side : stream;
  if (fork) {  // the side stream must not start before the caller's earlier work (previous batch) is done
    LG_CUDA(cudaEventRecord(s->ev_join, main_st));
    LG_CUDA(cudaStreamWaitEvent(s->side, s->ev_join, 0));
  }
  int rc = lg_batch_generate(s, stream, p->all_ids, p->all_labels, p->total_cap, p->batch_size, p->counter, b);
  if (rc) return rc;
  int first_pending = 0;  // first hop whose rows have not been gathered yet
  for (int hop = 0; hop <= s->n_hops; hop++) {
    if (hop > 0) {
      rc = lg_random_sample(s, stream, topo, hop, p->rng_kind, p->rng_seed, p->batch_id, p->stream_id, b, nullptr);
      if (rc) return rc;
    }
    if (!cache) continue;
    // gather fusion: the final counters equal the reference's (the last lookup op's snapshot wins);
    // only the NUMBER of gather launches changes
    const bool last = (hop == s->n_hops);
    if (!last && (s->fuse_gathers == 2 || (s->fuse_gathers == 1 && hop == 0))) continue;
    if (fork) {
      LG_CUDA(cudaEventRecord(s->ev_fork[hop], main_st));
      LG_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork[hop], 0));
    }
    rc = lg_feature_cache_lookup_range(s, gst, cache, hop * LG_INTRABATCH_CON + 1, first_pending, p->local_part, b,
                                       tier_rows);
    if (rc) return rc;
    first_pending = hop + 1;
  }
  if (pipelined) {  // no join: the next batch is sampled while this batch's last gather streams
    LG_CUDA(cudaEventRecord(ev_done, s->side));
  } else if (fork) {  // join: the batch is complete on the caller's stream
    LG_CUDA(cudaEventRecord(s->ev_join, s->side));
    LG_CUDA(cudaStreamWaitEvent(main_st, s->ev_join, 0));
  }
  return lg_io_complete(s, stream, p->mode, b, nullptr, nullptr);
}

static int run_batch_host_impl(lg_sampler* s, lg_stream_t stream_, const lg_topology* topo,
                               const lg_feature_cache* cache, const lg_batch_params* p, const int32_t* host_seed_ids,
                               const int32_t* host_seed_labels, const lg_batch* b, int32_t* host_node_counter,
                               int32_t* host_edge_counter, bool sync) {
  LG_REQUIRE(s && p && b && host_seed_ids && host_seed_labels && host_node_counter && host_edge_counter,
             "lg_run_batch_host: null argument");
  LG_REQUIRE(p->batch_size <= s->max_batch, "lg_run_batch_host: batch %d > max_batch %d", p->batch_size, s->max_batch);
  cudaStream_t st = (cudaStream_t)stream_;
  // seeds arrive from the host: stage them in the (reused) head of gid[1], which hop 1 does not read
  int32_t* d_seeds = s->gid[1];
  int32_t* d_labels = s->gid[1] + s->max_batch;
  const size_t nb = (size_t)p->batch_size * sizeof(int32_t);
  LG_CUDA(cudaMemcpyAsync(d_seeds, host_seed_ids, nb, cudaMemcpyHostToDevice, st));
  LG_CUDA(cudaMemcpyAsync(d_labels, host_seed_labels, nb, cudaMemcpyHostToDevice, st));
  lg_batch_params q = *p;
  q.all_ids = d_seeds;
  q.all_labels = d_labels;
  q.total_cap = p->batch_size + 1;  // strictly inside the set: no tail clipping
  q.counter = 0;
  int rc = lg_run_batch(s, stream_, topo, cache, &q, b, nullptr);
  if (rc) return rc;
  // what get_next reads back (training_backend/ipc_cuda_kernel.cu:194-195); in pipelined mode the counters
  // are final once the sampler (caller's stream) is done, the features once lg_batch_wait has passed
  rc = lg_batch_wait(s, stream_, b);
  if (rc) return rc;
  LG_CUDA(cudaMemcpyAsync(host_node_counter, b->node_counter, LG_COUNTER_SLOTS * sizeof(int32_t),
                          cudaMemcpyDeviceToHost, st));
  LG_CUDA(cudaMemcpyAsync(host_edge_counter, b->edge_counter, LG_COUNTER_SLOTS * sizeof(int32_t),
                          cudaMemcpyDeviceToHost, st));
  if (sync) LG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int lg_run_batch_host(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                                 const lg_feature_cache* cache, const lg_batch_params* p,
                                 const int32_t* host_seed_ids, const int32_t* host_seed_labels, const lg_batch* b,
                                 int32_t* host_node_counter, int32_t* host_edge_counter) {
  return run_batch_host_impl(s, stream, topo, cache, p, host_seed_ids, host_seed_labels, b, host_node_counter,
                             host_edge_counter, true);
}

extern "C" int lg_run_batch_host_async(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                                       const lg_feature_cache* cache, const lg_batch_params* p,
                                       const int32_t* host_seed_ids, const int32_t* host_seed_labels,
                                       const lg_batch* b, int32_t* host_node_counter, int32_t* host_edge_counter) {
  return run_batch_host_impl(s, stream, topo, cache, p, host_seed_ids, host_seed_labels, b, host_node_counter,
                             host_edge_counter, false);
}

// sampler.cu — k-hop neighbour sampler of the Legion data path for sm_100a.
//
// Replaces, per hop, the reference's chain  FindTopo (2 bcht probes) -> random_sample ->
// construct_graph -> counter_update (+ 2 blocking D2H copies)   [engine/operator_impl.cu:175-296,
// 400-499; cache/cache.cu:217-225]  with two or three launches and no host synchronisation:
//
//   sample_hop_kernel   tile of frontier entries -> row lookup (directory -> HBM shard / peer
//                       shard / host UVA), per-entry edge count min(deg, fanout), block scan + two-level
//                       prefix over the earlier tiles for the canonical edge offsets, with-replacement
//                       pick (Philox4x32-10 or the reference's minstd stream), edge emission in
//                       ascending slot order, RED.MIN of (kNewBit | first edge position) into the
//                       position map word of the sampled vertex.  For hop > 1 it also writes the
//                       previous hop's agg_src (construct_graph) — it reads those words anyway.
//   rank_kernel         first-occurrence flags -> scan -> batch-local ids in first-seen order, `ids`
//                       append, position-map publish (not after the last hop), the op's counter_update by
//                       the last CTA; optionally agg_src of the hop: the local id where this pass knows it,
//                       ~p_first for later occurrences of vertices that are new in the hop.
//   relabel_kernel      where rank wrote agg_src: agg_src[e] = agg_src[p_first] for the ~p_first entries.
//
// The kernels of a batch are chained with programmatic dependent launch (pdl_prologue, common.cuh).
//
// Dedup state ("position map": vertex -> kNewBit | first edge position while a hop is open, batch-local id
// afterwards), two layouts chosen per handle by the size of the graph (lg_sampler_create):
//   DENSE   one 32-bit word per vertex (the reference's position_map, engine/server.cu:224).  Every access is one
//           4-byte load, store or fire-and-forget RED.MIN: no hashing, no probes.  No accessed-bitmap and no
//           per-batch O(N) memset (engine/operator_impl.cu:151): words are released by an O(batch) pass at the
//           end of the batch (ClearPosMap, :542-548).  Best while 4N bytes stay L2-resident (products: 10 MB).
//   HASHED  O(batch) open-addressing table of (vertex << 32 | value) words, 2^k >= 1.5 x num_ids slots (32 MB for
//           B=8000, [25,10]), L2-resident whatever N is.  With a 534 MB (UK-Union) or 3.8 GB (Clueweb) dense map
//           every one of the ~10 M map accesses of a batch is a random DRAM access; the table keeps them in L2.
//           Insert-min is ONE returning atomicMin per probe on the packed word: an empty slot (all ones) or the
//           same key take the minimum directly; a larger resident key is displaced and carried to the next slot
//           (linear probing, no CAS, no pre-read).  The table is re-initialised by one 32 MB streaming fill per batch
//           (release_map, run by the batch's last kernel).
#include "common.cuh"
#include "sampler_state.cuh"
#include "scan.cuh"

#include <mutex>

using namespace lg;

namespace {

constexpr uint32_t kPmEmpty = 0xFFFFFFFFu;  // position-map word of a vertex not in the batch
constexpr uint32_t kNewBit = 0x80000000u;   // kNewBit | first edge position while a hop is open; final local ids are < 2^31
constexpr int kSlotUnroll = 5;              // neighbour reads in flight per thread in sample_hop_kernel
constexpr int kSlotUnrollHashed = 5;        // HASHED: 10 in flight was measured slower (registers): 0.195 vs 0.174 ms for hop 2 at UK-Union scale

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


// ------------------------------------------------------------------------------------------
// Position map, dense or hashed (see the header comment).  kPmEmpty is "not in the batch" in both layouts
// (the low word of an empty table slot is all ones too).
// ------------------------------------------------------------------------------------------
constexpr u64 kSlotEmpty = 0xFFFFFFFFFFFFFFFFull;
struct DedupMap {
  uint32_t* pm;   // DENSE: [num_nodes]
  u64* table;     // HASHED: [mask + 1]
  uint32_t mask;
};
__device__ __forceinline__ u64 atom_min_u64_hint(u64* p, u64 v, u64 pol) {
  u64 old;
  asm volatile("atom.relaxed.gpu.global.min.L2::cache_hint.u64 %0, [%1], %2, %3;" : "=l"(old) : "l"(p), "l"(v), "l"(pol) : "memory");
  return old;
}
__device__ __forceinline__ u64 ld_ca_u64_hint(const u64* p, u64 pol) {
  u64 v;
  asm volatile("ld.global.ca.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_u64_hint(u64* p, u64 v, u64 pol) {
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ u64 map_pack(uint32_t v, uint32_t val) { return ((u64)v << 32) | val; }
__device__ __forceinline__ uint32_t map_home(const DedupMap& m, uint32_t v) { return hash32(v) & m.mask; }
// HASHED insert-min, continued after the first probe (`old` = what the atomicMin at `slot` returned)
__device__ __forceinline__ void table_insert_finish(const DedupMap& m, uint32_t slot, u64 carry, u64 old, u64 pol) {
  while (true) {
    if (old == kSlotEmpty || (uint32_t)(old >> 32) == (uint32_t)(carry >> 32)) return;  // claimed / merged
    if (old > carry) carry = old;  // a larger key lived here: our word took its place, it moves on
    slot = (slot + 1) & m.mask;
    old = atom_min_u64_hint(m.table + slot, carry, pol);
  }
}
template <bool HASHED>
__device__ __forceinline__ void map_insert_min(const DedupMap& m, uint32_t v, uint32_t val, u64 pol) {
  if (HASHED) {
    const uint32_t slot = map_home(m, v);
    const u64 carry = map_pack(v, val);
    table_insert_finish(m, slot, carry, atom_min_u64_hint(m.table + slot, carry, pol), pol);
  } else {
    red_min_u32_hint(m.pm + v, val, pol);
  }
}
// HASHED lookup, continued after the first probe; returns the value (kPmEmpty if absent), *slot = where it lives
__device__ __forceinline__ uint32_t table_find_finish(const DedupMap& m, uint32_t v, uint32_t* slot, u64 cur, u64 pol) {
  while ((uint32_t)(cur >> 32) != v && cur != kSlotEmpty) {
    *slot = (*slot + 1) & m.mask;
    cur = ld_ca_u64_hint(m.table + *slot, pol);
  }
  return (uint32_t)cur;
}
// L1-cached (see the rank kernel for why a stale line is harmless)
template <bool HASHED>
__device__ __forceinline__ uint32_t map_lookup(const DedupMap& m, uint32_t v, u64 pol) {
  if (HASHED) {
    uint32_t slot = map_home(m, v);
    return table_find_finish(m, v, &slot, ld_ca_u64_hint(m.table + slot, pol), pol);
  }
  return ld_ca_u32_hint(m.pm + v, pol);
}

// ------------------------------------------------------------------------------------------
// batch_generate + op-0 counter_update (engine/operator_impl.cu:27-89,159-165)
// ------------------------------------------------------------------------------------------
struct BatchGenArgs {
  const int32_t* all_ids;
  const int32_t* all_labels;
  int32_t total_cap, size, stride, counter, hop_num;
  int32_t* ids;
  int32_t* labels;
  int32_t* nc;
  int32_t* ec;
  DedupMap map;
  u64* small;  // per-batch scan state (tickets + tile words), zeroed here
  int32_t small_words;
  int32_t l2;
};
// One pass over the launch's threads (t = global thread index of n_threads): works for any grid size.
template <bool HASHED>
__device__ __forceinline__ void batch_generate_body(const BatchGenArgs& g, int32_t t, int32_t n_threads) {
  const u64 keep = l2_policy((g.l2 & 4) ? 1 : 0);
  // reset of the per-batch scan state, a few KB: done here instead of a memset node so that the whole chain is
  // kernel -> kernel (the reference memsets an N/8-byte bitmap at this point, :151)
  for (int32_t i = t; i < g.small_words; i += n_threads) g.small[i] = 0ull;
  if (t < LG_COUNTER_SLOTS) {
    int32_t v = 0;
    if (t == 1 || t == LG_INTRABATCH_CON * 3) v = g.size;  // nc[1], nc[9]
    if (t == LG_INTRABATCH_CON * 3 - 1) v = g.hop_num;     // nc[8]
    g.nc[t] = v;
    g.ec[t] = 0;
  }
  for (int32_t idx = t; idx < g.size; idx += n_threads) {
    // offset of the batch in the set: batch_size * counter.  The reference passes the CLIPPED size of a tail batch as the
    // kernel's batch_size (:159-162), so its last batch re-reads seeds from the middle of the set (:40,:44);
    // lg_sampler_set_tail_mode(LG_TAIL_REFERENCE) keeps that stride (g.stride = g.size)
    const long long pos = (long long)g.stride * g.counter + idx;
    if (pos >= g.total_cap) {
      g.ids[idx] = -1;
      g.labels[idx] = -1;
      continue;
    }
    const int32_t v = g.all_ids[pos % g.total_cap];
    g.ids[idx] = v;
    g.labels[idx] = g.all_labels[pos % g.total_cap];
    if (v >= 0) map_insert_min<HASHED>(g.map, (uint32_t)v, (uint32_t)idx, keep);  // position_map (:51): local index = first position
  }
}
template <bool HASHED>
__global__ void __launch_bounds__(kBlock) batch_generate_kernel(const BatchGenArgs g) {
  pdl_prologue();
  batch_generate_body<HASHED>(g, (int32_t)(blockIdx.x * blockDim.x + threadIdx.x), (int32_t)(gridDim.x * blockDim.x));
}

// HASHED only: batch-local id of every seed (= its first position; duplicates share it).  The hashed insert moves
// resident words (a displaced key is in flight between two slots for a moment), so no kernel may look a vertex up
// while another CTA of the same launch inserts: the hop-1 sample kernel reads these ids instead of the table, and
// later hops read the previous hop's relabelled agg_src.
__global__ void __launch_bounds__(kBlock) seed_local_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ nc,
                                                            int32_t* __restrict__ seed_local, const DedupMap map, int32_t l2) {
  pdl_prologue();
  const u64 keep = l2_policy((l2 & 4) ? 1 : 0);
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld_counter(nc + 1)) return;
  const int32_t v = ids[i];
  seed_local[i] = v >= 0 ? (int32_t)map_lookup<true>(map, (uint32_t)v, keep) : 0;
}

// ------------------------------------------------------------------------------------------
// sample_hop_kernel
// ------------------------------------------------------------------------------------------
struct SampleHop {  // what changes from hop to hop
  const int32_t* frontier_prev;  // hop > 1: global ids of the previous hop's sampled sources
  int32_t* gid_out;              // this hop's sampled sources (global ids), hop-relative positions
  u64* tile_state;
  u64* anchors;                  // group words of the tile prefix
  HopState* hs;
  int32_t hop;
  int32_t fanout;
  uint32_t fanout_magic;  // ceil(2^32 / fanout): k / fanout as one multiply-high, exact for every k < 256 * fanout while
                          // fanout <= 4096 (k * (magic * fanout - 2^32) < 2^32); 0 = larger fan-out, plain division
  int32_t relabel_prev;   // also write the previous hop's agg_src (its construct_graph) from the position map
};
struct SampleArgs {
  lg_topology topo;
  const int32_t* seed_local;     // HASHED: batch-local ids of the seeds (seed_local_kernel)
  int32_t* ids;
  int32_t* agg_src;
  int32_t* agg_dst;
  int32_t* nc;
  int32_t* ec;
  DedupMap map;
  u64* edge_hot;
  int32_t* status;        // sticky status word of the handle (4 = edge_dst holds an id outside [0, num_nodes))
  int32_t persistent;     // the CTAs loop over tickets (grid smaller than the number of tiles)
  int32_t precheck;       // DENSE: L1-cached look at the map word before the RED.MIN (LG_RED_PRECHECK)
  uint32_t batch_id, stream_id, k0, k1;
  int32_t l2;  // lg_l2_hints()
  u64* trace;
  SampleHop h;
};

// Row of vertex v: (part, first edge, degree) through the topology directory (FindTopo, cache/cache.cu:217-225)
__device__ __forceinline__ void row_locate(const lg_topology& t, int32_t v, int32_t loc, int* part, long long* row) {
  *part = t.n_parts;
  *row = v;
  if (loc >= 0) {
    *part = loc / t.shard_rows;
    *row = loc - *part * t.shard_rows;
  }
}

template <int TILE_F>
struct SampleSmem {
  long long start[TILE_F];
  const int32_t* indices[TILE_F];
  int32_t deg[TILE_F];
  int32_t cnt[TILE_F];
  int32_t off[TILE_F];
  int32_t flocal[TILE_F];
  int32_t red[kBlock / 32];
};

// One tile of TILE_F frontier entries, by the whole CTA.  Returns false when `tile` is past the hop's last tile.
template <int TILE_F, int RNG, bool HASHED, int U_OVERRIDE = 0>
__device__ __forceinline__ bool sample_tile(const SampleArgs& a, const SampleHop& h, const int tile, SampleSmem<TILE_F>& sm) {
  static_assert(TILE_F <= kBlock, "one thread per frontier entry of the tile");
  long long* const s_start = sm.start;
  const int32_t** const s_indices = sm.indices;
  int32_t* const s_deg = sm.deg;
  int32_t* const s_cnt = sm.cnt;
  int32_t* const s_off = sm.off;
  int32_t* const s_flocal = sm.flocal;
  int32_t* const s_red = sm.red;

  constexpr int U = U_OVERRIDE ? U_OVERRIDE : (HASHED ? kSlotUnrollHashed : kSlotUnroll);  // slots in flight per thread
  const int tid = threadIdx.x;
  const u64 keep = l2_policy((a.l2 & 4) ? 1 : 0), once = l2_policy((a.l2 & 8) ? 2 : 0);
  const bool first_hop = (h.hop == 1);
  const int32_t ec0 = ld_counter(a.ec), ec1 = ld_counter(a.ec + 1);
  const int32_t F = first_hop ? ld_counter(a.nc + 1) : ec1;  // :201-206
  const int32_t prev_edge_off = ec0;
  const int32_t edge_base = ec0 + ec1;                       // :275
  const int tslot = (h.hop - 1) * 2;
  if (tid == 0) trace_mark(a.trace, tslot, tile, 0);
  const int n_tiles = (F + TILE_F - 1) / TILE_F;
  if (tile >= n_tiles) {
    if (n_tiles == 0 && tile == 0 && tid == 0) a.ec[2] = 0;
    return false;
  }
  const int32_t c = h.fanout;
  const int32_t i0 = tile * TILE_F;

  // 1. row lookup for the tile's frontier entries
  int32_t cnt = 0, fl = 0;
  bool live = false, patched = false;
  if (tid < TILE_F) {
    const int32_t i = i0 + tid;
    int32_t deg = 0;
    long long start = 0;
    const int32_t* ind = a.topo.indices[a.topo.n_parts];
    if (i < F) {
      const int32_t v = first_hop ? a.ids[i] : h.frontier_prev[i];
      if (v >= 0) {
        live = true;
        // batch-local index of the frontier vertex (position_map, :291-294): final since the previous op.  The load
        // is issued here and consumed after the prefix (s_flocal).
        if (HASHED) {  // never from the table while this launch inserts (see seed_local_kernel)
          fl = first_hop ? a.seed_local[i] : a.agg_src[prev_edge_off + i];
          if (fl < 0) {  // the previous hop's rank pass left ~p_first: the first occurrence's entry holds the id
            fl = a.agg_src[prev_edge_off + ~fl];  // (construct_graph of the previous hop, second half, fused)
            patched = true;
          }
        } else {
          fl = (int32_t)map_lookup<false>(a.map, (uint32_t)v, keep);
        }
        int part;
        long long row;
        row_locate(a.topo, v, a.topo.directory ? ld_nc_s32_hint(a.topo.directory + v, keep) : -1, &part, &row);
        const int64_t* ip = a.topo.indptr[part];
        start = ld_nc_s64_hint(ip + row, keep);
        deg = (int32_t)(ld_nc_s64_hint(ip + row + 1, keep) - start);  // :226 (int32 col_size)
        ind = a.topo.indices[part];
        cnt = deg < c ? deg : c;
        if (cnt < 0) cnt = 0;
        if (a.edge_hot && cnt > 0) atomicAdd(a.edge_hot + v, (u64)cnt);  // pre_sample :358, summed per entry
      }
    }
    s_start[tid] = start;
    s_deg[tid] = deg;
    s_cnt[tid] = cnt;
    s_indices[tid] = ind;
  }
  if (tid == 0) trace_mark(a.trace, tslot, tile, 1);

  // 2. exclusive scan of the per-entry edge counts inside the tile, 3. prefix over the earlier tiles
  int32_t total;
  const int32_t off = block_exclusive_scan(cnt, s_red, &total);
  if (tid < TILE_F) s_off[tid] = off;
  const int32_t base = block_exclusive_prefix(h.tile_state, h.anchors, tile, total, s_red);
  if (tid == 0 && tile == n_tiles - 1) a.ec[2] = base + total;  // E_h (:264)
  if (tid < TILE_F) {
    s_flocal[tid] = fl;
    // construct_graph of the previous hop, fused: every entry (dense), or the ~p_first entries (hashed)
    if (HASHED ? patched : (h.relabel_prev && live)) a.agg_src[prev_edge_off + i0 + tid] = fl;
  }
  __syncthreads();
  if (tid == 0) trace_mark(a.trace, tslot, tile, 2);

  // 4. one thread per slot: pick, emit, min-insert into the position map.  U independent neighbour
  //    reads are issued back to back before any of them is consumed (random 4-byte HBM/NVLink/PCIe accesses);
  //    the position-map update is a fire-and-forget RED, so nothing in this loop waits on a second round trip.
  const int n_slots = TILE_F * c;
  for (int k0 = tid; k0 < n_slots; k0 += kBlock * U) {
    int32_t w[U], p[U], fl[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = k0 + u * kBlock;
      p[u] = -1;
      if (k < n_slots) {
        const int t = (c == 1) ? k : (h.fanout_magic ? (int)__umulhi((uint32_t)k, h.fanout_magic) : k / c);
        const int j = k - t * c;
        if (j < s_cnt[t]) {  // :232  neighbor_offset >= col_size -> none
          const uint32_t slot = (uint32_t)(i0 + t) * (uint32_t)c + (uint32_t)j;
          const int32_t pick =
              pick_neighbor<RNG>(slot, s_deg[t], (uint32_t)h.hop, a.batch_id, a.stream_id, a.k0, a.k1);
          w[u] = ld_nc_s32_hint(s_indices[t] + s_start[t] + pick, once);  // :240-242
          // the reference drops an edge whose neighbour id is negative (:244) and reads out of bounds for ids >= N; here
          // the edge positions are fixed before the pick is read, so such a dataset is flagged (sticky status 4) and the
          // slot falls back to vertex 0 — every later access stays in bounds, the batch is marked invalid
          if ((uint32_t)w[u] >= (uint32_t)a.topo.num_nodes) {
            *a.status = 4;
            w[u] = 0;
          }
          p[u] = base + s_off[t] + j;
          fl[u] = s_flocal[t];
        }
      }
    }
    if (HASHED) {  // all first probes of the group in flight together, then the (rare) continuations
      u64 carry[U], old[U];
      uint32_t sl[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p[u] >= 0) {
          h.gid_out[p[u]] = w[u];
          a.agg_dst[edge_base + p[u]] = fl[u];  // construct_graph :292,294
          sl[u] = map_home(a.map, (uint32_t)w[u]);
          carry[u] = map_pack((uint32_t)w[u], kNewBit | (uint32_t)p[u]);
          if (a.precheck & 2) {  // L1-cached look first: a word that already holds this vertex with an earlier position needs no atomic
            const u64 cur = ld_ca_u64_hint(a.map.table + sl[u], keep);
            old[u] = ((uint32_t)(cur >> 32) == (uint32_t)w[u] && cur <= carry[u]) ? cur
                                                                                   : atom_min_u64_hint(a.map.table + sl[u], carry[u], keep);
          } else {
            old[u] = atom_min_u64_hint(a.map.table + sl[u], carry[u], keep);
          }
        }
      }
      // continuations (collision with another key, or a displaced word to carry on) advance in lockstep: every
      // pending chain of the thread issues its next atomic before any of them is waited for
      bool pend[U];
      bool any = false;
#pragma unroll
      for (int u = 0; u < U; u++) {
        pend[u] = false;
        if (p[u] >= 0 && !(old[u] == kSlotEmpty || (uint32_t)(old[u] >> 32) == (uint32_t)(carry[u] >> 32))) {
          pend[u] = true;
          any = true;
          if (old[u] > carry[u]) carry[u] = old[u];
          sl[u] = (sl[u] + 1) & a.map.mask;
        }
      }
      while (any) {
#pragma unroll
        for (int u = 0; u < U; u++)
          if (pend[u]) old[u] = atom_min_u64_hint(a.map.table + sl[u], carry[u], keep);
        any = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (pend[u]) {
            if (old[u] == kSlotEmpty || (uint32_t)(old[u] >> 32) == (uint32_t)(carry[u] >> 32)) {
              pend[u] = false;
            } else {
              any = true;
              if (old[u] > carry[u]) carry[u] = old[u];
              sl[u] = (sl[u] + 1) & a.map.mask;
            }
          }
        }
      }
    } else if (a.precheck & 1) {
      // a word that already holds an earlier position (or a final id) needs no RED: hub vertices are sampled thousands
      // of times per hop, and their REDs serialise on one L2 slice.  A stale L1 line only holds a LARGER value than
      // the word has now (values only decrease while a hop is open), so skipping on `cur <= val` is always right.
      uint32_t cur[U];
#pragma unroll
      for (int u = 0; u < U; u++) cur[u] = (p[u] >= 0) ? ld_ca_u32_hint(a.map.pm + w[u], keep) : 0u;
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p[u] >= 0) {
          h.gid_out[p[u]] = w[u];
          a.agg_dst[edge_base + p[u]] = fl[u];  // construct_graph :292,294
          const uint32_t val = kNewBit | (uint32_t)p[u];
          if (cur[u] > val) red_min_u32_hint(a.map.pm + w[u], val, keep);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p[u] >= 0) {
          h.gid_out[p[u]] = w[u];
          a.agg_dst[edge_base + p[u]] = fl[u];  // construct_graph :292,294
          red_min_u32_hint(a.map.pm + w[u], kNewBit | (uint32_t)p[u], keep);
        }
      }
    }
  }
  if (a.trace) {
    if (tid == 0) trace_mark(a.trace, tslot, tile, 3);  // thread 0 done
    __syncthreads();
    if (tid == 0) trace_mark(a.trace, tslot, tile, 4);  // whole CTA done
  }
  return true;
}

template <int TILE_F, int RNG, bool HASHED, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) sample_hop_kernel(const SampleArgs a) {
  __shared__ SampleSmem<TILE_F> sm;
  __shared__ int32_t s_tile;
  pdl_prologue();
  // One tile per CTA (grid = tiles), or — LG_SAMPLE_CTAS_PER_SM=k — a grid of 148 x k CTAs that keep claiming tiles: the
  // kernel then never holds more than k CTAs' worth of registers per SM, which is what leaves room for the gather's CTAs
  // on every SM while both run (64 registers x 256 threads x 4 CTAs is the whole register file).
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.h.hs->sample_ticket, 1);  // tiles are claimed in order (see block_exclusive_prefix)
    __syncthreads();
    // HASHED with a register cap (MINB >= 5: 47 / 40 registers): fewer slots in flight per thread, more CTAs per SM
    const bool more = sample_tile<TILE_F, RNG, HASHED, (HASHED && MINB >= 5) ? 3 : 0>(a, a.h, s_tile, sm);
    if (!a.persistent || !more) break;
    __syncthreads();  // the tile's shared memory and s_tile are free again
  }
}

// ------------------------------------------------------------------------------------------
// rank_kernel: first-occurrence flags -> scan -> batch-local ids in first-seen order, `ids` append, publish
// ------------------------------------------------------------------------------------------
struct RankHop {  // what changes from hop to hop
  const int32_t* gid;  // this hop's sampled sources
  u64* tile_state;
  u64* anchors;
  HopState* hs;
  int32_t* agg_src;  // non-null: also write this hop's agg_src (construct_graph) as far as this pass knows it
  int32_t hop;
};
struct RankArgs {
  int32_t* ids;
  int32_t* nc;
  int32_t* ec;
  DedupMap map;
  int32_t ids_cap;
  int32_t l2;
  int32_t persistent;
  int32_t* status;
  u64* trace;
  RankHop h;
};

// PUBLISH: write the final local ids back into the map — needed while a later hop will insert into it (its RED.MIN
// must lose against a final id) or look frontier vertices up; not after the last hop.
// agg_src (construct_graph, :283-296) is resolved here for two of three kinds of edges: the source was already in the
// batch before this hop (the map word IS its local id), or this edge is the source's first occurrence (it gets the id
// assigned below).  A later occurrence of a vertex that is new in this hop only knows the position p_first of the first
// one: it stores ~p_first, and relabel_kernel replaces it by agg_src[p_first] once every tile is done.
// One tile of kBlock * ITEMS edges, by the whole CTA.  Returns false when `tile` is past the hop's last tile.
template <int ITEMS, bool HASHED, bool PUBLISH>
__device__ __forceinline__ bool rank_tile(const RankArgs& a, const RankHop& h, const int tile, int32_t* s_red) {
  static_assert(ITEMS % 4 == 0, "edges are loaded as int4");
  constexpr int TILE = kBlock * ITEMS;
  const int tid = threadIdx.x;
  const u64 keep = l2_policy((a.l2 & 4) ? 1 : 0);
  const int32_t E = ld_counter(a.ec + 2);
  const int32_t node_base = ld_counter(a.nc) + ld_counter(a.nc + 1);  // :268
  const int32_t edge_base = ld_counter(a.ec) + ld_counter(a.ec + 1);  // :275 (the counters move on when every tile is done)
  const int tslot = (h.hop - 1) * 2 + 1;
  if (tid == 0) trace_mark(a.trace, tslot, tile, 0);
  const int n_tiles = (E + TILE - 1) / TILE;
  if (tile >= n_tiles) return false;
  {
    const int32_t p0 = tile * TILE + tid * ITEMS;  // ITEMS consecutive edges per thread
    // The edges are looked up in chunks of CH (all probes of a chunk in flight together) and only a bit per edge
    // survives the chunk.  HASHED: chunks of 8 — with the vertices, 64-bit table words and slots of all ITEMS edges
    // live at once the kernel needed 60-120 registers and its tiles ran in 1.65 waves.  DENSE: one chunk (every lookup
    // of the thread in flight together; chunks of 4 were measured 10 % slower per launch).
    constexpr int CH = HASHED ? ((ITEMS % 8 == 0) ? 8 : 4) : ITEMS;
    uint32_t mask = 0;
#pragma unroll 1
    for (int c = 0; c < ITEMS; c += CH) {
      int32_t w[CH];
      uint32_t q[CH];
#pragma unroll
      for (int k = 0; k < CH; k += 4) {  // the buffer is padded to a multiple of TILE, reads past E are discarded
        const int4 v = *reinterpret_cast<const int4*>(h.gid + p0 + c + k);
        w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
      }
      // L1-cached: a line fetched before the owner publishes holds kNewBit|p_first, one fetched after holds the
      // final id — neither can equal kNewBit|p for a non-owner, and an owner's word is only rewritten by itself
      if (HASHED) {
        u64 cur[CH];
        uint32_t sl[CH];
#pragma unroll
        for (int k = 0; k < CH; k++) {  // first probes in flight together
          sl[k] = map_home(a.map, (uint32_t)w[k]);
          cur[k] = (p0 + c + k < E) ? ld_ca_u64_hint(a.map.table + sl[k], keep) : 0ull;
        }
        // continuations (the home slot holds another key) advance in lockstep: every pending probe chain of the thread
        // issues its next load before any of them is waited for — finished one after the other, the chains of a thread's
        // 8 edges added up, and the slowest thread of the CTA (the block scan waits for it) took 12 us per tile
        // (profiles/r02v_rank_lockstep.md)
        unsigned pend = 0;
#pragma unroll
        for (int k = 0; k < CH; k++)
          if (p0 + c + k < E && (uint32_t)(cur[k] >> 32) != (uint32_t)w[k] && cur[k] != kSlotEmpty) pend |= 1u << k;
        while (pend) {
#pragma unroll
          for (int k = 0; k < CH; k++)
            if (pend & (1u << k)) {
              sl[k] = (sl[k] + 1) & a.map.mask;
              cur[k] = ld_ca_u64_hint(a.map.table + sl[k], keep);
            }
#pragma unroll
          for (int k = 0; k < CH; k++)
            if ((pend & (1u << k)) && ((uint32_t)(cur[k] >> 32) == (uint32_t)w[k] || cur[k] == kSlotEmpty)) pend &= ~(1u << k);
        }
#pragma unroll
        for (int k = 0; k < CH; k++) q[k] = (p0 + c + k < E) ? (uint32_t)cur[k] : 0u;
      } else {
#pragma unroll
        for (int k = 0; k < CH; k++) q[k] = (p0 + c + k < E) ? ld_ca_u32_hint(a.map.pm + w[k], keep) : 0u;
      }
#pragma unroll
      for (int k = 0; k < CH; k++) {
        const int32_t p = p0 + c + k;
        if (p < E) {
          if (q[k] == (kNewBit | (uint32_t)p)) mask |= 1u << (c + k);
          else if (h.agg_src)  // everything but the first occurrences is known now: the local id, or ~p_first
            h.agg_src[edge_base + p] = (q[k] < kNewBit) ? (int32_t)q[k] : ~(int32_t)(q[k] & ~kNewBit);
        }
      }
    }
    if (tid == 0) trace_mark(a.trace, tslot, tile, 1);
    int32_t total;
    const int32_t mine = block_exclusive_scan(__popc(mask), s_red, &total);
    if (tid == 0) trace_mark(a.trace, tslot, tile, 2);
    const int32_t excl = block_exclusive_prefix(h.tile_state, h.anchors, tile, total, s_red);
    if (tid == 0 && tile == n_tiles - 1) h.hs->new_nodes = excl + total;  // C_h (:263)
    if (tid == 0) trace_mark(a.trace, tslot, tile, 3);
    int32_t local = node_base + excl + mine;
    if (!PUBLISH) {
      // last hop: nothing goes back into the map, so the first occurrences only need their vertex ids again — re-read as
      // vectors up front (cached) and consumed by predicated straight-line code; the loop below re-reads one vertex per
      // iteration, a chain of up to ITEMS dependent loads (3.3 us per tile in the trace)
      int32_t wv[ITEMS];
#pragma unroll
      for (int k = 0; k < ITEMS; k += 4) {
        const int4 v = *reinterpret_cast<const int4*>(h.gid + p0 + k);
        wv[k] = v.x; wv[k + 1] = v.y; wv[k + 2] = v.z; wv[k + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < ITEMS; k++) {
        if (mask & (1u << k)) {
          if (local < a.ids_cap) a.ids[local] = wv[k];  // :270
          else *a.status = 1;
          if (h.agg_src) h.agg_src[edge_base + p0 + k] = local;
          local++;
        }
      }
      mask = 0;
    }
    while (mask) {  // first occurrences, in edge order; the vertex is re-read (coalesced, cached) rather than kept
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      const int32_t w = h.gid[p0 + k];
      if (local < a.ids_cap) a.ids[local] = w;  // :270
      else *a.status = 1;
      if (PUBLISH) {  // position_map :271
        if (HASHED) {
          uint32_t slot = map_home(a.map, (uint32_t)w);
          table_find_finish(a.map, (uint32_t)w, &slot, ld_ca_u64_hint(a.map.table + slot, keep), keep);
          st_u64_hint(a.map.table + slot, map_pack((uint32_t)w, (uint32_t)local), keep);
        } else {
          st_u32_hint(a.map.pm + w, (uint32_t)local, keep);
        }
      }
      if (h.agg_src) h.agg_src[edge_base + p0 + k] = local;
      local++;
    }
    if (tid == 0) trace_mark(a.trace, tslot, tile, 4);
  }
  return true;
}

// the op's counter_update (:69-82), by one thread once every tile of the hop is done
__device__ __forceinline__ void rank_counter_update(const RankArgs& a, const RankHop& h) {
  volatile int32_t* nc = a.nc;
  volatile int32_t* ec = a.ec;
  const int32_t E = ec[2];
  const int32_t C = (E > 0) ? *((volatile int32_t*)&h.hs->new_nodes) : 0;
  const int32_t nc0 = nc[0] + nc[1];
  nc[0] = nc0;
  nc[1] = C;
  nc[LG_INTRABATCH_CON * 2] = 0;
  nc[LG_INTRABATCH_CON * 2 + 1] = nc0 + C;
  const int32_t ec0 = ec[0] + ec[1];
  ec[0] = ec0;
  ec[1] = E;
  ec[2] = 0;
  nc[LG_INTRABATCH_CON * 3 + h.hop] = nc0 + C;
  ec[LG_INTRABATCH_CON * 3 + h.hop] = ec0 + E;
}

template <int ITEMS, bool HASHED, bool PUBLISH>
__global__ void __launch_bounds__(kBlock) rank_kernel(const RankArgs a) {
  __shared__ int32_t s_red[kBlock / 32];
  __shared__ int32_t s_tile, s_last;
  const int tid = threadIdx.x;
  pdl_prologue();
  for (;;) {  // one tile per CTA, or (a.persistent) a smaller grid that keeps claiming tiles — see sample_hop_kernel
    if (tid == 0) s_tile = atomicAdd(&a.h.hs->rank_ticket, 1);
    __syncthreads();
    const bool more = rank_tile<ITEMS, HASHED, PUBLISH>(a, a.h, s_tile, s_red);
    if (!a.persistent || !more) break;
    __syncthreads();
  }
  // the last CTA to finish updates the counters
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int32_t done = atomicAdd(&a.h.hs->rank_done, 1);
    s_last = (done == (int32_t)gridDim.x - 1);
  }
  __syncthreads();
  if (tid == 0) trace_mark(a.trace, (a.h.hop - 1) * 2 + 1, s_tile, 7);
  if (s_last && tid == 0) {
    __threadfence();
    rank_counter_update(a, a.h);
  }
}

// Release of the batch's position-map words (ClearPosMap, :542-548), as a device-side step that the last kernel of a
// batch can run itself (lg_run_batch): one kernel launch less per batch — every launch on the GPU stalls the gather
// streaming on the other stream for ~4 us, whatever its size (profiles/r01d_overlap.md).
struct ReleaseArgs {
  uint4* fill;  // non-null: streaming fill of n16 16-byte words with all ones (small dense maps, the hashed table)
  int64_t n16;
  uint32_t* pm;  // else non-null: O(batch) scatter over the batch's vertices
  const int32_t* ids;
  const int32_t* nc;
  int32_t l2;
};
__device__ __forceinline__ void release_map(const ReleaseArgs& r) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  if (r.fill) {
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    for (int64_t i = t; i < r.n16; i += nt) r.fill[i] = ones;
  } else if (r.pm) {
    const u64 keep = l2_policy((r.l2 & 4) ? 1 : 0);
    int32_t n = ld_counter(r.nc + LG_INTRABATCH_CON * 2 + 1);
    const int32_t seeds = ld_counter(r.nc + LG_INTRABATCH_CON * 3);
    if (seeds > n) n = seeds;  // before the first hop nc[7] is still 0
    for (int64_t i = t; i < n; i += nt) {
      const int32_t v = r.ids[i];
      if (v >= 0) st_u32_hint(r.pm + v, kPmEmpty, keep);
    }
  }
}

// construct_graph for the sources of the hop just ranked (:283-296), second half: the rank kernel left ~p_first in the
// agg_src entries of later occurrences of vertices that are new in this hop; the first occurrence's entry holds the id.
// Runs after the rank kernel's counter_update: (ec[0], ec[1]) = (offset, count) of the hop's edges.  Entries that are
// read (first occurrences, >= 0) are never written here.  For every hop but the last, lg_run_batch (dense layout) folds
// construct_graph into the next hop's sample kernel instead (which looks the same vertices up anyway).
// After the last hop nobody reads the position map again: lg_run_batch lets this kernel release it (`rel`).
__device__ __forceinline__ void relabel_body(int32_t* __restrict__ agg_src, const int32_t* __restrict__ ec, int64_t t,
                                             int64_t n_threads) {
  const int32_t off = ld_counter(ec), E = ld_counter(ec + 1);
  for (int64_t q0 = t * 4; q0 < E; q0 += n_threads * 4) {
    const int32_t p0 = (int32_t)q0;
    int32_t x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = (p0 + k < E) ? agg_src[off + p0 + k] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++)  // L1-cached: first occurrences of hub vertices are read thousands of times
      y[k] = (x[k] < 0) ? __ldca(agg_src + off + ~x[k]) : x[k];
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (x[k] < 0) agg_src[off + p0 + k] = y[k];
  }
}
__global__ void __launch_bounds__(kBlock) relabel_kernel(int32_t* __restrict__ agg_src, const int32_t* __restrict__ ec,
                                                         const ReleaseArgs rel) {
  pdl_prologue();
  release_map(rel);
  relabel_body(agg_src, ec, (int64_t)blockIdx.x * kBlock + threadIdx.x, (int64_t)gridDim.x * kBlock);
}

// ClearPosMap (:542-548) as its own launch (lg_io_complete of the op-by-op API): the position-map words of this
// batch's vertices go back to "not in the batch".  O(batch) scatter — the map itself is O(N) like the reference's, but
// never memset — or, for maps of a few MB and the hashed table, one streaming fill at store bandwidth (products: 9.8 MB
// against ~0.9 M random 4-byte stores).
__global__ void __launch_bounds__(kBlock) release_kernel(const ReleaseArgs rel) {
  pdl_prologue();
  release_map(rel);
}

// ------------------------------------------------------------------------------------------
// chain_kernel (dense layout): the whole sampler chain of one batch — batch_generate, (sample, rank) per hop, the last
// hop's relabel and the release of the position map — as ONE persistent launch, with grid barriers where the separate
// kernels had launch boundaries.  Why: every kernel launch on the GPU stalls the gather streaming on the other stream
// for ~3.8 us, whatever its size (profiles/r01d_overlap.md); 6 launches per batch become 1.
// 148 x 4 CTAs at <= 40 registers: exactly what fits an SM next to the gather's 12 single-warp CTAs (17 k registers,
// 130 KB of shared memory), so the grid is co-resident whether or not a gather is running and the barriers cannot wait on
// a CTA that has no room.  Tiles are claimed through the same tickets as in the separate kernels (a CTA finishes its tile
// before it claims the next, so the prefix spins still only wait on running CTAs).  Two chain kernels must never be
// partially resident at the same time (each would spin on CTAs the other leaves no room for): the host orders them with
// an event per device (chain_order), and the kernel lets its stream successor launch only when it enters its last phase.
// MEASURED AND NOT ADOPTED (opt-in, LG_CHAIN=1): bit-exact on every parity test, but the chain alone takes 0.141 ms
// against 0.108 ms for the PDL-chained kernels (40-register cap shared by all phases, a ticket round trip per tile, six
// grid barriers), and a persistent grid never lets an SM drain, so whichever of chain and gather arrives second waits for
// the other's shared-memory configuration: pipelined 31.6 M seeds/s (34.1 M with one carve-out for both) against 40.8 M.
// ------------------------------------------------------------------------------------------
constexpr int kChainCtasPerSm = 4;
struct ChainArgs {
  BatchGenArgs gen;
  SampleArgs smp;  // the fields common to all hops; the per-hop ones are patched in the kernel
  RankArgs rnk;
  ReleaseArgs rel;
  int32_t n_hops;
  int32_t fanout[LG_MAX_HOPS];
  uint32_t magic[LG_MAX_HOPS];
  int32_t tile_f[LG_MAX_HOPS];
  int32_t rank_items[LG_MAX_HOPS];
  u64* sample_state[LG_MAX_HOPS];
  u64* sample_groups[LG_MAX_HOPS];
  u64* rank_state[LG_MAX_HOPS];
  u64* rank_groups[LG_MAX_HOPS];
  HopState* hs;
  int32_t* gid[2];
  unsigned* bar;  // [0] arrivals, [1] generation: persistent across batches, self-resetting
};

// Grid barrier; `last` runs on one thread of the last CTA to arrive, before anybody is released.
template <typename F>
__device__ __forceinline__ void chain_barrier(unsigned* bar, unsigned n_ctas, F last) {
  __threadfence();  // every thread: its own stores and REDs are performed before the CTA arrives
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = *(volatile unsigned*)(bar + 1);  // read BEFORE arriving
    __threadfence();
    if (atomicAdd(bar, 1u) == n_ctas - 1u) {
      last();
      atomicExch(bar, 0u);
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*(volatile unsigned*)(bar + 1) == gen) __nanosleep(100);
    }
  }
  __syncthreads();
  __threadfence();  // every thread: invalidates this SM's L1 (CCTL.IVALL) — the next phase reads other CTAs' results
}

template <int RNG>
__global__ void __launch_bounds__(kBlock, 6) chain_kernel(const ChainArgs c) {
  __shared__ SampleSmem<256> sm;
  __shared__ int32_t s_tile;
  const int tid = threadIdx.x;
  const unsigned n_ctas = gridDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * kBlock + tid, n_threads = (int64_t)n_ctas * kBlock;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __threadfence();  // see pdl_prologue
  batch_generate_body<false>(c.gen, (int32_t)gtid, (int32_t)n_threads);
  chain_barrier(c.bar, n_ctas, [] {});
  for (int h = 0; h < c.n_hops; h++) {
    const bool last_hop = (h == c.n_hops - 1);
    {
      SampleHop sh;
      sh.frontier_prev = c.gid[(h + 1) & 1];
      sh.gid_out = c.gid[h & 1];
      sh.tile_state = c.sample_state[h];
      sh.anchors = c.sample_groups[h];
      sh.hs = c.hs + h;
      sh.hop = h + 1;
      sh.fanout = c.fanout[h];
      sh.fanout_magic = c.magic[h];
      sh.relabel_prev = h > 0;  // construct_graph of the previous hop rides along (its vertices are looked up anyway)
      const int tile_f = c.tile_f[h];
      for (;;) {
        __syncthreads();  // the previous tile's shared memory and s_tile are free
        if (tid == 0) s_tile = atomicAdd(&sh.hs->sample_ticket, 1);
        __syncthreads();
        const int tile = s_tile;
        bool more;
        switch (tile_f) {
          case 256: more = sample_tile<256, RNG, false>(c.smp, sh, tile, sm); break;
          case 128: more = sample_tile<128, RNG, false>(c.smp, sh, tile, reinterpret_cast<SampleSmem<128>&>(sm)); break;
          case 64: more = sample_tile<64, RNG, false>(c.smp, sh, tile, reinterpret_cast<SampleSmem<64>&>(sm)); break;
          default: more = sample_tile<32, RNG, false>(c.smp, sh, tile, reinterpret_cast<SampleSmem<32>&>(sm)); break;
        }
        if (!more) break;
      }
    }
    chain_barrier(c.bar, n_ctas, [] {});
    RankHop rh;
    rh.gid = c.gid[h & 1];
    rh.tile_state = c.rank_state[h];
    rh.anchors = c.rank_groups[h];
    rh.hs = c.hs + h;
    rh.hop = h + 1;
    rh.agg_src = last_hop ? c.smp.agg_src : nullptr;  // earlier hops: written by the next hop's sample phase
    const int items = c.rank_items[h];
    for (;;) {
      __syncthreads();
      if (tid == 0) s_tile = atomicAdd(&rh.hs->rank_ticket, 1);
      __syncthreads();
      const int tile = s_tile;
      bool more;
      if (items == 12)
        more = last_hop ? rank_tile<12, false, false>(c.rnk, rh, tile, sm.red) : rank_tile<12, false, true>(c.rnk, rh, tile, sm.red);
      else
        more = last_hop ? rank_tile<4, false, false>(c.rnk, rh, tile, sm.red) : rank_tile<4, false, true>(c.rnk, rh, tile, sm.red);
      if (!more) break;
    }
    chain_barrier(c.bar, n_ctas, [&] { rank_counter_update(c.rnk, rh); });
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the stream successor may become resident now
  release_map(c.rel);
  relabel_body(c.smp.agg_src, c.rnk.ec, gtid, n_threads);
}

// lg_batch_publish: the used part of one buffer set copied into another (ids, labels, COO, counters), one launch
struct PublishArgs {
  const int32_t *ids, *labels, *agg_src, *agg_dst, *nc, *ec;
  int32_t *o_ids, *o_labels, *o_agg_src, *o_agg_dst, *o_nc, *o_ec;
};
__global__ void __launch_bounds__(kBlock) publish_kernel(const PublishArgs a) {
  pdl_prologue();
  const int32_t hops = ld_counter(a.nc + LG_INTRABATCH_CON * 3 - 1);
  const int32_t B = ld_counter(a.nc + LG_INTRABATCH_CON * 3);
  const int32_t n = hops >= 0 && hops <= LG_MAX_HOPS ? ld_counter(a.nc + LG_INTRABATCH_CON * 3 + hops) : 0;
  const int32_t e = hops >= 0 && hops <= LG_MAX_HOPS ? ld_counter(a.ec + LG_INTRABATCH_CON * 3 + hops) : 0;
  const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x, nt = (int64_t)gridDim.x * kBlock;
  auto copy = [&](const int32_t* __restrict__ src, int32_t* __restrict__ dst, int32_t count) {
    const int32_t n4 = count >> 2;  // the buffers are cudaMalloc'ed: 16-byte aligned
    for (int64_t i = t; i < n4; i += nt) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(src)[i];
    for (int64_t i = (int64_t)n4 * 4 + t; i < count; i += nt) dst[i] = src[i];
  };
  copy(a.ids, a.o_ids, n);
  copy(a.agg_src, a.o_agg_src, e);
  copy(a.agg_dst, a.o_agg_dst, e);
  copy(a.labels, a.o_labels, B);
  if (t < LG_COUNTER_SLOTS) {
    a.o_nc[t] = a.nc[t];
    a.o_ec[t] = a.ec[t];
  }
}

// HotnessMeasure (cache/cache_impl.cuh:190-198) + max_ids_ (cache/cache.cu:59-61)
__global__ void __launch_bounds__(kBlock) hotness_measure_kernel(const int32_t* __restrict__ ids,
                                                                 const int32_t* __restrict__ nc, u64* node_hot,
                                                                 int32_t* max_ids) {
  pdl_prologue();
  const int32_t n = ld_counter(nc + LG_INTRABATCH_CON * 2 + 1);
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
// Tuning knobs of the sampler chain (environment, read once; defaults are the measured best, profiles/r01d_sampler_chain.md):
//   LG_SAMPLE_TILE   frontier entries per CTA of the long hops (32..256)
//   LG_SAMPLE_MINB   6 (default) = the 256-entry sample kernel of the dense layout is capped at 40 registers so that
//                    6 CTAs fit an SM: the 782 tiles of a 200 k-entry frontier run as one wave (592 slots otherwise)
//   LG_RANK_ITEMS    edges per thread of the long hops' rank kernel (4, 8, 12 or 16; default 12 dense, 16 hashed: one wave)
//   LG_PM_FILL_MB    dense position maps up to this size (default 16 MB) are released by a streaming fill instead of
//                    the O(batch) scatter (read when a handle is created)
//   LG_RED_PRECHECK  bit 0 (default on) = dense layout: L1-cached look at the map word before the RED.MIN; skips the RED when
//                    the word already holds an earlier position (hub vertices: sample hop 2 0.0915 -> 0.0827 ms);
//                    bit 1 = the same before the hashed layout's atomicMin (off: an earlier version measured +3 % on a
//                    hub-heavy 2.4 M-vertex graph, -5 % at UK-Union scale — the regime the hashed layout is for)
//   LG_SAMPLE_CTAS_PER_SM / LG_RANK_CTAS_PER_SM   (default 3 / 3; 0 = one tile per CTA) the long hops' sample and rank kernels
//                    run as grids of 148 x k CTAs that keep claiming tiles.  Alone the kernels get slower (hop 2 of the
//                    UK-Union shape 0.172 -> 0.199 ms), but they no longer fill the register files (64 registers x 256
//                    threads x 4 CTAs = all of it), so the gather's CTAs of the other batches in flight find room on every
//                    SM: pipelined +3.5 % (UK-Union shape) / +4.5 % (products shape), profiles/r02_persistent_grids.md
//   LG_CHAIN         1 = lg_run_batch runs the dense layout's sampler chain as ONE persistent kernel (chain_kernel; read
//                    when a handle is created).  Opt-in: bit-exact, but measured slower — 34.1 vs 40.8 M seeds/s at best,
//                    profiles/r01d_chain_kernel.md.  LG_CHAIN_CTAS: its CTAs per SM.
//   (having a hop's sample kernel also look up the rows of the vertices it samples, so that the next hop starts from two
//   coalesced arrays, was measured and rejected: hop 1 +5 us, hop 2 -2 us)
// (A forced shared-memory carve-out on these kernels, LG_CARVEOUT, was measured and rejected: profiles/r01b_overlap.md.)
struct SamplerTune {
  int sample_tile, sample_minb, rank_items, red_precheck, chain_ctas, sample_minb_hashed, sample_ctas_per_sm, rank_ctas_per_sm;
};
static const SamplerTune& sampler_tune() {
  static SamplerTune t = [] {
    SamplerTune x{0, 6, 0, 1, kChainCtasPerSm, 0, 3, 3};  // persistent grids of 3 CTAs per SM: profiles/r02_persistent_grids.md
    if (const char* e = getenv("LG_SAMPLE_CTAS_PER_SM")) x.sample_ctas_per_sm = atoi(e);
    if (const char* e = getenv("LG_RANK_CTAS_PER_SM")) x.rank_ctas_per_sm = atoi(e);
    if (const char* e = getenv("LG_SAMPLE_MINB_HASHED")) x.sample_minb_hashed = atoi(e);
    if (const char* e = getenv("LG_SAMPLE_TILE")) x.sample_tile = atoi(e);
    if (const char* e = getenv("LG_SAMPLE_MINB")) x.sample_minb = atoi(e);
    if (const char* e = getenv("LG_RANK_ITEMS")) x.rank_items = atoi(e);
    if (const char* e = getenv("LG_RED_PRECHECK")) x.red_precheck = atoi(e);
    if (const char* e = getenv("LG_CHAIN_CTAS")) x.chain_ctas = atoi(e);
    return x;
  }();
  return t;
}

// PDL on the handle's chain: LG_PDL if set, else on for the dense layout only.  Next to the gather the hashed kernels
// (multi-wave, L1-hungry) lost 2-7 % with it, the dense ones gained 4-10 % (profiles/r01d_sampler_chain.md).
static bool pdl_on(const lg_sampler* s, int except_bit = 0) {
  const int v = lg_pdl();
  if (v < 0) return !s->hashed;
  return (v & 1) && !(v & except_bit);
}

extern "C" int64_t lg_num_ids(int32_t batch_size, const int32_t* fanout, int32_t n_hops) {
  int64_t tot = batch_size, per = batch_size;  // engine/server.cu:187-199
  for (int i = 0; i < n_hops; i++) {
    per *= fanout[i];
    tot += per;
  }
  return tot;
}

// see SampleHop::fanout_magic
static uint32_t fanout_magic(int32_t fanout) {
  if (fanout > 4096) return 0u;
  return (uint32_t)(((1ull << 32) + (uint64_t)fanout - 1) / (uint64_t)fanout);
}

static int pick_tile_f(int64_t frontier_max) {
  // a tile is one CTA: prefer many small tiles for short frontiers (latency), 256-entry tiles for long ones
  if (frontier_max >= 128ll * kSMs * 4) return 256;
  if (frontier_max >= 64ll * kSMs * 4) return 128;
  if (frontier_max >= 64ll * kSMs) return 64;
  return 32;
}
static int pick_tile_f_tuned(int64_t frontier_max) {
  const int t = pick_tile_f(frontier_max), o = sampler_tune().sample_tile;
  return (t == 256 && (o == 32 || o == 64 || o == 128 || o == 256)) ? o : t;
}
static int pick_rank_items(int64_t edges_max, bool hashed) {
  // all tiles of a hop resident at once: 4 edges per thread for short hops; for long ones 12 (dense: 32 registers,
  // 652 tiles for 2 M edges) or 16 (hashed: 58 registers, 4 CTAs per SM = 592 slots for 489 tiles)
  if (edges_max <= 4ll * kBlock * kSMs * 5) return 4;
  const int o = sampler_tune().rank_items;
  if (o == 4 || o == 8 || o == 12 || o == 16) return o;
  // persistent grid (LG_RANK_CTAS_PER_SM, default 3 CTAs per SM = 444): 8 edges per thread = 977 tiles of 2048 edges for a
  // 2 M-edge hop, 2.2 rounds of small tiles instead of 1.1 rounds of large ones (profiles/r02_persistent_grids.md)
  if (sampler_tune().rank_ctas_per_sm > 0) return 8;
  return hashed ? 16 : 12;
}

static int sampler_init(lg_sampler* s, int32_t device, int32_t max_batch, const int32_t* fanout, int32_t n_hops,
                        int64_t num_nodes);
extern "C" int lg_sampler_destroy(lg_sampler* s);

extern "C" int lg_sampler_create(int32_t device, int32_t max_batch, const int32_t* fanout, int32_t n_hops,
                                 int64_t num_nodes, lg_sampler** out) {
  LG_REQUIRE(out && fanout, "lg_sampler_create: null argument");
  LG_REQUIRE(n_hops >= 1 && n_hops <= LG_MAX_HOPS, "lg_sampler_create: n_hops %d outside [1,%d]", n_hops, LG_MAX_HOPS);
  LG_REQUIRE(max_batch >= 1, "lg_sampler_create: max_batch %d", max_batch);
  LG_REQUIRE(num_nodes >= 1 && num_nodes < (1ll << 31), "lg_sampler_create: num_nodes %lld outside [1, 2^31)",
             (long long)num_nodes);
  LG_CUDA(cudaSetDevice(device));
  {  // L2 set-aside for the accesses that carry an evict_last policy (position map / dedup table, directories, indptr):
     // LG_L2_PERSIST_MB, default 0 = the driver's default (no set-aside).  Measured: see profiles/r02_l2_persist.md
    static const int mb = [] {
      const char* e = getenv("LG_L2_PERSIST_MB");
      return e ? atoi(e) : 0;
    }();
    if (mb > 0) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)mb << 20);
      cudaGetLastError();
    }
  }
  lg_sampler* s = new lg_sampler();
  memset(s, 0, sizeof(*s));
  const int rc = sampler_init(s, device, max_batch, fanout, n_hops, num_nodes);
  if (rc) {  // nothing of a half-built handle survives (every member is null or owned)
    lg_sampler_destroy(s);
    return rc;
  }
  *out = s;
  return 0;
}

static int sampler_init(lg_sampler* s, int32_t device, int32_t max_batch, const int32_t* fanout, int32_t n_hops,
                        int64_t num_nodes) {
  s->device = device;
  s->max_batch = max_batch;
  s->n_hops = n_hops;
  s->num_nodes = num_nodes;
  s->gather_variant = LG_GATHER_AUTO;
  s->slots_per_hop[0] = max_batch;
  for (int h = 0; h < n_hops; h++) {
    LG_REQUIRE(fanout[h] >= 1 && fanout[h] <= 65535, "lg_sampler_create: fanout[%d]=%d", h, fanout[h]);
    s->fanout[h] = fanout[h];
    s->slots_per_hop[h + 1] = s->slots_per_hop[h] * fanout[h];
  }
  s->num_ids = lg_num_ids(max_batch, fanout, n_hops);
  LG_REQUIRE(s->num_ids < (1ll << 31), "lg_sampler_create: num_ids %lld does not fit int32", (long long)s->num_ids);
  int64_t smax = s->slots_per_hop[n_hops];
  if (smax < 2ll * max_batch) smax = 2ll * max_batch;  // head of gid[1] doubles as the e2e seed staging area
  smax = (smax + kBlock * 48 - 1) / (kBlock * 48) * (kBlock * 48);  // whole rank tiles (4, 8, 12 or 16 edges per thread): vector loads never leave the buffer
  for (int b = 0; b < 2; b++) {
    LG_CUDA(cudaMalloc(&s->gid[b], (size_t)smax * sizeof(int32_t)));
    LG_CUDA(cudaMemset(s->gid[b], 0, (size_t)smax * sizeof(int32_t)));
  }
  // position map.  DENSE: one word per vertex (the reference's position_map, engine/server.cu:224), 0xFFFFFFFF =
  // absent; HASHED: 2^k >= 1.5 x num_ids packed words.  Dense while the map can stay L2-resident next to the
  // gather's stream (LG_DENSE_MAX_MB, default 48 MB of the 126 MB L2); LG_DEDUP=dense|hash overrides.
  {
    int64_t dense_max_mb = 48;
    if (const char* e = getenv("LG_DENSE_MAX_MB")) dense_max_mb = atoll(e);
    s->hashed = (num_nodes * 4 > dense_max_mb * (1ll << 20)) ? 1 : 0;
    if (const char* e = getenv("LG_DEDUP")) s->hashed = (e[0] == 'h' || e[0] == 'H') ? 1 : 0;
  }
  if (s->hashed) {
    uint64_t slots = 1024;
    while (slots < (uint64_t)s->num_ids * 3 / 2) slots <<= 1;
    LG_REQUIRE(slots <= (1ull << 31), "lg_sampler_create: dedup table of %llu slots", (unsigned long long)slots);
    s->table_mask = (uint32_t)(slots - 1);
    LG_CUDA(cudaMalloc(&s->table, (size_t)slots * sizeof(u64)));
    LG_CUDA(cudaMemset(s->table, 0xFF, (size_t)slots * sizeof(u64)));
  } else {
    const size_t pm_words = ((size_t)num_nodes + 3) & ~(size_t)3;  // whole 16-byte words for the streaming fill
    LG_CUDA(cudaMalloc(&s->pm, pm_words * sizeof(uint32_t)));
    LG_CUDA(cudaMemset(s->pm, 0xFF, pm_words * sizeof(uint32_t)));
  }
  // small per-batch state
  int64_t bytes = sizeof(HopState) * LG_MAX_HOPS;
  bytes = (bytes + 255) & ~255ll;
  int64_t off_state[2][LG_MAX_HOPS], off_anchor[2][LG_MAX_HOPS];
  for (int h = 0; h < n_hops; h++) {
    int tf = pick_tile_f_tuned(s->slots_per_hop[h]);
    s->sample_tile_f[h] = tf;
    s->sample_tiles[h] = (int32_t)((s->slots_per_hop[h] + tf - 1) / tf);
    s->rank_items[h] = pick_rank_items(s->slots_per_hop[h + 1], s->hashed != 0);
    const int64_t rank_tile = (int64_t)kBlock * s->rank_items[h];
    s->rank_tiles[h] = (int32_t)((s->slots_per_hop[h + 1] + rank_tile - 1) / rank_tile);
    off_state[0][h] = bytes;
    bytes += (int64_t)s->sample_tiles[h] * 8;
    bytes = (bytes + 127) & ~127ll;
    off_anchor[0][h] = bytes;
    bytes += (int64_t)(s->sample_tiles[h] / kGroup + 1) * kGroupStride * 8;
    off_state[1][h] = bytes;
    bytes += (int64_t)s->rank_tiles[h] * 8;
    bytes = (bytes + 127) & ~127ll;
    off_anchor[1][h] = bytes;
    bytes += (int64_t)(s->rank_tiles[h] / kGroup + 1) * kGroupStride * 8;
  }
  s->small_bytes = bytes;
  LG_CUDA(cudaMalloc(&s->small, (size_t)bytes));
  LG_CUDA(cudaMemset(s->small, 0, (size_t)bytes));
  s->hs = (HopState*)s->small;
  for (int h = 0; h < n_hops; h++) {
    s->sample_state[h] = (u64*)(s->small + off_state[0][h]);
    s->sample_anchor[h] = (u64*)(s->small + off_anchor[0][h]);
    s->rank_state[h] = (u64*)(s->small + off_state[1][h]);
    s->rank_anchor[h] = (u64*)(s->small + off_anchor[1][h]);
  }
  LG_CUDA(cudaMalloc(&s->status, sizeof(int32_t)));
  LG_CUDA(cudaMemset(s->status, 0, sizeof(int32_t)));
  if (const char* e = getenv("LG_CHAIN")) s->chain = atoi(e) != 0;
  s->pm_fill_mb = 16;
  if (const char* e = getenv("LG_PM_FILL_MB")) s->pm_fill_mb = atoi(e);
  LG_CUDA(cudaMalloc(&s->chain_bar, 2 * sizeof(unsigned)));
  LG_CUDA(cudaMemset(s->chain_bar, 0, 2 * sizeof(unsigned)));
  if (s->hashed) LG_CUDA(cudaMalloc(&s->seed_local, (size_t)max_batch * sizeof(int32_t)));
  {
    const char* e = getenv("LG_GATHER_DYNAMIC");
    s->gather_chunk = 4;
    s->gather_static_pct = 0;
    if (const char* p = getenv("LG_GATHER_STATIC_PCT")) s->gather_static_pct = atoi(p) < 0 ? 0 : (atoi(p) > 100 ? 100 : atoi(p));
    if (e && atoi(e) != 0) {  // opt-in; LG_GATHER_DYNAMIC = tiles per claim (1 = one claim per tile)
      s->gather_chunk = atoi(e) > 0 ? atoi(e) : 4;
      LG_CUDA(cudaMalloc(&s->gather_ticket, 2 * sizeof(int32_t)));
      LG_CUDA(cudaMemset(s->gather_ticket, 0, 2 * sizeof(int32_t)));
    }
  }
  LG_CUDA(cudaMallocHost(&s->pinned_seeds, (size_t)max_batch * 2 * sizeof(int32_t)));
  // (stream priorities were measured: the side stream at the highest priority 41.6 -> 40.0 M seeds/s, the callers'
  // sampling streams at the highest priority 41.5-42.9 -> 40.2 M: any asymmetry loses, profiles/r01d_overlap.md)
  LG_CUDA(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  for (int h = 0; h <= LG_MAX_HOPS; h++) LG_CUDA(cudaEventCreateWithFlags(&s->ev_fork[h], cudaEventDisableTiming));
  LG_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  LG_CUDA(cudaEventCreateWithFlags(&s->ev_clear, cudaEventDisableTiming));
  s->overlap = 1;
  s->fuse_gathers = 1;
  return 0;
}

extern "C" int lg_sampler_destroy(lg_sampler* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->pm);
  cudaFree(s->table);
  cudaFree(s->seed_local);
  cudaFree(s->gather_ticket);
  cudaFree(s->gid[0]);
  cudaFree(s->gid[1]);
  cudaFree(s->small);
  cudaFree(s->status);
  cudaFree(s->chain_bar);
  cudaFreeHost(s->pinned_seeds);
  if (s->side) cudaStreamDestroy(s->side);  // a half-built handle (lg_sampler_create failed) has null members
  for (int h = 0; h <= LG_MAX_HOPS; h++)
    if (s->ev_fork[h]) cudaEventDestroy(s->ev_fork[h]);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->ev_clear) cudaEventDestroy(s->ev_clear);
  for (int i = 0; i < s->n_done; i++) cudaEventDestroy(s->done[i].ev);
  delete s;
  return 0;
}

extern "C" int lg_sampler_reset(lg_sampler* s, lg_stream_t stream) {
  LG_REQUIRE(s, "null sampler");
  if (s->hashed)
    LG_CUDA(cudaMemsetAsync(s->table, 0xFF, ((size_t)s->table_mask + 1) * sizeof(u64), (cudaStream_t)stream));
  else
    LG_CUDA(cudaMemsetAsync(s->pm, 0xFF, (size_t)s->num_nodes * sizeof(uint32_t), (cudaStream_t)stream));
  LG_CUDA(cudaMemsetAsync(s->status, 0, sizeof(int32_t), (cudaStream_t)stream));
  s->pm_dirty = 0;
  return 0;
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

extern "C" int lg_sampler_set_lazy_relabel(lg_sampler* s, int32_t mode) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(mode == 0 || mode == 1, "lazy relabel mode %d", mode);
  s->lazy_relabel = mode;
  return 0;
}

extern "C" int lg_sampler_set_tail_mode(lg_sampler* s, int32_t mode) {
  LG_REQUIRE(s, "null sampler");
  LG_REQUIRE(mode == LG_TAIL_EXACT || mode == LG_TAIL_REFERENCE, "tail mode %d", mode);
  s->tail_reference = mode;
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

extern "C" int lg_sampler_status_async(lg_sampler* s, lg_stream_t stream, int32_t* pinned_host_status) {
  LG_REQUIRE(s && pinned_host_status, "null argument");
  LG_CUDA(cudaMemcpyAsync(pinned_host_status, s->status, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

extern "C" int32_t lg_sampler_dedup_layout(const lg_sampler* s) { return s ? s->hashed : -1; }

extern "C" int64_t lg_sampler_scratch_bytes(const lg_sampler* s) {
  if (!s) return 0;
  const int64_t map_bytes = s->hashed ? ((int64_t)s->table_mask + 1) * 8 : s->num_nodes * 4;
  return map_bytes + 2 * s->slots_per_hop[s->n_hops] * 4 + s->small_bytes + 4;
}

// how the map of this handle is released after batch `b`
static ReleaseArgs release_args(const lg_sampler* s, const lg_batch* b) {
  ReleaseArgs r;
  memset(&r, 0, sizeof(r));
  r.l2 = lg_l2_hints();
  if (s->hashed) {  // O(batch)-sized table, L2-resident: one streaming fill instead of an O(batch) random clear
    r.fill = (uint4*)s->table;
    r.n16 = ((int64_t)s->table_mask + 1) / 2;
  } else if (s->num_nodes * 4 <= (int64_t)s->pm_fill_mb * (1ll << 20)) {
    r.fill = (uint4*)s->pm;
    r.n16 = (s->num_nodes + 3) / 4;
  } else {
    r.pm = s->pm;
    r.ids = b->ids;
    r.nc = b->node_counter;
  }
  return r;
}
// bookkeeping after a release was enqueued on `st` (as its own launch or inside the batch's last kernel)
static int released(lg_sampler* s, cudaStream_t st) {
  if (s->hashed) {
    s->table_clean = 1;
  } else {
    // the next batch's inserts must not overtake this release when the host enqueues them on another stream
    LG_CUDA(cudaEventRecord(s->ev_clear, st));
    s->clear_stream = st;
    s->clear_recorded = 1;
  }
  s->pm_dirty = 0;
  return 0;
}
// the position-map words of the batch last generated into `b` go back to "absent" (ClearPosMap, :542-548)
static int clear_position_map(lg_sampler* s, cudaStream_t st, const lg_batch* b) {
  if (!s->pm_dirty) return 0;  // already released by the batch's last kernel (lg_run_batch)
  if (s->hashed) {  // the table is re-initialised by the next lg_batch_generate instead
    s->pm_dirty = 0;
    return 0;
  }
  LG_CUDA(lg_launch_opt(pdl_on(s, 2), release_kernel, kSMs * 8, kBlock, 0, st, release_args(s, b)));
  return released(s, st);
}
static DedupMap map_of(const lg_sampler* s) {
  DedupMap m;
  m.pm = s->pm;
  m.table = s->table;
  m.mask = s->table_mask;
  return m;
}

static int32_t clipped_batch_size(int32_t total_cap, int32_t batch_size, int32_t counter) {
  const long long done = (long long)batch_size * ((long long)counter + 1);
  int32_t size = (done >= total_cap) ? (int32_t)(total_cap - (long long)batch_size * counter) : batch_size;  // :159
  return size < 0 ? 0 : size;
}
static BatchGenArgs batch_gen_args(const lg_sampler* s, const int32_t* all_ids, const int32_t* all_labels, int32_t total_cap,
                                   int32_t size, int32_t batch_size, int32_t counter, const lg_batch* b) {
  BatchGenArgs g;
  g.stride = s->tail_reference ? size : batch_size;
  g.all_ids = all_ids;
  g.all_labels = all_labels;
  g.total_cap = total_cap;
  g.size = size;
  g.counter = counter;
  g.hop_num = s->n_hops;
  g.ids = b->ids;
  g.labels = b->labels;
  g.nc = b->node_counter;
  g.ec = b->edge_counter;
  g.map = map_of(s);
  g.small = (u64*)s->small;
  g.small_words = (int32_t)(s->small_bytes / 8);
  g.l2 = lg_l2_hints();
  return g;
}

extern "C" int lg_batch_generate(lg_sampler* s, lg_stream_t stream_, const int32_t* all_ids,
                                 const int32_t* all_labels, int32_t total_cap, int32_t batch_size, int32_t counter,
                                 const lg_batch* b) {
  LG_REQUIRE(s && b && all_ids && all_labels, "lg_batch_generate: null argument");
  LG_REQUIRE(batch_size <= s->max_batch, "lg_batch_generate: batch %d > max_batch %d", batch_size, s->max_batch);
  LG_REQUIRE(b->num_ids >= s->num_ids, "lg_batch_generate: batch buffers hold %d ids, need %lld", b->num_ids,
             (long long)s->num_ids);
  cudaStream_t st = (cudaStream_t)stream_;
  // a previous batch that never reached lg_io_complete still owns position-map words: release them first
  if (s->pm_dirty) {
    int rc = clear_position_map(s, st, &s->dirty_batch);
    if (rc) return rc;
  }
  if (s->clear_recorded && s->clear_stream != st) LG_CUDA(cudaStreamWaitEvent(st, s->ev_clear, 0));
  if (s->hashed && !s->table_clean)  // not already re-initialised by the previous batch's last kernel
    LG_CUDA(lg_launch_opt(pdl_on(s, 4), release_kernel, kSMs * 8, kBlock, 0, st, release_args(s, b)));
  s->table_clean = 0;
  const int32_t size = clipped_batch_size(total_cap, batch_size, counter);
  const BatchGenArgs g = batch_gen_args(s, all_ids, all_labels, total_cap, size, batch_size, counter, b);
  int grid = size > 0 ? (size + kBlock - 1) / kBlock : 1;
  // the kernel also zeroes the per-batch scan state (tickets + tile words): enough threads for one pass
  const int grid_small = (g.small_words + kBlock - 1) / kBlock;
  if (grid < grid_small) grid = grid_small < 64 ? grid_small : 64;
  if (s->hashed) {
    LG_CUDA(lg_launch_opt(pdl_on(s, 4), batch_generate_kernel<true>, grid, kBlock, 0, st, g));
    LG_CUDA(lg_launch_opt(pdl_on(s), seed_local_kernel, grid, kBlock, 0, st, (const int32_t*)b->ids, (const int32_t*)b->node_counter,
                          s->seed_local, map_of(s), lg_l2_hints()));
  } else {
    LG_CUDA(lg_launch_opt(pdl_on(s, 4), batch_generate_kernel<false>, grid, kBlock, 0, st, g));
  }
  s->pm_dirty = 1;
  s->dirty_batch = *b;
  return 0;
}

template <int RNG, bool HASHED>
static cudaError_t launch_sample(bool pdl, int tile_f, int grid, cudaStream_t st, const SampleArgs& a) {
  switch (tile_f) {
    case 256:
      if (!HASHED && sampler_tune().sample_minb == 6)
        return lg_launch_opt(pdl, sample_hop_kernel<256, RNG, false, 6>, grid, kBlock, 0, st, a);
      if (HASHED && sampler_tune().sample_minb_hashed == 6)
        return lg_launch_opt(pdl, sample_hop_kernel<256, RNG, HASHED, 6>, grid, kBlock, 0, st, a);
      if (HASHED && sampler_tune().sample_minb_hashed == 5)
        return lg_launch_opt(pdl, sample_hop_kernel<256, RNG, HASHED, 5>, grid, kBlock, 0, st, a);
      return lg_launch_opt(pdl, sample_hop_kernel<256, RNG, HASHED, 0>, grid, kBlock, 0, st, a);
    case 128: return lg_launch_opt(pdl, sample_hop_kernel<128, RNG, HASHED, 0>, grid, kBlock, 0, st, a);
    case 64: return lg_launch_opt(pdl, sample_hop_kernel<64, RNG, HASHED, 0>, grid, kBlock, 0, st, a);
    default: return lg_launch_opt(pdl, sample_hop_kernel<32, RNG, HASHED, 0>, grid, kBlock, 0, st, a);
  }
}
template <bool HASHED, bool PUBLISH>
static cudaError_t launch_rank(bool pdl, int items, int grid, cudaStream_t st, const RankArgs& r) {
  switch (items) {
    case 16: return lg_launch_opt(pdl, rank_kernel<16, HASHED, PUBLISH>, grid, kBlock, 0, st, r);
    case 12: return lg_launch_opt(pdl, rank_kernel<12, HASHED, PUBLISH>, grid, kBlock, 0, st, r);
    case 8: return lg_launch_opt(pdl, rank_kernel<8, HASHED, PUBLISH>, grid, kBlock, 0, st, r);
    default: return lg_launch_opt(pdl, rank_kernel<4, HASHED, PUBLISH>, grid, kBlock, 0, st, r);
  }
}

// one hop: sample + rank (+ the hop's own relabel pass unless the caller folds it into the next hop's sample kernel)
static int sample_hop(lg_sampler* s, cudaStream_t st, const lg_topology* topo, int32_t hop, int32_t rng_kind,
                      uint64_t rng_seed, uint32_t batch_id, uint32_t stream_id, const lg_batch* b,
                      unsigned long long* edge_hotness, bool relabel_prev, bool relabel_own, bool release) {
  const int h = hop - 1;
  // HASHED: the next hop never reads the table (see seed_local_kernel) but the previous hop's agg_src, which its rank
  // pass always writes; entries left as ~p_first are patched by relabel_kernel (relabel_own) or by the next hop's
  // sample kernel (lg_run_batch: one launch less per hop)
  SampleArgs a;
  a.topo = *topo;
  a.h.frontier_prev = s->gid[(h + 1) & 1];
  a.seed_local = s->seed_local;
  a.h.gid_out = s->gid[h & 1];
  a.ids = b->ids;
  a.agg_src = b->agg_src;
  a.agg_dst = b->agg_dst;
  a.nc = b->node_counter;
  a.ec = b->edge_counter;
  a.map = map_of(s);
  a.h.tile_state = s->sample_state[h];
  a.h.anchors = s->sample_anchor[h];
  a.h.hs = s->hs + h;
  a.edge_hot = (u64*)edge_hotness;
  a.h.hop = hop;
  a.h.fanout = s->fanout[h];
  a.h.fanout_magic = fanout_magic(s->fanout[h]);
  a.h.relabel_prev = relabel_prev ? 1 : 0;
  a.precheck = sampler_tune().red_precheck;
  a.status = s->status;
  int sample_grid = s->sample_tiles[h];
  a.persistent = 0;
  if (sampler_tune().sample_ctas_per_sm > 0 && sample_grid > kSMs * sampler_tune().sample_ctas_per_sm) {
    sample_grid = kSMs * sampler_tune().sample_ctas_per_sm;
    a.persistent = 1;
  }
  a.batch_id = batch_id;
  a.stream_id = stream_id;
  a.k0 = (uint32_t)rng_seed;
  a.k1 = (uint32_t)(rng_seed >> 32);
  a.l2 = lg_l2_hints();
  a.trace = s->trace;
  if (rng_kind == LG_RNG_MINSTD) {
    if (s->hashed) LG_CUDA((launch_sample<LG_RNG_MINSTD, true>(pdl_on(s), s->sample_tile_f[h], sample_grid, st, a)));
    else LG_CUDA((launch_sample<LG_RNG_MINSTD, false>(pdl_on(s), s->sample_tile_f[h], sample_grid, st, a)));
  } else {
    if (s->hashed) LG_CUDA((launch_sample<LG_RNG_PHILOX, true>(pdl_on(s), s->sample_tile_f[h], sample_grid, st, a)));
    else LG_CUDA((launch_sample<LG_RNG_PHILOX, false>(pdl_on(s), s->sample_tile_f[h], sample_grid, st, a)));
  }
  RankArgs r;
  r.h.gid = s->gid[h & 1];
  r.ids = b->ids;
  r.nc = b->node_counter;
  r.ec = b->edge_counter;
  r.map = map_of(s);
  r.h.agg_src = (relabel_own || s->hashed) ? b->agg_src : nullptr;
  r.h.tile_state = s->rank_state[h];
  r.h.anchors = s->rank_anchor[h];
  r.h.hs = s->hs + h;
  r.h.hop = hop;
  r.ids_cap = b->num_ids;
  r.l2 = lg_l2_hints();
  r.status = s->status;
  r.trace = s->trace;
  int rank_grid = s->rank_tiles[h];
  r.persistent = 0;
  if (sampler_tune().rank_ctas_per_sm > 0 && rank_grid > kSMs * sampler_tune().rank_ctas_per_sm) {
    rank_grid = kSMs * sampler_tune().rank_ctas_per_sm;
    r.persistent = 1;
  }
  const bool publish = hop < s->n_hops;  // a later hop inserts into / reads the map
  if (s->hashed) {
    if (publish) LG_CUDA((launch_rank<true, true>(pdl_on(s), s->rank_items[h], rank_grid, st, r)));
    else LG_CUDA((launch_rank<true, false>(pdl_on(s), s->rank_items[h], rank_grid, st, r)));
  } else {
    if (publish) LG_CUDA((launch_rank<false, true>(pdl_on(s), s->rank_items[h], rank_grid, st, r)));
    else LG_CUDA((launch_rank<false, false>(pdl_on(s), s->rank_items[h], rank_grid, st, r)));
  }
  if (relabel_own) {
    int64_t grid = (s->slots_per_hop[hop] + kBlock * 4 - 1) / (kBlock * 4);
    {  // grid-stride kernel: LG_RELABEL_CTAS_PER_SM caps its footprint like the sample / rank kernels' (0 = one pass per thread)
      static const int cap = [] {
        const char* e = getenv("LG_RELABEL_CTAS_PER_SM");
        return e ? atoi(e) : 0;
      }();
      if (cap > 0 && grid > (int64_t)kSMs * cap) grid = (int64_t)kSMs * cap;
    }
    ReleaseArgs rel;
    memset(&rel, 0, sizeof(rel));
    if (release) rel = release_args(s, b);
    LG_CUDA(lg_launch_opt(pdl_on(s), relabel_kernel, (int)grid, kBlock, 0, st, b->agg_src, (const int32_t*)b->edge_counter, rel));
    if (release) {
      int rc = released(s, st);
      if (rc) return rc;
    }
  }
  return 0;
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
  LG_REQUIRE(topo->num_nodes <= s->num_nodes, "lg_random_sample: topology has %lld vertices, sampler was created for %lld",
             (long long)topo->num_nodes, (long long)s->num_nodes);
  const bool lazy = s->lazy_relabel != 0, last = hop == s->n_hops;
  return sample_hop(s, (cudaStream_t)stream_, topo, hop, rng_kind, rng_seed, batch_id, stream_id, b, edge_hotness,
                    /*relabel_prev=*/lazy && hop > 1, /*relabel_own=*/!lazy || last, /*release=*/lazy && last);
}

extern "C" int lg_io_submit(lg_sampler*, lg_stream_t, int32_t, const lg_batch*) { return 0; }

extern "C" int lg_batch_publish(lg_stream_t stream, const lg_batch* from, const lg_batch* to) {
  LG_REQUIRE(from && to && from->ids && to->ids && from->node_counter && to->node_counter, "lg_batch_publish: null argument");
  LG_REQUIRE(to->num_ids >= from->num_ids, "lg_batch_publish: destination holds %d ids, source %d", to->num_ids, from->num_ids);
  PublishArgs a{from->ids, from->labels, from->agg_src, from->agg_dst, from->node_counter, from->edge_counter,
                to->ids,   to->labels,   to->agg_src,   to->agg_dst,   to->node_counter,   to->edge_counter};
  LG_CUDA(lg_launch_opt(lg_pdl() != 0, publish_kernel, kSMs * 4, kBlock, 0, (cudaStream_t)stream, a));
  return 0;
}

extern "C" int lg_io_complete(lg_sampler* s, lg_stream_t stream_, int32_t mode, const lg_batch* b,
                              unsigned long long* node_hotness, int32_t* max_ids) {
  LG_REQUIRE(s && b, "lg_io_complete: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  if (mode == LG_TRAINMODE && node_hotness) {  // :558
    LG_CUDA(lg_launch_opt(pdl_on(s), hotness_measure_kernel, kSMs * 2, kBlock, 0, st, (const int32_t*)b->ids,
                      (const int32_t*)b->node_counter, (u64*)node_hotness, max_ids));
  }
  // ClearPosMap: the reference clears in train mode only (its bitmap makes stale entries harmless in the
  // other modes); the position map is the only dedup state here, so it is released in every mode
  return clear_position_map(s, st, b);
}

// ---- chain_kernel launch (lg_run_batch, dense layout) ----
// Chain kernels of one device are totally ordered by an event (see chain_kernel): a chain on another stream — a second
// runner in flight — only starts when the previous one has completed.
namespace {
struct ChainOrder {
  std::mutex mu;
  cudaEvent_t ev[LG_MAX_DEVICE * 2] = {};
  cudaStream_t last[LG_MAX_DEVICE * 2] = {};
  bool recorded[LG_MAX_DEVICE * 2] = {};
} g_chain_order;
}  // namespace

static bool chain_eligible(const lg_sampler* s, const lg_feature_cache* cache) {
  if (!s->chain || s->hashed || s->device < 0 || s->device >= LG_MAX_DEVICE * 2) return false;
  if (cache && s->fuse_gathers != 2) return false;  // gathers between the hops need the hop boundaries
  for (int h = 0; h < s->n_hops; h++)
    if (s->rank_items[h] != 4 && s->rank_items[h] != 12) return false;
  return true;
}

static int launch_chain(lg_sampler* s, cudaStream_t st, const lg_topology* topo, const lg_batch_params* p, const lg_batch* b) {
  LG_REQUIRE(p->batch_size <= s->max_batch, "lg_run_batch: batch %d > max_batch %d", p->batch_size, s->max_batch);
  LG_REQUIRE(b->num_ids >= s->num_ids, "lg_run_batch: batch buffers hold %d ids, need %lld", b->num_ids, (long long)s->num_ids);
  LG_REQUIRE(p->all_ids && p->all_labels, "lg_run_batch: null seed arrays");
  if (s->pm_dirty) {  // a previous batch that never reached lg_io_complete still owns position-map words
    int rc = clear_position_map(s, st, &s->dirty_batch);
    if (rc) return rc;
  }
  if (s->clear_recorded && s->clear_stream != st) LG_CUDA(cudaStreamWaitEvent(st, s->ev_clear, 0));
  ChainArgs c;
  memset(&c, 0, sizeof(c));
  const int32_t size = clipped_batch_size(p->total_cap, p->batch_size, p->counter);
  c.gen = batch_gen_args(s, p->all_ids, p->all_labels, p->total_cap, size, p->batch_size, p->counter, b);
  SampleArgs& a = c.smp;
  a.topo = *topo;
  a.seed_local = nullptr;
  a.ids = b->ids;
  a.agg_src = b->agg_src;
  a.agg_dst = b->agg_dst;
  a.nc = b->node_counter;
  a.ec = b->edge_counter;
  a.map = map_of(s);
  a.edge_hot = nullptr;
  a.precheck = sampler_tune().red_precheck;
  a.status = s->status;
  a.batch_id = p->batch_id;
  a.stream_id = p->stream_id;
  a.k0 = (uint32_t)p->rng_seed;
  a.k1 = (uint32_t)(p->rng_seed >> 32);
  a.l2 = lg_l2_hints();
  a.trace = s->trace;
  RankArgs& r = c.rnk;
  r.ids = b->ids;
  r.nc = b->node_counter;
  r.ec = b->edge_counter;
  r.map = map_of(s);
  r.ids_cap = b->num_ids;
  r.l2 = lg_l2_hints();
  r.status = s->status;
  r.trace = s->trace;
  c.rel = release_args(s, b);
  c.n_hops = s->n_hops;
  for (int h = 0; h < s->n_hops; h++) {
    c.fanout[h] = s->fanout[h];
    c.magic[h] = fanout_magic(s->fanout[h]);
    c.tile_f[h] = s->sample_tile_f[h];
    c.rank_items[h] = s->rank_items[h];
    c.sample_state[h] = s->sample_state[h];
    c.sample_groups[h] = s->sample_anchor[h];
    c.rank_state[h] = s->rank_state[h];
    c.rank_groups[h] = s->rank_anchor[h];
  }
  c.hs = s->hs;
  c.gid[0] = s->gid[0];
  c.gid[1] = s->gid[1];
  c.bar = s->chain_bar;
  {
    std::lock_guard<std::mutex> lock(g_chain_order.mu);
    const int d = s->device;
    if (!g_chain_order.ev[d]) LG_CUDA(cudaEventCreateWithFlags(&g_chain_order.ev[d], cudaEventDisableTiming));
    if (g_chain_order.recorded[d] && g_chain_order.last[d] != st) LG_CUDA(cudaStreamWaitEvent(st, g_chain_order.ev[d], 0));
    int per_sm = sampler_tune().chain_ctas;
    if (per_sm < 1 || per_sm > 6) per_sm = kChainCtasPerSm;
    const int grid = kSMs * per_sm;
    if (p->rng_kind == LG_RNG_MINSTD) LG_CUDA(lg_launch_opt(pdl_on(s), chain_kernel<LG_RNG_MINSTD>, grid, kBlock, 0, st, c));
    else LG_CUDA(lg_launch_opt(pdl_on(s), chain_kernel<LG_RNG_PHILOX>, grid, kBlock, 0, st, c));
    LG_CUDA(cudaEventRecord(g_chain_order.ev[d], st));
    g_chain_order.last[d] = st;
    g_chain_order.recorded[d] = true;
  }
  s->dirty_batch = *b;
  s->pm_dirty = 1;
  return released(s, st);  // the kernel's last phase releases the position map
}

extern "C" int lg_run_batch(lg_sampler* s, lg_stream_t stream, const lg_topology* topo,
                            const lg_feature_cache* cache, const lg_batch_params* p, const lg_batch* b,
                            unsigned long long* tier_rows) {
  LG_REQUIRE(s && topo && p && b, "lg_run_batch: null argument");
  LG_REQUIRE(p->rng_kind == LG_RNG_MINSTD || p->rng_kind == LG_RNG_PHILOX, "lg_run_batch: rng_kind %d", p->rng_kind);
  LG_REQUIRE(topo->n_parts >= 0 && topo->n_parts <= LG_MAX_DEVICE && topo->indptr[topo->n_parts] &&
                 topo->indices[topo->n_parts], "lg_run_batch: bad topology descriptor");
  LG_REQUIRE(topo->num_nodes <= s->num_nodes, "lg_run_batch: topology has %lld vertices, sampler was created for %lld",
             (long long)topo->num_nodes, (long long)s->num_nodes);
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
  const bool chain = chain_eligible(s, cache);  // the whole sampler chain as one persistent kernel
  int rc = chain ? launch_chain(s, main_st, topo, p, b)
                 : lg_batch_generate(s, stream, p->all_ids, p->all_labels, p->total_cap, p->batch_size, p->counter, b);
  if (rc) return rc;
  int first_pending = 0;  // first hop whose rows have not been gathered yet
  for (int hop = 0; hop <= s->n_hops; hop++) {
    if (hop > 0 && !chain) {
      // the relabel pass of hop h rides in the sample kernel of hop h+1; only the last hop runs its own
      rc = sample_hop(s, main_st, topo, hop, p->rng_kind, p->rng_seed, p->batch_id, p->stream_id, b, nullptr,
                      /*relabel_prev=*/hop > 1, /*relabel_own=*/hop == s->n_hops, /*release=*/hop == s->n_hops);
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
  q.total_cap = p->batch_size;  // the staged seeds ARE the batch
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

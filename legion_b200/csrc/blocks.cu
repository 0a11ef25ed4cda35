// blocks.cu — block construction: COO of a message-flow block -> CSC.
//
// The reference trainer rebuilds a DGL block from the COO edge list on every step
// (training_backend/legion_graphsage.py:66-79: create_unitgraph_from_coo(2, num_src, num_dst, src, dst, 'coo',
// row_sorted=True) although the list is not sorted) and DGL converts it to CSC for the SpMM of every layer.
// lg_block_csc* produce that CSC straight from the CUDA-IPC buffers:
//   indptr[d] .. indptr[d+1]  = the in-edges of destination d (batch-local index), in COO order (stable)
//   indices[k]                = batch-local source of the k-th in-edge
//   eids[k]                   = its position in the COO (DGL edge id), optional
// Stable order makes the result a pure function of the COO, so it is checked bit-exactly against the oracle.
//
// No library sort, and no sort of the EDGES at all.  The sampler emits the edges of a hop grouped by frontier entry
// (engine/operator_impl.cu:208-258: slot idx = entry * fanout + j; here in ascending slot order), so agg_dst is a
// sequence of RUNS of equal destinations, <= fanout edges each: 2.2 M edges of a [25,10] batch are ~0.2 M runs.
// The runs are what gets grouped by destination — a hand-written stable LSD radix sort over (destination, run start),
// 2 passes of <= 10 bits (3 beyond 1 M destinations) — and the edges are then copied run by run:
//   runs_hist_kernel     tile of 4096 edges: run heads (dst[e] != dst[e-1]), the start of the run after each head,
//                        per-tile digit histogram of the heads
//   digit_scan_kernel    exclusive scan of the digit-major [digit][tile] histogram (tiles + cross-tile prefix, scan.cuh)
//   runs_scatter_kernel  pass 1: the tile's heads again, stable in-tile rank per digit, write (dst, start) in digit order
//   pairs_hist_kernel / pairs_scatter_kernel   pass 2 (and 3) over the dense (dst, start) arrays
//   expand_kernel        tile of 256 sorted runs: run lengths, block scan + cross-tile prefix -> output offset of every
//                        run; indptr for the destinations the tile starts; the tile's edges copied in output order
// All blocks of a batch (H cumulative blocks, training_backend/ipc_cuda_kernel.cu:218-231) go through the SAME six
// launches (pass 2's histogram is counted by pass 1's scatter) (lg_block_csc_batch): a CTA finds its block from its index.  The launches are chained with programmatic
// dependent launch like the sampler's.
#include "common.cuh"
#include "scan.cuh"

using namespace lg;

namespace {

constexpr int kItems = 16;                  // edges (runs) per thread of a tile
constexpr int kTile = kBlock * kItems;      // 4096
constexpr int kMaxDigitBits = 10;          // per-warp digit counters of a tile: 8 x 1024 x 4 B of shared memory
constexpr int kMaxBins = 1 << kMaxDigitBits;
constexpr int kExpItems = 1;                // runs per thread of an expansion tile: small tiles, so that the copy work of
constexpr int kExpTile = kBlock * kExpItems;  // a hop-1 block (8000 runs of 25 edges) still spreads over the SMs
constexpr int kMaxBlocks = LG_MAX_HOPS;

struct PrefixRef {  // cross-tile prefix state (scan.cuh), zeroed per call
  int32_t* ticket;
  u64* tile_state;
  u64* anchors;
};

struct CscBlock {
  const int32_t* src;
  const int32_t* dst;
  int64_t n_edges;            // host-known sizes, or upper bounds when the two device words below are set
  int32_t num_dst;
  const int32_t* n_edges_dev; // lg_block_csc_batch: the block's sizes live in the batch's counters on the device
  const int32_t* num_dst_dev;
  int32_t shift, bits;        // digit of the current pass
  int32_t n_tiles;            // edge tiles (= histogram columns)
  int32_t tile0;              // first CTA of this block in the edge-tile kernels
  int32_t scan_tiles, scan_tile0;
  int32_t* hist;              // [bins][n_tiles] digit-major, scanned in place
  int32_t* hist_next;         // pass 1's histogram, counted by the scatter of pass 0 (global atomics: where a run lands
  int32_t next_shift, next_bits;  // tells its tile of the next pass); null / 0 when there is no second pass
  int32_t* key[2];            // (dst, start) of the runs, double buffered
  int32_t* val[2];
  int32_t* n_runs;            // device: number of runs R (written by the first scan)
  int32_t* next_start;        // for a run head e: the start of the next run in COO order (n_edges after the last)
  PrefixRef scan[3], expand;
  int32_t* indptr;
  int32_t* indices;
  int32_t* eids;
};
struct CscMulti {
  CscBlock b[kMaxBlocks];
  int32_t n;
};

// the block a CTA of an edge-tile (or scan-tile) kernel works on, with its device-side sizes resolved
__device__ __forceinline__ CscBlock pick_block(const CscMulti& m, bool scan_grid, int* tile) {
  int k = 0;
#pragma unroll 1
  for (int i = 1; i < m.n; i++)
    if ((int)blockIdx.x >= (scan_grid ? m.b[i].scan_tile0 : m.b[i].tile0)) k = i;
  CscBlock a = m.b[k];
  *tile = (int)blockIdx.x - (scan_grid ? a.scan_tile0 : a.tile0);
  if (a.n_edges_dev) a.n_edges = *a.n_edges_dev;
  if (a.num_dst_dev) a.num_dst = *a.num_dst_dev;
  return a;
}

__device__ __forceinline__ int32_t digit_of(int32_t key, int shift, int bits) { return (key >> shift) & ((1 << bits) - 1); }

// Stable rank of every item of a tile among the tile's items with the same digit.  Item order = (warp, round, lane):
// warp w owns the 512 consecutive items [w * 512, (w + 1) * 512), round r of it the 32 items at r * 32.  Each warp keeps
// its own running count per digit (s_wcnt[w][digit]), so the 16 rounds need no block-wide synchronisation; one scan over
// the 8 warps per digit then turns the counts into bases.  s_wcnt must be zero on entry.
constexpr int kWarps = kBlock / 32;
__device__ __forceinline__ int tile_item(int r) { return (threadIdx.x >> 5) * (kItems * 32) + r * 32 + (threadIdx.x & 31); }
template <typename F>
__device__ __forceinline__ void tile_stable_ranks(int32_t (*s_wcnt)[kMaxBins], int bins, const int32_t (&dig)[kItems],
                                                  const bool (&valid)[kItems], F&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t local[kItems];
#pragma unroll
  for (int r = 0; r < kItems; r++) {  // unrolled: dig / valid / local stay in registers
    const unsigned act = __ballot_sync(0xffffffffu, valid[r]);
    local[r] = 0;
    if (valid[r]) {
      const unsigned peers = __match_any_sync(act, dig[r]);
      const int below = __popc(peers & ((1u << lane) - 1u));
      const int32_t base = s_wcnt[warp][dig[r]];  // every lane of a peer group reads the word before the leader bumps it
      __syncwarp(act);
      if (below == 0) s_wcnt[warp][dig[r]] = base + __popc(peers);
      local[r] = base + below;
    }
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < bins; d += kBlock) {  // counts of the warps -> exclusive bases, per digit
    int32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const int32_t c = s_wcnt[w][d];
      s_wcnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; r++)
    if (valid[r]) emit(r, s_wcnt[warp][dig[r]] + local[r]);
}

// ---- pass 1, first half: run heads of an edge tile + their digit histogram; also next_start ----
__global__ void __launch_bounds__(kBlock) runs_hist_kernel(const CscMulti m) {
  __shared__ int32_t s_hist[kMaxBins];
  __shared__ uint32_t s_head[kTile / 32];  // bit x = edge e0 + x starts a run
  __shared__ int32_t s_after;              // first run start at or beyond the end of the tile
  pdl_prologue();
  int tile;
  const CscBlock a = pick_block(m, false, &tile);
  const int tid = threadIdx.x, lane = tid & 31, bins = 1 << a.bits;
  for (int i = tid; i < bins; i += kBlock) s_hist[i] = 0;
  __syncthreads();
  const int64_t e0 = (int64_t)tile * kTile;
  int32_t d[kItems], dp[kItems];
#pragma unroll
  for (int r = 0; r < kItems; r++) {  // striped: edge e0 + r * kBlock + tid; every load issued before the first use
    const int64_t e = e0 + (int64_t)r * kBlock + tid;
    d[r] = e < a.n_edges ? a.dst[e] : -1;
    dp[r] = (e > 0 && e < a.n_edges) ? a.dst[e - 1] : -1;
  }
#pragma unroll
  for (int r = 0; r < kItems; r++) {
    const int64_t e = e0 + (int64_t)r * kBlock + tid;
    const bool head = e < a.n_edges && (e == 0 || d[r] != dp[r]);
    if (head) atomicAdd(&s_hist[digit_of(d[r], a.shift, a.bits)], 1);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_head[(r * kBlock + tid) >> 5] = hm;
  }
  if (tid == 0) {  // the run that is open at the end of the tile ends where the destination changes (or at n_edges)
    int64_t e = e0 + kTile;
    if (e < a.n_edges) {
      const int32_t d_last = a.dst[e - 1];
      while (e < a.n_edges && a.dst[e] == d_last) e++;
    }
    s_after = (int32_t)(e < a.n_edges ? e : a.n_edges);
  }
  __syncthreads();
  // next_start of every head: the next set bit of the tile's mask, else s_after
  for (int x = tid; x < kTile; x += kBlock) {
    if (!((s_head[x >> 5] >> (x & 31)) & 1u)) continue;
    int w = x >> 5;
    uint32_t hm = (x & 31) == 31 ? 0u : (s_head[w] & (0xFFFFFFFFu << ((x & 31) + 1)));
    while (hm == 0u && ++w < kTile / 32) hm = s_head[w];
    a.next_start[e0 + x] = hm ? (int32_t)(e0 + w * 32 + (__ffs(hm) - 1)) : s_after;
  }
  for (int i = tid; i < bins; i += kBlock) a.hist[(size_t)i * a.n_tiles + tile] = s_hist[i];
}

// ---- exclusive scan (in place) of the digit-major histogram: tiles of 4096 words, cross-tile prefix from scan.cuh ----
__global__ void __launch_bounds__(kBlock) digit_scan_kernel(const CscMulti m, int pass) {
  __shared__ int32_t s_red[kBlock / 32];
  __shared__ int32_t s_tile;
  pdl_prologue();
  int unused;
  const CscBlock a = pick_block(m, true, &unused);
  const PrefixRef& ps = a.scan[pass];
  const int tid = threadIdx.x;
  const int64_t n = (int64_t)(1 << a.bits) * a.n_tiles;
  if (tid == 0) s_tile = atomicAdd(ps.ticket, 1);  // tiles in order: the prefix only waits on running CTAs
  __syncthreads();
  const int tile = s_tile;
  const int n_tiles = (int)((n + kTile - 1) / kTile);
  if (tile >= n_tiles) return;
  const int64_t i0 = (int64_t)tile * kTile + (int64_t)tid * kItems;
  int32_t v[kItems], sum = 0;
#pragma unroll
  for (int k = 0; k < kItems; k += 4) {
    if (i0 + k + 3 < n) {
      const int4 q = *reinterpret_cast<const int4*>(a.hist + i0 + k);
      v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) v[k + j] = (i0 + k + j < n) ? a.hist[i0 + k + j] : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < kItems; k++) sum += v[k];
  int32_t total;
  const int32_t mine = block_exclusive_scan(sum, s_red, &total);
  const int32_t base = block_exclusive_prefix(ps.tile_state, ps.anchors, tile, total, s_red);
  int32_t run = base + mine;
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    if (i0 + k < n) a.hist[i0 + k] = run;
    run += v[k];
  }
  if (pass == 0 && tile == n_tiles - 1 && tid == kBlock - 1) *a.n_runs = run;  // every head was counted once: R
}

// ---- pass 1, second half: the tile's heads in COO order, stable rank per digit, scatter (dst, start) ----
__global__ void __launch_bounds__(kBlock) runs_scatter_kernel(const CscMulti m) {
  __shared__ int32_t s_wcnt[kWarps][kMaxBins];
  pdl_prologue();
  int tile;
  const CscBlock a = pick_block(m, false, &tile);
  const int tid = threadIdx.x, bins = 1 << a.bits;
  for (int w = 0; w < kWarps; w++)
    for (int i = tid; i < bins; i += kBlock) s_wcnt[w][i] = 0;
  const int64_t e0 = (int64_t)tile * kTile;
  int32_t dig[kItems], key[kItems], dp[kItems];
  bool valid[kItems];
#pragma unroll
  for (int r = 0; r < kItems; r++) {  // item order (warp, round, lane) = COO order; loads first
    const int64_t e = e0 + tile_item(r);
    key[r] = e < a.n_edges ? a.dst[e] : -1;
    dp[r] = (e > 0 && e < a.n_edges) ? a.dst[e - 1] : -1;
  }
#pragma unroll
  for (int r = 0; r < kItems; r++) {
    const int64_t e = e0 + tile_item(r);
    valid[r] = e < a.n_edges && (e == 0 || key[r] != dp[r]);
    dig[r] = valid[r] ? digit_of(key[r], a.shift, a.bits) : 0;
  }
  __syncthreads();
  tile_stable_ranks(s_wcnt, bins, dig, valid, [&](int r, int32_t rank) {
    const int32_t pos = a.hist[(size_t)dig[r] * a.n_tiles + tile] + rank;
    a.key[0][pos] = key[r];
    a.val[0][pos] = (int32_t)(e0 + tile_item(r));
    if (a.hist_next) atomicAdd(&a.hist_next[(size_t)digit_of(key[r], a.next_shift, a.next_bits) * a.n_tiles + pos / kTile], 1);
  });
}

// ---- later passes over the dense (dst, start) arrays ----
__global__ void __launch_bounds__(kBlock) pairs_hist_kernel(const CscMulti m, int in) {
  __shared__ int32_t s_hist[kMaxBins];
  pdl_prologue();
  int tile;
  const CscBlock a = pick_block(m, false, &tile);
  const int tid = threadIdx.x, bins = 1 << a.bits;
  for (int i = tid; i < bins; i += kBlock) s_hist[i] = 0;
  __syncthreads();
  const int32_t R = *a.n_runs;
  const int64_t i0 = (int64_t)tile * kTile;
  if (i0 < R) {
    for (int k = tid; k < kTile; k += kBlock)
      if (i0 + k < R) atomicAdd(&s_hist[digit_of(a.key[in][i0 + k], a.shift, a.bits)], 1);
  }
  __syncthreads();
  for (int i = tid; i < bins; i += kBlock) a.hist[(size_t)i * a.n_tiles + tile] = s_hist[i];
}
__global__ void __launch_bounds__(kBlock) pairs_scatter_kernel(const CscMulti m, int in) {
  __shared__ int32_t s_wcnt[kWarps][kMaxBins];
  pdl_prologue();
  int tile;
  const CscBlock a = pick_block(m, false, &tile);
  const int tid = threadIdx.x, bins = 1 << a.bits;
  const int32_t R = *a.n_runs;
  const int64_t i0 = (int64_t)tile * kTile;
  if (i0 >= R) return;
  for (int w = 0; w < kWarps; w++)
    for (int i = tid; i < bins; i += kBlock) s_wcnt[w][i] = 0;
  int32_t dig[kItems], key[kItems], val[kItems];
  bool valid[kItems];
#pragma unroll
  for (int r = 0; r < kItems; r++) {
    const int64_t i = i0 + tile_item(r);
    valid[r] = i < R;
    key[r] = valid[r] ? a.key[in][i] : 0;
    val[r] = valid[r] ? a.val[in][i] : 0;
    dig[r] = digit_of(key[r], a.shift, a.bits);
  }
  __syncthreads();
  tile_stable_ranks(s_wcnt, bins, dig, valid, [&](int r, int32_t rank) {
    const int32_t pos = a.hist[(size_t)dig[r] * a.n_tiles + tile] + rank;
    a.key[in ^ 1][pos] = key[r];
    a.val[in ^ 1][pos] = val[r];
  });
}

// ---- runs sorted by destination -> offsets, indptr, edges ----
__global__ void __launch_bounds__(kBlock) expand_kernel(const CscMulti m, int in) {
  __shared__ int32_t s_red[kBlock / 32];
  __shared__ int32_t s_tile;
  __shared__ int32_t s_off[kExpTile + 1];
  __shared__ int32_t s_start[kExpTile];
  pdl_prologue();
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int bi = 0; bi < m.n; bi++) {
    CscBlock a = m.b[bi];
    if (a.n_edges_dev) a.n_edges = *a.n_edges_dev;
    if (a.num_dst_dev) a.num_dst = *a.num_dst_dev;
    const int32_t R = (a.n_tiles > 0 && a.n_edges > 0) ? *a.n_runs : 0;
    const int n_tiles = (R + kExpTile - 1) / kExpTile;
    if (R == 0) {  // no edge: every destination is empty
      for (int64_t d = (int64_t)blockIdx.x * kBlock + tid; d <= a.num_dst; d += (int64_t)gridDim.x * kBlock) a.indptr[d] = 0;
      continue;
    }
    for (;;) {  // persistent over the tiles, claimed in order: the prefix only waits on running CTAs
      __syncthreads();
      if (tid == 0) s_tile = atomicAdd(a.expand.ticket, 1);
      __syncthreads();
      const int tile = s_tile;
      if (tile >= n_tiles) break;
      const int64_t i0 = (int64_t)tile * kExpTile;
      // blocked order: thread t owns runs i0 + t*kExpItems ..
      int32_t len[kExpItems], sum = 0;
#pragma unroll
      for (int k = 0; k < kExpItems; k++) {
        const int64_t i = i0 + (int64_t)tid * kExpItems + k;
        len[k] = 0;
        int32_t start = 0;
        if (i < R) {
          start = a.val[in][i];
          len[k] = a.next_start[start] - start;
        }
        s_start[tid * kExpItems + k] = start;
        sum += len[k];
      }
      int32_t total;
      const int32_t mine = block_exclusive_scan(sum, s_red, &total);
      const int32_t base = block_exclusive_prefix(a.expand.tile_state, a.expand.anchors, tile, total, s_red);
      int32_t off = base + mine;
#pragma unroll
      for (int k = 0; k < kExpItems; k++) {
        s_off[tid * kExpItems + k] = off;
        off += len[k];
      }
      if (tid == kBlock - 1) s_off[kExpTile] = off;
      __syncthreads();
      // indptr: run i starts the destinations (key[i-1], key[i]]; the last run closes (key[R-1], num_dst]
      for (int k = tid; k < kExpTile; k += kBlock) {
        const int64_t i = i0 + k;
        if (i >= R) break;
        const int32_t cur = a.key[in][i];
        const int32_t prev = i == 0 ? -1 : a.key[in][i - 1];
        for (int32_t d = prev + 1; d <= cur && d <= a.num_dst; d++) a.indptr[d] = s_off[k];
        if (i == R - 1)
          for (int32_t d = cur + 1; d <= a.num_dst; d++) a.indptr[d] = s_off[k + 1];
      }
      // edges in output order: position q belongs to the run found by bisection of the tile's offsets
      const int32_t q0 = s_off[0], q1 = s_off[kExpTile];
      const int32_t* __restrict__ src = a.src;
      int32_t* __restrict__ out_idx = a.indices;
      int32_t* __restrict__ out_eid = a.eids;
      constexpr int U = 8;  // positions per thread and step: their source loads are in flight together
      for (int32_t qb = q0 + tid; qb < q1; qb += kBlock * U) {
        int32_t e[U], v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int32_t q = qb + u * kBlock;
          e[u] = -1;
          if (q < q1) {
            int lo = 0, hi = kExpTile;  // largest k with s_off[k] <= q (empty runs share an offset: the last one wins)
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (s_off[mid] <= q) lo = mid; else hi = mid;
            }
            e[u] = s_start[lo] + (q - s_off[lo]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = e[u] >= 0 ? src[e[u]] : 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int32_t q = qb + u * kBlock;
          if (e[u] >= 0) {
            out_idx[q] = v[u];
            if (out_eid) out_eid[q] = e[u];
          }
        }
      }
    }
  }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~255ll; }
inline int64_t prefix_bytes(int64_t tiles) {
  return align256(128 + ((tiles * 8 + 127) & ~127ll) + (tiles / kGroup + 1) * kGroupStride * 8);
}
PrefixRef prefix_at(char* w, int64_t off, int64_t tiles) {
  PrefixRef p;
  p.ticket = (int32_t*)(w + off);
  p.tile_state = (u64*)(w + off + 128);
  p.anchors = (u64*)(w + off + 128 + ((tiles * 8 + 127) & ~127ll));
  return p;
}
struct Layout {
  int64_t key[2], val[2], next_start, hist, hist_b, small, scan[3], expand, total;
  int64_t scan_tiles_max, exp_tiles_max;
  int32_t n_tiles;
};
Layout layout_of(int64_t max_edges) {
  Layout l;
  l.n_tiles = (int32_t)((max_edges + kTile - 1) / kTile);
  if (l.n_tiles < 1) l.n_tiles = 1;
  int64_t o = 0;
  for (int b = 0; b < 2; b++) { l.key[b] = o; o += align256(max_edges * 4); }
  for (int b = 0; b < 2; b++) { l.val[b] = o; o += align256(max_edges * 4); }
  l.next_start = o; o += align256((max_edges + 1) * 4);
  l.hist = o; o += align256((int64_t)kMaxBins * l.n_tiles * 4);
  // zeroed per call from here on: pass 1's histogram (filled with atomics), [n_runs], the prefix states of the (up to 3)
  // histogram scans and of the expansion
  l.hist_b = o; o += align256((int64_t)kMaxBins * l.n_tiles * 4);
  l.small = o; o += 256;
  l.scan_tiles_max = ((int64_t)kMaxBins * l.n_tiles + kTile - 1) / kTile + 1;
  l.exp_tiles_max = (max_edges + kExpTile - 1) / kExpTile + 1;
  for (int p = 0; p < 3; p++) { l.scan[p] = o; o += prefix_bytes(l.scan_tiles_max); }
  l.expand = o; o += prefix_bytes(l.exp_tiles_max);
  l.total = o;
  return l;
}

struct BlockSpec {  // one block of a call
  const int32_t* src;
  const int32_t* dst;
  int64_t n_edges;
  int32_t num_dst;
  const int32_t* n_edges_dev;
  const int32_t* num_dst_dev;
  int32_t* indptr;
  int32_t* indices;
  int32_t* eids;
};

int csc_impl(cudaStream_t st, int n_blocks, const BlockSpec* spec, void* workspace, int64_t workspace_bytes, bool pdl) {
  LG_REQUIRE(n_blocks >= 1 && n_blocks <= kMaxBlocks, "lg_block_csc: %d blocks", n_blocks);
  char* w = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  int64_t need = 256;
  for (int b = 0; b < n_blocks; b++) need += layout_of(spec[b].n_edges).total;
  LG_REQUIRE(workspace && workspace_bytes >= need, "lg_block_csc: workspace of %lld bytes, need %lld", (long long)workspace_bytes,
             (long long)need);
  CscMulti m;
  memset(&m, 0, sizeof(m));
  m.n = n_blocks;
  int key_bits[kMaxBlocks], passes = 1;
  int64_t o = 0, edge_tiles = 0, exp_tiles = 0, max_dst_all = 0;
  bool any_edges = false;
  for (int b = 0; b < n_blocks; b++) {
    const BlockSpec& s = spec[b];
    LG_REQUIRE(s.indptr && s.num_dst >= 0, "lg_block_csc: null indptr / negative num_dst");
    LG_REQUIRE(s.n_edges >= 0 && s.n_edges < (1ll << 31), "lg_block_csc: n_edges %lld", (long long)s.n_edges);
    LG_REQUIRE(s.n_edges == 0 || (s.src && s.dst && s.indices), "lg_block_csc: null argument");
    const Layout l = layout_of(s.n_edges);
    CscBlock& a = m.b[b];
    a.src = s.src;
    a.dst = s.dst;
    a.n_edges = s.n_edges;
    a.num_dst = s.num_dst;
    a.n_edges_dev = s.n_edges_dev;
    a.num_dst_dev = s.num_dst_dev;
    a.n_tiles = s.n_edges > 0 ? l.n_tiles : 0;
    a.tile0 = (int32_t)edge_tiles;
    edge_tiles += a.n_tiles;
    for (int k = 0; k < 2; k++) {
      a.key[k] = (int32_t*)(w + o + l.key[k]);
      a.val[k] = (int32_t*)(w + o + l.val[k]);
    }
    a.next_start = (int32_t*)(w + o + l.next_start);
    a.hist = (int32_t*)(w + o + l.hist);
    a.hist_next = (int32_t*)(w + o + l.hist_b);
    a.n_runs = (int32_t*)(w + o + l.small);
    for (int p = 0; p < 3; p++) a.scan[p] = prefix_at(w, o + l.scan[p], l.scan_tiles_max);
    a.expand = prefix_at(w, o + l.expand, l.exp_tiles_max);
    a.indptr = s.indptr;
    a.indices = s.indices;
    a.eids = s.eids;
    LG_CUDA(cudaMemsetAsync(w + o + l.hist_b, 0, (size_t)(l.total - l.hist_b), st));
    o += l.total;
    // digits: the significant bits of the destinations, split evenly over the fewest passes of <= 10 bits
    key_bits[b] = 1;
    while (key_bits[b] < 31 && (1ll << key_bits[b]) < (int64_t)s.num_dst) key_bits[b]++;
    const int pb = (key_bits[b] + kMaxDigitBits - 1) / kMaxDigitBits;
    if (s.n_edges > 0 && pb > passes) passes = pb;
    if (s.n_edges > 0) any_edges = true;
    exp_tiles += (s.n_edges + kExpTile - 1) / kExpTile;
    if (s.num_dst > max_dst_all) max_dst_all = s.num_dst;
  }
  int in = 0;
  int32_t* hist_a[kMaxBlocks];
  int32_t* hist_b[kMaxBlocks];
  for (int b = 0; b < n_blocks; b++) {
    hist_a[b] = m.b[b].hist;
    hist_b[b] = m.b[b].hist_next;
  }
  auto digit = [&](int b, int p, int32_t* shift, int32_t* bits) {
    const int per = (key_bits[b] + passes - 1) / passes;
    *shift = p * per;
    *bits = key_bits[b] - p * per;
    if (*bits > per) *bits = per;
    if (*bits < 1) *bits = 1;
  };
  if (any_edges) {
    for (int p = 0; p < passes; p++) {
      int64_t scan_grid = 0;
      for (int b = 0; b < n_blocks; b++) {  // every block runs the same number of passes; a narrow key just has narrow digits
        CscBlock& a = m.b[b];
        digit(b, p, &a.shift, &a.bits);
        a.hist = (p == 1) ? hist_b[b] : hist_a[b];  // pass 1's histogram was counted by the scatter of pass 0
        a.hist_next = nullptr;
        a.next_shift = a.next_bits = 0;
        if (p == 0 && passes > 1) {
          a.hist_next = hist_b[b];
          digit(b, 1, &a.next_shift, &a.next_bits);
        }
        a.scan_tile0 = (int32_t)scan_grid;
        a.scan_tiles = a.n_tiles > 0 ? (int32_t)(((int64_t)(1 << a.bits) * a.n_tiles + kTile - 1) / kTile) : 0;
        scan_grid += a.scan_tiles;
      }
      if (p == 0) {
        LG_CUDA(lg_launch_opt(pdl, runs_hist_kernel, (int)edge_tiles, kBlock, 0, st, m));
        LG_CUDA(lg_launch_opt(pdl, digit_scan_kernel, (int)scan_grid, kBlock, 0, st, m, p));
        LG_CUDA(lg_launch_opt(pdl, runs_scatter_kernel, (int)edge_tiles, kBlock, 0, st, m));
        in = 0;
      } else {  // at most as many runs as edges: the same tile count covers them
        if (p >= 2) LG_CUDA(lg_launch_opt(pdl, pairs_hist_kernel, (int)edge_tiles, kBlock, 0, st, m, in));
        LG_CUDA(lg_launch_opt(pdl, digit_scan_kernel, (int)scan_grid, kBlock, 0, st, m, p));
        LG_CUDA(lg_launch_opt(pdl, pairs_scatter_kernel, (int)edge_tiles, kBlock, 0, st, m, in));
        in ^= 1;
      }
    }
  }
  // persistent over the run tiles; every CTA that starts finishes its tile before it claims another, so any grid works
  int64_t grid = any_edges ? exp_tiles : (max_dst_all + 1 + kBlock - 1) / kBlock;
  if (grid > (int64_t)kSMs * 8) grid = kSMs * 8;
  if (grid < 1) grid = 1;
  LG_CUDA(lg_launch_opt(pdl, expand_kernel, (int)grid, kBlock, 0, st, m, in));
  return 0;
}

}  // namespace

extern "C" int lg_block_csc_workspace(int64_t max_edges, int64_t* bytes) {
  LG_REQUIRE(bytes && max_edges >= 0 && max_edges < (1ll << 31), "lg_block_csc_workspace: bad argument");
  *bytes = layout_of(max_edges).total + 512;
  return 0;
}

extern "C" int lg_block_csc(lg_stream_t stream, const int32_t* agg_src, const int32_t* agg_dst, int64_t n_edges,
                            int32_t num_dst, int32_t* indptr, int32_t* indices, int32_t* eids, void* workspace,
                            int64_t workspace_bytes) {
  const BlockSpec s = {agg_src, agg_dst, n_edges, num_dst, nullptr, nullptr, indptr, indices, eids};
  return csc_impl((cudaStream_t)stream, 1, &s, workspace, workspace_bytes, lg_pdl() != 0);
}

extern "C" int lg_block_csc_batch_workspace(int32_t n_hops, const int64_t* max_edges, int64_t* bytes) {
  LG_REQUIRE(bytes && max_edges && n_hops >= 1 && n_hops <= LG_MAX_HOPS, "lg_block_csc_batch_workspace: bad argument");
  int64_t tot = 512;
  for (int h = 0; h < n_hops; h++) {
    LG_REQUIRE(max_edges[h] > 0 && max_edges[h] < (1ll << 31), "lg_block_csc_batch_workspace: max_edges[%d]", h);
    tot += layout_of(max_edges[h]).total;
  }
  *bytes = tot;
  return 0;
}

// Every block of a batch (block h = edges [0, ec[9+h]), destinations [0, nc[9+h-1])) in ONE set of launches and without
// any host-side knowledge of the sizes: they are read from the batch's counters on the device, so the call can be
// enqueued behind the sampling ops of a batch that has not run yet (the server builds the blocks next to the gather).
extern "C" int lg_block_csc_batch(lg_stream_t stream, const lg_batch* batch, int32_t n_hops, const int64_t* max_edges,
                                  const int32_t* max_dst, int32_t* const* indptr, int32_t* const* indices,
                                  int32_t* const* eids, void* workspace, int64_t workspace_bytes) {
  LG_REQUIRE(batch && n_hops >= 1 && n_hops <= LG_MAX_HOPS && max_edges && max_dst && indptr && indices,
             "lg_block_csc_batch: bad argument");
  BlockSpec s[kMaxBlocks];
  for (int h = 1; h <= n_hops; h++) {
    LG_REQUIRE(max_edges[h - 1] > 0 && max_dst[h - 1] > 0, "lg_block_csc_batch: bounds of block %d must be positive", h);
    s[h - 1] = {batch->agg_src, batch->agg_dst, max_edges[h - 1], max_dst[h - 1],
                batch->edge_counter + LG_INTRABATCH_CON * 3 + h, batch->node_counter + LG_INTRABATCH_CON * 3 + h - 1,
                indptr[h - 1], indices[h - 1], eids ? eids[h - 1] : nullptr};
  }
  return csc_impl((cudaStream_t)stream, n_hops, s, workspace, workspace_bytes, lg_pdl() != 0);
}

// gather.cu — unified-cache lookup + feature gather for sm_100a.
//
// Replaces  UnifiedCache::FindFeat (bcht probe kernel + blocking D2H, cache/cache.cu:180-215)  and
// multiGPU_feat_cache_lookup (one thread per float, 4-byte loads, 64-bit div+mod per element,
// 32 CTAs: cache/cache_impl.cuh:239-272; cache/cache.cu:726-748)  with one launch per op:
// the location of a row (local HBM shard / peer HBM shard over NVLink / backing matrix in host
// pinned memory or HBM) is decoded ONCE per row from a dense int32 directory, rows move as
// 16-byte vectors.  Two data movers:
//   LDG  — warp per row, R rows in flight per warp, ld.global.nc.L1::no_allocate.v4 -> st.global.v4
//   TMA  — one warp per CTA; every lane issues a cp.async.bulk (row -> shared memory stage,
//          mbarrier complete_tx) — or, for 4 rows that all live in the local shard, one lane issues ONE
//          cp.async.bulk.tensor tile::gather4 over the shard's tensor map —, lane 0 drains a finished stage
//          with ONE bulk store of the tile's contiguous destination rows.  No registers touch the payload
//          (SASS: UBLKCP, UTMALDG.2D.GATHER4).
// Values are moved bit-for-bit (no arithmetic), so the output is bit-exact by construction.
#include <cuda.h>

#include "common.cuh"
#include "sampler_state.cuh"

using namespace lg;

namespace {

struct GatherArgs {
  lg_feature_cache cache;
  const int32_t* ids;  // row r reads ids[off + r]
  float* dst;          // row r writes dst[(off + r) * dim ...]
  int32_t* nc;         // non-null: (off, cnt) come from the counter protocol for `hop`
  int64_t off, cnt;    // used when nc == null
  int64_t dst_rows;    // capacity of dst in rows
  int32_t hop;         // op_id / 3
  int32_t hop_lo;      // first hop whose rows this launch moves (== hop unless the batch's gathers are fused)
  int32_t op_slot;     // op_id % 3 (snapshot slot, engine/operator_impl.cu:83-85)
  int32_t local_part;
  u64* tier;           // [3] local / peer / miss rows, may be null
  int32_t* status;
  int32_t l2;          // lg_l2_hints()
  int32_t* ticket;     // [2] zeroed: tiles are claimed dynamically (robust to CTAs that start late or run slowed down next to
                       // another kernel); null = static round-robin over the grid
  int32_t chunk;       // consecutive tiles per chunk (static order) / per claim (dynamic)
  int32_t static_pct;  // dynamic: share of a CTA's fair share that keeps the static order (no atomics)
  int32_t use_g4;      // TMA mover: groups of 4 rows that all live in the local shard move as ONE tile::gather4 tensor copy
  alignas(64) CUtensorMap tmap;  // 2-D view of shard[local_part] ([rows][dim] fp32, box = one row) for those copies
};

__device__ __forceinline__ void row_range(const GatherArgs& a, int64_t* off, int64_t* cnt) {
  if (a.nc) {
    // rows of hop h are [nc[9+h-1], nc[9+h]) — identical to (nc[0], nc[1]) right after op 3h
    // (engine/operator_impl.cu:65-81) but immutable afterwards, so a later hop may already run.
    int32_t lo = a.hop == 0 ? 0 : ld_counter(a.nc + LG_INTRABATCH_CON * 3 + a.hop - 1);
    int32_t hi = ld_counter(a.nc + LG_INTRABATCH_CON * 3 + a.hop);
    int32_t first = a.hop_lo == 0 ? 0 : ld_counter(a.nc + LG_INTRABATCH_CON * 3 + a.hop_lo - 1);
    *off = first;
    *cnt = hi - first;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the op's counter_update (:83-85): snapshot of hop `hop`
      a.nc[a.op_slot * 2] = lo;
      a.nc[a.op_slot * 2 + 1] = hi - lo;
    }
  } else {
    *off = a.off;
    *cnt = a.cnt;
  }
}

// location decode: returns the source row pointer or nullptr (row is skipped), tier in *t
__device__ __forceinline__ const float* locate(const GatherArgs& a, int32_t id, int* t) {
  *t = -1;
  if (id < 0) return nullptr;  // -1 padding of a tail batch (cache_impl.cuh:263-264)
  if (a.cache.flags & LG_CACHE_IDENTITY) {  // every row is resident at its own index: no directory
    *t = 0;
    return a.cache.shard[a.local_part] + (int64_t)(id % a.cache.num_nodes) * a.cache.dim;
  }
  int32_t gidx = LG_CACHEMISS_FLAG;
  if (a.cache.directory && id < a.cache.num_nodes) gidx = a.cache.directory[id];
  if (gidx < 0) {  // miss -> backing matrix (cache_impl.cuh:262-266)
    *t = 2;
    if (!a.cache.backing) {  // fully-cached deployment without a backing matrix: a miss is an error
      *a.status = 3;
      return nullptr;
    }
    return a.cache.backing + (int64_t)(id % a.cache.num_nodes) * a.cache.dim;
  }
  int32_t didx = gidx / a.cache.shard_rows;  // cache_impl.cuh:259-260
  int32_t fidx = gidx - didx * a.cache.shard_rows;
  *t = (didx == a.local_part) ? 0 : 1;
  return a.cache.shard[didx] + (int64_t)fidx * a.cache.dim;
}

__device__ __forceinline__ float4 ld_nc_v4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_v4(float* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void tier_flush(const GatherArgs& a, int32_t t0, int32_t t1, int32_t t2, int lane) {
  if (!a.tier) return;
  t0 = warp_sum(t0);
  t1 = warp_sum(t1);
  t2 = warp_sum(t2);
  if (lane == 0) {
    if (t0) atomicAdd(a.tier + 0, (u64)t0);
    if (t1) atomicAdd(a.tier + 1, (u64)t1);
    if (t2) atomicAdd(a.tier + 2, (u64)t2);
  }
}

// ---------------- LDG mover: warp per R consecutive rows ----------------
// (an L2 eviction-policy variant — evict-first output stores, evict-last loads for the hottest ranks —
//  was measured and lost 4-12 %: profiles/r01_gather_sweep_v1.txt)
template <int R, bool VEC4>
__global__ void __launch_bounds__(256) gather_ldg_kernel(const GatherArgs a) {
  int64_t off, cnt;
  row_range(a, &off, &cnt);
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int dim = a.cache.dim;
  int32_t t0 = 0, t1 = 0, t2 = 0;
  for (int64_t r0 = warp_global * R; r0 < cnt; r0 += n_warps * R) {
    const float* src = nullptr;
    if (lane < R && r0 + lane < cnt) {
      int64_t row = off + r0 + lane;
      if (row < a.dst_rows) {
        int t;
        src = locate(a, a.ids[row], &t);
        t0 += (t == 0);
        t1 += (t == 1);
        t2 += (t == 2);
      } else {
        *a.status = 2;
      }
    }
    if (VEC4) {
      const int d4 = dim >> 2;
      for (int c0 = 0; c0 < d4; c0 += 32) {
        const int c = c0 + lane;
        float4 v[R];
        const float* sk[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
          sk[k] = (const float*)__shfl_sync(0xffffffffu, (u64)src, k);
          if (sk[k] && c < d4) v[k] = ld_nc_v4(sk[k] + 4 * c);
        }
#pragma unroll
        for (int k = 0; k < R; k++)
          if (sk[k] && c < d4) st_v4(a.dst + (off + r0 + k) * dim + 4 * c, v[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) {
        const float* s = (const float*)__shfl_sync(0xffffffffu, (u64)src, k);
        if (!s) continue;
        float* d = a.dst + (off + r0 + k) * dim;
        for (int c = lane; c < dim; c += 32) d[c] = __ldg(s + c);
      }
    }
  }
  tier_flush(a, t0, t1, t2, lane);
}

// ---------------- TMA mover: bulk async copies through shared memory ----------------
// The tile counter's atomicAdd as inline PTX: written as `if (lane == 0) c = atomicAdd(...)`, nvcc placed the broadcast of
// the result (SHFL) directly behind the ATOMG — ptxas aggregates an `atom.add` over the active lanes (VOTE / POPC / elected
// ATOMG / SHFL of the result), also for inline PTX, so the warp waited for every claim's round trip although the value is
// only needed a whole chunk later.  `atom.inc` (wrap bound 2^31-1: a plain increment here) is not aggregated: the result
// stays in lane 0's register until claim() reads it.
__device__ __forceinline__ int32_t atom_inc_deferred(int32_t* p) {
  uint32_t old;
  asm volatile("atom.relaxed.gpu.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "l"(p) : "memory");
  return (int32_t)old;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, u64 pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
// four rows of the 2-D table behind `tmap`, picked by row index, land back to back in shared memory (sm_100 TMA gather mode)
__device__ __forceinline__ void tensor_gather4(uint32_t dst_smem, const CUtensorMap* tmap, int32_t r0, int32_t r1, int32_t r2,
                                               int32_t r3, uint32_t bar, u64 pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(dst_smem),
      "l"(tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes, u64 pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem), "r"(bytes),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Rows per stage (tile) of one single-warp CTA.  Every row is its own bulk copy, and the compiler serialises a
// per-thread cp.async.bulk over the active lanes (ELECT / R2UR x3 / UBLKCP / loop: ~8 dependent-latency instructions
// per row), so a 32-row tile keeps its warp busy issuing for >1000 cycles.  Alone that is still 2x faster than DRAM,
// but next to ANY other resident warp (one spinning warp per SM is enough, profiles/r01b_overlap.md) the issue loop
// becomes the bottleneck and the gather loses a third of its bandwidth.  Smaller tiles spread the same rows in flight
// over more warps (8-row tiles: 16 warps per SM, 4 per scheduler), so the issue chains interleave.
// Pipeline: iteration `it` fills stage it % STAGES with tile(it) and drains tile(it - LAG),
// LAG = STAGES - 2: LAG tiles of row loads in flight, one stage being stored, one being refilled.
template <int STAGES, int kTmaRows>
__global__ void __maxnreg__(40) gather_tma_kernel(const __grid_constant__ GatherArgs a) {
  static_assert(STAGES >= 3, "need a stage in store and a stage in refill besides the loads in flight");
  constexpr int LAG = STAGES - 2;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) u64 bars[STAGES];
  __shared__ unsigned s_valid[STAGES];
  int64_t off, cnt;
  row_range(a, &off, &cnt);
  const int lane = threadIdx.x;
  const u64 pol_ld = l2_policy((a.l2 & 1) ? 2 : 0), pol_st = l2_policy((a.l2 & 2) ? 2 : 0), keep = l2_policy((a.l2 & 4) ? 1 : 0);
  const uint32_t row_bytes = (uint32_t)a.cache.dim * 4u;
  const uint32_t stage_bytes = row_bytes * kTmaRows;
  const int64_t n_tiles = (cnt + kTmaRows - 1) / kTmaRows;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  int32_t t0 = 0, t1 = 0, t2 = 0;
  __shared__ int s_tileidx[STAGES];

  // Tile order.  A CTA works on chunks of a.chunk consecutive tiles (consecutive ids, consecutive destination rows).
  // Static (default, chunk = 1): chunk k of CTA b is b + k * gridDim.x.  Dynamic (a.ticket): a CTA's first chunk is its
  // block index, every later one comes from a global counter whose atomic is issued a whole chunk BEFORE its result is
  // needed — a CTA that starts late (its SM still busy with another kernel's CTAs) or whose warp gets fewer issue slots
  // simply ends up moving fewer tiles instead of holding the whole launch back with a fixed share.  Claims are monotone:
  // once one fails, every later one would fail too.  Measured (profiles/r02t_gather_tile_order.md): alone the launch goes
  // from 0.90-0.92 to 0.96 of the HBM peak with 4 tiles per claim, next to the sampler's kernels it is slower than the
  // static order, which therefore stays the default; one claim per tile is bound by the counter (one address takes
  // ~0.35 G atomics/s).
  // (tile indices are 32-bit: a launch moves fewer than 2^31 ROWS, checked by the host)
  const int nt = (int)n_tiles, grid = (int)gridDim.x;
  int chunk = a.chunk > 0 ? a.chunk : 1;
  if (chunk > nt / grid) chunk = nt / grid > 0 ? nt / grid : 1;  // few tiles: every CTA gets some
  // dynamic: the first k_static rounds of chunks (a.static_pct % of a CTA's fair share) keep the static order and cost
  // no atomics; only the rest of the launch is handed out by the counter
  int k_static = 1;
  if (a.ticket && a.static_pct > 0) k_static = (int)((int64_t)((nt + chunk - 1) / chunk / grid) * a.static_pct / 100);
  if (k_static < 1) k_static = 1;
  int static_left = a.ticket ? k_static - 1 : 0x7fffffff;
  int ch_index = blockIdx.x;  // index of the chunk being worked on (static order) / base of the counter's chunks (dynamic)
  int ch_next = 0;            // next tile of the chunk being worked on, relative to it
  int32_t ch_pending = 0;     // dynamic, lane 0: the counter value of the chunk after this one (in flight until it is used)
  if (a.ticket && static_left == 0 && lane == 0) ch_pending = atom_inc_deferred(a.ticket);
  auto claim = [&]() -> int {
    if (ch_next == chunk) {
      ch_next = 0;
      if (static_left > 0) {
        ch_index += grid;
        if (--static_left == 0 && lane == 0) ch_pending = atom_inc_deferred(a.ticket);  // (never reached without a counter)
      } else {
        ch_index = k_static * grid + __shfl_sync(0xffffffffu, ch_pending, 0);
        if (lane == 0 && ch_index * chunk < nt) ch_pending = atom_inc_deferred(a.ticket);
      }
    }
    const int t = ch_index * chunk + ch_next++;
    return t < nt ? t : -1;
  };

  // The id -> directory -> pointer chain is software-pipelined so that no iteration waits on it:
  // ids are loaded two tiles ahead, directory entries one tile ahead, pointers are formed at use.
  auto load_id = [&](int tile) -> int32_t {
    const int64_t r = (int64_t)tile * kTmaRows + lane;
    if (tile < 0 || lane >= kTmaRows || r >= cnt) return -1;
    if (off + r >= a.dst_rows) {
      *a.status = 2;
      return -1;
    }
    return a.ids[off + r];
  };
  const bool identity = (a.cache.flags & LG_CACHE_IDENTITY) != 0;
  auto load_loc = [&](int32_t id) -> int32_t {
    if (identity) return id;  // the row index IS the vertex id (see form_ptr)
    if (id < 0 || !a.cache.directory || id >= a.cache.num_nodes) return LG_CACHEMISS_FLAG;
    return ld_nc_s32_hint(a.cache.directory + id, keep);
  };
  // *lrow: the row's index inside the local shard, or -1 when it lives elsewhere (peer shard, backing matrix)
  auto form_ptr = [&](int32_t id, int32_t gidx, int32_t* lrow) -> const float* {
    *lrow = -1;
    if (id < 0) return nullptr;  // -1 padding / out of range rows are skipped (cache_impl.cuh:263-264)
    if (identity) {
      t0++;
      *lrow = (int32_t)(id % a.cache.num_nodes);
      return a.cache.shard[a.local_part] + (int64_t)*lrow * a.cache.dim;
    }
    if (gidx < 0) {              // miss -> backing matrix (cache_impl.cuh:262-266)
      t2++;
      if (!a.cache.backing) {
        *a.status = 3;
        return nullptr;
      }
      return a.cache.backing + (int64_t)(id % a.cache.num_nodes) * a.cache.dim;
    }
    const int32_t didx = gidx / a.cache.shard_rows;  // cache_impl.cuh:259-260
    const int32_t fidx = gidx - didx * a.cache.shard_rows;
    if (didx == a.local_part) {
      t0++;
      *lrow = fidx;
    } else {
      t1++;
    }
    return a.cache.shard[didx] + (int64_t)fidx * a.cache.dim;
  };
  int tile0 = claim();
  int tile1 = tile0 >= 0 ? claim() : -1;
  int32_t id0 = load_id(tile0), id1 = load_id(tile1);
  int32_t loc0 = load_loc(id0);
  int n_issued = 0;  // tiles whose loads this CTA issued; the drain runs LAG iterations behind
  for (int it = 0; tile0 >= 0 || it < n_issued + LAG; it++) {
    const int tile2 = tile1 >= 0 ? claim() : -1;
    const int32_t id2 = load_id(tile2);
    const int32_t loc1 = load_loc(id1);
    int32_t lrow;
    const float* src = form_ptr(id0, loc0, &lrow);
    if (tile0 >= 0) {
      const int s = (int)(it % STAGES);
      // stage s was read by the store of tile(it - STAGES), issued two iterations ago: only the
      // store issued in the previous iteration may still be reading shared memory
      bulk_wait_read<1>();
      __syncwarp();
      const unsigned valid = __ballot_sync(0xffffffffu, src != nullptr);
      const uint32_t bar = smem_u32(&bars[s]);
      if (lane == 0) {
        s_valid[s] = valid;
        s_tileidx[s] = tile0;
        mbar_expect_tx(bar, row_bytes * (uint32_t)__popc(valid));
      }
      __syncwarp();
      const uint32_t stage = smem_u32(smem + (size_t)s * stage_bytes);
      if (a.use_g4) {
        // a group of 4 consecutive rows of the tile that all live in the local shard is ONE tensor copy issued by the
        // group's first lane (a quarter of the per-row UBLKCP issue chains); any other group moves row by row
        const unsigned lmask = __ballot_sync(0xffffffffu, lrow >= 0);
#pragma unroll
        for (int g = 0; g < kTmaRows / 4; g++) {
          if (((lmask >> (4 * g)) & 0xFu) == 0xFu) {
            const int32_t r0 = __shfl_sync(0xffffffffu, lrow, 4 * g), r1 = __shfl_sync(0xffffffffu, lrow, 4 * g + 1),
                          r2 = __shfl_sync(0xffffffffu, lrow, 4 * g + 2), r3 = __shfl_sync(0xffffffffu, lrow, 4 * g + 3);
            if (lane == 4 * g) tensor_gather4(stage + (uint32_t)(4 * g) * row_bytes, &a.tmap, r0, r1, r2, r3, bar, pol_ld);
          } else if ((lane >> 2) == g && src) {
            bulk_g2s(stage + (uint32_t)lane * row_bytes, src, row_bytes, bar, pol_ld);
          }
        }
      } else if (src) {
        bulk_g2s(stage + (uint32_t)lane * row_bytes, src, row_bytes, bar, pol_ld);
      }
      n_issued = it + 1;
    }
    tile0 = tile1;
    tile1 = tile2;
    id0 = id1;
    id1 = id2;
    loc0 = loc1;
    const int dt = it - LAG;
    if (dt >= 0 && dt < n_issued) {
      const int s = (int)(dt % STAGES);
      const uint32_t parity = (uint32_t)((dt / STAGES) & 1);
      const int64_t tile = s_tileidx[s];
      const int64_t row0 = off + tile * kTmaRows;
      const unsigned valid = s_valid[s];
      mbar_wait(smem_u32(&bars[s]), parity);
      const int64_t left = cnt - tile * kTmaRows;
      const int rows_here = left < kTmaRows ? (int)left : kTmaRows;
      const unsigned want = (rows_here >= 32) ? 0xffffffffu : ((1u << rows_here) - 1u);  // lanes >= kTmaRows never hold a row
      if (valid == want) {
        // every row of the tile is present: the destination rows are contiguous -> ONE bulk store
        if (lane == 0) bulk_s2g(a.dst + row0 * a.cache.dim, smem_u32(smem + (size_t)s * stage_bytes),
                                row_bytes * (uint32_t)rows_here, pol_st);
      } else if (valid & (1u << lane)) {  // holes (-1 padded seeds): row-wise stores
        bulk_s2g(a.dst + (row0 + lane) * a.cache.dim,
                 smem_u32(smem + (size_t)s * stage_bytes + (size_t)lane * row_bytes), row_bytes, pol_st);
      }
      bulk_commit();  // every lane commits one (possibly empty) group per drained tile
      __syncwarp();
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tier_flush(a, t0, t1, t2, lane);
  if (a.ticket && lane == 0) {  // the last CTA out re-arms the counters for the next launch on this stream
    __threadfence();  // this CTA's claims (also the one it never looked at) are performed before it counts as done
    if (atomicAdd(a.ticket + 1, 1) == (int32_t)gridDim.x - 1) {
      __threadfence();
      a.ticket[0] = 0;
      a.ticket[1] = 0;
    }
  }
}

// tuning knobs (environment, read once): LG_LDG_R rows in flight per warp {4,8}, LG_LDG_CTAS resident CTAs per SM assumed
// when sizing the LDG grid; TMA mover: LG_TMA_ROWS rows per tile {8,16,32}, LG_TMA_STAGES {3,4,6}, LG_GATHER_SMEM_KB shared
// memory the gather may hold per SM, LG_TMA_CTAS cap on CTAs per SM in units of 32-row tiles, LG_GATHER_CARVEOUT (%).
// Defaults (profiles/r01d_gather_footprint.md): 8-row tiles, 3 stages, 130 KB per SM.  The gather runs next to the
// sampler's kernels, and whatever shared memory it configures on an SM is taken from the L1 of the sampler CTAs on that
// SM: with 32-row tiles of 512-byte rows (4 CTAs x 48 KB = the whole array) the hashed sampler slowed down by 40 % while
// co-resident.  8-row tiles on 10-12 single-warp CTAs keep as many bytes in flight with 130 KB (alone: 0.81 of the HBM peak
// instead of 0.85 at D=128, unchanged at D=100) and leave 124 KB of L1: UK-Union shape 16.4 -> 18.9 M seeds/s.
struct Tune {
  int ldg_r, tma_stages, ldg_ctas, tma_ctas, tma_rows, carveout, smem_kb, chunk;
};
static const Tune& tune() {
  static Tune t = [] {
    Tune x{8, 3, 8, 8, 8, -1, 130, 1};  // LDG: R=8 rows per warp (profiles/r01_gather_sweep_v3.txt)
    if (const char* e = getenv("LG_LDG_R")) x.ldg_r = atoi(e);
    if (const char* e = getenv("LG_TMA_STAGES")) x.tma_stages = atoi(e);
    if (const char* e = getenv("LG_LDG_CTAS")) x.ldg_ctas = atoi(e);
    if (const char* e = getenv("LG_TMA_CTAS")) x.tma_ctas = atoi(e);  // cap in units of 32-row tiles
    if (const char* e = getenv("LG_TMA_ROWS")) x.tma_rows = atoi(e);
    if (const char* e = getenv("LG_GATHER_SMEM_KB")) x.smem_kb = atoi(e);
    if (const char* e = getenv("LG_GATHER_CARVEOUT")) x.carveout = atoi(e);
    if (const char* e = getenv("LG_GATHER_CHUNK")) x.chunk = atoi(e) > 0 ? atoi(e) : 1;  // tiles per chunk, static order
    return x;
  }();
  return t;
}

// ---- tile::gather4: the tensor map over the local shard -------------------------------------------------------------
// LG_GATHER4: 0 = per-row bulk copies only, 1 (default) = tensor gathers where they apply.  They apply when the local shard can be
// described as a 2-D fp32 tensor whose 4-row groups keep shared memory 128-byte aligned (dim % 8 == 0), a row fits one box
// (dim <= 256 elements) and the table has fewer than 2^31 rows.  The driver entry point comes from cudaGetDriverEntryPoint
// (no link-time dependency on libcuda, like the VMM calls in runtime.cu).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
static int gather4_mode() {  // read per launch (a getenv, ~0.1 us) so that a process can switch it, e.g. the tests
  const char* e = getenv("LG_GATHER4");
  return e ? atoi(e) : 1;
}
static int gather4_promotion() {
  static int m = [] {
    const char* e = getenv("LG_GATHER4_PROMO");  // 0 none, 1 64 B, 2 128 B, 3 256 B
    return e ? atoi(e) : 2;
  }();
  return m;
}
// fills a->tmap / a->use_g4; never fails the launch: without a tensor map the rows move one bulk copy each
static void setup_gather4(GatherArgs* a) {
  a->use_g4 = 0;
  memset(&a->tmap, 0, sizeof(a->tmap));
  if (gather4_mode() == 0) return;
  const lg_feature_cache& c = a->cache;
  const bool identity = (c.flags & LG_CACHE_IDENTITY) != 0;
  if (a->local_part < 0 || a->local_part >= LG_MAX_DEVICE) return;
  if (!identity && (!c.directory || a->local_part >= c.n_parts)) return;
  const float* base = c.shard[a->local_part];
  const int64_t rows = identity ? c.num_nodes : (int64_t)c.shard_rows;
  if (!base || rows <= 0 || rows >= (1ll << 31) || c.dim % 8 != 0 || c.dim > 256 || ((uintptr_t)base & 15)) return;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return;
  struct Cached {
    const float* base;
    int64_t rows;
    int32_t dim, promo;
    CUtensorMap map;
  };
  static thread_local Cached cache[8] = {};
  static thread_local int next = 0;
  const int promo = gather4_promotion();
  for (const Cached& k : cache)
    if (k.base == base && k.rows == rows && k.dim == c.dim && k.promo == promo) {
      a->tmap = k.map;
      a->use_g4 = 1;
      return;
    }
  const cuuint64_t gdim[2] = {(cuuint64_t)c.dim, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)c.dim * 4};  // bytes between rows
  const cuuint32_t box[2] = {(cuuint32_t)c.dim, 1};       // gather mode: one row per index, four indices per copy
  const cuuint32_t estride[2] = {1, 1};
  const CUtensorMapL2promotion l2p = promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                     : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                  : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  CUtensorMap m;
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return;
  Cached& slot = cache[next++ % 8];
  slot.base = base;
  slot.rows = rows;
  slot.dim = c.dim;
  slot.promo = promo;
  slot.map = m;
  a->tmap = m;
  a->use_g4 = 1;
}

template <int STAGES, int ROWS>
int launch_tma(cudaStream_t st, const GatherArgs& a, int64_t max_rows) {
  const size_t smem = (size_t)a.cache.dim * 4 * ROWS * STAGES;
  LG_CUDA(cudaFuncSetAttribute(gather_tma_kernel<STAGES, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas_per_sm = (int)(((size_t)tune().smem_kb * 1024) / (smem + 1024));
  const int cap = tune().tma_ctas * (32 / ROWS);  // the cap is stated for 32-row tiles: same bytes in flight per SM
  if (ctas_per_sm > cap) ctas_per_sm = cap;
  if (ctas_per_sm > 24) ctas_per_sm = 24;  // leave CTA slots (32 per SM) to the sampler
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  {  // carve-out: exactly what the resident CTAs need (the SM rounds up to its next configuration), or LG_GATHER_CARVEOUT
    int pct = tune().carveout;
    if (pct < 0) pct = (int)((100 * (size_t)ctas_per_sm * (smem + 1024) + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if (pct >= 0)
      LG_CUDA(cudaFuncSetAttribute(gather_tma_kernel<STAGES, ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  int64_t tiles = (max_rows + ROWS - 1) / ROWS;
  int64_t grid = (int64_t)kSMs * ctas_per_sm;
  if (tiles < grid) grid = tiles > 0 ? tiles : 1;
  gather_tma_kernel<STAGES, ROWS><<<(int)grid, 32, smem, st>>>(a);
  LG_LAUNCH_OK();
  return 0;
}
template <int STAGES>
int launch_tma_rows(cudaStream_t st, const GatherArgs& a, int64_t max_rows, int rows) {
  switch (rows) {
    case 8: return launch_tma<STAGES, 8>(st, a, max_rows);
    case 16: return launch_tma<STAGES, 16>(st, a, max_rows);
    default: return launch_tma<STAGES, 32>(st, a, max_rows);
  }
}

template <int R>
int launch_ldg(cudaStream_t st, const GatherArgs& a, int64_t max_rows, bool vec_ok) {
  int64_t warps = (max_rows + R - 1) / R;
  int64_t grid = (warps + 7) / 8;
  const int64_t cap = (int64_t)kSMs * tune().ldg_ctas;  // resident CTAs per SM x 148, grid-stride beyond
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (vec_ok)
    gather_ldg_kernel<R, true><<<(int)grid, 256, 0, st>>>(a);
  else
    gather_ldg_kernel<R, false><<<(int)grid, 256, 0, st>>>(a);
  LG_LAUNCH_OK();
  return 0;
}

int launch_gather(cudaStream_t st, GatherArgs a, int variant, int64_t max_rows) {
  const int dim = a.cache.dim;
  setup_gather4(&a);
  LG_REQUIRE(max_rows < (1ll << 31), "gather: %lld rows in one launch (tile indices are 32-bit)", (long long)max_rows);
  const bool vec_ok = (dim % 4 == 0) && (((uintptr_t)a.dst & 15) == 0) && (((uintptr_t)a.cache.backing & 15) == 0);
  const Tune& t = tune();
  if (variant == LG_GATHER_AUTO) variant = LG_GATHER_TMA;  // falls through to LDG when rows are not 16-byte multiples
  int rows = (t.tma_rows == 8 || t.tma_rows == 16) ? t.tma_rows : 32;
  while (rows > 8 && (size_t)dim * 4 * rows * 3 > 200 * 1024) rows >>= 1;  // very wide rows: smaller tiles
  if (variant == LG_GATHER_TMA && vec_ok && (size_t)dim * 4 * rows * 3 <= 200 * 1024) {
    int stages = t.tma_stages;
    while (stages > 3 && (size_t)dim * 4 * rows * stages > 200 * 1024) stages--;
    switch (stages) {
      case 4: return launch_tma_rows<4>(st, a, max_rows, rows);
      case 6: return launch_tma_rows<6>(st, a, max_rows, rows);
      default: return launch_tma_rows<3>(st, a, max_rows, rows);
    }
  }
  if (t.ldg_r == 4) return launch_ldg<4>(st, a, max_rows, vec_ok);
  return launch_ldg<8>(st, a, max_rows, vec_ok);
}

int check_cache(const lg_feature_cache* c) {
  LG_REQUIRE(c, "feature cache descriptor is null");
  LG_REQUIRE(c->dim > 0, "feature cache: dim %d", c->dim);
  if (c->flags & LG_CACHE_IDENTITY) {
    LG_REQUIRE(!c->directory && c->n_parts >= 1, "feature cache: LG_CACHE_IDENTITY takes no directory and needs the shard");
    return 0;
  }
  LG_REQUIRE(c->backing || c->directory, "feature cache: neither a backing matrix nor a directory");
  LG_REQUIRE(c->n_parts >= 0 && c->n_parts <= LG_MAX_DEVICE, "feature cache: n_parts %d", c->n_parts);
  LG_REQUIRE(!c->directory || c->shard_rows > 0, "feature cache: directory without shard_rows");
  return 0;
}

}  // namespace

extern "C" int lg_feature_cache_lookup(lg_sampler* s, lg_stream_t stream, const lg_feature_cache* cache,
                                       int32_t op_id, int32_t local_part, const lg_batch* b,
                                       unsigned long long* tier_rows) {
  return lg_feature_cache_lookup_range(s, stream, cache, op_id, op_id / LG_INTRABATCH_CON, local_part, b, tier_rows);
}

// rows of hops [first_hop, op_id/3] in one launch (fused gathers of lg_run_batch)
extern "C" int lg_feature_cache_lookup_range(lg_sampler* s, lg_stream_t stream, const lg_feature_cache* cache, int32_t op_id,
                                  int32_t first_hop, int32_t local_part, const lg_batch* b,
                                  unsigned long long* tier_rows) {
  LG_REQUIRE(s && b, "lg_feature_cache_lookup: null argument");
  int rc = check_cache(cache);
  if (rc) return rc;
  LG_REQUIRE(op_id % LG_INTRABATCH_CON == 1, "lg_feature_cache_lookup: op_id %d is not a lookup op", op_id);
  LG_REQUIRE(b->features, "lg_feature_cache_lookup: features buffer is null");
  GatherArgs a;
  a.cache = *cache;
  a.ids = b->ids;
  a.dst = b->features;
  a.nc = b->node_counter;
  a.off = 0;
  a.cnt = 0;
  a.dst_rows = b->feature_rows;
  a.hop = op_id / LG_INTRABATCH_CON;
  LG_REQUIRE(first_hop >= 0 && first_hop <= op_id / LG_INTRABATCH_CON, "lg_feature_cache_lookup_range: first_hop %d", first_hop);
  a.hop_lo = first_hop;
  a.op_slot = op_id % LG_INTRABATCH_CON;
  a.local_part = local_part;
  a.tier = (u64*)tier_rows;
  a.status = s->status;
  a.l2 = lg_l2_hints();
  a.ticket = s->gather_ticket;
  a.chunk = s->gather_ticket ? s->gather_chunk : tune().chunk;
  a.static_pct = s->gather_static_pct;
  LG_REQUIRE(a.hop <= s->n_hops, "lg_feature_cache_lookup: op_id %d beyond %d hops", op_id, s->n_hops);
  int64_t max_rows = 0;
  for (int h = first_hop; h <= a.hop; h++) max_rows += s->slots_per_hop[h];
  return launch_gather((cudaStream_t)stream, a, s->gather_variant, max_rows);
}

extern "C" int lg_gather_rows(lg_stream_t stream, const lg_feature_cache* cache, const int32_t* ids, int64_t n,
                              float* dst, int32_t local_part, int32_t variant, unsigned long long* tier_rows) {
  int rc = check_cache(cache);
  if (rc) return rc;
  LG_REQUIRE(ids && dst, "lg_gather_rows: null argument");
  if (n <= 0) return 0;
  // no handle to own a status word: one per (thread, device), zeroed when it is created
  static thread_local int32_t* dummy_by_device[LG_MAX_DEVICE * 2] = {};
  int dev = 0;
  LG_CUDA(cudaGetDevice(&dev));
  LG_REQUIRE(dev >= 0 && dev < LG_MAX_DEVICE * 2, "lg_gather_rows: device %d", dev);
  if (!dummy_by_device[dev]) {
    LG_CUDA(cudaMalloc(&dummy_by_device[dev], sizeof(int32_t)));
    LG_CUDA(cudaMemset(dummy_by_device[dev], 0, sizeof(int32_t)));
  }
  int32_t* dummy_status = dummy_by_device[dev];
  GatherArgs a;
  a.cache = *cache;
  a.ids = ids;
  a.dst = dst;
  a.nc = nullptr;
  a.off = 0;
  a.cnt = n;
  a.dst_rows = n;
  a.hop = 0;
  a.hop_lo = 0;
  a.op_slot = 0;
  a.local_part = local_part;
  a.tier = (u64*)tier_rows;
  a.status = dummy_status;
  a.l2 = lg_l2_hints();
  a.ticket = nullptr;  // no handle to own a counter: static tile order
  a.chunk = tune().chunk;
  a.static_pct = 0;
  return launch_gather((cudaStream_t)stream, a, variant, n);
}

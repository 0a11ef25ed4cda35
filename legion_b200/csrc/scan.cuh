// scan.cuh — block-level scan and the cross-tile prefix shared by the sampler kernels (sampler.cu) and the block
// builder (blocks.cu).  All kernels that use these run kBlock = 256 threads per CTA.
#pragma once
#include "common.cuh"

namespace lg {

constexpr int kBlock = 256;

// ------------------------------------------------------------------------------------------
// Exclusive prefix of a tile's aggregate over all earlier tiles, computed by the WHOLE block in one L2 round trip.
// Two levels: a tile posts (1<<32 | aggregate) into its own word AND adds the same packed value to the word of its group
// of kGroup consecutive tiles (count in the high half, sum in the low half: a group is complete when the count reaches
// kGroup).  A tile then sums the words of the complete groups before its own (one load per group) and the words of the
// earlier tiles of its own group (< kGroup loads) — all loads in flight together.  Group words sit on their own 128-byte
// lines.  The first version summed every predecessor's word directly: with one wave of ~800 tiles that is ~300 k polls on
// ~50 cache lines, and the lines of the lowest tiles were served for 8 us (p50) to 27 us (max) per kernel
// (profiles/r01d_sampler_chain.md).  Tiles are claimed through an atomic ticket, so every predecessor is already
// running: the spins cannot deadlock.
// state: [n_tiles] tile words; groups: [n_tiles / kGroup + 1] x kGroupStride words; both zeroed per batch.
// ------------------------------------------------------------------------------------------
constexpr int kGroup = 32;
constexpr int kGroupStride = 16;  // u64 words per group word: one 128-byte line each
__device__ __forceinline__ void red_add_u64(u64* p, u64 v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int32_t block_exclusive_prefix(u64* state, u64* groups, int tile, int32_t aggregate,
                                                          int32_t* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tile / kGroup, r = tile - g * kGroup;
  if (tid == 0) {
    const u64 word = (1ull << 32) | (uint32_t)aggregate;
    st_relaxed(state + tile, word);
    red_add_u64(groups + (size_t)g * kGroupStride, word);
  }
  int32_t sum = 0;
  // earlier tiles of the own group: threads of the last warp (so that the group loads below start on warp 0)
  if (warp == kBlock / 32 - 1 && lane < r) {
    const u64* p = state + g * kGroup + lane;
    u64 s = ld_relaxed(p);
    while ((s >> 32) == 0ull) {
      __nanosleep(40);
      s = ld_relaxed(p);
    }
    sum = (int32_t)(uint32_t)s;
  }
  // complete groups before the own one
  for (int j0 = 0; j0 < g; j0 += kBlock - 32) {
    const int j = j0 + tid;
    if (tid < kBlock - 32 && j < g) {
      const u64* p = groups + (size_t)j * kGroupStride;
      u64 s = ld_relaxed(p);
      while ((s >> 32) != (u64)kGroup) {
        __nanosleep(40);
        s = ld_relaxed(p);
      }
      sum += (int32_t)(uint32_t)s;
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  int32_t excl = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; w++) excl += s_red[w];
  __syncthreads();  // s_red may be reused by the caller
  return excl;
}

// exclusive scan of one int per thread over the block; returns the thread's exclusive prefix, *total = block sum
__device__ __forceinline__ int32_t block_exclusive_scan(int32_t v, int32_t* s_red, int32_t* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) s_red[warp] = inc;
  __syncthreads();
  int32_t before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; w++) {
    const int32_t x = s_red[w];
    if (w < warp) before += x;
    tot += x;
  }
  __syncthreads();
  *total = tot;
  return before + inc - v;
}


}  // namespace lg

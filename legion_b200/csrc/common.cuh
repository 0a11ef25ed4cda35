// common.cuh — shared device helpers of liblegion_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/legion_b200.h"
#include "../../include/legion_b200_debug.h"

int lg_set_error(const char* fmt, ...);
int lg_l2_hints();  // LG_L2_HINTS bitmask (see the L2 eviction-priority helpers below)
int lg_pdl();  // LG_PDL: -1 unset (per-handle default: on for the dense position map, off for the hashed one — measured),
               // else a bitmask: bit 0 launch the per-batch kernel chain with programmatic stream serialization;
               // bit 1 except the position-map release; bit 2 except batch_generate
int lg_chain_carveout();  // LG_CARVEOUT: preferred shared-memory carve-out (%) of the sampler-chain kernels, -1 = driver's choice
void lg_apply_carveout(const void* kernel);
void lg_count_launch();  // lg_debug_launch_count

// kernel launch of the per-batch chain: cudaLaunchKernelEx, with the PDL attribute when enabled
template <typename... KArgs, typename... Args>
static inline cudaError_t lg_launch_opt(bool allow_pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem,
                                        cudaStream_t st, Args... args) {
  if (lg_chain_carveout() >= 0) lg_apply_carveout((const void*)kernel);
  lg_count_launch();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (allow_pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define LG_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return lg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
  } while (0)
#define LG_LAUNCH_OK()        \
  do {                        \
    lg_count_launch();        \
    LG_CUDA(cudaGetLastError()); \
  } while (0)
#define LG_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) return lg_set_error(__VA_ARGS__); \
  } while (0)

typedef unsigned long long u64;

namespace lg {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- Philox4x32-10 (Random123).  ctr = (slot, hop, batch, stream), key = seed ----
__device__ __forceinline__ uint32_t philox_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

// ---- thrust::minstd_rand seed 1, discard(idx), one draw:  x = 48271^(idx+1) mod (2^31-1)
//      (reference stream, engine/operator_impl.cu:235-238).  Mersenne fold instead of %. ----
__device__ __forceinline__ uint32_t mod_m31(u64 x) {
  x = (x & 0x7FFFFFFFull) + (x >> 31);
  x = (x & 0x7FFFFFFFull) + (x >> 31);
  return (x >= 0x7FFFFFFFull) ? (uint32_t)(x - 0x7FFFFFFFull) : (uint32_t)x;
}
__device__ __forceinline__ uint32_t minstd_x(uint32_t idx) {
  u64 mult = 48271ull, acc = 1ull;
  u64 z = (u64)idx + 1ull;  // discard(idx) then engine() == 48271^(idx+1)
  while (z) {
    if (z & 1ull) acc = mod_m31(acc * mult);
    z >>= 1;
    mult = mod_m31(mult * mult);
  }
  return (uint32_t)acc;
}
__device__ __forceinline__ int32_t pick_minstd(uint32_t idx, int32_t deg) {
  double r = (double)(minstd_x(idx) - 1u);
  r = __ddiv_rn(r, 2147483646.0);  // uniform_real_distribution: /(1 + (max-min))
  return (int32_t)__dmul_rn(r, (double)deg);
}

template <int RNG>
__device__ __forceinline__ int32_t pick_neighbor(uint32_t slot, int32_t deg, uint32_t hop, uint32_t batch_id,
                                                 uint32_t stream_id, uint32_t k0, uint32_t k1) {
  if (RNG == LG_RNG_MINSTD) return pick_minstd(slot, deg);
  uint32_t r = philox_word0(slot, hop, batch_id, stream_id, k0, k1);
  return (int32_t)__umulhi(r, (uint32_t)deg);
}

// ---- programmatic dependent launch (PDL).  Every kernel of the per-batch chain starts with pdl_prologue():
//      griddepcontrol.wait blocks until every kernel before it in the stream has completed and its memory operations
//      are performed; then the NEXT kernel of the stream is allowed to become resident
//      (griddepcontrol.launch_dependents) — it blocks in its own wait until this grid is done, so stream order holds
//      for memory and only the launch latency between dependent kernels disappears.
//      The fence.acq_rel.gpu after the wait is the ACQUIRE side of that hand-over in the PTX memory model: loads that
//      follow an acquire fence may not return values older than what the synchronisation made visible, which on this
//      hardware means the SM's L1 is invalidated (SASS: CCTL.IVALL) — a PDL grid is dispatched, and its launch-time
//      invalidation happens, while its parent still runs and keeps filling L1 on the same SM.  Without the fence a
//      plain load hit such a sector (seen on the GPU: hotness_measure_kernel read the previous hop's nc[7]).
//      Belt and braces on top of the fence: the words a kernel's control flow depends on — the batch's counters — are
//      read with ld.relaxed.gpu (ld_counter below: always served by L2, never by a possibly stale L1 line); the
//      position-map lookups keep their L1-cached loads (hub words are re-read thousands of times per kernel) and rely
//      on the acquire fence alone.  Without the launch attribute (LG_PDL=0, or a kernel launched with <<<>>>) wait and
//      launch_dependents are no-ops and the kernel boundary itself invalidates L1. ----
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __threadfence();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// counters written by an earlier kernel of the chain (node_counter / edge_counter words): L2-coherent load
__device__ __forceinline__ int32_t ld_counter(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- relaxed gpu-scope 64-bit accesses (cross-CTA flags and dedup-table words) ----
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- L2 eviction priorities (createpolicy + .L2::cache_hint).  The small random-access arrays of the sampler
//      (position map, directories, indptr) are worth keeping in the 126 MB L2 while the gather streams ~0.8 GB per
//      batch through it; the gather's rows and output have no reuse.  kind: 0 normal, 1 evict_last, 2 evict_first.
//      LG_L2_HINTS (environment, read once) selects which accesses carry a hint:
//      bit 0 gather row loads evict_first, bit 1 gather output stores evict_first, bit 2 sampler arrays + feature
//      directory evict_last, bit 3 neighbour (indices) reads evict_first. ----
__device__ __forceinline__ u64 l2_policy(int kind) {
  u64 p;
  if (kind == 1)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ld_ca_u32_hint(const uint32_t* p, u64 pol) {
  uint32_t v;
  asm volatile("ld.global.ca.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t ld_nc_s32_hint(const int32_t* p, u64 pol) {
  int32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ long long ld_nc_s64_hint(const int64_t* p, u64 pol) {
  long long v;
  asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_u32_hint(uint32_t* p, uint32_t v, u64 pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_min_u32_hint(uint32_t* p, uint32_t v, u64 pol) {
  asm volatile("red.relaxed.gpu.global.min.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ uint32_t hash32(uint32_t k) {  // murmur3 fmix32
  k ^= k >> 16; k *= 0x85EBCA6Bu; k ^= k >> 13; k *= 0xC2B2AE35u; k ^= k >> 16;
  return k;
}

__device__ __forceinline__ int32_t warp_sum(int32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int32_t warp_incl_scan(int32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace lg

/*
 * legion_b200_synth.h — deterministic synthetic datasets in Legion's on-disk/in-memory layout
 * (dataset/README.md:3-10: edge_src int64[N+1] CSR offsets, edge_dst int32[E], features
 * fp32[N x D], labels int32[N]).  Bench/test tooling of liblegion_b200.so, generated directly
 * in device memory so the paper-scale shapes (legion_server.py:41-88) never cross PCIe.
 * Every value is a pure function of (seed, vertex, position): legion_b200/synth.py holds the
 * numpy restatement used on CPU-only boxes and checked bit-for-bit against these kernels.
 *
 *   deg(v)      = min(dmax, floor(dmin / sqrt(1 - u(seed, v))))            power-law tail
 *   nbr(v, k)   = perm(floor(N * u^3))  with  u = u(seed', v*2^21 + k)      skewed popularity
 *   perm(r)     = (r * 2654435761 + 12345) mod N                            scatters hot ids
 *   feat(v, c)  = bits(h(seed'', v*D + c)) & 0xBFFFFFFF  as fp32            always finite
 */
#ifndef LEGION_B200_SYNTH_H_
#define LEGION_B200_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* indptr[0..N]: degrees then exclusive offsets (device). indptr[N] = E after the call. */
int lg_synth_indptr(void* stream, int64_t num_nodes, double dmin, int32_t dmax, uint64_t seed,
                    int64_t* indptr);
/* indices[E] for the offsets produced above */
int lg_synth_indices(void* stream, int64_t num_nodes, const int64_t* indptr, uint64_t seed,
                     int32_t* indices);
/* rows [row0, row0+rows) of the feature matrix into out[rows x dim] */
int lg_synth_features(void* stream, int64_t row0, int64_t rows, int32_t dim, uint64_t seed,
                      float* out);
int lg_synth_labels(void* stream, int64_t num_nodes, int32_t classes, int32_t* labels);
/* out[r,:] = feat(ids[r], :) for r < n; rows with ids[r] < 0 are zero-filled.  Used to check gathered rows at
 * scales where the feature matrix is never materialised in vertex order. */
int lg_synth_feature_rows(void* stream, const int32_t* ids, int64_t n, int32_t dim, uint64_t seed,
                          float* out);
/* a cache shard generated in place: shard[r,:] = feat(order[r*kg + j], :) (FeatFillUp, cache/cache_impl.cuh:183-188,
 * without a backing matrix); ranks >= num_nodes are zero-filled */
int lg_synth_feature_shard(void* stream, const int32_t* order, int64_t cap, int32_t kg, int32_t j,
                           int32_t dim, int64_t num_nodes, uint64_t seed, float* shard);
/* the same for the hybrid placement (lg_fill_feature_shard_hybrid): rows < rep hold ranks 0..rep-1 on every part */
int lg_synth_feature_shard_hybrid(void* stream, const int32_t* order, int64_t cap, int32_t kg, int64_t rep, int32_t j,
                                  int32_t dim, int64_t num_nodes, uint64_t seed, float* shard);
#ifdef __cplusplus
}
#endif
#endif

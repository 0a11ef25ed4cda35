/*
 * synth_oracle.c — host-side (OpenMP) twin of the synthetic dataset functions.
 *
 * TEST INFRASTRUCTURE ONLY (part of liblegion_oracle.so, see legion_oracle.c).  The CPU arm of
 * bench.py (`--impl reference`, `cpu_baseline`) builds its dataset with these functions so that no
 * product code (liblegion_b200.so) runs in that arm; the tests check them bit-for-bit against
 * legion_b200/synth.py (numpy) and, on the GPU box, against legion_b200/csrc/synth.cu.
 *
 * The dataset layout is Legion's: edge_src int64[N+1] CSR offsets, edge_dst int32[E], features
 * fp32[N x D], labels int32[N] (reference dataset/README.md:3-10; loaded by
 * sampling_server/src/storage/storage_management.cu:100-164).  The functions themselves have no
 * reference counterpart (the reference ships no generator): see include/legion_b200_synth.h.
 *
 * Built with -ffp-contract=off: every double operation rounds once, like numpy and the __d*_rn
 * intrinsics of the device generator.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
static inline uint64_t hash2(uint64_t seed, uint64_t a) { return mix64(seed ^ mix64(a + 0x9E3779B97F4A7C15ull)); }
static inline double unit(uint64_t h) { return (double)(h >> 11) * 1.1102230246251565e-16; /* 2^-53 */ }

/* deg(v) = min(dmax, floor(dmin / sqrt(1 - u(seed, v)))); indptr[0..n] = exclusive offsets, returns E */
int64_t lgo_synth_indptr(int64_t n, double dmin, int32_t dmax, uint64_t seed, int64_t* indptr) {
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  int64_t* part = (int64_t*)calloc((size_t)nt + 1, sizeof(int64_t));
#pragma omp parallel num_threads(nt)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    int64_t sum = 0;
    for (int64_t v = lo; v < hi; v++) {
      const double x = 1.0 - unit(hash2(seed, (uint64_t)v));
      long long d = (long long)(dmin / sqrt(x));
      if (d > dmax) d = dmax;
      indptr[v + 1] = d; /* degree, turned into an offset below */
      sum += d;
    }
    part[t + 1] = sum;
#pragma omp barrier
#pragma omp single
    {
      for (int i = 0; i < nt; i++) part[i + 1] += part[i];
      indptr[0] = 0;
    }
    int64_t run = part[t];
    for (int64_t v = lo; v < hi; v++) {
      run += indptr[v + 1];
      indptr[v + 1] = run;
    }
  }
  free(part);
  return indptr[n];
}

/* nbr(v, k) = perm(floor(N * u^3)), u = u(seed ^ S2, v * 2^21 + k), perm(r) = (r * 2654435761 + 12345) mod N */
void lgo_synth_indices(int64_t n, const int64_t* indptr, uint64_t seed, int32_t* indices) {
  const uint64_t s2 = seed ^ 0xA5A5A5A55A5A5A5Aull;
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t v = 0; v < n; v++) {
    const int64_t b = indptr[v], e = indptr[v + 1];
    for (int64_t k = 0; k < e - b; k++) {
      const double u = unit(hash2(s2, ((uint64_t)v << 21) + (uint64_t)k));
      const double t = (u * u) * u;
      long long r = (long long)(t * (double)n);
      if (r >= n) r = n - 1;
      indices[b + k] = (int32_t)(((uint64_t)r * 2654435761ull + 12345ull) % (uint64_t)n);
    }
  }
}

/* feat(v, c) = bits(h(seed ^ S3, v * D + c)) & 0xBFFFFFFF as fp32: rows [row0, row0 + rows) */
void lgo_synth_features(int64_t row0, int64_t rows, int32_t dim, uint64_t seed, float* out) {
  const uint64_t s3 = seed ^ 0xFEA7FEA7FEA7FEA7ull;
  const int64_t total = rows * dim;
  uint32_t* o = (uint32_t*)out;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; i++) o[i] = (uint32_t)hash2(s3, (uint64_t)(row0 * dim + i)) & 0xBFFFFFFFu;
}

/* out[r, :] = feat(ids[r], :); rows with ids[r] < 0 are zero-filled (lg_synth_feature_rows) */
void lgo_synth_feature_rows(const int32_t* ids, int64_t n, int32_t dim, uint64_t seed, float* out) {
  const uint64_t s3 = seed ^ 0xFEA7FEA7FEA7FEA7ull;
  uint32_t* o = (uint32_t*)out;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; r++) {
    const int64_t v = ids[r];
    for (int32_t c = 0; c < dim; c++)
      o[r * dim + c] = v < 0 ? 0u : ((uint32_t)hash2(s3, (uint64_t)(v * dim + c)) & 0xBFFFFFFFu);
  }
}

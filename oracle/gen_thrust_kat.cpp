// gen_thrust_kat.cpp — known-answer vectors for the reference's neighbour pick, produced by
// the REAL Thrust headers of the CUDA toolkit (the third-party code engine/operator_impl.cu
// :9-10,235-238 calls; not vendored under /root/reference).  Test infrastructure only.
//   g++ -O2 -std=c++17 -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP -I/usr/local/cuda/include \
//       oracle/gen_thrust_kat.cpp -o oracle/_ref/gen_thrust_kat && oracle/_ref/gen_thrust_kat > tests/golden/minstd_pick.json
#include <thrust/random/linear_congruential_engine.h>
#include <thrust/random/uniform_int_distribution.h>
#include <cstdint>
#include <cstdio>

static int pick(uint64_t idx, int deg) {
  // the three statements of engine/operator_impl.cu:235-238
  thrust::minstd_rand engine;
  engine.discard(idx);
  thrust::uniform_int_distribution<> dist(0, deg - 1);
  return dist(engine);
}

int main() {
  // deterministic sweep: small/edge slots, fan-out boundaries, the survey's 7 vectors, a
  // pseudo-random spread of (idx, deg) up to the Clueweb slot count
  std::printf("{\"source\": \"thrust (CUDA toolkit headers) minstd_rand + uniform_int_distribution<int>\",\n \"vectors\": [\n");
  bool first = true;
  auto emit = [&](uint64_t idx, int deg) {
    std::printf("%s  [%llu, %d, %d]", first ? "" : ",\n", (unsigned long long)idx, deg, pick(idx, deg));
    first = false;
  };
  const uint64_t idxs[] = {0, 1, 2, 3, 24, 25, 26, 199999, 200000, 1999999, 2207999, 5999999, 7327999,
                           2147483645ull, 2147483646ull, 2147483647ull, 4294967295ull};
  const int degs[] = {1, 2, 3, 7, 10, 25, 33, 1000, 65536, 1000000, 2147483647};
  for (uint64_t i : idxs)
    for (int d : degs) emit(i, d);
  uint64_t s = 0x1e910;
  for (int k = 0; k < 400; k++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t idx = (s >> 33) % 7328000ull;
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    int deg = 1 + (int)((s >> 33) % 50000ull);
    emit(idx, deg);
  }
  std::printf("\n ]}\n");
  return 0;
}

/*
 * legion_b200_debug.h — diagnostics of liblegion_b200.so.  NOT part of the drop-in ABI (legion_b200.h): nothing a
 * Legion host needs is declared here; bench.py and scripts/ use these to count launches and to trace the kernels.
 */
#ifndef LEGION_B200_DEBUG_H_
#define LEGION_B200_DEBUG_H_
#include "legion_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* number of kernels this library has launched from this process so far (every launch site counts itself);
 * reset != 0 zeroes the counter after reading it.  bench.py reports it as gpu_launches. */
long long lg_debug_launch_count(int32_t reset);
/* per-tile phase timestamps of the sampler kernels (globaltimer ns, clock64) written to a device buffer of
 * lg_debug_trace_words() u64 words, layout [(hop-1)*2 + kernel][tile < 2048][phase < 8][2]; NULL = off */
int lg_debug_set_trace(lg_sampler* s, unsigned long long* device_buf);
int64_t lg_debug_trace_words(void);
/* `ctas` x `threads` spinning for `cycles` SM clocks without touching memory */
int lg_debug_spin(lg_stream_t stream, int32_t ctas, int32_t threads, int64_t cycles);
#ifdef __cplusplus
}
#endif
#endif

/*
 * legion_b200_ext.h — optional side channel between the B200 sampling server and the B200 trainer extension.
 *
 * The reference wire (`simpleIPCshm`: int32 steps[3] + 8 x 2 x 7 CUDA-IPC handles = 7,180 bytes,
 * sampling_server/src/engine/ipc_service.cu:28-31 == training_backend/ipc_cuda_kernel.cu:30-33) has no spare byte and is
 * left untouched.  This SECOND POSIX shm segment carries what the trainer's get_next otherwise fetches with two blocking
 * cudaMemcpy per batch (training_backend/ipc_cuda_kernel.cu:194-195): the 16 + 16 counter words of the batch, copied to
 * host memory by the server behind the batch's last kernel and complete before sem_post(sem_w).
 *
 * Compatibility: a reference trainer never opens the segment; a B200 trainer extension talking to a reference server
 * finds no segment (or one whose sequence numbers do not advance) and falls back to the two copies.
 */
#ifndef LEGION_B200_EXT_H_
#define LEGION_B200_EXT_H_
#include <stdint.h>

#define LG_EXT_SHM_NAME "legionB200ext"
#define LG_EXT_MAGIC 0x3032424Cu /* "LB20" */
#define LG_EXT_VERSION 1u
#define LG_EXT_MAX_DEVICE 8
#define LG_EXT_SLOTS 2
#define LG_EXT_MAX_HOPS 6

typedef struct lg_ext_shm {
  uint32_t magic, version;
  int32_t n_gpus, reserved;
  /* node_counter[16] | edge_counter[16] of the batch in (gpu, slot) */
  volatile int32_t counters[LG_EXT_MAX_DEVICE][LG_EXT_SLOTS][32];
  /* number of batches handed off through (gpu, slot) so far; written after the counters, before sem_post */
  volatile uint32_t seq[LG_EXT_MAX_DEVICE][LG_EXT_SLOTS];
  /* Blocks as CSC, built by the server next to the gather (LEGION_EMIT_CSC=1; csc_hops = 0 otherwise): per (gpu, slot)
   * and block h = 1..csc_hops three more CUDA-IPC buffers — indptr int32[nc[9+h-1] + 1] over the block's destinations,
   * indices int32[ec[9+h]] = batch-local sources, eids int32[ec[9+h]] = positions in the COO — the same cumulative blocks
   * the trainer cuts out of agg_src/agg_dst (training_backend/ipc_cuda_kernel.cu:218-231), grouped by destination. */
  int32_t csc_hops, csc_reserved;
  unsigned char csc_handle[LG_EXT_MAX_DEVICE][LG_EXT_SLOTS][LG_EXT_MAX_HOPS][3][64];
} lg_ext_shm;

#endif

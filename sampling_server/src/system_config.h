// system_config.h — constants of the wire contract shared with the trainer
// (reference: sampling_server/src/include/system_config.cuh:47-57).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/legion_b200.h"

#define INTERBATCH_CON LG_INTERBATCH_CON
#define INTRABATCH_CON LG_INTRABATCH_CON
#define MAX_DEVICE LG_MAX_DEVICE
#define MEMORY_USAGE LG_MEMORY_USAGE
#define TRAINMODE LG_TRAINMODE
#define VALIDMODE LG_VALIDMODE
#define TESTMODE LG_TESTMODE
#define CACHEMISS_FLAG LG_CACHEMISS_FLAG

// every failure is fatal, like the reference's cudaCheckError (engine/operator_impl.cu:16-24)
#define LGCHECK(call)                                                                          \
  do {                                                                                         \
    if ((call) != 0) {                                                                         \
      std::fprintf(stderr, "legion failure %s:%d: '%s'\n", __FILE__, __LINE__, lg_last_error()); \
      std::exit(EXIT_FAILURE);                                                                 \
    }                                                                                          \
  } while (0)

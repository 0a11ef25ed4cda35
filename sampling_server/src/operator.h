// operator.h — operator plug-in API (reference: sampling_server/src/engine/operator.h:4-28).
// Same struct, same abstract class, same five factories; the bodies call the C ABI of
// liblegion_b200.so instead of the reference's extern "C" kernels drivers.
#pragma once
#include "../../include/legion_b200.h"

struct OpParams {
  int device_id;
  lg_stream_t stream;
  lg_event_t event;
  void* memorypool;
  void* cache;
  void* graph;
  void* feature;
  void* env;
  int neighbor_count;
  bool is_presc;
  bool in_memory;
  int hop_num;
};

class Operator {
 public:
  virtual ~Operator() {}
  virtual void run(OpParams* params) = 0;
};

Operator* NewBatchGenerateOP(int op_id);
Operator* NewRandomSampleOP(int op_id);
Operator* NewCacheLookupOP(int op_id);
Operator* NewSSDIOSubmitOP(int op_id);
Operator* NewSSDIOCompleteOP(int op_id);

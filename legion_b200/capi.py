"""ctypes binding of include/legion_b200.h + include/legion_b200_synth.h.

The product path: there is no CPU fallback.  If liblegion_b200.so is missing this module raises
on first use instead of silently doing the work some other way.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblegion_b200.so")

MAX_DEVICE = 8
COUNTER_SLOTS = 16
CACHEMISS_FLAG = -2
RNG_MINSTD, RNG_PHILOX = 0, 1
GATHER_AUTO, GATHER_LDG, GATHER_TMA = 0, 1, 2
TRAINMODE, VALIDMODE, TESTMODE = 0, 1, 2
TAIL_EXACT, TAIL_REFERENCE = 0, 1
CACHE_IDENTITY = 1

vp = C.c_void_p


class Topology(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("shard_rows", C.c_int32), ("num_nodes", C.c_int64),
                ("indptr", vp * (MAX_DEVICE + 1)), ("indices", vp * (MAX_DEVICE + 1)), ("directory", vp)]


class FeatureCache(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("shard_rows", C.c_int32), ("dim", C.c_int32), ("flags", C.c_int32),
                ("num_nodes", C.c_int64), ("shard", vp * MAX_DEVICE), ("backing", vp), ("directory", vp)]


class Batch(C.Structure):
    _fields_ = [("ids", vp), ("features", vp), ("labels", vp), ("agg_src", vp), ("agg_dst", vp),
                ("node_counter", vp), ("edge_counter", vp), ("feature_rows", C.c_int64), ("num_ids", C.c_int32),
                ("reserved", C.c_int32)]


class BatchParams(C.Structure):
    _fields_ = [("all_ids", vp), ("all_labels", vp), ("total_cap", C.c_int32), ("batch_size", C.c_int32),
                ("counter", C.c_int32), ("mode", C.c_int32), ("rng_kind", C.c_int32), ("batch_id", C.c_uint32),
                ("stream_id", C.c_uint32), ("local_part", C.c_int32), ("rng_seed", C.c_uint64)]


_PROTOS = {
    # name: (restype, argtypes)
    "lg_last_error": (C.c_char_p, []),
    "lg_version": (C.c_int, []),
    "lg_num_ids": (C.c_int64, [C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "lg_sampler_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.POINTER(vp)]),
    "lg_sampler_destroy": (C.c_int, [vp]),
    "lg_sampler_reset": (C.c_int, [vp, vp]),
    "lg_sampler_scratch_bytes": (C.c_int64, [vp]),
    "lg_sampler_dedup_layout": (C.c_int32, [vp]),
    "lg_sampler_set_gather_variant": (C.c_int, [vp, C.c_int32]),
    "lg_sampler_set_overlap": (C.c_int, [vp, C.c_int32]),
    "lg_sampler_set_tail_mode": (C.c_int, [vp, C.c_int32]),
    "lg_sampler_set_lazy_relabel": (C.c_int, [vp, C.c_int32]),
    "lg_sampler_set_gather_fusion": (C.c_int, [vp, C.c_int32]),
    # include/legion_b200_debug.h
    "lg_debug_launch_count": (C.c_longlong, [C.c_int32]),
    "lg_debug_spin": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int64]),
    "lg_debug_set_trace": (C.c_int, [vp, vp]),
    "lg_debug_trace_words": (C.c_int64, []),
    "lg_batch_wait": (C.c_int, [vp, vp, C.POINTER(Batch)]),
    "lg_sampler_status": (C.c_int, [vp, vp, C.POINTER(C.c_int32)]),
    "lg_sampler_status_async": (C.c_int, [vp, vp, vp]),
    "lg_batch_generate": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Batch)]),
    "lg_random_sample": (C.c_int, [vp, vp, C.POINTER(Topology), C.c_int32, C.c_int32, C.c_uint64, C.c_uint32,
                                   C.c_uint32, C.POINTER(Batch), vp]),
    "lg_feature_cache_lookup": (C.c_int, [vp, vp, C.POINTER(FeatureCache), C.c_int32, C.c_int32, C.POINTER(Batch), vp]),
    "lg_feature_cache_lookup_range": (C.c_int, [vp, vp, C.POINTER(FeatureCache), C.c_int32, C.c_int32, C.c_int32,
                                              C.POINTER(Batch), vp]),
    "lg_batch_publish": (C.c_int, [vp, C.POINTER(Batch), C.POINTER(Batch)]),
    "lg_io_submit": (C.c_int, [vp, vp, C.c_int32, C.POINTER(Batch)]),
    "lg_io_complete": (C.c_int, [vp, vp, C.c_int32, C.POINTER(Batch), vp, vp]),
    "lg_run_batch": (C.c_int, [vp, vp, C.POINTER(Topology), C.POINTER(FeatureCache), C.POINTER(BatchParams),
                               C.POINTER(Batch), vp]),
    "lg_run_batch_host": (C.c_int, [vp, vp, C.POINTER(Topology), C.POINTER(FeatureCache), C.POINTER(BatchParams),
                                    vp, vp, C.POINTER(Batch), vp, vp]),
    "lg_run_batch_host_async": (C.c_int, [vp, vp, C.POINTER(Topology), C.POINTER(FeatureCache), C.POINTER(BatchParams),
                                          vp, vp, C.POINTER(Batch), vp, vp]),
    "lg_gather_rows": (C.c_int, [vp, C.POINTER(FeatureCache), vp, C.c_int64, vp, C.c_int32, C.c_int32, vp]),
    "lg_hotness_accumulate": (C.c_int, [vp, vp, vp, C.c_int64]),
    "lg_hotness_rank": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, C.POINTER(C.c_int64)]),
    "lg_place_features": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int64, vp]),
    "lg_place_topology": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp]),
    "lg_fill_feature_shard": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp]),
    "lg_place_features_hybrid": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp]),
    "lg_fill_feature_shard_hybrid": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp]),
    "lg_topo_shard_indptr": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp]),
    "lg_topo_shard_fill": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp, vp, vp]),
    "lg_fill_i32": (C.c_int, [vp, vp, C.c_int32, C.c_int64]),
    "lg_cost_model": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "lg_cost_model_saturating": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "lg_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "lg_set_device": (C.c_int, [C.c_int32]),
    "lg_enable_peer_access": (C.c_int, [C.c_int32]),
    "lg_device_alloc": (C.c_int, [C.POINTER(vp), C.c_int64]),
    "lg_device_free": (C.c_int, [vp]),
    "lg_vmm_round_up": (C.c_int64, [C.c_int64]),
    "lg_vmm_alloc": (C.c_int, [C.c_int64, C.POINTER(vp), C.POINTER(C.c_int32)]),
    "lg_vmm_import": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(vp)]),
    "lg_vmm_free": (C.c_int, [vp]),
    "lg_host_alloc_mapped": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.c_int64]),
    "lg_host_free": (C.c_int, [vp]),
    "lg_host_register": (C.c_int, [vp, C.c_int64, C.POINTER(vp)]),
    "lg_host_unregister": (C.c_int, [vp]),
    "lg_ipc_export": (C.c_int, [vp, C.c_char * 64]),
    "lg_ipc_open": (C.c_int, [C.c_char * 64, C.POINTER(vp)]),
    "lg_ipc_close": (C.c_int, [vp]),
    "lg_stream_create": (C.c_int, [C.POINTER(vp)]),
    "lg_stream_destroy": (C.c_int, [vp]),
    "lg_stream_synchronize": (C.c_int, [vp]),
    "lg_memcpy_h2d": (C.c_int, [vp, vp, C.c_int64, vp]),
    "lg_memcpy_d2h": (C.c_int, [vp, vp, C.c_int64, vp]),
    "lg_memcpy_d2d": (C.c_int, [vp, vp, C.c_int64, vp]),
    "lg_memset_async": (C.c_int, [vp, C.c_int32, C.c_int64, vp]),
    "lg_event_create": (C.c_int, [C.POINTER(vp)]),
    "lg_event_destroy": (C.c_int, [vp]),
    "lg_event_record": (C.c_int, [vp, vp]),
    "lg_event_query": (C.c_int, [vp, C.POINTER(C.c_int32)]),
    "lg_event_synchronize": (C.c_int, [vp]),
    "lg_stream_wait_event": (C.c_int, [vp, vp]),
    "lg_device_mem_info": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lg_block_csc_workspace": (C.c_int, [C.c_int64, C.POINTER(C.c_int64)]),
    "lg_block_csc": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp, C.c_int64]),
    "lg_block_csc_batch_workspace": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lg_block_csc_batch": (C.c_int, [vp, C.POINTER(Batch), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(vp),
                                     C.POINTER(vp), C.POINTER(vp), vp, C.c_int64]),
    # include/legion_b200_synth.h
    "lg_synth_indptr": (C.c_int, [vp, C.c_int64, C.c_double, C.c_int32, C.c_uint64, vp]),
    "lg_synth_indices": (C.c_int, [vp, C.c_int64, vp, C.c_uint64, vp]),
    "lg_synth_features": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int32, C.c_uint64, vp]),
    "lg_synth_labels": (C.c_int, [vp, C.c_int64, C.c_int32, vp]),
    "lg_synth_feature_rows": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_uint64, vp]),
    "lg_synth_feature_shard": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, vp]),
    "lg_synth_feature_shard_hybrid": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, vp]),
}

_lib = None


class LegionError(RuntimeError):
    pass


def load():
    """dlopen liblegion_b200.so; fail loudly if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LegionError(
            f"{LIB_PATH} is missing: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            f"g.build()'`). legion_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LegionError(load().lg_last_error().decode())


def declared_symbols():
    return sorted(_PROTOS)

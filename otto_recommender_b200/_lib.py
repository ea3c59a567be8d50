"""ctypes binding of libottocov.so (C ABI: include/ottocov.h).  Fails loudly when the library is
missing -- there is no pure-Python or CPU path behind it."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, os.environ.get("OTTOCOV_SO_NAME", "libottocov.so"))     # tuning builds only
_lib = None

K_FAMILIES = 11
HOST, DEVICE = 0, 1
ORDER_KEY, ORDER_COUNT_DESC = 0, 1


class OttocovError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ottocov error {code}: {msg}")
        self.code = code


class Spec(Structure):
    _fields_ = [("type_this", c_int32), ("next_mask", c_uint32), ("window", c_int64), ("pair_budget", c_int64),
                ("min_count", c_uint32), ("flags", c_uint32), ("dt_min", c_int64), ("dt_max", c_int64)]


class EventsInfo(Structure):
    _fields_ = [("n_rows_in", c_int64), ("n_events", c_int64), ("n_by_type", c_int64 * 3),
                ("session_min", c_int32), ("session_max", c_int32), ("ts_min", c_int32), ("ts_max", c_int32),
                ("aid_max", c_int32), ("aid_bits", c_int32), ("was_sorted", c_int32)]


class CountInfo(Structure):
    _fields_ = [("n_pairs", c_int64), ("n_unique", c_int64), ("n_chunks", c_int32), ("sort_passes", c_int32),
                ("fused", c_int32), ("reserved", c_int32), ("h2d_bytes", c_int64)]


class XPlan(Structure):
    """ottocov_xplan: stripe capacities + byte layout of a rank's receive area (fused expansion + exchange)."""
    _fields_ = [("n_ranks", c_int32), ("aid_bits", c_int32), ("bucket_bits", c_int32), ("sub_bits", c_int32),
                ("rest_passes", c_int32), ("reserved", c_int32), ("stripe_cap", c_int64), ("mirror_cap", c_int64),
                ("off_counts", c_int64), ("off_status", c_int64), ("off_hist", c_int64), ("off_keys", c_int64),
                ("off_mstatus", c_int64), ("off_mkeys", c_int64), ("off_mcnt", c_int64), ("total_bytes", c_int64)]


class KernelStat(Structure):
    _fields_ = [("launches", c_int64), ("ms", c_double), ("algo_bytes", c_double)]


# every symbol include/ottocov.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ottocov_version": (c_int, []),
    "ottocov_create": (c_int, [c_int, POINTER(c_void_p)]),
    "ottocov_destroy": (c_int, [c_void_p]),
    "ottocov_last_error": (c_char_p, [c_void_p]),
    "ottocov_set_stream": (c_int, [c_void_p, c_void_p]),
    "ottocov_synchronize": (c_int, [c_void_p]),
    "ottocov_trim": (c_int, [c_void_p]),
    "ottocov_memory_info": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "ottocov_set_profiling": (c_int, [c_void_p, c_int]),
    "ottocov_kernel_stats": (c_int, [c_void_p, POINTER(KernelStat), c_int]),
    "ottocov_kernel_family_name": (c_char_p, [c_int]),
    "ottocov_load_events": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int]),
    "ottocov_count_parts": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                    POINTER(c_int64), POINTER(Spec), c_int, POINTER(c_void_p)]),
    "ottocov_get_events_info": (c_int, [c_void_p, POINTER(EventsInfo)]),
    "ottocov_count": (c_int, [c_void_p, POINTER(Spec), POINTER(c_void_p)]),
    "ottocov_expand_prepare": (c_int, [c_void_p, POINTER(Spec), POINTER(c_int64), POINTER(c_int)]),
    "ottocov_expand_run": (c_int, [c_void_p, c_int, c_void_p, c_void_p, POINTER(c_int), POINTER(c_int64)]),
    "ottocov_push_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int, POINTER(c_uint64)]),
    "ottocov_xplan_make": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int64, POINTER(XPlan)]),
    "ottocov_expand_scatter": (c_int, [c_void_p, POINTER(XPlan), c_int, POINTER(c_uint64)]),
    "ottocov_reduce_received": (c_int, [c_void_p, POINTER(XPlan), c_uint64, c_uint32, c_int, POINTER(c_void_p), POINTER(c_int64)]),
    "ottocov_mirror_push": (c_int, [c_void_p, POINTER(XPlan), c_int, c_void_p, POINTER(c_uint64)]),
    "ottocov_mirror_collect": (c_int, [c_void_p, POINTER(XPlan), c_int, c_void_p, c_uint64, POINTER(c_void_p), POINTER(c_int64)]),
    "ottocov_reduce_pairs": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_uint32, c_int, c_int, POINTER(c_void_p)]),
    "ottocov_table_mirror": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_void_p)]),
    "ottocov_get_count_info": (c_int, [c_void_p, POINTER(CountInfo)]),
    "ottocov_table_from_arrays": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, POINTER(c_void_p)]),
    "ottocov_table_from_packed": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, POINTER(c_void_p)]),
    "ottocov_table_free": (c_int, [c_void_p, c_void_p]),
    "ottocov_table_rows": (c_int, [c_void_p, POINTER(c_int64)]),
    "ottocov_table_total": (c_int, [c_void_p, c_void_p, POINTER(c_int64)]),
    "ottocov_table_merge": (c_int, [c_void_p, POINTER(c_void_p), c_int, POINTER(c_void_p)]),
    "ottocov_table_filter": (c_int, [c_void_p, c_void_p, c_uint32, POINTER(c_void_p)]),
    "ottocov_table_fetch": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_int, POINTER(c_int64)]),
    "ottocov_table_device_ptrs": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p)]),
    "ottocov_table_topk": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int64)]),
    "ottocov_topk_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int]),
    "ottocov_topk_lookup": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ottocov_table_partition": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, POINTER(c_int64)]),
    "ottocov_hash_dest": (c_uint32, [c_uint32, c_uint32]),
    "ottocov_sort_u64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int]),
    "ottocov_key_mix": (c_uint64, [c_int, c_uint32, c_uint32]),
    "ottocov_key_unmix": (c_uint64, [c_int, c_uint64]),
    "ottocov_count_features": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, POINTER(c_int64)]),
    "ottocov_count_features_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int64, c_int]),
    "ottocov_count_weighted": (c_int, [c_void_p, POINTER(Spec), POINTER(c_void_p)]),
    "ottocov_wtable_rows": (c_int, [c_void_p, POINTER(c_int64)]),
    "ottocov_wtable_free": (c_int, [c_void_p, c_void_p]),
    "ottocov_wtable_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, POINTER(c_int64)]),
    "ottocov_wtable_topk": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                    POINTER(c_int64)]),
    "ottocov_count_popularity": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int32, c_int,
                                         POINTER(c_int64)]),
    "ottocov_popularity_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int]),
}


def library_path() -> str:
    return _SO


def load_library():
    """Load libottocov.so; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise OttocovError(-1, f"{_SO} not found: build it with `python -m otto_recommender_b200.build` "
                               "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ctypes.CDLL(_SO)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

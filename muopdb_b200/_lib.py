"""ctypes binding of libmuopdb_gpu.so (include/muopdb_gpu.h).  No fallback: if the CUDA library is missing or no
device is present every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmuopdb_gpu.so")

OK = 0
ERR_INVALID_ARG, ERR_OUT_OF_RANGE, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
L2, DOT = 0, 1
QUANT_NONE, QUANT_PQ = 0, 1
HOST, DEVICE = 0, 1
K_COARSE, K_SELECT, K_QUANTIZE, K_SCAN, K_FINALIZE, K_HNSW, K_MERGE, K_OTHER, K_FALLBACK = range(9)
KERNEL_CLASS_NAMES = ["coarse", "select", "quantize", "scan", "finalize", "hnsw", "merge", "other", "fallback"]


class MuopdbGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[mgpu {code}] {msg}")
        self.code = code


class InvalidArgument(MuopdbGpuError, ValueError):
    pass


class OutOfRange(MuopdbGpuError, ValueError):
    """Where the reference panics (e.g. num_probes == 0 or > num_clusters, ivf/block_based/index.rs:158)."""


class Unsupported(MuopdbGpuError, NotImplementedError):
    pass


class NoDevice(MuopdbGpuError):
    pass


_ERR = {ERR_INVALID_ARG: InvalidArgument, ERR_OUT_OF_RANGE: OutOfRange, ERR_UNSUPPORTED: Unsupported,
        ERR_NO_DEVICE: NoDevice}

_vp, _u8p, _u32p, _u64p, _f32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses

# name -> (restype, argtypes); every symbol include/muopdb_gpu.h declares
SIGNATURES = {
    "mgpu_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mgpu_destroy": (None, [C.c_void_p]),
    "mgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "mgpu_version": (C.c_char_p, []),
    "mgpu_sync": (C.c_int, [C.c_void_p]),
    "mgpu_stream": (C.c_void_p, [C.c_void_p]),
    "mgpu_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_stream_signal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_device_sm_count": (C.c_int, [C.c_void_p]),
    "mgpu_timer_start": (C.c_int, [C.c_void_p]),
    "mgpu_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mgpu_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mgpu_profile_reset": (C.c_int, [C.c_void_p]),
    "mgpu_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "mgpu_launch_count": (C.c_uint64, [C.c_void_p]),
    "mgpu_last_kernel": (C.c_char_p, [C.c_void_p, C.c_int]),
    "mgpu_coarse_band_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "mgpu_distance_batch": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _f32p, C.c_int]),
    "mgpu_distance_batch_lanes": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _f32p, C.c_int]),
    "mgpu_pq_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, C.c_int, C.POINTER(C.c_void_p)]),
    "mgpu_pq_destroy": (None, [C.c_void_p]),
    "mgpu_pq_quantize_batch": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _u8p, C.c_int]),
    "mgpu_pq_distance_batch": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_uint64, _f32p, C.c_int]),
    "mgpu_pq_original_vector": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, _f32p, C.c_int]),
    "mgpu_ivf_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _f32p, _u64p, _u32p, C.c_int, C.c_int, C.c_void_p,
                                  _vp, C.c_int, C.c_uint64, _vp, C.POINTER(C.c_void_p)]),
    "mgpu_ivf_destroy": (None, [C.c_void_p]),
    "mgpu_ivf_num_vectors": (C.c_uint64, [C.c_void_p]),
    "mgpu_ivf_num_clusters": (C.c_uint32, [C.c_void_p]),
    "mgpu_ivf_invalidate": (C.c_int, [C.c_void_p, _u32p, C.c_uint32]),
    "mgpu_ivf_is_invalidated": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_int)]),
    "mgpu_ivf_invalidate_docs": (C.c_int, [C.c_void_p, _vp, C.c_uint32, _u8p, C.POINTER(C.c_uint32)]),
    "mgpu_ivf_is_doc_invalidated": (C.c_int, [C.c_void_p, _vp, C.POINTER(C.c_int)]),
    "mgpu_ivf_get_point_id": (C.c_int, [C.c_void_p, _vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]),
    "mgpu_ivf_get_doc_ids": (C.c_int, [C.c_void_p, _u32p, C.c_uint32, _vp]),
    "mgpu_ivf_get_vectors": (C.c_int, [C.c_void_p, _u32p, C.c_uint32, _vp]),
    "mgpu_ivf_coarse": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, _u32p, _f32p, C.c_int]),
    "mgpu_ivf_scan": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, _u32p, C.c_uint32, _u32p, C.c_uint32, _u32p, _f32p, _u32p, C.c_int]),
    "mgpu_ivf_scan_remap": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, _u32p, C.c_uint32, _u32p, C.c_uint32, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_ivf_search_submit": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p, C.POINTER(C.c_uint64)]),
    "mgpu_search_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mgpu_ivf_search": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_ivf_search_filtered": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint64, _vp, _f32p, _u32p,
                                           C.c_int]),
    "mgpu_ivf_scan_remap_filtered": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, _u32p, C.c_uint32, _u32p, C.c_uint32, _u32p, C.c_uint64,
                                               _vp, _f32p, _u32p, C.c_int]),
    "mgpu_ivf_last_scan_bytes": (C.c_uint64, [C.c_void_p]),
    "mgpu_ivf_last_scan_rows": (C.c_uint64, [C.c_void_p]),
    "mgpu_ivf_last_scan_fallbacks": (C.c_uint64, [C.c_void_p]),
    "mgpu_ivf_assign": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u32p, C.c_int]),
    "mgpu_kmeans_assign": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _f32p, C.c_uint32, C.c_uint32, C.c_int, _f32p, _u32p, _f32p, C.c_int]),
    "mgpu_hnsw_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _u32p, C.c_uint64, _u32p, C.c_uint64, _u64p, C.c_uint64,
                                   _u64p, C.c_int, C.c_int, C.c_void_p, _vp, C.c_int, C.c_uint64, _vp, C.POINTER(C.c_void_p)]),
    "mgpu_hnsw_destroy": (None, [C.c_void_p]),
    "mgpu_hnsw_search": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p, _u64p, C.c_int]),
    "mgpu_hnsw_search_submit": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p, C.POINTER(C.c_uint64)]),
    "mgpu_spann_search_submit": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _vp, _f32p, _u32p,
                                           C.POINTER(C.c_uint64)]),
    "mgpu_spann_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mgpu_spann_destroy": (None, [C.c_void_p]),
    "mgpu_spann_search": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_spann_search_filtered": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _u32p,
                                             C.c_uint64, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_batcher_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "mgpu_batcher_create_spann": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                            C.POINTER(C.c_void_p)]),
    "mgpu_batcher_destroy": (None, [C.c_void_p]),
    "mgpu_batcher_search": (C.c_int, [C.c_void_p, _f32p, _vp, _f32p, _u32p]),
    "mgpu_batcher_search_filtered": (C.c_int, [C.c_void_p, _f32p, _u32p, _vp, _f32p, _u32p]),
    "mgpu_batcher_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_merge_topk": (C.c_int, [C.c_void_p, _vp, _f32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mgpu_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mgpu_comm_destroy": (C.c_int, [C.c_void_p]),
    "mgpu_shard_ivf_search": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _vp, _f32p, _u32p, C.c_int]),
    "mgpu_shard_ivf_search_submit": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _vp, _f32p, _u32p, C.POINTER(C.c_uint64)]),
    "mgpu_shard_spann_search": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int, _vp, _f32p,
                                          _u32p, C.c_int]),
    "mgpu_shard_spann_search_submit": (C.c_int, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int, _vp,
                                                 _f32p, _u32p, C.POINTER(C.c_uint64)]),
    "mgpu_shard_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "mgpu_shard_allgather_merge": (C.c_int, [C.c_void_p, _vp, _f32p, _u32p, C.c_uint32, C.c_uint32, _vp, _f32p, _u32p]),
    "mgpu_ef_decode": (C.c_int64, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "mgpu_pq_load": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "mgpu_ivf_load": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mgpu_hnsw_load": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                 C.POINTER(C.c_void_p)]),
    "mgpu_user_index_info_decode": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_user_index_info_encode": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_user_index_info_read": (C.c_int64, [C.c_char_p, C.c_void_p, C.c_uint64]),
    "mgpu_spann_load_user": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mgpu_hnsw_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgpu_hnsw_copy_graph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(make -C muopdb_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def check(status: int, ctx=None):
    if status == OK:
        return
    msg = ""
    if ctx:
        m = load().mgpu_last_error(ctx)
        msg = m.decode() if m else ""
    raise _ERR.get(status, MuopdbGpuError)(status, msg or "error")

"""ctypes wrapper around oracle/liboracle.so (oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package (muopdb_b200) must never
import this module.  Parity status: pinned by tests/test_oracle_golden.py against the reference's own
known-answer tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

FLAT, PQ = 0, 1
L2, DOT = 0, 1
IMPL_SCALAR, IMPL_SIMD, IMPL_STREAMING = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


def _declare(L):
    for name in ("orc_l2_squared", "orc_l2", "orc_l2_scalar", "orc_dot", "orc_dot_scalar"):
        f = getattr(L, name)
        f.restype = C.c_float
        f.argtypes = [_f32p, _f32p, C.c_uint64]
    L.orc_lane_conforming.restype = C.c_float
    L.orc_lane_conforming.argtypes = [_f32p, _f32p, C.c_uint64, C.c_int, C.c_int]
    L.orc_distance_batch.restype = None
    L.orc_distance_batch.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _f32p]
    L.orc_pq_quantize.restype = None
    L.orc_pq_quantize.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _u8p]
    L.orc_pq_quantize_batch.restype = None
    L.orc_pq_quantize_batch.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, C.c_uint64, _u8p]
    L.orc_pq_original_vector.restype = None
    L.orc_pq_original_vector.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _f32p]
    L.orc_pq_distance.restype = C.c_float
    L.orc_pq_distance.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p, C.c_int, C.c_int]
    L.orc_merge_topk.restype = C.c_uint32
    L.orc_merge_topk.argtypes = [_u64p, _f32p, C.c_uint32, C.c_uint32, _u64p, _f32p]
    L.orc_ivf_new.restype = C.c_void_p
    L.orc_ivf_new.argtypes = [C.c_uint32, C.c_uint32, _f32p, _u64p, _u32p, C.c_int, C.c_int, C.c_void_p,
                              C.c_uint64, _u64p, _f32p, C.c_uint32, C.c_uint32]
    L.orc_ivf_free.restype = None
    L.orc_ivf_free.argtypes = [C.c_void_p]
    L.orc_ivf_set_hoist_quantize.restype = None
    L.orc_ivf_set_hoist_quantize.argtypes = [C.c_void_p, C.c_int]
    L.orc_ivf_invalidate.restype = None
    L.orc_ivf_invalidate.argtypes = [C.c_void_p, _u32p, C.c_uint32]
    L.orc_ivf_find_nearest_centroids.restype = C.c_int
    L.orc_ivf_find_nearest_centroids.argtypes = [C.c_void_p, _f32p, C.c_uint32, _u32p, _f32p]
    L.orc_ivf_search_with_centroids.restype = C.c_int
    L.orc_ivf_search_with_centroids.argtypes = [C.c_void_p, _f32p, _u32p, C.c_uint32, C.c_uint32, _u32p, _f32p]
    L.orc_ivf_search_with_centroids_and_remap.restype = C.c_int
    L.orc_ivf_search_with_centroids_and_remap.argtypes = [C.c_void_p, _f32p, _u32p, C.c_uint32, C.c_uint32, _u64p, _f32p]
    L.orc_ivf_search.restype = C.c_int
    L.orc_ivf_search.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, _u64p, _f32p]
    L.orc_ivf_search_batch.restype = None
    L.orc_ivf_search_batch.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _f32p, _i32p, C.c_int]
    L.orc_ivf_search_batch_f.restype = None
    L.orc_ivf_search_batch_f.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint64, _u64p, _f32p,
                                         _i32p, C.c_int]
    L.orc_spann_search_batch_f.restype = None
    L.orc_spann_search_batch_f.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_float, _u32p, C.c_uint64, _u64p, _f32p, _i32p, C.c_int]
    L.orc_ivf_assign.restype = None
    L.orc_ivf_assign.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _u32p, _u32p]
    L.orc_kmeans_assign.restype = None
    L.orc_kmeans_assign.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint32, C.c_uint32, C.c_int, _f32p, _u32p, _f32p]
    L.orc_lane_conforming_batch.restype = None
    L.orc_lane_conforming_batch.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _f32p]
    L.orc_hnsw_new.restype = C.c_void_p
    L.orc_hnsw_new.argtypes = [C.c_uint32, C.c_uint32, _u32p, _u32p, _u64p, C.c_uint64, _u64p, C.c_int, C.c_int,
                               C.c_void_p, C.c_uint64, _u64p, _f32p, C.c_uint32, C.c_uint32]
    L.orc_hnsw_free.restype = None
    L.orc_hnsw_free.argtypes = [C.c_void_p]
    L.orc_hnsw_entry_point.restype = C.c_uint32
    L.orc_hnsw_entry_point.argtypes = [C.c_void_p]
    L.orc_hnsw_ann_search.restype = C.c_int
    L.orc_hnsw_ann_search.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, _u64p, _f32p, _u32p, _u64p]
    L.orc_hnsw_search_batch.restype = None
    L.orc_hnsw_search_batch.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _f32p, _i32p, _u64p, C.c_int]
    L.orc_spann_search.restype = C.c_int
    L.orc_spann_search.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _u64p, _f32p]
    L.orc_spann_search_batch.restype = None
    L.orc_spann_search_batch.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_float, _u64p, _f32p, _i32p, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_num_threads.argtypes = []


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u128_to_pairs(doc_ids) -> np.ndarray | None:
    """list/array of python ints (u128) or (n,2) u64 -> (n,2) u64 (lo,hi)."""
    if doc_ids is None:
        return None
    a = np.asarray(doc_ids)
    if a.dtype == np.uint64 and a.ndim == 2:
        return np.ascontiguousarray(a)
    out = np.empty((len(doc_ids), 2), dtype=np.uint64)
    for i, d in enumerate(doc_ids):
        d = int(d)
        out[i, 0] = d & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = d >> 64
    return out


def pairs_to_u128(p: np.ndarray) -> list[int]:
    return [int(lo) | (int(hi) << 64) for lo, hi in np.asarray(p, dtype=np.uint64).reshape(-1, 2)]


# ---- distances -------------------------------------------------------------------------------
def l2_squared(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_l2_squared(_p(a, _f32p), _p(b, _f32p), a.size))


def l2(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_l2(_p(a, _f32p), _p(b, _f32p), a.size))


def l2_scalar(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_l2_scalar(_p(a, _f32p), _p(b, _f32p), a.size))


def dot(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot(_p(a, _f32p), _p(b, _f32p), a.size))


def dot_scalar(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot_scalar(_p(a, _f32p), _p(b, _f32p), a.size))


def lane_conforming(a, b, lanes, metric=L2):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_lane_conforming(_p(a, _f32p), _p(b, _f32p), a.size, lanes, metric))


def distance_batch(A, B, metric=L2, squared=False):
    A, B = _f32(A), _f32(B)
    out = np.empty((A.shape[0], B.shape[0]), dtype=np.float32)
    lib().orc_distance_batch(_p(A, _f32p), A.shape[0], _p(B, _f32p), B.shape[0], A.shape[1], metric, int(squared),
                             _p(out, _f32p))
    return out


# ---- product quantizer -----------------------------------------------------------------------
class ProductQuantizer:
    """rs/quantization/src/pq/mod.rs:23-286 (codebook is an input)."""

    def __init__(self, dimension, subvector_dimension, num_bits, codebook, metric=L2):
        assert dimension % subvector_dimension == 0
        self.dimension, self.subvector_dimension, self.num_bits = dimension, subvector_dimension, num_bits
        self.codebook = _f32(codebook).reshape(-1)
        assert self.codebook.size == dimension * (1 << num_bits)
        self.metric = metric

    def quantized_dimension(self):
        return self.dimension // self.subvector_dimension

    def quantize(self, v):
        v = _f32(v)
        if v.ndim == 1:
            out = np.empty(self.quantized_dimension(), dtype=np.uint8)
            lib().orc_pq_quantize(_p(self.codebook, _f32p), self.dimension, self.subvector_dimension, self.num_bits,
                                  _p(v, _f32p), _p(out, _u8p))
            return out
        out = np.empty((v.shape[0], self.quantized_dimension()), dtype=np.uint8)
        lib().orc_pq_quantize_batch(_p(self.codebook, _f32p), self.dimension, self.subvector_dimension, self.num_bits,
                                    _p(v, _f32p), v.shape[0], _p(out, _u8p))
        return out

    def original_vector(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        out = np.empty(self.dimension, dtype=np.float32)
        lib().orc_pq_original_vector(_p(self.codebook, _f32p), self.dimension, self.subvector_dimension, self.num_bits,
                                     _p(codes, _u8p), _p(out, _f32p))
        return out

    def distance(self, a, b, impl=IMPL_STREAMING):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        return float(lib().orc_pq_distance(_p(self.codebook, _f32p), self.dimension, self.subvector_dimension,
                                           self.num_bits, _p(a, _u8p), _p(b, _u8p), impl, self.metric))


# ---- merge -----------------------------------------------------------------------------------
def merge_topk(doc_ids, scores, k):
    d = _u128_to_pairs(doc_ids)
    s = _f32(scores)
    od = np.empty((max(k, 1), 2), dtype=np.uint64)
    os_ = np.empty(max(k, 1), dtype=np.float32)
    n = lib().orc_merge_topk(_p(d, _u64p), _p(s, _f32p), len(s), k, _p(od, _u64p), _p(os_, _f32p))
    return pairs_to_u128(od[:n]), os_[:n].copy()


# ---- IVF -------------------------------------------------------------------------------------
class Ivf:
    """BlockBasedIvf<Q> over in-memory arrays (rs/index/src/ivf/block_based/index.rs)."""

    def __init__(self, centroids, list_offsets, list_ids, rows, doc_ids=None, pq: ProductQuantizer | None = None,
                 metric=L2, dim=None):
        self.centroids = _f32(centroids)
        self.nlist, self.dim = self.centroids.shape
        self.list_offsets = np.ascontiguousarray(list_offsets, dtype=np.uint64)
        self.list_ids = np.ascontiguousarray(list_ids, dtype=np.uint32)
        self.pq = pq
        if pq is None:
            self.rows = _f32(rows)
            quant = FLAT
        else:
            self.rows = np.ascontiguousarray(rows, dtype=np.uint8)
            quant = PQ
        self.n = self.rows.shape[0]
        self.doc_ids = _u128_to_pairs(doc_ids)
        cb = pq.codebook if pq else None
        self.h = lib().orc_ivf_new(self.dim, self.nlist, _p(self.centroids, _f32p), _p(self.list_offsets, _u64p),
                                   _p(self.list_ids, _u32p), quant, pq.metric if pq else metric,
                                   self.rows.ctypes.data_as(C.c_void_p), self.n, _p(self.doc_ids, _u64p),
                                   _p(cb, _f32p), pq.subvector_dimension if pq else 0, pq.num_bits if pq else 0)

    def __del__(self):
        try:
            lib().orc_ivf_free(self.h)
        except Exception:
            pass

    def set_hoist_quantize(self, on: bool):
        lib().orc_ivf_set_hoist_quantize(self.h, int(on))

    def invalidate_batch(self, point_ids):
        a = np.ascontiguousarray(point_ids, dtype=np.uint32)
        lib().orc_ivf_invalidate(self.h, _p(a, _u32p), a.size)

    def find_nearest_centroids(self, q, nprobe, with_dist=False):
        q = _f32(q)
        ids = np.empty(max(nprobe, 1), dtype=np.uint32)
        ds = np.empty(max(nprobe, 1), dtype=np.float32)
        r = lib().orc_ivf_find_nearest_centroids(self.h, _p(q, _f32p), nprobe, _p(ids, _u32p), _p(ds, _f32p))
        if r < 0:
            raise ValueError("num_probes out of range (reference panics)")
        return (ids[:r].copy(), ds[:r].copy()) if with_dist else ids[:r].copy()

    def search_with_centroids(self, q, cids, k):
        q = _f32(q)
        cids = np.ascontiguousarray(cids, dtype=np.uint32)
        pids = np.empty(max(k, 1), dtype=np.uint32)
        ds = np.empty(max(k, 1), dtype=np.float32)
        n = lib().orc_ivf_search_with_centroids(self.h, _p(q, _f32p), _p(cids, _u32p), cids.size, k, _p(pids, _u32p),
                                                _p(ds, _f32p))
        return pids[:n].copy(), ds[:n].copy()

    def search_with_centroids_and_remap(self, q, cids, k):
        q = _f32(q)
        cids = np.ascontiguousarray(cids, dtype=np.uint32)
        od = np.empty((max(k, 1), 2), dtype=np.uint64)
        os_ = np.empty(max(k, 1), dtype=np.float32)
        n = lib().orc_ivf_search_with_centroids_and_remap(self.h, _p(q, _f32p), _p(cids, _u32p), cids.size, k,
                                                          _p(od, _u64p), _p(os_, _f32p))
        return pairs_to_u128(od[:n]), os_[:n].copy()

    def search(self, q, k, nprobe):
        q = _f32(q)
        od = np.empty((max(k, 1), 2), dtype=np.uint64)
        os_ = np.empty(max(k, 1), dtype=np.float32)
        n = lib().orc_ivf_search(self.h, _p(q, _f32p), k, nprobe, _p(od, _u64p), _p(os_, _f32p))
        if n < 0:
            raise ValueError("num_probes out of range (reference panics)")
        return pairs_to_u128(od[:n]), os_[:n].copy()

    def search_batch(self, Q, k, nprobe, nthreads=0, filter_bits=None):
        """-> doc_ids (B,k,2) u64, scores (B,k) f32, counts (B,) i32.  filter_bits: planner filter as point-id bitmaps,
        (words,) shared by all queries or (B, words) one per query (index.rs:212-226)."""
        Q = _f32(Q)
        B = Q.shape[0]
        od = np.zeros((B, max(k, 1), 2), dtype=np.uint64)
        os_ = np.zeros((B, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(B, dtype=np.int32)
        if filter_bits is None:
            lib().orc_ivf_search_batch(self.h, _p(Q, _f32p), B, k, nprobe, _p(od, _u64p), _p(os_, _f32p), _p(cnt, _i32p),
                                       nthreads)
        else:
            fb = np.ascontiguousarray(filter_bits, dtype=np.uint32)
            stride = fb.shape[1] if fb.ndim == 2 else 0
            lib().orc_ivf_search_batch_f(self.h, _p(Q, _f32p), B, k, nprobe, _p(fb, _u32p), stride, _p(od, _u64p),
                                         _p(os_, _f32p), _p(cnt, _i32p), nthreads)
        return od, os_, cnt


def ivf_assign(X, centroids, max_clusters=1, threshold=0.1):
    """IvfBuilder::find_nearest_centroids + acceptance rule (ivf/builder.rs:268-329)."""
    X, centroids = _f32(X), _f32(centroids)
    n, dim = X.shape
    cids = np.empty((n, max_clusters), dtype=np.uint32)
    cnt = np.empty(n, dtype=np.uint32)
    lib().orc_ivf_assign(_p(X, _f32p), n, _p(centroids, _f32p), centroids.shape[0], dim, max_clusters,
                         float(threshold), _p(cids, _u32p), _p(cnt, _u32p))
    return cids, cnt


def build_posting_lists(X, centroids, max_clusters=1, threshold=0.1):
    """build_posting_lists (ivf/builder.rs:292-366): lists of ascending point ids -> (offsets, ids)."""
    cids, cnt = ivf_assign(X, centroids, max_clusters, threshold)
    nlist = centroids.shape[0]
    lists = [[] for _ in range(nlist)]
    for pid in range(X.shape[0]):
        for j in range(cnt[pid]):
            lists[cids[pid, j]].append(pid)
    offsets = np.zeros(nlist + 1, dtype=np.uint64)
    for c in range(nlist):
        offsets[c + 1] = offsets[c] + len(lists[c])
    ids = np.array([p for l in lists for p in sorted(l)], dtype=np.uint32)
    return offsets, ids


def kmeans_assign(X, centroids, penalties=None, metric=L2, with_costs=False):
    """Assignment step of KMeansBuilder::run_lloyd (kmeans_builder.rs:199-221), calculator chosen by dimension (:126-136)."""
    X, centroids = _f32(X), _f32(centroids)
    out = np.empty(X.shape[0], dtype=np.uint32)
    costs = np.empty(X.shape[0], dtype=np.float32)
    pen = _f32(penalties) if penalties is not None else None
    lib().orc_kmeans_assign(_p(X, _f32p), X.shape[0], _p(centroids, _f32p), centroids.shape[0], X.shape[1], metric,
                            _p(pen, _f32p), _p(out, _u32p), _p(costs, _f32p))
    return (out, costs) if with_costs else out


def lane_conforming_batch(A, B, lanes, metric=L2):
    """All pairs of LaneConformingDistanceCalculator<LANES, D>::calculate_squared (lane_conforming.rs:16-27)."""
    A, B = _f32(A), _f32(B)
    out = np.empty((A.shape[0], B.shape[0]), dtype=np.float32)
    lib().orc_lane_conforming_batch(_p(A, _f32p), A.shape[0], _p(B, _f32p), B.shape[0], A.shape[1], lanes, metric, _p(out, _f32p))
    return out


# ---- HNSW ------------------------------------------------------------------------------------
class Hnsw:
    """BlockBasedHnsw<Q> over the graph arrays of the on-disk format."""

    def __init__(self, num_layers, edges, points, edge_offsets, level_offsets, rows, doc_ids=None,
                 pq: ProductQuantizer | None = None, metric=L2):
        self.num_layers = int(num_layers)
        self.edges = np.ascontiguousarray(edges, dtype=np.uint32)
        self.points = np.ascontiguousarray(points, dtype=np.uint32)
        if self.points.size == 0:
            self.points = np.zeros(1, dtype=np.uint32)
        self.edge_offsets = np.ascontiguousarray(edge_offsets, dtype=np.uint64)
        self.level_offsets = np.ascontiguousarray(level_offsets, dtype=np.uint64)
        assert self.level_offsets.size == self.num_layers + 1
        self.pq = pq
        if pq is None:
            self.rows = _f32(rows)
            quant, dim = FLAT, self.rows.shape[1]
        else:
            self.rows = np.ascontiguousarray(rows, dtype=np.uint8)
            quant, dim = PQ, pq.dimension
        self.dim = dim
        self.n = self.rows.shape[0]
        self.doc_ids = _u128_to_pairs(doc_ids)
        cb = pq.codebook if pq else None
        self.h = lib().orc_hnsw_new(dim, self.num_layers, _p(self.edges, _u32p), _p(self.points, _u32p),
                                    _p(self.edge_offsets, _u64p), self.edge_offsets.size, _p(self.level_offsets, _u64p),
                                    quant, pq.metric if pq else metric, self.rows.ctypes.data_as(C.c_void_p), self.n,
                                    _p(self.doc_ids, _u64p), _p(cb, _f32p), pq.subvector_dimension if pq else 0,
                                    pq.num_bits if pq else 0)

    def __del__(self):
        try:
            lib().orc_hnsw_free(self.h)
        except Exception:
            pass

    def entry_point(self):
        return int(lib().orc_hnsw_entry_point(self.h))

    def ann_search(self, q, k, ef, with_stats=False, with_point_ids=False):
        q = _f32(q)
        od = np.empty((max(k, 1), 2), dtype=np.uint64)
        os_ = np.empty(max(k, 1), dtype=np.float32)
        pids = np.empty(max(k, 1), dtype=np.uint32)
        st = np.zeros(2, dtype=np.uint64)
        n = lib().orc_hnsw_ann_search(self.h, _p(q, _f32p), k, ef, _p(od, _u64p), _p(os_, _f32p), _p(pids, _u32p),
                                      _p(st, _u64p))
        res = [pairs_to_u128(od[:n]), os_[:n].copy()]
        if with_point_ids:
            res.append(pids[:n].copy())
        if with_stats:
            res.append((int(st[0]), int(st[1])))
        return tuple(res)

    def search_batch(self, Q, k, ef, nthreads=0):
        Q = _f32(Q)
        B = Q.shape[0]
        od = np.zeros((B, max(k, 1), 2), dtype=np.uint64)
        os_ = np.zeros((B, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(B, dtype=np.int32)
        st = np.zeros((B, 2), dtype=np.uint64)
        lib().orc_hnsw_search_batch(self.h, _p(Q, _f32p), B, k, ef, _p(od, _u64p), _p(os_, _f32p), _p(cnt, _i32p),
                                    _p(st, _u64p), nthreads)
        return od, os_, cnt, st


# ---- SPANN -----------------------------------------------------------------------------------
class Spann:
    """Spann<Q>::search (rs/index/src/spann/index.rs:211-266)."""

    def __init__(self, centroids: Hnsw, posting_lists: Ivf):
        self.centroids, self.posting_lists = centroids, posting_lists

    def search(self, q, top_k, ef_construction, num_explored_centroids=None, centroid_distance_ratio=0.1):
        q = _f32(q)
        ne = top_k if num_explored_centroids is None else num_explored_centroids
        od = np.empty((max(top_k, 1), 2), dtype=np.uint64)
        os_ = np.empty(max(top_k, 1), dtype=np.float32)
        n = lib().orc_spann_search(self.centroids.h, self.posting_lists.h, _p(q, _f32p), top_k, ef_construction, ne,
                                   float(centroid_distance_ratio), _p(od, _u64p), _p(os_, _f32p))
        if n < 0:
            return None
        return pairs_to_u128(od[:n]), os_[:n].copy()

    def search_batch(self, Q, top_k, ef_construction, num_explored_centroids=None, centroid_distance_ratio=0.1,
                     nthreads=0, filter_bits=None):
        Q = _f32(Q)
        B = Q.shape[0]
        ne = top_k if num_explored_centroids is None else num_explored_centroids
        od = np.zeros((B, max(top_k, 1), 2), dtype=np.uint64)
        os_ = np.zeros((B, max(top_k, 1)), dtype=np.float32)
        cnt = np.zeros(B, dtype=np.int32)
        if filter_bits is None:
            lib().orc_spann_search_batch(self.centroids.h, self.posting_lists.h, _p(Q, _f32p), B, top_k, ef_construction,
                                         ne, float(centroid_distance_ratio), _p(od, _u64p), _p(os_, _f32p),
                                         _p(cnt, _i32p), nthreads)
        else:
            fb = np.ascontiguousarray(filter_bits, dtype=np.uint32)
            stride = fb.shape[1] if fb.ndim == 2 else 0
            lib().orc_spann_search_batch_f(self.centroids.h, self.posting_lists.h, _p(Q, _f32p), B, top_k, ef_construction,
                                           ne, float(centroid_distance_ratio), _p(fb, _u32p), stride, _p(od, _u64p),
                                           _p(os_, _f32p), _p(cnt, _i32p), nthreads)
        return od, os_, cnt


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ---- seeded builders of synthetic index inputs (builders.c) -------------------------------------
def _declare_builders(L):
    L.orc_kmeans.restype = None
    L.orc_kmeans.argtypes = [_f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _f32p]
    L.orc_hnsw_build.restype = C.c_void_p
    L.orc_hnsw_build.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    L.orc_graph_sizes.restype = None
    L.orc_graph_sizes.argtypes = [C.c_void_p, _u64p]
    L.orc_graph_copy.restype = None
    L.orc_graph_copy.argtypes = [C.c_void_p, _u32p, _u32p, _u64p, _u64p]
    L.orc_graph_free.restype = None
    L.orc_graph_free.argtypes = [C.c_void_p]


def kmeans(X, k, iters=10, seed=1234):
    X = _f32(X)
    L = lib()
    _declare_builders(L)
    out = np.empty((k, X.shape[1]), dtype=np.float32)
    L.orc_kmeans(_p(X, _f32p), X.shape[0], X.shape[1], k, iters, seed, _p(out, _f32p))
    return out


def train_pq_codebook(X, dsub, nbits, iters=10, seed=1234):
    """Per-subspace k-means -> codebook laid out [subspace][centroid][dsub] (pq/mod.rs:155-167)."""
    X = _f32(X)
    dim = X.shape[1]
    m, K = dim // dsub, 1 << nbits
    cb = np.empty((m, K, dsub), dtype=np.float32)
    for s in range(m):
        cb[s] = kmeans(np.ascontiguousarray(X[:, s * dsub:(s + 1) * dsub]), K, iters, seed + s)
    return cb.reshape(-1)


def hnsw_build(X, max_neighbors=32, max_layer=10, ef_construction=100, seed=1234):
    """-> dict(num_layers, edges, points, edge_offsets, level_offsets) in the a14 array layout."""
    X = _f32(X)
    L = lib()
    _declare_builders(L)
    g = L.orc_hnsw_build(_p(X, _f32p), X.shape[0], X.shape[1], max_neighbors, max_layer, ef_construction, seed)
    sz = np.zeros(4, dtype=np.uint64)
    L.orc_graph_sizes(g, _p(sz, _u64p))
    nl, ne, npnt, neo = (int(v) for v in sz)
    edges = np.zeros(max(ne, 1), dtype=np.uint32)
    points = np.zeros(max(npnt, 1), dtype=np.uint32)
    eo = np.zeros(neo, dtype=np.uint64)
    lo = np.zeros(nl + 1, dtype=np.uint64)
    L.orc_graph_copy(g, _p(edges, _u32p), _p(points, _u32p), _p(eo, _u64p), _p(lo, _u64p))
    L.orc_graph_free(g)
    return dict(num_layers=nl, edges=edges[:ne], points=points[:npnt], edge_offsets=eo, level_offsets=lo)

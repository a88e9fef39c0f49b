"""muopdb_b200 -- B200-native (sm_100a) batched ANN search path for MuopDB.

Host-side mirror of the reference's interfaces for the search hot path, over the C ABI of include/muopdb_gpu.h:

    reference (Rust)                                              here
    utils::distance::l2::L2DistanceCalculator                     L2DistanceCalculator
    utils::distance::dot_product::DotProductDistanceCalculator    DotProductDistanceCalculator
    quantization::noq::NoQuantizer<D>                             NoQuantizer
    quantization::pq::ProductQuantizer<D>                         ProductQuantizer
    index::ivf::block_based::index::BlockBasedIvf<Q>              BlockBasedIvf
    index::hnsw::block_based::index::BlockBasedHnsw<Q>            BlockBasedHnsw
    index::spann::index::Spann<Q>                                 Spann
    config::search_params::SearchParams                           SearchParams
    index::utils::{IdWithScore, SearchResult}                     IdWithScore, SearchResult

Every method that computes runs hand-written CUDA kernels; there is no CPU fallback (a missing library or device
raises).  Single-query methods keep the reference's names and argument meaning; `*_batch` methods are the batched
entry points the GPU path is built for.  Array arguments may be numpy arrays (host) or torch CUDA tensors (device).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (DEVICE, DOT, HOST, L2, QUANT_NONE, QUANT_PQ, InvalidArgument, MuopdbGpuError, NoDevice, OutOfRange,
                   Unsupported)

__all__ = ["Context", "default_context", "L2DistanceCalculator", "DotProductDistanceCalculator",
           "LaneConformingDistanceCalculator", "kmeans_assign", "UserIndexInfo", "MultiSpannIndex", "NoQuantizer",
           "ProductQuantizer", "BlockBasedIvf", "BlockBasedHnsw", "Spann", "SearchParams", "IdWithScore", "SearchResult",
           "Planner", "MicroBatcher", "merge_topk", "assign_to_centroids", "elias_fano_decode", "L2", "DOT", "MuopdbGpuError", "OutOfRange", "InvalidArgument", "Unsupported",
           "NoDevice"]


# ---- buffers ---------------------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Buf:
    """Pointer + residency of an array argument (numpy host array or torch CUDA tensor)."""

    def __init__(self, x, dtype, shape=None, allow_none=False):
        self.keep = None
        if x is None:
            assert allow_none
            self.ptr, self.mem = None, None
            return
        if _is_torch(x):
            import torch
            tdt = {np.float32: torch.float32, np.uint8: torch.uint8, np.uint32: torch.int32, np.int32: torch.int32,
                   np.uint64: torch.int64}[dtype]
            if x.dtype != tdt and not (dtype == np.uint32 and x.dtype in (torch.int32, torch.uint32)) \
                    and not (dtype == np.uint64 and x.dtype in (torch.int64, torch.uint64)):
                raise InvalidArgument(_lib.ERR_INVALID_ARG, f"tensor dtype {x.dtype} != {tdt}")
            if not x.is_contiguous():
                x = x.contiguous()
            self.keep = x
            self.ptr = x.data_ptr()
            self.mem = DEVICE if x.is_cuda else HOST
            self.shape = tuple(x.shape)
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            self.keep = a
            self.ptr = a.ctypes.data
            self.mem = HOST
            self.shape = a.shape
        if shape is not None:
            a_shape = self.shape
            for want, got in zip(shape, a_shape):
                if want is not None and want != got:
                    raise InvalidArgument(_lib.ERR_INVALID_ARG, f"shape {a_shape} does not match {shape}")


def _pairs_to_ints(p: np.ndarray) -> list:
    p = np.asarray(p, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in p]


def _ints_to_pairs(ids) -> Optional[np.ndarray]:
    if ids is None:
        return None
    a = np.asarray(ids)
    if a.dtype == np.uint64 and a.ndim == 2 and a.shape[1] == 2:
        return np.ascontiguousarray(a)
    if a.dtype != object and a.ndim == 1 and a.dtype.kind in "iu":
        out = np.zeros((a.shape[0], 2), dtype=np.uint64)
        out[:, 0] = a.astype(np.uint64)
        return out
    out = np.empty((len(ids), 2), dtype=np.uint64)
    for i, d in enumerate(ids):
        d = int(d)
        out[i, 0] = d & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = d >> 64
    return out


# ---- context ---------------------------------------------------------------------------------------------------------
class Context:
    """One mgpu_ctx: a device, a stream and a workspace.  Calls on one context are serialised."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.mgpu_init(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self.lib.mgpu_sync(self.h), self.h)

    def order(self, mem):
        """Context manager around one C-ABI call with DEVICE buffers: the library stream is non-blocking, so torch's current
        stream and the library stream are ordered explicitly -- before the call the library stream waits for torch's stream
        (inputs computed by earlier torch ops, the zero-fill of freshly allocated outputs), after it torch's stream waits for
        the library streams (so `.cpu()`, indexing, ... see the results).  No-op for host buffers and when torch's current
        stream IS the library stream.  With shard_overlap(True) the trailing wait is skipped (it would chain consecutive
        sharded calls through torch's stream); results are then complete after sync()."""
        return _StreamOrder(self, mem == DEVICE)

    @property
    def sm_count(self) -> int:
        return self.lib.mgpu_device_sm_count(self.h)

    @property
    def stream(self) -> int:
        return self.lib.mgpu_stream(self.h)

    def timer_start(self):
        _lib.check(self.lib.mgpu_timer_start(self.h), self.h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        _lib.check(self.lib.mgpu_timer_stop(self.h, C.byref(ms)), self.h)
        return ms.value

    def profile_enable(self, on=True, classes=None):
        """classes: iterable of kernel classes (_lib.K_*) to bracket with events; None = all (when `on`)."""
        v = int(bool(on)) if classes is None or not on else -sum(1 << c for c in classes)
        _lib.check(self.lib.mgpu_profile_enable(self.h, v), self.h)

    def profile_reset(self):
        _lib.check(self.lib.mgpu_profile_reset(self.h), self.h)

    def profile_get(self, kernel_class: int):
        ms, n = C.c_float(), C.c_uint64()
        _lib.check(self.lib.mgpu_profile_get(self.h, kernel_class, C.byref(ms), C.byref(n)), self.h)
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.mgpu_launch_count(self.h))

    def coarse_band_stats(self, reset: bool = False):
        """(centroids re-scored exactly in the tensor-core selection's uncertain band, queries) since the last reset."""
        out = (C.c_uint64 * 2)()
        _lib.check(self.lib.mgpu_coarse_band_stats(self.h, out, 1 if reset else 0), self.h)
        return int(out[0]), int(out[1])

    def last_kernel(self, kernel_class: int) -> str:
        """Name of the kernel launched last in a class (which of the alternative scan / HNSW kernels served the call)."""
        return (self.lib.mgpu_last_kernel(self.h, kernel_class) or b"").decode()

    # multi-GPU (one process per GPU)
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.mgpu_comm_init(self.h, nranks, rank, C.addressof(buf)), self.h)

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.load().mgpu_comm_unique_id(C.addressof(buf)))
        return bytes(buf)

    _overlap = False

    def shard_overlap(self, on: bool = True):
        """mgpu_shard_overlap: device-buffer sharded calls return with their result exchange still in flight on the
        exchange stream (the next batch's kernels run next to it); `sync()` completes both streams."""
        _lib.check(self.lib.mgpu_shard_overlap(self.h, int(bool(on))), self.h)
        self._overlap = bool(on)

    def shard_allgather_merge(self, local_doc_ids, local_scores, local_counts, B, k, out_doc_ids, out_scores, out_counts):
        """All device tensors: doc ids (B,k,2) int64, scores (B,k) f32, counts (B,) int32."""
        _lib.check(self.lib.mgpu_shard_allgather_merge(self.h, local_doc_ids.data_ptr(), local_scores.data_ptr(),
                                                       local_counts.data_ptr(), B, k, out_doc_ids.data_ptr(),
                                                       out_scores.data_ptr(), out_counts.data_ptr()), self.h)


class _StreamOrder:
    def __init__(self, ctx, active):
        self.ctx, self.active, self.stream = ctx, active, None

    def __enter__(self):
        if self.active:
            import torch
            st = torch.cuda.current_stream(self.ctx.device).cuda_stream
            if st != (self.ctx.stream or 0):
                self.stream = st
                _lib.check(self.ctx.lib.mgpu_stream_wait(self.ctx.h, st), self.ctx.h)
        return self

    def __exit__(self, et, ev, tb):
        if self.active and self.stream is not None and et is None and not self.ctx._overlap:
            _lib.check(self.ctx.lib.mgpu_stream_signal(self.ctx.h, self.stream), self.ctx.h)
        return False


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- DistanceCalculator (rs/utils/src/lib.rs:17-40) ---------------------------------------------------------------------
class _DistanceCalculator:
    METRIC = L2

    @classmethod
    def calculate_batch(cls, A, B, squared=False, ctx: Optional[Context] = None):
        """All pairs: out[i, j] = calculate(A[i], B[j]) (calculate_squared when squared=True)."""
        ctx = ctx or default_context()
        a, b = _Buf(A, np.float32), _Buf(B, np.float32)
        if len(a.shape) != 2 or len(b.shape) != 2 or a.shape[1] != b.shape[1]:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "calculate_batch expects (nA, dim) and (nB, dim)")
        if a.mem != b.mem:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "A and B must live in the same memory space")
        nA, dim = a.shape
        nB = b.shape[0]
        if a.mem == DEVICE:
            import torch
            out = torch.empty((nA, nB), dtype=torch.float32, device=A.device)
            optr = out.data_ptr()
        else:
            out = np.empty((nA, nB), dtype=np.float32)
            optr = out.ctypes.data
        with ctx.order(a.mem):
            _lib.check(ctx.lib.mgpu_distance_batch(ctx.h, a.ptr, nA, b.ptr, nB, dim, cls.METRIC, int(squared), optr, a.mem), ctx.h)
        return out

    @classmethod
    def calculate(cls, a, b) -> float:
        """DistanceCalculator::calculate(a, b) -> f32."""
        a = np.asarray(a, dtype=np.float32).reshape(1, -1)
        b = np.asarray(b, dtype=np.float32).reshape(1, -1)
        return float(cls.calculate_batch(a, b)[0, 0])

    @classmethod
    def calculate_squared(cls, a, b) -> float:
        """CalculateSquared::calculate_squared(a, b) -> f32."""
        a = np.asarray(a, dtype=np.float32).reshape(1, -1)
        b = np.asarray(b, dtype=np.float32).reshape(1, -1)
        return float(cls.calculate_batch(a, b, squared=True)[0, 0])


class L2DistanceCalculator(_DistanceCalculator):
    """rs/utils/src/distance/l2.rs:17-100"""
    METRIC = L2


class DotProductDistanceCalculator(_DistanceCalculator):
    """rs/utils/src/distance/dot_product.rs:7-99 (negated dot; calculate_squared forwards to calculate)"""
    METRIC = DOT


class LaneConformingDistanceCalculator:
    """rs/utils/src/distance/lane_conforming.rs:9-28: LaneConformingDistanceCalculator<LANES, D> -- only CalculateSquared;
    the dimension must be a multiple of LANES."""

    def __init__(self, lanes: int, distance=L2DistanceCalculator):
        self.lanes, self.calculator, self.METRIC = lanes, distance, distance.METRIC

    def calculate_squared_batch(self, A, B, ctx: Optional[Context] = None):
        ctx = ctx or default_context()
        a, b = _Buf(A, np.float32), _Buf(B, np.float32)
        if len(a.shape) != 2 or len(b.shape) != 2 or a.shape[1] != b.shape[1] or a.mem != b.mem:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "calculate_squared_batch expects (nA, dim) and (nB, dim) in one memory space")
        nA, dim = a.shape
        nB = b.shape[0]
        if a.mem == DEVICE:
            import torch
            out = torch.empty((nA, nB), dtype=torch.float32, device=A.device)
            optr = out.data_ptr()
        else:
            out = np.empty((nA, nB), dtype=np.float32)
            optr = out.ctypes.data
        with ctx.order(a.mem):
            _lib.check(ctx.lib.mgpu_distance_batch_lanes(ctx.h, a.ptr, nA, b.ptr, nB, dim, self.METRIC, self.lanes, optr, a.mem), ctx.h)
        return out

    def calculate_squared(self, a, b) -> float:
        a = np.asarray(a, dtype=np.float32).reshape(1, -1)
        b = np.asarray(b, dtype=np.float32).reshape(1, -1)
        return float(self.calculate_squared_batch(a, b)[0, 0])


def kmeans_assign(X, centroids, penalties=None, distance=L2DistanceCalculator, with_costs=False, ctx: Optional[Context] = None):
    """Assignment step of KMeansBuilder::run_lloyd (rs/utils/src/kmeans_builder/kmeans_builder.rs:199-221): per row the first
    minimum of T::calculate_squared(x, c) + penalties[c], T chosen by the dimension (kmeans_builder.rs:126-136).
    -> labels (n,) [, costs (n,)]"""
    ctx = ctx or default_context()
    x, c = _Buf(X, np.float32), _Buf(centroids, np.float32)
    p = _Buf(penalties, np.float32, allow_none=True)
    if x.mem != c.mem or (p.mem is not None and p.mem != x.mem):
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "X, centroids and penalties must live in the same memory space")
    n, dim = x.shape
    if c.shape[1] != dim or (p.ptr is not None and p.shape[0] != c.shape[0]):
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "shape mismatch")
    if x.mem == DEVICE:
        import torch
        labels = torch.zeros((n,), dtype=torch.int32, device=X.device)
        costs = torch.zeros((n,), dtype=torch.float32, device=X.device)
        lp, cp = labels.data_ptr(), costs.data_ptr()
    else:
        labels = np.zeros((n,), dtype=np.uint32)
        costs = np.zeros((n,), dtype=np.float32)
        lp, cp = labels.ctypes.data, costs.ctypes.data
    with ctx.order(x.mem):
        _lib.check(ctx.lib.mgpu_kmeans_assign(ctx.h, x.ptr, n, c.ptr, c.shape[0], dim, distance.METRIC, p.ptr, lp, cp, x.mem), ctx.h)
    return (labels, costs) if with_costs else labels


# ---- Quantizers (rs/quantization/src/quantization.rs:6-38) ----------------------------------------------------------------
class NoQuantizer:
    """rs/quantization/src/noq/mod.rs:14-60"""
    QUANT = QUANT_NONE

    def __init__(self, dimension: int, distance=L2DistanceCalculator):
        self.dimension = dimension
        self.calculator = distance
        self.metric = distance.METRIC
        self.handle = None

    def quantize(self, value):
        return np.array(value, dtype=np.float32, copy=True)

    def quantized_dimension(self) -> int:
        return self.dimension

    def original_vector(self, quantized):
        return np.array(quantized, dtype=np.float32, copy=True)

    def distance(self, query, point) -> float:
        return self.calculator.calculate(query, point)


class ProductQuantizer:
    """rs/quantization/src/pq/mod.rs:23-286.  The codebook ([subspace][centroid][dsub] f32) is an input."""
    QUANT = QUANT_PQ

    def __init__(self, dimension: int, subvector_dimension: int, num_bits: int, codebook, distance=L2DistanceCalculator,
                 ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.dimension, self.subvector_dimension, self.num_bits = dimension, subvector_dimension, num_bits
        self.metric = distance.METRIC
        cb = np.ascontiguousarray(codebook, dtype=np.float32).reshape(-1)
        if subvector_dimension <= 0 or dimension % subvector_dimension != 0:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "Dimensions are not valid")  # pq/mod.rs:41-46
        if cb.size != dimension * (1 << num_bits):
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "codebook size does not match dimension * 2^num_bits")
        self.codebook = cb
        h = C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_pq_create(self.ctx.h, dimension, subvector_dimension, num_bits, cb.ctypes.data,
                                               self.metric, C.byref(h)), self.ctx.h)
        self.handle = h

    @classmethod
    def read(cls, dir: str, distance=L2DistanceCalculator, ctx: Optional[Context] = None):
        """Quantizer::read / ProductQuantizerReader::read (pq/mod.rs:52-136): yaml config + raw f32 codebook."""
        import os
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.metric = distance.METRIC
        h = C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_pq_load(self.ctx.h, dir.encode(), self.metric, C.byref(h)), self.ctx.h)
        self.handle = h
        cfg = {}
        for line in open(os.path.join(dir, "product_quantizer_config.yaml")):
            if ":" in line:
                k, v = line.split(":", 1)
                cfg[k.strip()] = int(v)
        self.dimension, self.subvector_dimension, self.num_bits = cfg["dimension"], cfg["subvector_dimension"], cfg["num_bits"]
        self.codebook = np.fromfile(os.path.join(dir, "codebook"), dtype="<f4")
        return self

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.mgpu_pq_destroy(self.handle)
        except Exception:
            pass

    def quantized_dimension(self) -> int:
        return self.dimension // self.subvector_dimension

    def quantize(self, value):
        """Quantizer::quantize; accepts one vector or a batch (n, dim)."""
        single = (not _is_torch(value)) and np.ndim(value) == 1
        x = _Buf(np.asarray(value, dtype=np.float32).reshape(1, -1) if single else value, np.float32, (None, self.dimension))
        n = x.shape[0]
        m = self.quantized_dimension()
        if x.mem == DEVICE:
            import torch
            out = torch.empty((n, m), dtype=torch.uint8, device=value.device)
            optr = out.data_ptr()
        else:
            out = np.empty((n, m), dtype=np.uint8)
            optr = out.ctypes.data
        with self.ctx.order(x.mem):
            _lib.check(self.ctx.lib.mgpu_pq_quantize_batch(self.handle, x.ptr, n, optr, x.mem), self.ctx.h)
        return out[0] if single else out

    def original_vector(self, quantized):
        """pq/mod.rs:184-200 -- the codebook centroids a code word names; one code word (m,) or a batch (n, m)."""
        single = (not _is_torch(quantized)) and np.ndim(quantized) == 1
        q = _Buf(np.asarray(quantized, dtype=np.uint8).reshape(1, -1) if single else quantized, np.uint8, (None, self.quantized_dimension()))
        n = q.shape[0]
        if q.mem == DEVICE:
            import torch
            out = torch.empty((n, self.dimension), dtype=torch.float32, device=quantized.device)
            optr = out.data_ptr()
        else:
            out = np.empty((n, self.dimension), dtype=np.float32)
            optr = out.ctypes.data
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_pq_original_vector(self.handle, q.ptr, n, optr, q.mem), self.ctx.h)
        return out[0] if single else out

    def distance(self, a, b):
        """Quantizer::distance(a, b, StreamingSIMD); a and b are code words (m,) or batches (n, m) -> (n,)."""
        single = np.ndim(a) == 1
        a2 = np.ascontiguousarray(np.asarray(a, dtype=np.uint8).reshape(-1, self.quantized_dimension()))
        b2 = np.ascontiguousarray(np.asarray(b, dtype=np.uint8).reshape(-1, self.quantized_dimension()))
        out = np.empty(a2.shape[0], dtype=np.float32)
        _lib.check(self.ctx.lib.mgpu_pq_distance_batch(self.handle, a2.ctypes.data, b2.ctypes.data, a2.shape[0],
                                                       out.ctypes.data, HOST), self.ctx.h)
        return float(out[0]) if single else out


# ---- results --------------------------------------------------------------------------------------------------------------
class IdWithScore(NamedTuple):
    """rs/index/src/utils.rs:89-93"""
    doc_id: int
    score: float


@dataclass
class SearchResult:
    """rs/index/src/utils.rs:152-155 (stats.num_pages_accessed has no meaning for an HBM-resident index: always 0)"""
    id_with_scores: list
    num_pages_accessed: int = 0


@dataclass
class SearchParams:
    """rs/config/src/search_params.rs:2-34"""
    top_k: int
    ef_construction: int
    record_pages: bool = False
    num_explored_centroids: Optional[int] = None
    centroid_distance_ratio: float = 0.1

    def explored(self) -> int:
        return self.top_k if self.num_explored_centroids is None else self.num_explored_centroids


class Planner:
    """Host mirror of rs/index/src/query/planner.rs as the search path sees it: `planner.plan_with_ids(scanned_ids)`
    (planner.rs:43-60) yields the scanned point ids that satisfy the request's DocumentFilter, and scan_posting_list keeps only
    those (ivf/block_based/index.rs:212-226).  Here the filter is already evaluated to its point-id set; the GPU scan tests
    one bit per row.  Pass one Planner for a whole batch, or a list with one Planner (or None = no filter) per query."""

    def __init__(self, allowed_point_ids=None, bitmap=None):
        if (allowed_point_ids is None) == (bitmap is None):
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "give either the allowed point ids or a ready bitmap")
        self._ids = None if allowed_point_ids is None else np.unique(np.asarray(allowed_point_ids, dtype=np.uint64))
        self._bitmap = None if bitmap is None else np.ascontiguousarray(bitmap, dtype=np.uint32)

    def bitmap(self, num_vectors: int) -> np.ndarray:
        words = (num_vectors + 31) // 32
        if self._bitmap is not None:
            if self._bitmap.size < words:
                raise InvalidArgument(_lib.ERR_INVALID_ARG, "filter bitmap shorter than the index")
            return self._bitmap[:words]
        bits = np.zeros(words, dtype=np.uint32)
        ids = self._ids[self._ids < num_vectors].astype(np.int64)
        np.bitwise_or.at(bits, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
        return bits


def _filter_arg(planner, B: int, num_vectors: int, device_like):
    """-> (pointer, stride_words, keepalive) for the *_filtered entry points; (None, 0, None) when there is no filter."""
    if planner is None:
        return None, 0, None
    words = (num_vectors + 31) // 32
    if isinstance(planner, Planner):
        bits, stride = planner.bitmap(num_vectors), 0
    else:
        if len(planner) != B:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "one planner per query expected")
        bits = np.full((B, words), 0xFFFFFFFF, dtype=np.uint32)
        for b, p in enumerate(planner):
            if p is not None:
                bits[b] = p.bitmap(num_vectors)
        stride = words
    bits = np.ascontiguousarray(bits)
    if device_like is not None:
        import torch
        t = torch.from_numpy(bits.view(np.int32)).to(device_like.device)
        return t.data_ptr(), stride, t
    return bits.ctypes.data, stride, bits


class BatchResult(NamedTuple):
    doc_ids: object   # (B, k, 2) uint64 (numpy) / int64 (torch): (lo, hi) of each u128 doc id
    scores: object    # (B, k) float32
    counts: object    # (B,) number of valid entries per query (UINT32_MAX encodes None for Spann)

    def to_results(self):
        d, s, c = (np.asarray(x.cpu()) if _is_torch(x) else x for x in (self.doc_ids, self.scores, self.counts))
        d = d.view(np.uint64) if d.dtype != np.uint64 else d
        out = []
        for b in range(len(c)):
            n = int(np.uint32(c[b]))
            if n == 0xFFFFFFFF:
                out.append(None)
                continue
            ids = _pairs_to_ints(d[b, :n])
            out.append(SearchResult([IdWithScore(i, float(x)) for i, x in zip(ids, s[b, :n])]))
        return out


def _alloc_out(B, k, device_like, u128=True):
    k = max(int(k), 1)
    if device_like is not None:
        import torch
        ids = torch.zeros((B, k, 2) if u128 else (B, k), dtype=torch.int64 if u128 else torch.int32, device=device_like.device)
        scores = torch.zeros((B, k), dtype=torch.float32, device=device_like.device)
        counts = torch.zeros((B,), dtype=torch.int32, device=device_like.device)
        return ids, scores, counts, ids.data_ptr(), scores.data_ptr(), counts.data_ptr()
    ids = np.zeros((B, k, 2) if u128 else (B, k), dtype=np.uint64 if u128 else np.uint32)
    scores = np.zeros((B, k), dtype=np.float32)
    counts = np.zeros((B,), dtype=np.uint32)
    return ids, scores, counts, ids.ctypes.data, scores.ctypes.data, counts.ctypes.data


# ---- BlockBasedIvf<Q> (rs/index/src/ivf/block_based/index.rs) ---------------------------------------------------------------
class BlockBasedIvf:
    """HBM-resident IVF index.  Inputs are the arrays BlockBasedIvf::new reads from the `index` and `vectors` files
    (ivf/writer.rs:300-353): centroids, posting lists (offsets + ascending point ids), rows by point id, doc ids."""

    def __init__(self, centroids, list_offsets, list_point_ids, rows, quantizer, doc_ids=None, ctx: Optional[Context] = None):
        self.ctx = ctx or getattr(quantizer, "ctx", None) or default_context()
        self.quantizer = quantizer
        cent = np.ascontiguousarray(centroids, dtype=np.float32)
        if cent.ndim != 2:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "centroids must be (nlist, dim)")
        self.nlist, self.dim = cent.shape
        lo = np.ascontiguousarray(list_offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(list_point_ids, dtype=np.uint32)
        if lo.size != self.nlist + 1:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "list_offsets must have nlist + 1 entries")
        rdt = np.uint8 if quantizer.QUANT == QUANT_PQ else np.float32
        r = _Buf(rows, rdt, (None, quantizer.quantized_dimension()))
        n = r.shape[0]
        docs = _ints_to_pairs(doc_ids)
        if docs is not None and docs.shape[0] != n:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "doc_ids must have one entry per row")
        h = C.c_void_p()
        with self.ctx.order(r.mem):
            _lib.check(self.ctx.lib.mgpu_ivf_create(self.ctx.h, self.dim, self.nlist, cent.ctypes.data, lo.ctypes.data,
                                                    ids.ctypes.data if ids.size else None, quantizer.QUANT, quantizer.metric,
                                                    quantizer.handle, r.ptr, r.mem, n,
                                                    docs.ctypes.data if docs is not None else None, C.byref(h)), self.ctx.h)
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.mgpu_ivf_destroy(self.handle)
        except Exception:
            pass

    @classmethod
    def new(cls, base_directory: str, quantizer, index_offset: int = 0, vector_offset: int = 0, ctx: Optional[Context] = None):
        """BlockBasedIvf::new / new_with_offset (index.rs:50-138): open `{base}/index` + `{base}/vectors` written by the
        reference's IvfWriter (ivf/writer.rs:46-353); Elias-Fano posting lists are decoded on the host."""
        self = cls.__new__(cls)
        self.ctx = ctx or getattr(quantizer, "ctx", None) or default_context()
        self.quantizer = quantizer
        h = C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_ivf_load(self.ctx.h, base_directory.encode(), index_offset, vector_offset, quantizer.QUANT,
                                              quantizer.metric, quantizer.handle, C.byref(h)), self.ctx.h)
        self.handle = h
        self.nlist = int(self.ctx.lib.mgpu_ivf_num_clusters(h))
        self.dim = quantizer.dimension
        return self

    def num_clusters(self) -> int:
        return int(self.ctx.lib.mgpu_ivf_num_clusters(self.handle))

    def num_vectors(self) -> int:
        return int(self.ctx.lib.mgpu_ivf_num_vectors(self.handle))

    # -- invalidation.  The reference's methods are keyed by DOC id (index.rs:417-459); the scan tests point ids.
    def invalidate_points(self, point_ids: Sequence[int]):
        """invalid_point_ids.insert for a batch of point ids (what invalidate does after resolving the doc id)."""
        a = np.ascontiguousarray(point_ids, dtype=np.uint32)
        _lib.check(self.ctx.lib.mgpu_ivf_invalidate(self.handle, a.ctypes.data, a.size), self.ctx.h)

    def is_point_invalidated(self, point_id: int) -> bool:
        o = C.c_int()
        _lib.check(self.ctx.lib.mgpu_ivf_is_invalidated(self.handle, point_id, C.byref(o)), self.ctx.h)
        return bool(o.value)

    def invalidate_batch(self, doc_ids) -> list:
        """index.rs:439-452: -> the doc ids that were successfully (newly) invalidated, in input order."""
        d = _ints_to_pairs(list(doc_ids) if not isinstance(doc_ids, np.ndarray) else doc_ids)
        n = d.shape[0]
        ok = np.zeros(max(n, 1), dtype=np.uint8)
        num = C.c_uint32()
        _lib.check(self.ctx.lib.mgpu_ivf_invalidate_docs(self.handle, d.ctypes.data, n, ok.ctypes.data, C.byref(num)), self.ctx.h)
        ints = _pairs_to_ints(d)
        return [ints[i] for i in range(n) if ok[i]]

    def invalidate(self, doc_id: int) -> bool:
        """index.rs:417-429: true if the document was found and newly marked invalid."""
        return len(self.invalidate_batch([int(doc_id)])) == 1

    def is_invalidated(self, doc_id: int) -> bool:
        """index.rs:454-459"""
        d = _ints_to_pairs([int(doc_id)])
        o = C.c_int()
        _lib.check(self.ctx.lib.mgpu_ivf_is_doc_invalidated(self.handle, d.ctypes.data, C.byref(o)), self.ctx.h)
        return bool(o.value)

    # -- accessors (index.rs:350-384,469-471)
    def get_point_id(self, doc_id: int) -> Optional[int]:
        d = _ints_to_pairs([int(doc_id)])
        found, pid = C.c_int(), C.c_uint32()
        _lib.check(self.ctx.lib.mgpu_ivf_get_point_id(self.handle, d.ctypes.data, C.byref(found), C.byref(pid)), self.ctx.h)
        return int(pid.value) if found.value else None

    def get_doc_ids(self, point_ids: Sequence[int]) -> list:
        a = np.ascontiguousarray(point_ids, dtype=np.uint32)
        out = np.zeros((max(a.size, 1), 2), dtype=np.uint64)
        _lib.check(self.ctx.lib.mgpu_ivf_get_doc_ids(self.handle, a.ctypes.data, a.size, out.ctypes.data), self.ctx.h)
        return _pairs_to_ints(out[:a.size])

    def get_doc_id(self, point_id: int) -> int:
        return self.get_doc_ids([point_id])[0]

    def get_vectors(self, point_ids: Sequence[int]) -> np.ndarray:
        a = np.ascontiguousarray(point_ids, dtype=np.uint32)
        qd = self.quantizer.quantized_dimension()
        out = np.zeros((max(a.size, 1), qd), dtype=np.uint8 if self.quantizer.QUANT == QUANT_PQ else np.float32)
        _lib.check(self.ctx.lib.mgpu_ivf_get_vectors(self.handle, a.ctypes.data, a.size, out.ctypes.data), self.ctx.h)
        return out[:a.size]

    def get_vector(self, point_id: int) -> np.ndarray:
        """index.rs:372-384: the stored (quantized) row of a point."""
        return self.get_vectors([point_id])[0]

    # -- batched entry points
    def find_nearest_centroids_batch(self, Q, num_probes: int, with_distances=False):
        q = _Buf(Q, np.float32, (None, self.dim))
        B = q.shape[0]
        p = max(num_probes, 1)
        if q.mem == DEVICE:
            import torch
            ids = torch.zeros((B, p), dtype=torch.int32, device=Q.device)
            ds = torch.zeros((B, p), dtype=torch.float32, device=Q.device)
            ip, dp = ids.data_ptr(), ds.data_ptr()
        else:
            ids = np.zeros((B, p), dtype=np.uint32)
            ds = np.zeros((B, p), dtype=np.float32)
            ip, dp = ids.ctypes.data, ds.ctypes.data
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_ivf_coarse(self.handle, q.ptr, B, num_probes, ip, dp, q.mem), self.ctx.h)
        return (ids, ds) if with_distances else ids

    def search_with_centroids_batch(self, Q, centroid_ids, k: int, counts=None, remap=True, planner=None):
        """search_with_centroids(_and_remap) for a batch; centroid_ids (B, P), counts (B,) optional."""
        q = _Buf(Q, np.float32, (None, self.dim))
        B = q.shape[0]
        pr = _Buf(centroid_ids, np.uint32, (B, None))
        pc = _Buf(counts, np.uint32, (B,), allow_none=True)
        if pr.mem != q.mem or (pc.mem is not None and pc.mem != q.mem):
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "queries and probe lists must live in the same memory space")
        P = pr.shape[1]
        ids, scores, cnt, ip, sp, cp = _alloc_out(B, k, Q if q.mem == DEVICE else None, u128=remap)
        if planner is not None:
            if not remap:
                raise Unsupported(_lib.ERR_UNSUPPORTED, "the planner filter is only wired into the remapping search (index.rs:298-332)")
            fp, fstride, _keep = _filter_arg(planner, B, self.num_vectors(), Q if q.mem == DEVICE else None)
            with self.ctx.order(q.mem):
                _lib.check(self.ctx.lib.mgpu_ivf_scan_remap_filtered(self.handle, q.ptr, B, pr.ptr, P, pc.ptr, k, fp, fstride, ip, sp, cp,
                                                                     q.mem), self.ctx.h)
            if _keep is not None and q.mem == DEVICE:
                self.ctx.sync()  # the bitmap tensor must outlive the asynchronous scan
            return BatchResult(ids, scores, cnt)
        f = self.ctx.lib.mgpu_ivf_scan_remap if remap else self.ctx.lib.mgpu_ivf_scan
        with self.ctx.order(q.mem):
            _lib.check(f(self.handle, q.ptr, B, pr.ptr, P, pc.ptr, k, ip, sp, cp, q.mem), self.ctx.h)
        return BatchResult(ids, scores, cnt)

    def search_batch(self, Q, k: int, num_probes: int, out=None, planner=None) -> BatchResult:
        """BlockBasedIvf::search for a batch of queries: coarse scoring + list scan + remap, all on the GPU.
        planner: Option<Arc<Planner>> of index.rs:396-403 (one for the batch or one per query)."""
        q = _Buf(Q, np.float32, (None, self.dim))
        B = q.shape[0]
        if out is None:
            ids, scores, cnt, ip, sp, cp = _alloc_out(B, k, Q if q.mem == DEVICE else None)
        else:
            ids, scores, cnt = out
            ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        if planner is not None:
            fp, fstride, _keep = _filter_arg(planner, B, self.num_vectors(), Q if q.mem == DEVICE else None)
            with self.ctx.order(q.mem):
                _lib.check(self.ctx.lib.mgpu_ivf_search_filtered(self.handle, q.ptr, B, k, num_probes, fp, fstride, ip, sp, cp, q.mem),
                           self.ctx.h)
            if _keep is not None and q.mem == DEVICE:
                self.ctx.sync()  # the bitmap tensor must outlive the asynchronous scan
            return BatchResult(ids, scores, cnt)
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_ivf_search(self.handle, q.ptr, B, k, num_probes, ip, sp, cp, q.mem), self.ctx.h)
        return BatchResult(ids, scores, cnt)

    def shard_search_batch(self, Q, k: int, num_probes: int, out=None, shared_codebook: bool = True) -> BatchResult:
        """Sharded BlockBasedIvf::search (mgpu_shard_ivf_search): this rank's shard is searched for the replicated batch
        Q, the per-shard top-k are all-gathered and merged (snapshot.rs:60-61); collective over the context's ranks."""
        q = _Buf(Q, np.float32, (None, self.dim))
        B = q.shape[0]
        if out is None:
            ids, scores, cnt, ip, sp, cp = _alloc_out(B, k, Q if q.mem == DEVICE else None)
        else:
            ids, scores, cnt = out
            ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_shard_ivf_search(self.handle, q.ptr, B, k, num_probes, 1 if shared_codebook else 0, ip, sp, cp,
                                                          q.mem), self.ctx.h)
        return BatchResult(ids, scores, cnt)

    def shard_search_batch_submit(self, Q, k: int, num_probes: int, out, shared_codebook: bool = True) -> int:
        """Pipelined shard_search_batch over page-locked HOST buffers (mgpu_shard_ivf_search_submit); `search_wait(ticket)`
        completes it.  Collective: every rank submits the same batches in the same order."""
        q = _Buf(Q, np.float32, (None, self.dim))
        if q.mem != HOST:
            raise ValueError("shard_search_batch_submit takes host buffers")
        ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        t = C.c_uint64(0)
        _lib.check(self.ctx.lib.mgpu_shard_ivf_search_submit(self.handle, q.ptr, q.shape[0], k, num_probes, 1 if shared_codebook else 0,
                                                             ip, sp, cp, C.byref(t)), self.ctx.h)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (Q, out)
        return int(t.value)

    def search_batch_submit(self, Q, k: int, num_probes: int, out) -> int:
        """Pipelined search_batch over page-locked HOST buffers (mgpu_ivf_search_submit): returns a ticket at once; `out`
        = (ids, scores, counts) is valid after `search_wait(ticket)`.  Two batches may be in flight, so the next batch's
        upload and the previous one's download overlap the kernels."""
        q = _Buf(Q, np.float32, (None, self.dim))
        if q.mem != HOST:
            raise ValueError("search_batch_submit takes host buffers (device buffers are asynchronous already)")
        ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        t = C.c_uint64(0)
        _lib.check(self.ctx.lib.mgpu_ivf_search_submit(self.handle, q.ptr, q.shape[0], k, num_probes, ip, sp, cp, C.byref(t)),
                   self.ctx.h)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (Q, out)  # keep the buffers alive until the wait
        return int(t.value)

    def search_wait(self, ticket: int) -> None:
        _lib.check(self.ctx.lib.mgpu_search_wait(self.ctx.h, ticket), self.ctx.h)
        getattr(self, "_inflight", {}).pop(ticket, None)

    def last_scan_rows(self) -> int:
        return int(self.ctx.lib.mgpu_ivf_last_scan_rows(self.handle))

    def last_scan_fallbacks(self) -> int:
        return int(self.ctx.lib.mgpu_ivf_last_scan_fallbacks(self.handle))

    def last_scan_bytes(self) -> int:
        return int(self.ctx.lib.mgpu_ivf_last_scan_bytes(self.handle))

    # -- the reference's single-query methods
    def find_nearest_centroids(self, vector, num_probes: int) -> list:
        """index.rs:147-163"""
        ids = self.find_nearest_centroids_batch(np.asarray(vector, dtype=np.float32).reshape(1, -1), num_probes)
        return [int(x) for x in ids[0]]

    def search_with_centroids_and_remap(self, query, nearest_centroid_ids, k: int, planner: Optional[Planner] = None) -> SearchResult:
        """index.rs:298-332"""
        c = np.asarray(nearest_centroid_ids, dtype=np.uint32).reshape(1, -1)
        if c.size == 0:
            return SearchResult([])
        r = self.search_with_centroids_batch(np.asarray(query, dtype=np.float32).reshape(1, -1), c, k, planner=planner)
        return r.to_results()[0]

    def search(self, query, k: int, num_probes: int, planner: Optional[Planner] = None) -> Optional[SearchResult]:
        """index.rs:396-412"""
        return self.search_batch(np.asarray(query, dtype=np.float32).reshape(1, -1), k, num_probes, planner=planner).to_results()[0]


# ---- BlockBasedHnsw<Q> (rs/index/src/hnsw/block_based/index.rs) --------------------------------------------------------------
class BlockBasedHnsw:
    """HBM-resident HNSW graph in the array layout of the `hnsw/index` file (graph_storage.rs:122-193)."""

    def __init__(self, num_layers, edges, points, edge_offsets, level_offsets, rows, quantizer, doc_ids=None,
                 ctx: Optional[Context] = None):
        self.ctx = ctx or getattr(quantizer, "ctx", None) or default_context()
        self.quantizer = quantizer
        e = np.ascontiguousarray(edges, dtype=np.uint32)
        p = np.ascontiguousarray(points, dtype=np.uint32)
        eo = np.ascontiguousarray(edge_offsets, dtype=np.uint64)
        lo = np.ascontiguousarray(level_offsets, dtype=np.uint64)
        if lo.size != num_layers + 1:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "level_offsets must have num_layers + 1 entries")
        rdt = np.uint8 if quantizer.QUANT == QUANT_PQ else np.float32
        r = _Buf(rows, rdt, (None, quantizer.quantized_dimension()))
        n = r.shape[0]
        self.dim = quantizer.dimension
        docs = _ints_to_pairs(doc_ids)
        h = C.c_void_p()
        with self.ctx.order(r.mem):
            _lib.check(self.ctx.lib.mgpu_hnsw_create(self.ctx.h, self.dim, int(num_layers), e.ctypes.data if e.size else None,
                                                     e.size, p.ctypes.data if p.size else None, p.size, eo.ctypes.data, eo.size,
                                                     lo.ctypes.data, quantizer.QUANT, quantizer.metric, quantizer.handle, r.ptr,
                                                     r.mem, n, docs.ctypes.data if docs is not None else None, C.byref(h)),
                       self.ctx.h)
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.mgpu_hnsw_destroy(self.handle)
        except Exception:
            pass

    @classmethod
    def new(cls, base_directory: str, quantizer, index_offset: int = 0, vector_offset: int = 0, ctx: Optional[Context] = None):
        """BlockBasedHnsw::new / new_with_offsets (hnsw/block_based/index.rs:69-140): `{base}/hnsw/index` +
        `{base}/hnsw/vector_storage` written by the reference's HnswWriter (hnsw/writer.rs:43-265)."""
        self = cls.__new__(cls)
        self.ctx = ctx or getattr(quantizer, "ctx", None) or default_context()
        self.quantizer = quantizer
        self.dim = quantizer.dimension
        h = C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_hnsw_load(self.ctx.h, base_directory.encode(), index_offset, vector_offset, self.dim,
                                               quantizer.QUANT, quantizer.metric, quantizer.handle, C.byref(h)), self.ctx.h)
        self.handle = h
        return self

    def graph_arrays(self):
        """The resident graph sections (edges, points, edge_offsets, level_offsets) + entry point, read back from HBM."""
        sz = np.zeros(6, dtype=np.uint64)
        _lib.check(self.ctx.lib.mgpu_hnsw_info(self.handle, sz.ctypes.data), self.ctx.h)
        nl, ne, npnt, neo, n, ep = (int(v) for v in sz)
        edges = np.zeros(max(ne, 1), dtype=np.uint32)
        points = np.zeros(max(npnt, 1), dtype=np.uint32)
        eo = np.zeros(neo, dtype=np.uint64)
        lo = np.zeros(nl + 1, dtype=np.uint64)
        _lib.check(self.ctx.lib.mgpu_hnsw_copy_graph(self.handle, edges.ctypes.data, points.ctypes.data, eo.ctypes.data,
                                                     lo.ctypes.data), self.ctx.h)
        return dict(num_layers=nl, edges=edges[:ne], points=points[:npnt], edge_offsets=eo, level_offsets=lo, n=n, entry_point=ep)

    def ann_search_batch(self, Q, k: int, ef: int, with_stats=False):
        q = _Buf(Q, np.float32, (None, self.dim))
        B = q.shape[0]
        ids, scores, cnt, ip, sp, cp = _alloc_out(B, k, Q if q.mem == DEVICE else None)
        stats, stp = None, None
        if with_stats:
            if q.mem == DEVICE:
                import torch
                stats = torch.zeros((B, 2), dtype=torch.int64, device=Q.device)
                stp = stats.data_ptr()
            else:
                stats = np.zeros((B, 2), dtype=np.uint64)
                stp = stats.ctypes.data
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_hnsw_search(self.handle, q.ptr, B, k, ef, ip, sp, cp, stp, q.mem), self.ctx.h)
        r = BatchResult(ids, scores, cnt)
        return (r, stats) if with_stats else r

    def ann_search_batch_submit(self, Q, k: int, ef: int, out) -> int:
        """Pipelined ann_search_batch over page-locked HOST buffers (mgpu_hnsw_search_submit); `search_wait(ticket)` completes
        it (ticket 0: nothing was in flight, `out` is already valid)."""
        q = _Buf(Q, np.float32, (None, self.dim))
        if q.mem != HOST:
            raise ValueError("ann_search_batch_submit takes host buffers (device buffers are asynchronous already)")
        ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        t = C.c_uint64(0)
        _lib.check(self.ctx.lib.mgpu_hnsw_search_submit(self.handle, q.ptr, q.shape[0], k, ef, ip, sp, cp, C.byref(t)), self.ctx.h)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (Q, out)  # keep the buffers alive until the wait
        return int(t.value)

    def search_wait(self, ticket: int) -> None:
        if ticket:
            _lib.check(self.ctx.lib.mgpu_search_wait(self.ctx.h, ticket), self.ctx.h)
        getattr(self, "_inflight", {}).pop(ticket, None)

    def ann_search(self, query, k: int, ef: int) -> SearchResult:
        """index.rs:159-210"""
        return self.ann_search_batch(np.asarray(query, dtype=np.float32).reshape(1, -1), k, ef).to_results()[0]


# ---- Spann<Q> (rs/index/src/spann/index.rs) ---------------------------------------------------------------------------------
class Spann:
    def __init__(self, centroids: BlockBasedHnsw, posting_lists: BlockBasedIvf):
        self.ctx = posting_lists.ctx
        self.centroids, self.posting_lists = centroids, posting_lists
        h = C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_spann_create(self.ctx.h, centroids.handle, posting_lists.handle, C.byref(h)), self.ctx.h)
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.mgpu_spann_destroy(self.handle)
        except Exception:
            pass

    @classmethod
    def read(cls, base_directory: str, quantizer, ctx: Optional[Context] = None):
        """SpannReader::read (spann/reader.rs:75-76): `{base}/centroids` is an HNSW over the centroids with a NoQuantizer,
        `{base}/ivf` the posting lists with quantizer Q."""
        import os
        ctx = ctx or getattr(quantizer, "ctx", None) or default_context()
        centroids = BlockBasedHnsw.new(os.path.join(base_directory, "centroids"), NoQuantizer(quantizer.dimension), ctx=ctx)
        lists = BlockBasedIvf.new(os.path.join(base_directory, "ivf"), quantizer, ctx=ctx)
        return cls(centroids, lists)

    def search_batch(self, Q, params: SearchParams, planner=None) -> BatchResult:
        q = _Buf(Q, np.float32, (None, self.posting_lists.dim))
        B = q.shape[0]
        ids, scores, cnt, ip, sp, cp = _alloc_out(B, params.top_k, Q if q.mem == DEVICE else None)
        if planner is not None:
            fp, fstride, _keep = _filter_arg(planner, B, self.posting_lists.num_vectors(), Q if q.mem == DEVICE else None)
            with self.ctx.order(q.mem):
                _lib.check(self.ctx.lib.mgpu_spann_search_filtered(self.handle, q.ptr, B, params.top_k, params.ef_construction,
                                                                   params.explored(), float(params.centroid_distance_ratio), fp,
                                                                   fstride, ip, sp, cp, q.mem), self.ctx.h)
            if _keep is not None and q.mem == DEVICE:
                self.ctx.sync()
            return BatchResult(ids, scores, cnt)
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_spann_search(self.handle, q.ptr, B, params.top_k, params.ef_construction,
                                                      params.explored(), float(params.centroid_distance_ratio), ip, sp, cp,
                                                      q.mem), self.ctx.h)
        return BatchResult(ids, scores, cnt)

    def shard_search_batch(self, Q, params: SearchParams, out=None, shared_codebook: bool = True) -> BatchResult:
        """Config 5 (mgpu_shard_spann_search): Spann::search on this rank's doc-shard for the replicated batch Q, then the
        all-gather + (score, doc_id) merge over the context's ranks.  Collective."""
        q = _Buf(Q, np.float32, (None, self.posting_lists.dim))
        B = q.shape[0]
        if out is None:
            ids, scores, cnt, ip, sp, cp = _alloc_out(B, params.top_k, Q if q.mem == DEVICE else None)
        else:
            ids, scores, cnt = out
            ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        with self.ctx.order(q.mem):
            _lib.check(self.ctx.lib.mgpu_shard_spann_search(self.handle, q.ptr, B, params.top_k, params.ef_construction, params.explored(),
                                                            float(params.centroid_distance_ratio), 1 if shared_codebook else 0, ip, sp, cp,
                                                            q.mem), self.ctx.h)
        return BatchResult(ids, scores, cnt)

    def search_batch_submit(self, Q, params: SearchParams, out) -> int:
        """Pipelined search_batch over page-locked HOST buffers (mgpu_spann_search_submit); `search_wait(ticket)` completes it."""
        q = _Buf(Q, np.float32, (None, self.posting_lists.dim))
        if q.mem != HOST:
            raise ValueError("search_batch_submit takes host buffers (device buffers are asynchronous already)")
        ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        t = C.c_uint64(0)
        _lib.check(self.ctx.lib.mgpu_spann_search_submit(self.handle, q.ptr, q.shape[0], params.top_k, params.ef_construction,
                                                         params.explored(), float(params.centroid_distance_ratio), ip, sp, cp,
                                                         C.byref(t)), self.ctx.h)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (Q, out)
        return int(t.value)

    def shard_search_batch_submit(self, Q, params: SearchParams, out, shared_codebook: bool = True) -> int:
        """Pipelined form over page-locked HOST buffers; `search_wait(ticket)` completes it."""
        q = _Buf(Q, np.float32, (None, self.posting_lists.dim))
        if q.mem != HOST:
            raise ValueError("shard_search_batch_submit takes host buffers")
        ip, sp, cp = (x.data_ptr() if _is_torch(x) else x.ctypes.data for x in out)
        t = C.c_uint64(0)
        _lib.check(self.ctx.lib.mgpu_shard_spann_search_submit(self.handle, q.ptr, q.shape[0], params.top_k, params.ef_construction,
                                                               params.explored(), float(params.centroid_distance_ratio),
                                                               1 if shared_codebook else 0, ip, sp, cp, C.byref(t)), self.ctx.h)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (Q, out)
        return int(t.value)

    def search_wait(self, ticket: int) -> None:
        _lib.check(self.ctx.lib.mgpu_search_wait(self.ctx.h, ticket), self.ctx.h)
        getattr(self, "_inflight", {}).pop(ticket, None)

    def invalidate(self, doc_id: int) -> bool:
        """spann/index.rs: Spann::invalidate forwards to the posting lists."""
        return self.posting_lists.invalidate(doc_id)

    def invalidate_batch(self, doc_ids) -> list:
        return self.posting_lists.invalidate_batch(doc_ids)

    def is_invalidated(self, doc_id: int) -> bool:
        return self.posting_lists.is_invalidated(doc_id)

    def search(self, query, params: SearchParams, planner: Optional[Planner] = None) -> Optional[SearchResult]:
        """spann/index.rs:211-266"""
        return self.search_batch(np.asarray(query, dtype=np.float32).reshape(1, -1), params, planner=planner).to_results()[0]


# ---- multi-user SPANN files (rs/index/src/multi_spann) ------------------------------------------------------------------------
class _UserIndexInfoC(C.Structure):
    _fields_ = [("user_lo", C.c_uint64), ("user_hi", C.c_uint64)] + [(n, C.c_uint64) for n in (
        "centroid_vector_offset", "centroid_vector_len", "centroid_index_offset", "centroid_index_len", "ivf_vectors_offset",
        "ivf_vectors_len", "ivf_raw_vectors_offset", "ivf_raw_vectors_len", "ivf_index_offset", "ivf_index_len",
        "ivf_pq_codebook_offset", "ivf_pq_codebook_len")]


@dataclass
class UserIndexInfo:
    """rs/index/src/multi_spann/user_index_info.rs:4-18"""
    user_id: int
    centroid_vector_offset: int = 0
    centroid_vector_len: int = 0
    centroid_index_offset: int = 0
    centroid_index_len: int = 0
    ivf_vectors_offset: int = 0
    ivf_vectors_len: int = 0
    ivf_raw_vectors_offset: int = 0
    ivf_raw_vectors_len: int = 0
    ivf_index_offset: int = 0
    ivf_index_len: int = 0
    ivf_pq_codebook_offset: int = 0
    ivf_pq_codebook_len: int = 0

    def _c(self) -> _UserIndexInfoC:
        c = _UserIndexInfoC()
        c.user_lo, c.user_hi = self.user_id & 0xFFFFFFFFFFFFFFFF, self.user_id >> 64
        for n, _ in _UserIndexInfoC._fields_[2:]:
            setattr(c, n, getattr(self, n))
        return c

    @classmethod
    def _from_c(cls, c: _UserIndexInfoC):
        return cls(int(c.user_lo) | (int(c.user_hi) << 64), **{n: int(getattr(c, n)) for n, _ in _UserIndexInfoC._fields_[2:]})

    def to_le_bytes(self) -> bytes:
        """user_index_info.rs:26-42"""
        buf = (C.c_uint8 * 112)()
        c = self._c()
        _lib.check(_lib.load().mgpu_user_index_info_encode(C.addressof(c), C.addressof(buf)))
        return bytes(buf)

    @classmethod
    def from_le_bytes(cls, data: bytes):
        """user_index_info.rs:59-83"""
        buf = (C.c_uint8 * 112).from_buffer_copy(data[:112])
        c = _UserIndexInfoC()
        _lib.check(_lib.load().mgpu_user_index_info_decode(C.addressof(buf), C.addressof(c)))
        return cls._from_c(c)

    @classmethod
    def read_table(cls, path: str) -> dict:
        """All records of a `user_index_info` file (odht table image) -> {user_id: UserIndexInfo}."""
        lib = _lib.load()
        n = lib.mgpu_user_index_info_read(path.encode(), None, 0)
        if n < 0:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, f"{path}: not a user_index_info table")
        arr = (_UserIndexInfoC * max(n, 1))()
        lib.mgpu_user_index_info_read(path.encode(), C.addressof(arr), n)
        out = {}
        for i in range(n):
            u = cls._from_c(arr[i])
            out[u.user_id] = u
        return out


class MultiSpannIndex:
    """rs/index/src/multi_spann/index.rs: one Spann per user inside shared files, opened lazily by user id
    (get_or_create_index :100-128) at the byte offsets of `{base}/user_index_info`."""

    def __init__(self, base_directory: str, num_features: int, quantizer_kind: int = QUANT_NONE, distance=L2DistanceCalculator,
                 ctx: Optional[Context] = None):
        import os
        self.base, self.dim, self.quant, self.metric = base_directory, num_features, quantizer_kind, distance.METRIC
        self.ctx = ctx or default_context()
        self.user_index_infos = UserIndexInfo.read_table(os.path.join(base_directory, "user_index_info"))
        self.user_to_spann = {}

    def user_ids(self) -> list:
        return sorted(self.user_index_infos)

    def get_or_create_index(self, user_id: int) -> "Spann":
        if user_id in self.user_to_spann:
            return self.user_to_spann[user_id]
        info = self.user_index_infos.get(user_id)
        if info is None:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "User not found")   # index.rs:108
        c = info._c()
        pq, hn, iv, sp = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.ctx.lib.mgpu_spann_load_user(self.ctx.h, self.base.encode(), C.addressof(c), self.dim, self.quant, self.metric,
                                                     C.byref(pq), C.byref(hn), C.byref(iv), C.byref(sp)), self.ctx.h)
        s = Spann.__new__(Spann)
        s.ctx, s.handle = self.ctx, sp
        noq = NoQuantizer(self.dim)
        hnsw = BlockBasedHnsw.__new__(BlockBasedHnsw)
        hnsw.ctx, hnsw.quantizer, hnsw.dim, hnsw.handle = self.ctx, noq, self.dim, hn
        ivf = BlockBasedIvf.__new__(BlockBasedIvf)
        ivf.ctx, ivf.handle, ivf.dim = self.ctx, iv, self.dim
        ivf.nlist = int(self.ctx.lib.mgpu_ivf_num_clusters(iv))
        if self.quant == QUANT_PQ:
            q = ProductQuantizer.__new__(ProductQuantizer)
            q.ctx, q.handle, q.metric, q.dimension = self.ctx, pq, self.metric, self.dim
            import os
            cfg = {}
            for line in open(os.path.join(self.base, "ivf", "quantizer", "product_quantizer_config.yaml")):
                if ":" in line:
                    k, v = line.split(":", 1)
                    cfg[k.strip()] = int(v)
            q.subvector_dimension, q.num_bits = cfg["subvector_dimension"], cfg["num_bits"]
            q.codebook = None
            ivf.quantizer = q
        else:
            ivf.quantizer = noq
        s.centroids, s.posting_lists = hnsw, ivf
        self.user_to_spann[user_id] = s
        return s

    def search_for_user(self, user_id: int, query, params: SearchParams, planner: Optional[Planner] = None) -> Optional[SearchResult]:
        """multi_spann/index.rs: search_for_user -> Spann::search of that user's index (None for an unknown user)."""
        if user_id not in self.user_index_infos:
            return None
        return self.get_or_create_index(user_id).search(query, params, planner)


# ---- micro-batcher ------------------------------------------------------------------------------------------------------------
class MicroBatcher:
    """Per-request front door: `search(query)` has the reference's single-query shape
    (BlockBasedIvf::search index.rs:396-412 / Spann::search spann/index.rs:211-266) and may be called from many threads;
    a native worker thread groups concurrent calls into one batched GPU search (ctypes releases the GIL while a call waits)."""

    def __init__(self, index, k: int, num_probes: int = 0, params: Optional[SearchParams] = None, max_batch: int = 1024,
                 max_wait_us: int = 200):
        self.index, self.ctx = index, index.ctx
        h = C.c_void_p()
        if isinstance(index, Spann):
            p = params or SearchParams(k, 100)
            self.k = p.top_k
            self.dim = index.posting_lists.dim
            self.num_vectors = index.posting_lists.num_vectors()
            _lib.check(self.ctx.lib.mgpu_batcher_create_spann(index.handle, max_batch, max_wait_us, p.top_k, p.ef_construction,
                                                              p.explored(), float(p.centroid_distance_ratio), C.byref(h)), self.ctx.h)
        else:
            self.k, self.dim, self.num_vectors = k, index.dim, index.num_vectors()
            _lib.check(self.ctx.lib.mgpu_batcher_create(index.handle, max_batch, max_wait_us, k, num_probes, C.byref(h)), self.ctx.h)
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.mgpu_batcher_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, query, planner: Optional[Planner] = None) -> Optional[SearchResult]:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        if q.size != self.dim:
            raise InvalidArgument(_lib.ERR_INVALID_ARG, "query has the wrong dimension")
        ids = np.zeros((1, self.k, 2), dtype=np.uint64)
        scores = np.zeros((1, self.k), dtype=np.float32)
        cnt = np.zeros(1, dtype=np.uint32)
        if planner is None:
            st = self.ctx.lib.mgpu_batcher_search(self.handle, q.ctypes.data, ids.ctypes.data, scores.ctypes.data, cnt.ctypes.data)
        else:
            bits = np.ascontiguousarray(planner.bitmap(self.num_vectors))
            st = self.ctx.lib.mgpu_batcher_search_filtered(self.handle, q.ctypes.data, bits.ctypes.data, ids.ctypes.data,
                                                           scores.ctypes.data, cnt.ctypes.data)
        _lib.check(st, self.ctx.h)
        return BatchResult(ids, scores, cnt).to_results()[0]

    def stats(self) -> dict:
        s = (C.c_uint64 * 4)()
        _lib.check(self.ctx.lib.mgpu_batcher_stats(self.handle, C.addressof(s)), self.ctx.h)
        return {"queries": s[0], "batches": s[1], "largest_batch": s[2], "full_batches": s[3]}


# ---- merge / assignment ---------------------------------------------------------------------------------------------------
def elias_fano_decode(payload: bytes) -> np.ndarray:
    """Decode one Elias-Fano posting list as written by EliasFano::write (rs/compression/src/elias_fano/ef.rs:197-215)."""
    buf = np.frombuffer(payload, dtype=np.uint8)
    if buf.size < 32:
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "payload shorter than the Elias-Fano header")
    n = int(np.frombuffer(payload[:8], dtype="<u8")[0])
    out = np.zeros(max(n, 1), dtype=np.uint64)
    r = _lib.load().mgpu_ef_decode(buf.ctypes.data, buf.size, out.ctypes.data, n)
    if r < 0:
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "malformed Elias-Fano payload")
    return out[:n]


def merge_topk(doc_ids, scores, counts, k: int, ctx: Optional[Context] = None) -> BatchResult:
    """Snapshot merge (collection/snapshot.rs:49-63,79-108): doc_ids (S,B,k,2), scores (S,B,k), counts (S,B)."""
    ctx = ctx or default_context()
    d, s, c = _Buf(doc_ids, np.uint64), _Buf(scores, np.float32), _Buf(counts, np.uint32)
    S, B = s.shape[0], s.shape[1]
    kk = s.shape[2]
    if kk != k:
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "partial results must have stride k")
    ids, sc, cnt, ip, sp, cp = _alloc_out(B, k, doc_ids if d.mem == DEVICE else None)
    with ctx.order(d.mem):
        _lib.check(ctx.lib.mgpu_merge_topk(ctx.h, d.ptr, s.ptr, c.ptr, S, B, k, ip, sp, cp, d.mem), ctx.h)
    return BatchResult(ids, sc, cnt)


def assign_to_centroids(X, centroids, max_clusters_per_vector=1, distance_threshold=0.1, ctx: Optional[Context] = None):
    """IvfBuilder::find_nearest_centroids + acceptance rule (ivf/builder.rs:268-329) for a batch of vectors.
    -> (cids (n, max_clusters) padded with UINT32_MAX, counts (n,))"""
    ctx = ctx or default_context()
    x, c = _Buf(X, np.float32), _Buf(centroids, np.float32)
    if x.mem != c.mem:
        raise InvalidArgument(_lib.ERR_INVALID_ARG, "X and centroids must live in the same memory space")
    n, dim = x.shape
    r = max_clusters_per_vector
    if x.mem == DEVICE:
        import torch
        cids = torch.zeros((n, r), dtype=torch.int32, device=X.device)
        cnt = torch.zeros((n,), dtype=torch.int32, device=X.device)
        ip, cp = cids.data_ptr(), cnt.data_ptr()
    else:
        cids = np.zeros((n, r), dtype=np.uint32)
        cnt = np.zeros((n,), dtype=np.uint32)
        ip, cp = cids.ctypes.data, cnt.ctypes.data
    with ctx.order(x.mem):
        _lib.check(ctx.lib.mgpu_ivf_assign(ctx.h, x.ptr, n, c.ptr, c.shape[0], dim, r, float(distance_threshold), ip, cp, x.mem), ctx.h)
    return cids, cnt

"""Doc-sharded multi-GPU search: one process per GPU, each rank holds one shard (its own coarse centroids, posting
lists and doc-id table); queries are replicated; every rank returns the merged top-k.

Semantic model: rs/aggregator/src/aggregator.rs:81-132 (independent shards "{index}--{shard_id}", results unioned) with the
leaf ordering of rs/index/src/collection/snapshot.rs:49-63 (sort by (score, doc_id), truncate to k) -- the aggregator's
descending, untruncated sort (aggregator.rs:135) is NOT followed.

Exchange step: one NCCL all-gather of the per-shard (doc_id, score, count) blocks over NVLink followed by the merge kernel
(mgpu_shard_allgather_merge).  torch.distributed is only the rendezvous (it ships the NCCL unique id).
"""
from __future__ import annotations

import numpy as np


def shard_of(doc_ids_lo: np.ndarray, world: int) -> np.ndarray:
    """Shard assignment: doc_id mod world (on the low 64 bits; ids are u128 (lo, hi) pairs)."""
    return (np.asarray(doc_ids_lo, dtype=np.uint64) % np.uint64(world)).astype(np.int64)


def split_probes(nprobe: int, nlist: int, world: int):
    """Per-shard (nlist, nprobe) keeping the probed fraction of the collection constant."""
    return max(nlist // world, 1), max(nprobe // world, 1)


def encode_slice(B: int, world: int, rank: int):
    """Queries rank `rank` encodes when the codebook is shared (mgpu_shard_ivf_search, api.cu): (first, count).  Every rank
    all-gathers ceil(B / world) code rows, so rank r's slice starts at r * ceil(B / world); the last slices may be short or
    empty when B is not a multiple of the world size."""
    per = (B + world - 1) // world
    lo = min(B, rank * per)
    return lo, min(B, lo + per) - lo


def init_comm(ctx, dist=None):
    """Create the NCCL communicator of `ctx` across the torch.distributed world (any backend is fine for the
    rendezvous)."""
    import torch
    import torch.distributed as dist_mod
    dist = dist or dist_mod
    world, rank = dist.get_world_size(), dist.get_rank()
    from . import Context
    obj = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx.comm_init(world, rank, obj[0])
    return world, rank

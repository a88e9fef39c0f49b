"""N > 1 host logic on CPU: world_size-2 gloo processes, each owning one doc-shard (doc_id mod N), exchange their per-shard
top-k and merge by (score, doc_id) -- the semantics the NCCL all-gather + merge kernel implements on GPUs
(muopdb_b200/sharding.py).  The per-shard search here is the CPU oracle (test infrastructure)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from muopdb_b200 import sharding
from tests import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = synth.clustered(1200, 24, n_blobs=6, seed=5)
        docs = synth.doc_ids_for(len(X), seed=6, wide=True)
        Q = X[:16] + 0.01
        k = 7
        mine = sharding.shard_of(docs[:, 0], world) == rank
        Xs, ds = X[mine], docs[mine]
        nlist, nprobe = sharding.split_probes(8, 8, world)  # probe every list of the shard => exact per-shard top-k
        cents = O.kmeans(Xs, nlist, 5, seed=rank + 1)
        offs, ids = O.build_posting_lists(Xs, cents)
        ivf = O.Ivf(cents, offs, ids, Xs, doc_ids=ds)
        od, os_, oc = ivf.search_batch(Q, k, nprobe)
        # rendezvous plumbing used by sharding.init_comm: rank 0's 128-byte id reaches every rank
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        assert obj[0] == bytes(range(128))
        gathered = [None] * world
        dist.all_gather_object(gathered, (od, os_, oc))
        merged = []
        for b in range(len(Q)):
            dl, sl = [], []
            for gd, gs, gc in gathered:
                dl += [int(lo) | (int(hi) << 64) for lo, hi in gd[b, :gc[b]]]
                sl += gs[b, :gc[b]].tolist()
            merged.append(O.merge_topk(dl, sl, k))
        if rank == 0:
            ret["merged"] = [(m[0], m[1].tolist()) for m in merged]
    finally:
        dist.destroy_process_group()


def test_two_rank_doc_sharding_equals_global_topk():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    X = synth.clustered(1200, 24, n_blobs=6, seed=5)
    docs = synth.doc_ids_for(len(X), seed=6, wide=True)
    Q = X[:16] + 0.01
    ids128 = [int(lo) | (int(hi) << 64) for lo, hi in docs]
    for b, (got_ids, got_scores) in enumerate(ret["merged"]):
        ref = sorted((O.l2(Q[b], X[i]), ids128[i]) for i in range(len(X)))[:7]
        assert got_ids == [d for _, d in ref]
        assert got_scores == [float(np.float32(s)) for s, _ in ref]


def test_shard_helpers():
    assert sharding.split_probes(64, 4096, 8) == (512, 8)
    assert sharding.split_probes(64, 4096, 1) == (4096, 64)
    lo = np.arange(10, dtype=np.uint64)
    assert sharding.shard_of(lo, 4).tolist() == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]


def test_encode_slices_tile_the_batch():
    """Split query encode of the sharded search: the ranks' slices are disjoint, ordered and cover the batch, with the gathered
    buffer holding query q at row q (slices are ceil(B / world) apart)."""
    from muopdb_b200 import sharding
    for B in (0, 1, 7, 203, 1024, 8192):
        for world in (1, 2, 3, 8):
            per = (B + world - 1) // world
            covered = []
            for r in range(world):
                lo, cnt = sharding.encode_slice(B, world, r)
                assert lo == min(B, r * per) and 0 <= cnt <= per
                covered.extend(range(lo, lo + cnt))
            assert covered == list(range(B)), (B, world)

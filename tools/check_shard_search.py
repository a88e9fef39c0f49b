#!/usr/bin/env python
"""Multi-rank check of the sharded query path (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_shard_search.py
Every rank builds its doc-shard (doc_id mod N) of one seeded collection with a shared PQ codebook and compares
  (a) mgpu_shard_ivf_search (split query encode + code all-gather + local search + result all-gather + merge), host and
      device buffers, with
  (b) the CPU oracle: per-shard oracle searches merged with the leaf ordering of snapshot.rs:60-61;
and the same for mgpu_shard_spann_search (config 5: centroid HNSW -> ratio prune -> PQ lists per shard), including the
overlapped device-buffer mode (mgpu_shard_overlap) and the pipelined host-buffer form.
Bit-exact doc ids and scores are required on every rank.  Rank 0 prints one JSON line (kept under profiles/)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import muopdb_b200 as M  # noqa: E402
import oracle as O  # noqa: E402
import synth  # noqa: E402
from muopdb_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = M.Context(local)
    sharding.init_comm(ctx)
    dim, n, k, nprobe, B = 256, 12000, 10, 4, 203   # B not a multiple of the world size: the last encode slice is ragged
    X = synth.clustered(n, dim, n_blobs=14, seed=5)
    docs = np.arange(n, dtype=np.uint64) * 3 + 7
    cb = O.train_pq_codebook(X[:3000], 8, 8, iters=3, seed=2)
    rng = np.random.default_rng(9)
    Q = (X[rng.integers(0, n, B)] + 0.02 * rng.standard_normal((B, dim))).astype(np.float32)
    shards = []
    for r in range(world):
        sel = sharding.shard_of(docs, world) == r
        Xs, ds = X[sel], docs[sel]
        cents = O.kmeans(Xs, 16, iters=4, seed=1 + r)
        offsets, ids = O.build_posting_lists(Xs, cents)
        opq = O.ProductQuantizer(dim, 8, 8, cb)
        codes = opq.quantize(Xs)
        pairs = np.zeros((len(ds), 2), dtype=np.uint64)
        pairs[:, 0] = ds
        shards.append((cents, offsets, ids, codes, pairs, opq))
    # oracle: every shard, then merge
    per = [O.Ivf(c, o, i, cd, doc_ids=p, pq=q).search_batch(Q, k, nprobe) for (c, o, i, cd, p, q) in shards]
    merged = []
    for b in range(B):
        d = np.concatenate([np.asarray(pr[0][b, :int(pr[2][b])]).reshape(-1, 2) for pr in per])
        sc = np.concatenate([np.asarray(pr[1][b, :int(pr[2][b])], dtype=np.float32) for pr in per])
        merged.append(O.merge_topk(d, sc, k))   # (list of u128 ints, f32 scores) ordered by (score, doc_id)
    c, o, i, cd, p, _ = shards[rank]
    givf = M.BlockBasedIvf(c, o, i, cd, M.ProductQuantizer(dim, 8, 8, cb, ctx=ctx), doc_ids=p, ctx=ctx)
    ok = True

    def same(r, dev, tag=""):
        ctx.sync()   # device-buffer calls are asynchronous on the library stream
        ids_, sc, cn = ((x.cpu().numpy() if dev else np.asarray(x)) for x in (r.doc_ids, r.scores, r.counts))
        bad = 0
        for b in range(B):
            md, ms = merged[b]
            nn = len(ms)
            got = [int(lo) | (int(hi) << 64) for lo, hi in ids_[b, :nn].view(np.uint64).reshape(-1, 2)]
            g = int(cn[b]) == nn and got == [int(x) for x in md] and \
                np.array_equal(sc[b, :nn].view(np.uint32), np.asarray(ms, dtype=np.float32).view(np.uint32))
            if not g:
                if bad == 0:
                    print(f"[rank {rank}] {tag} dev={dev} query {b}: count {int(cn[b])} vs {nn}\n  got {got[:6]} {sc[b, :6]}\n  ref {[int(x) for x in md][:6]} {ms[:6]}", flush=True)
                bad += 1
        if bad:
            print(f"[rank {rank}] {tag} dev={dev}: {bad}/{B} queries differ", flush=True)
        return bad == 0

    for dev in (False, True):
        ok &= same(givf.shard_search_batch(torch.from_numpy(Q).cuda() if dev else Q, k, nprobe, shared_codebook=True), dev, 'split')
    # pipelined submit/wait over pinned buffers, two batches in flight
    Qp = torch.from_numpy(Q).pin_memory()
    outs = [(torch.zeros((B, k, 2), dtype=torch.int64).pin_memory(), torch.zeros((B, k), dtype=torch.float32).pin_memory(),
             torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(3)]
    tks = [givf.shard_search_batch_submit(Qp, k, nprobe, o) for o in outs]
    for t in tks:
        givf.search_wait(t)
    for o in outs:
        ok &= same(M.BatchResult(o[0].numpy(), o[1].numpy(), o[2].numpy()), False, 'pipelined')
    # and without the split encode (every rank encodes every query): same answer
    ok &= same(givf.shard_search_batch(Q, k, nprobe, shared_codebook=False), False, 'nosplit')
    # overlapped device-buffer mode: three calls back to back, their exchanges in flight next to the following searches
    ctx.shard_overlap(True)
    Qd = torch.from_numpy(Q).cuda()
    rs = [givf.shard_search_batch(Qd, k, nprobe, shared_codebook=True) for _ in range(3)]
    for r in rs:
        ok &= same(r, True, 'overlap')
    ctx.shard_overlap(False)
    ivf_ok = ok

    # ---- config 5: SPANN per shard (centroid HNSW with NoQuantizer<L2> over the shard's centroids + the PQ lists above)
    ne, ratio, ef = 6, 0.4, 40
    graphs = [O.hnsw_build(sh[0], 8, 2, 60, seed=3 + r) for r, sh in enumerate(shards)]
    per = []
    for (c2, o2, i2, cd2, p2, q2), g in zip(shards, graphs):
        osp = O.Spann(O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], c2),
                      O.Ivf(c2, o2, i2, cd2, doc_ids=p2, pq=q2))
        per.append(osp.search_batch(Q, k, ef, ne, ratio))
    merged.clear()
    n_none = 0
    for b in range(B):
        live = [pr for pr in per if int(pr[2][b]) >= 0]          # count -1 == None (spann/index.rs:229-231)
        n_none += len(per) - len(live)
        if live:
            d = np.concatenate([np.asarray(pr[0][b, :int(pr[2][b])]).reshape(-1, 2) for pr in live])
            sc = np.concatenate([np.asarray(pr[1][b, :int(pr[2][b])], dtype=np.float32) for pr in live])
            merged.append(O.merge_topk(d, sc, k))
        else:
            merged.append(([], np.zeros(0, dtype=np.float32)))
    g = graphs[rank]
    ghn = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], c, M.NoQuantizer(dim), ctx=ctx)
    gsp = M.Spann(ghn, givf)
    params = M.SearchParams(k, ef, False, ne, ratio)
    sp_ok = True
    for dev in (False, True):
        sp_ok &= same(gsp.shard_search_batch(torch.from_numpy(Q).cuda() if dev else Q, params), dev, 'spann')
    sp_ok &= same(gsp.shard_search_batch(Q, params, shared_codebook=False), False, 'spann-nosplit')
    tks = [gsp.shard_search_batch_submit(Qp, params, o) for o in outs]
    for t in tks:
        gsp.search_wait(t)
    for o in outs:
        sp_ok &= same(M.BatchResult(o[0].numpy(), o[1].numpy(), o[2].numpy()), False, 'spann-pipelined')
    ctx.shard_overlap(True)
    rs = [gsp.shard_search_batch(Qd, params) for _ in range(3)]
    for r in rs:
        sp_ok &= same(r, True, 'spann-overlap')
    ctx.shard_overlap(False)
    ok = ivf_ok and sp_ok
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("shard search check:", "OK" if int(t.item()) == 1 else "MISMATCH", f"(world {world}, B {B})")
        print(json.dumps({"check": "sharded IVF + SPANN search vs per-shard oracle + oracle merge", "world": world, "queries": B,
                          "k": k, "modes": ["host", "device", "pipelined", "no split encode", "overlapped exchange"],
                          "ivf_ok_rank0": bool(ivf_ok), "spann_ok_rank0": bool(sp_ok), "all_ranks_ok": int(t.item()) == 1,
                          "gpu": torch.cuda.get_device_name(local)}))
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

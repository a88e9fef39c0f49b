#!/usr/bin/env python
"""One rank's share of the N-GPU bench on ONE GPU: shard 0 of `--nshards` (doc_id mod N), nlist 4096/N, nprobe 64/N, the
replicated global batch of 1024*N queries -- the shape at which per-query fixed work (LUT build, top-k merge, re-rank)
dominates the scan.  No collective runs: this isolates the local part of mgpu_shard_ivf_search so that
kernel work at the shard shape can be iterated on a 1-GPU box.
Usage: python tools/bench_shard_shape.py [--nshards 8] [--steps 10]   (one JSON line)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--nshards", type=int, default=8)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--check", type=int, default=64, help="queries compared bit for bit with the oracle")
    a = p.parse_args()
    import torch
    import muopdb_b200 as M
    from muopdb_b200 import _lib
    import oracle as O

    class A:
        pass
    args = A()
    args.n, args.dim, args.nlist, args.nprobe, args.batch, args.k, args.dsub, args.seed = 1_000_000, 768, 4096, 64, 1024, 10, 8, 1234
    dev = torch.device("cuda", 0)
    ctx = M.default_context(0)
    S = a.nshards
    col = bench.make_collection(args, dev, shard=0, nshards=S)
    nprobe, B, k = max(args.nprobe // S, 1), args.batch * S, args.k
    cb = col["codebook"].cpu().numpy()
    pq = M.ProductQuantizer(args.dim, args.dsub, 8, cb, ctx=ctx)
    codes = pq.quantize(col["X"])
    ctx.sync()
    docs = np.zeros((col["docs"].shape[0], 2), dtype=np.uint64)
    docs[:, 0] = col["docs"].cpu().numpy().astype(np.uint64)
    cents = col["centroids"].cpu().numpy()
    offs = col["offsets"].cpu().numpy().astype(np.uint64)
    ids = col["list_ids"].cpu().numpy().astype(np.uint32)
    ivf = M.BlockBasedIvf(cents, offs, ids, codes, pq, doc_ids=docs, ctx=ctx)
    Q = col["Q"]
    nb = Q.shape[0] // B
    out = (torch.zeros((B, k, 2), dtype=torch.int64, device=dev), torch.zeros((B, k), dtype=torch.float32, device=dev),
           torch.zeros((B,), dtype=torch.int32, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(ctx.stream)

    def step(i):
        ivf.search_batch(Q[(i % nb) * B:(i % nb + 1) * B], k, nprobe, out=out)

    for i in range(3):
        step(i)
    ctx.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    with torch.cuda.stream(ext):
        for i in range(a.steps):
            flush.zero_()
            ev[i][0].record(ext)
            step(i)
            ev[i][1].record(ext)
    ctx.sync(); torch.cuda.synchronize()
    ms = sum(x.elapsed_time(y) for x, y in ev) / a.steps
    ctx.profile_reset(); ctx.profile_enable(True)
    with torch.cuda.stream(ext):
        for i in range(a.steps):
            flush.zero_()
            step(i)
    ctx.sync(); torch.cuda.synchronize()
    ctx.profile_enable(False)
    prof = {n: ctx.profile_get(c)[0] / a.steps for c, n in enumerate(_lib.KERNEL_CLASS_NAMES)}
    rows = ivf.last_scan_rows()
    # parity of a sample against the oracle
    step(0); ctx.sync()
    ns = a.check
    opq = O.ProductQuantizer(args.dim, args.dsub, 8, cb)
    oivf = O.Ivf(cents, offs, ids, codes.cpu().numpy(), doc_ids=docs, pq=opq)
    od, os_, oc = oivf.search_batch(Q[:ns].cpu().numpy(), k, nprobe)
    ok = np.array_equal(od, out[0][:ns].cpu().numpy().view(np.uint64)) and \
        np.array_equal(os_.view(np.uint32), out[1][:ns].cpu().numpy().view(np.uint32))
    print(json.dumps({"shape": f"shard 0 of {S}: {col['X'].shape[0]} rows, nlist {col['nlist']}, nprobe {nprobe}, batch {B}",
                      "ms_per_step": ms, "qps_per_rank_equiv": B / (ms / 1e3), "kernel_ms_per_step": prof, "rows_per_launch": rows, "fallback_queries": ivf.last_scan_fallbacks(),
                      "parity_sample_ok": bool(ok)}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line:
  C2  IVF flat-L2 1M x 768, nlist 4096, nprobe 32, batch 1024
  C4  HNSW M=32 ef_search=128, 1M x 768, batch 256
  C5s SPANN (centroid HNSW + PQ lists) on one 1.25M x 768 shard, batch 1024 (the per-GPU unit of config 5)
Each leg: QPS with device-resident queries (CUDA events on the library stream), algorithmic bytes -> fraction of the
measured HBM roofline, and a bit-exact parity check of a query sample against the CPU oracle.
Usage: python tools/bench_configs.py [c2] [c4] [c5s] [--n N]   (writes one JSON line per leg)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic collection generator)


class A:
    pass


def args_for(n, nlist, nprobe, batch):
    a = A()
    a.n, a.dim, a.nlist, a.nprobe, a.batch, a.k, a.dsub, a.seed = n, 768, nlist, nprobe, batch, 10, 8, 1234
    return a


def timed(ctx, fn, steps, flush):
    import torch
    ext = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        fn(0)
    ctx.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(ext):
        for i in range(steps):
            flush.zero_()
            ev[i][0].record(ext)
            fn(i)
            ev[i][1].record(ext)
    ctx.sync()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / steps


def peak():
    return bench.measured_peak_gbs()[0]


def leg_c2(n):
    import torch
    import muopdb_b200 as M
    import oracle as O
    a = args_for(n, 4096, 32, 1024)
    dev = torch.device("cuda", 0)
    ctx = M.default_context(0)
    col = bench.make_collection(a, dev)
    docs = np.zeros((n, 2), dtype=np.uint64)
    docs[:, 0] = np.arange(n, dtype=np.uint64)
    cents = col["centroids"].cpu().numpy()
    offs = col["offsets"].cpu().numpy().astype(np.uint64)
    ids = col["list_ids"].cpu().numpy().astype(np.uint32)
    ivf = M.BlockBasedIvf(cents, offs, ids, col["X"], M.NoQuantizer(768), doc_ids=docs, ctx=ctx)
    Q = col["Q"]
    B = a.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = (torch.zeros((B, 10, 2), dtype=torch.int64, device=dev), torch.zeros((B, 10), dtype=torch.float32, device=dev),
           torch.zeros((B,), dtype=torch.int32, device=dev))
    nb = Q.shape[0] // B
    ctx.profile_reset(); ctx.profile_enable(True)
    ms = timed(ctx, lambda i: ivf.search_batch(Q[(i % nb) * B:(i % nb + 1) * B], 10, a.nprobe, out=out), 10, flush)
    ctx.profile_enable(False)
    scan_ms, launches = ctx.profile_get(3)
    byts = ivf.last_scan_bytes()
    ivf.search_batch(Q[:B], 10, a.nprobe, out=out)
    ctx.sync()
    gt = bench.exact_topk(col["X"], col["docs"], Q[:B], 10)
    recall = (out[0][:, :, 0][:, :, None] == gt[:, None, :]).any(-1).float().mean().item() * 1.0
    recall = (out[0][:, :, 0][:, :, None] == gt[:, None, :]).any(-1).float().sum().item() / (B * 10)
    # parity sample vs oracle
    oivf = O.Ivf(cents, offs, ids, col["X"].cpu().numpy(), doc_ids=docs)
    ns = 32
    t0 = time.perf_counter()
    od, os_, oc = oivf.search_batch(Q[:ns].cpu().numpy(), 10, a.nprobe)
    cpu_s = time.perf_counter() - t0
    ok = np.array_equal(od, out[0][:ns].cpu().numpy().view(np.uint64)) and np.array_equal(os_.view(np.uint32), out[1][:ns].cpu().numpy().view(np.uint32))
    ach = byts / (scan_ms / launches / 1e3) / 1e9
    return {"config": "C2 IVF flat-L2 %dx768 nlist=4096 nprobe=32 batch=1024" % n, "qps": B / (ms / 1e3), "ms_per_batch": ms,
            "recall_at_10": recall, "scan_ms": scan_ms / launches, "algorithmic_bytes": byts, "roofline_GBs": ach, "roofline_frac": ach / peak(),
            "cpu_oracle_qps_%d_threads" % O.num_threads(): ns / cpu_s, "parity_sample_ok": bool(ok)}


def knn_graph(X, Msz, chunk=4096):
    """Approximate index input: per-point nearest neighbours by brute force (torch matmul; setup only)."""
    import torch
    n = X.shape[0]
    xn = (X * X).sum(1)
    out = torch.empty((n, Msz), dtype=torch.int32, device=X.device)
    for i in range(0, n, chunk):
        q = X[i:i + chunk]
        best_d = torch.full((q.shape[0], Msz + 1), float("inf"), device=X.device)
        best_i = torch.zeros((q.shape[0], Msz + 1), dtype=torch.int64, device=X.device)
        for j in range(0, n, 262144):
            d = xn[None, j:j + 262144] - 2.0 * (q @ X[j:j + 262144].T)
            dd, ii = torch.topk(d, min(Msz + 1, d.shape[1]), dim=1, largest=False)
            cd, ci = torch.cat([best_d, dd], 1), torch.cat([best_i, ii + j], 1)
            sel = torch.topk(cd, Msz + 1, dim=1, largest=False)
            best_d, best_i = sel.values, torch.gather(ci, 1, sel.indices)
        rows = torch.arange(i, i + q.shape[0], device=X.device)[:, None]
        keep = best_i != rows
        # drop self (first column normally) keeping order
        idx = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :Msz]
        out[i:i + q.shape[0]] = torch.gather(best_i, 1, idx).to(torch.int32)
    return out


def build_hnsw_arrays(X, Msz=32, seed=7):
    """HNSW-format graph arrays (hnsw/writer.rs layout) from brute-force kNN per layer + a few random long edges on layer 0;
    layer membership by the reference's level rule floor(-ln(u)/ln(M)) (hnsw/builder.rs:332-337), seeded."""
    import torch
    n = X.shape[0]
    g = torch.Generator(device=X.device); g.manual_seed(seed)
    u = torch.rand(n, generator=g, device=X.device).clamp_(1e-9, 1.0)
    level = torch.floor(-torch.log(u) / np.log(Msz)).to(torch.int64).clamp_(0, 6)
    top = int(level.max().item())
    layers_pts = [torch.nonzero(level >= l).flatten() for l in range(top + 1)]
    edges, points, edge_offsets, level_offsets = [], [], [0], [0]
    cur = 0
    for l in range(top, -1, -1):
        pts = layers_pts[l]
        m = pts.shape[0]
        kk = min(Msz, max(m - 1, 1))
        if m > 1:
            nb = knn_graph(X[pts], kk)
            nb = pts[nb.long()].to(torch.int32)
            if l == 0:  # a few random long-range edges for navigability
                rnd = torch.randint(0, n, (n, 4), generator=g, device=X.device, dtype=torch.int32)
                nb[:, -4:] = rnd
        else:
            nb = torch.zeros((m, 0), dtype=torch.int32, device=X.device)
        deg = nb.shape[1]
        if l > 0:
            points.append(pts.to(torch.int32).cpu().numpy())
        edges.append(nb.reshape(-1).cpu().numpy())
        offs = cur + deg * np.arange(1, m + 1, dtype=np.uint64)
        edge_offsets.extend(offs.tolist())
        cur += deg * m
        level_offsets.append(level_offsets[-1] + m)
    return dict(num_layers=top + 1, edges=np.concatenate(edges).astype(np.uint32),
                points=np.concatenate(points).astype(np.uint32) if points else np.zeros(0, np.uint32),
                edge_offsets=np.array(edge_offsets, dtype=np.uint64), level_offsets=np.array(level_offsets, dtype=np.uint64))


def leg_c4(n):
    import torch
    import muopdb_b200 as M
    import oracle as O
    a = args_for(n, 64, 8, 256)
    dev = torch.device("cuda", 0)
    ctx = M.default_context(0)
    col = bench.make_collection(a, dev)
    X, Q = col["X"], col["Q"]
    t0 = time.perf_counter()
    g = build_hnsw_arrays(X)
    build_s = time.perf_counter() - t0
    hn = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X, M.NoQuantizer(768), ctx=ctx)
    B = 256
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nb = Q.shape[0] // B
    res = {}

    def step(i):
        res["r"], res["st"] = hn.ann_search_batch(Q[(i % nb) * B:(i % nb + 1) * B], 10, 128, with_stats=True)

    ms = timed(ctx, step, 10, flush)
    step(0); ctx.sync()
    st = res["st"].cpu().numpy()
    deg = 32
    byts = int(st[:, 0].sum()) * 768 * 4 + int(st[:, 1].sum()) * (16 + 4 * deg)
    gt = bench.exact_topk(X, col["docs"], Q[:B], 10)
    got = res["r"].doc_ids[:, :, 0]
    recall = (got[:, :, None] == gt[:, None, :]).any(-1).float().sum().item() / (B * 10)
    oh = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X.cpu().numpy())
    ns = 16
    t0 = time.perf_counter()
    od, os_, oc, ost = oh.search_batch(Q[:ns].cpu().numpy(), 10, 128)
    cpu_s = time.perf_counter() - t0
    ok = np.array_equal(od, res["r"].doc_ids[:ns].cpu().numpy().view(np.uint64)) and \
        np.array_equal(os_.view(np.uint32), res["r"].scores[:ns].cpu().numpy().view(np.uint32)) and np.array_equal(ost, st[:ns].astype(np.uint64))
    ach = byts / (ms / 1e3) / 1e9
    return {"config": "C4 HNSW M=32 ef=128 %dx768 batch=256 (kNN-built graph, %d layers, build %.0fs)" % (n, g["num_layers"], build_s),
            "qps": B / (ms / 1e3), "ms_per_batch": ms, "recall_at_10": recall, "dist_evals_per_query": float(st[:, 0].mean()),
            "expansions_per_query": float(st[:, 1].mean()), "algorithmic_bytes": byts, "roofline_GBs": ach, "roofline_frac": ach / peak(),
            "cpu_oracle_qps_%d_threads" % O.num_threads(): ns / cpu_s, "parity_sample_ok": bool(ok)}


def leg_c5s(n):
    import torch
    import muopdb_b200 as M
    import oracle as O
    a = args_for(n, 4096, 64, 1024)
    dev = torch.device("cuda", 0)
    ctx = M.default_context(0)
    col = bench.make_collection(a, dev)
    X, Q = col["X"], col["Q"]
    cents = col["centroids"]
    g = build_hnsw_arrays(cents, Msz=32, seed=3)
    cb = col["codebook"].cpu().numpy()
    pq = M.ProductQuantizer(768, 8, 8, cb, ctx=ctx)
    codes = pq.quantize(X); ctx.sync()
    docs = np.zeros((n, 2), dtype=np.uint64); docs[:, 0] = np.arange(n, dtype=np.uint64)
    cn = cents.cpu().numpy()
    offs = col["offsets"].cpu().numpy().astype(np.uint64); ids = col["list_ids"].cpu().numpy().astype(np.uint32)
    ivf = M.BlockBasedIvf(cn, offs, ids, codes, pq, doc_ids=docs, ctx=ctx)
    hn = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cn, M.NoQuantizer(768), ctx=ctx)
    sp = M.Spann(hn, ivf)
    params = M.SearchParams(10, 128, False, 64, 1e9)  # 64 explored centroids, no ratio pruning
    B = 1024
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nb = Q.shape[0] // B
    res = {}

    def step(i):
        res["r"] = sp.search_batch(Q[(i % nb) * B:(i % nb + 1) * B], params)

    ms = timed(ctx, step, 10, flush)
    step(0); ctx.sync()
    gt = bench.exact_topk(X, col["docs"], Q[:B], 10)
    recall = (res["r"].doc_ids[:, :, 0][:, :, None] == gt[:, None, :]).any(-1).float().sum().item() / (B * 10)
    opq = O.ProductQuantizer(768, 8, 8, cb)
    oivf = O.Ivf(cn, offs, ids, codes.cpu().numpy(), doc_ids=docs, pq=opq)
    ohn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cn)
    ns = 32
    t0 = time.perf_counter()
    od, os_, oc = O.Spann(ohn, oivf).search_batch(Q[:ns].cpu().numpy(), 10, 128, 64, 1e9)
    cpu_s = time.perf_counter() - t0
    ok = np.array_equal(od, res["r"].doc_ids[:ns].cpu().numpy().view(np.uint64)) and \
        np.array_equal(os_.view(np.uint32), res["r"].scores[:ns].cpu().numpy().view(np.uint32))
    return {"config": "C5 shard: SPANN (centroid HNSW M=32 ef=128 + PQ m=96 lists) %dx768, 64 explored centroids, batch=1024" % n,
            "qps": B / (ms / 1e3), "ms_per_batch": ms, "recall_at_10": recall,
            "cpu_oracle_qps_%d_threads" % O.num_threads(): ns / cpu_s, "parity_sample_ok": bool(ok)}


if __name__ == "__main__":
    legs = [x for x in sys.argv[1:] if x in ("c2", "c4", "c5s")] or ["c2", "c4", "c5s"]
    n = 1_000_000
    if "--n" in sys.argv:
        n = int(sys.argv[sys.argv.index("--n") + 1])
    for leg in legs:
        r = {"c2": leg_c2, "c4": leg_c4, "c5s": lambda nn: leg_c5s(nn if "--n" in sys.argv else 1_250_000)}[leg](n)
        print(json.dumps(r), flush=True)

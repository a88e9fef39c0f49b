#!/usr/bin/env python
"""Where does a config-5 step go?  Host enqueue time vs device time of the centroid HNSW search and of the whole Spann search
on one 1.25M-row shard.  Usage: python tools/prof_spann.py [--n 1250000]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import muopdb_b200 as M
    from muopdb_b200 import _lib
    sys.argv = [sys.argv[0], "--config", "c5"] + sys.argv[1:]
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    ctx = M.default_context(0)
    col = bench.make_collection(args, dev, grow=True)
    idx = bench.build_indices(args, col, ctx, M, 1)
    sp = idx["objs"][0]
    Q = col["Q"][:args.batch]
    params = M.SearchParams(10, 128, False, 64, 1e9)
    ext = torch.cuda.ExternalStream(ctx.stream)
    out = {}
    for name, fn in (("hnsw_centroids_k64_ef128", lambda: sp.centroids.ann_search_batch(Q, 64, 128)),
                     ("spann_search", lambda: sp.search_batch(Q, params))):
        with torch.cuda.stream(ext):
            for _ in range(3):
                fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record(ext)
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            host = time.perf_counter() - t0
            e1.record(ext)
        ctx.sync()
        out[name] = {"device_ms_per_call": e0.elapsed_time(e1) / 10, "host_enqueue_ms_per_call": host * 100}
    ctx.profile_reset(); ctx.profile_enable(True)
    with torch.cuda.stream(ext):
        for _ in range(5):
            sp.search_batch(Q, params)
    ctx.sync()
    ctx.profile_enable(False)
    out["kernel_ms"] = {n: ctx.profile_get(c)[0] / 5 for c, n in enumerate(_lib.KERNEL_CLASS_NAMES)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- QPS @ recall@10 of the batched ANN search path (BASELINE.json metric) on N B200s.

Default workload (BASELINE.json configs[2], "C3"): 1M x 768 f32 vectors, IVF nlist 4096, PQ m=96 (dsub 8) 8 bit, nprobe 64,
batch 1024, k 10; synthetic seeded Gaussian-blob data.  One "step" = one batch of 1024 queries per GPU
through coarse scoring -> query quantize -> PQ LUT scan -> exact re-rank -> doc-id remap.

  value     whole-job QPS, queries already resident in HBM: CUDA events around the K back-to-back steps, max over ranks
  e2e       the same metric through the host-buffer C-ABI call, pipelined (mgpu_*_search_submit / mgpu_search_wait): pinned
            host queries in, results out, every copy inside the wall-clock timed region -- nothing subtracted
  roofline  the dominant kernel (posting-list scan / HNSW beam search): algorithmic bytes / event-timed kernel time vs the
            measured HBM peak
  cpu_baseline  the C oracle (port of the reference's Rust path) on the host cores, bounded sample, with a bit-for-bit
            parity check of the GPU result on that sample

L2: no flush kernel runs inside the timed regions.  Consecutive steps search different resident COPIES of the index
(enough copies that two steps' inputs exceed the 126 MB L2), so a step never finds its codes in L2 because the previous step
left them there; reuse inside a step (queries sharing lists) is the kernel's own.

--config  c3 (default) | c2 flat IVF 1M x 768, nprobe 32 | c4 HNSW M=32 ef=128, batch 256 | c5 SPANN (centroid HNSW + PQ lists),
          10M x 768 doc-sharded over the ranks (1.25M rows per rank; at N = 1: one shard)

N > 1 (torchrun): doc-sharded (doc_id mod N).  c3/c2: the 1M index is split N ways, every rank scans its shard for the
replicated global batch of 1024*N queries with nprobe/N of its nlist/N lists ("weak": per-GPU scan work is constant in N);
c5: every rank holds one 1.25M shard with its own 4096 lists and centroid graph.  One NCCL all-gather of the per-shard top-k +
merge kernel gives every rank the merged result; the exchange of batch i runs on its own stream next to batch i+1's
kernels.  Parity at N > 1: every rank runs the oracle on ITS shard for a query sample, rank 0 merges the per-shard oracle
results with the oracle's merge (snapshot.rs:49-63) and compares with the GPU's merged answer bit for bit.

--impl reference times the oracle alone on all host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QPS @ recall@10 on 1M x 768 IVF-PQ, batch=1024"
UNIT = "queries/s"
L2_BYTES = 126 << 20

CONFIGS = {
    # name: (metric string, n, nlist, nprobe, batch, quantizer)
    "c3": ("QPS @ recall@10 on 1M x 768 IVF-PQ, batch=1024", 1_000_000, 4096, 64, 1024, "pq"),
    "c2": ("QPS @ recall@10 on 1M x 768 IVF flat-L2, nlist=4096 nprobe=32, batch=1024", 1_000_000, 4096, 32, 1024, "flat"),
    "c4": ("QPS @ recall@10 on 1M x 768 HNSW M=32 ef_search=128, batch=256", 1_000_000, 64, 8, 256, "hnsw"),
    "c5": ("QPS @ recall@10 on 10M x 768 SPANN-PQ doc-sharded over 8 GPUs (1.25M rows/GPU), batch=1024", 1_250_000, 4096, 64, 1024, "spann"),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    p.add_argument("--n", type=int, default=0, help="rows (c5: rows per shard); 0 = the config's")
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nlist", type=int, default=0)
    p.add_argument("--nprobe", type=int, default=0)
    p.add_argument("--batch", type=int, default=0)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--dsub", type=int, default=8)
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU-baseline sample")
    p.add_argument("--ref-queries-per-step", type=int, default=0, help="--impl reference: queries per step (0 = 4 x threads)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--parity-queries", type=int, default=64, help="N > 1: queries of the per-shard oracle parity sample")
    a = p.parse_args()
    metric, n, nlist, nprobe, batch, quant = CONFIGS[a.config]
    a.metric, a.quant = metric, quant
    a.n, a.nlist, a.nprobe, a.batch = a.n or n, a.nlist or nlist, a.nprobe or nprobe, a.batch or batch
    return a


# ---- synthetic collection (setup, untimed; torch is only the generator/plumbing here) ---------------------------------------
def make_collection(args, device, shard=0, nshards=1, grow=False):
    """Seeded blobs -> coarse centroids (Lloyd on a sample) -> posting lists -> PQ codebook.  Returns torch tensors on
    `device`.  Posting lists / codebooks are INPUTS of the search path, any valid index will do.
    grow=False: a collection of args.n rows split `nshards` ways (nlist / nshards lists per shard).
    grow=True (config 5): args.n rows PER SHARD, collection = args.n * nshards rows, args.nlist lists per shard; only this
    rank's rows are ever materialised (row i of shard s is global row i * nshards + s, drawn from its own seeded stream)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(args.seed)
    N, D = args.n, args.dim
    # low intrinsic dimension (latent 32-d mixture of 2048 blobs, embedded by a fixed random map, plus a little isotropic
    # noise): i.i.d. 768-d noise would make every neighbour equidistant and recall@10 meaningless
    n_blobs, latent = 2048, 32
    W = torch.randn((latent, D), generator=g, device=device) / latent ** 0.5
    centers = torch.randn((n_blobs, latent), generator=g, device=device) * 2.0

    def draw(n):
        lab = torch.randint(0, n_blobs, (n,), generator=g, device=device)
        out = torch.empty((n, D), device=device)
        for i in range(0, n, 131072):
            j = min(n, i + 131072)
            z = centers[lab[i:j]] + torch.randn((j - i, latent), generator=g, device=device)
            out[i:j] = z @ W + 0.02 * torch.randn((j - i, D), generator=g, device=device)
        return out

    if grow:
        # queries and the codebook training sample come from the common stream (every rank holds the same ones), this
        # shard's rows from its own
        nq_total = args.batch * 4
        Q = draw(nq_total)
        train = draw(10000)
        g.manual_seed(args.seed + 7919 * (shard + 1))
        X = draw(N)
        doc_lo = torch.arange(N, device=device, dtype=torch.int64) * nshards + shard
        g.manual_seed(args.seed + 1)
    else:
        X = draw(N)
        # queries: fresh draws from the same mixture
        nq_total = args.batch * max(nshards, 1) * 4
        Q = draw(nq_total)
        doc_lo = torch.arange(N, device=device, dtype=torch.int64)
        train = None
    # PQ codebook shared by all shards: per-subspace Lloyd on 10 000 sampled rows (reference default
    # product_quantization_num_training_rows, rs/config/src/collection.rs:190)
    m, K = D // args.dsub, 256
    if train is None:
        train = X[torch.randperm(N, generator=g, device=device)[:10000]]
    samp = train.reshape(10000, m, args.dsub).permute(1, 0, 2).contiguous()
    cb = samp[:, torch.randperm(10000, generator=g, device=device)[:K]].clone()  # (m, K, dsub)
    for _ in range(8):
        d = torch.cdist(samp, cb)  # (m, 10000, K)
        a = d.argmin(dim=2)
        for s0 in range(0, m, 16):
            oh = torch.nn.functional.one_hot(a[s0:s0 + 16], K).to(torch.float32)  # (16, 10000, K)
            cnt = oh.sum(1)  # (16, K)
            newc = torch.einsum("snk,snd->skd", oh, samp[s0:s0 + 16])
            mask = cnt > 0
            cb[s0:s0 + 16][mask] = (newc[mask] / cnt[mask].unsqueeze(-1))
    # shard
    if nshards > 1 and not grow:
        keep = (doc_lo % nshards) == shard
        Xs, docs = X[keep].contiguous(), doc_lo[keep].contiguous()
    else:
        Xs, docs = X, doc_lo
    del X
    nlist = args.nlist if grow else max(args.nlist // nshards, 1)
    ns = Xs.shape[0]
    # coarse centroids: Lloyd on a sample, matmul-form assignment (setup only)
    perm = torch.randperm(ns, generator=g, device=device)
    S = Xs[perm[:min(ns, 262144)]]
    C = S[:nlist].clone()

    def assign(A, Cm):
        out = torch.empty(A.shape[0], dtype=torch.int64, device=device)
        cn = (Cm * Cm).sum(1)
        for i in range(0, A.shape[0], 65536):
            a = A[i:i + 65536]
            out[i:i + 65536] = (cn[None, :] - 2.0 * (a @ Cm.T)).argmin(dim=1)
        return out

    for _ in range(6):
        a = assign(S, C)
        sums = torch.zeros_like(C).index_add_(0, a, S)
        cnt = torch.bincount(a, minlength=nlist).to(torch.float32)
        nz = cnt > 0
        C[nz] = sums[nz] / cnt[nz].unsqueeze(1)
    a = assign(Xs, C)
    order = torch.argsort(a, stable=True)  # ascending point id inside every list (ivf/builder.rs:339-341)
    counts = torch.bincount(a, minlength=nlist)
    offsets = torch.zeros(nlist + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    return dict(X=Xs, Q=Q, docs=docs, centroids=C, offsets=offsets, list_ids=order.to(torch.int32), codebook=cb.reshape(-1),
                nlist=nlist)


def exact_topk(X, docs, Q, k):
    """brute-force fp32 L2 ground truth for recall (setup; torch)."""
    import torch
    xn = (X * X).sum(1)
    best_d = torch.full((Q.shape[0], k), float("inf"), device=X.device)
    best_i = torch.zeros((Q.shape[0], k), dtype=torch.int64, device=X.device)
    for i in range(0, X.shape[0], 262144):
        d = xn[None, i:i + 262144] - 2.0 * (Q @ X[i:i + 262144].T)
        dd, ii = torch.topk(d, k, dim=1, largest=False)
        cat_d = torch.cat([best_d, dd], 1)
        cat_i = torch.cat([best_i, docs[i:i + 262144][ii]], 1)
        sel = torch.topk(cat_d, k, dim=1, largest=False)
        best_d, best_i = sel.values, torch.gather(cat_i, 1, sel.indices)
    return best_i


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region of this bench is
    tens of milliseconds long, so the sampler polls NVML directly (~1 ms per sample) instead of `nvidia-smi -lms` (one
    sample per 100 ms); nvidia-smi is the fallback when the NVML binding is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons = [], [], set()
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD_MS", "1")) / 1e3   # NVML poll period

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for nme, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = [float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))]
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(shape_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at this workload shape, from the
    committed `ncu --set full` summary (profiles/ncu_traffic.json: {shape_key: bytes}); null when that shape was never
    captured (a profiler cannot run inside the timed bench)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(shape_key)
    except Exception:
        return None


# ---- CPU reference arm --------------------------------------------------------------------------------------------------------
def cpu_leg(search, Qh, seconds, per_call, keep):
    """Times `search` (the oracle) on successive slices of Qh (cycling) until `seconds` of CPU wall time are used; the results
    of the first `keep` queries are returned for the parity check."""
    done, t_used, i = 0, 0.0, 0
    res = []
    search(Qh[:min(per_call, Qh.shape[0])])  # warm-up (page faults, thread pool)
    while t_used < seconds or i < keep:
        lo = i % Qh.shape[0]
        q = Qh[lo:lo + per_call]
        t0 = time.perf_counter()
        r = search(q)
        t_used += time.perf_counter() - t0
        if i < keep:
            res.append(r)
        done += q.shape[0]
        i += per_call
        if t_used >= seconds and i >= keep:
            break
        if t_used > 6 * seconds:
            break
    return done, t_used, res


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



# ---- the GPU arm ---------------------------------------------------------------------------------------------------------------
def _pairs(docs_t):
    p = np.zeros((docs_t.shape[0], 2), dtype=np.uint64)
    p[:, 0] = docs_t.cpu().numpy().astype(np.uint64)
    return p


def build_indices(args, col, ctx, M, ncopies):
    """-> dict(search objects per copy, oracle factory inputs).  `ncopies` resident copies of the rank's index."""
    cfg = args.quant
    dev_codes = None
    pq = None
    cents = col["centroids"].cpu().numpy()
    offs = col["offsets"].cpu().numpy().astype(np.uint64)
    ids = col["list_ids"].cpu().numpy().astype(np.uint32)
    docs_pairs = _pairs(col["docs"])
    out = {"cents": cents, "offs": offs, "ids": ids, "docs_pairs": docs_pairs, "graph": None, "cgraph": None}
    if cfg in ("pq", "spann"):
        pq = M.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy(), ctx=ctx)
        dev_codes = pq.quantize(col["X"])  # device -> device, bit-exact with the reference's quantize
        ctx.sync()
        out["codes"] = dev_codes
    out["pq"] = pq
    objs = []
    if cfg == "hnsw":
        g = build_hnsw_arrays(col["X"])
        out["graph"] = g
        for _ in range(ncopies):
            objs.append(M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], col["X"],
                                         M.NoQuantizer(args.dim), ctx=ctx))
        out["objs"] = objs
        return out
    if cfg == "spann":
        out["cgraph"] = build_hnsw_arrays(col["centroids"], Msz=32, seed=3)
    for _ in range(ncopies):
        if cfg == "flat":
            ivf = M.BlockBasedIvf(cents, offs, ids, col["X"], M.NoQuantizer(args.dim), doc_ids=docs_pairs, ctx=ctx)
        else:
            ivf = M.BlockBasedIvf(cents, offs, ids, dev_codes, pq, doc_ids=docs_pairs, ctx=ctx)
        if cfg == "spann":
            g = out["cgraph"]
            hn = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents,
                                  M.NoQuantizer(args.dim), ctx=ctx)
            objs.append(M.Spann(hn, ivf))
        else:
            objs.append(ivf)
    out["objs"] = objs
    return out


def build_oracle(args, col, idx):
    """The CPU oracle over this rank's arrays (test infrastructure; only the cpu_baseline / parity legs call this)."""
    import oracle as O
    cfg = args.quant
    if cfg == "hnsw":
        g = idx["graph"]
        return O, O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], col["X"].cpu().numpy())
    if cfg == "flat":
        return O, O.Ivf(idx["cents"], idx["offs"], idx["ids"], col["X"].cpu().numpy(), doc_ids=idx["docs_pairs"])
    opq = O.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy())
    oivf = O.Ivf(idx["cents"], idx["offs"], idx["ids"], idx["codes"].cpu().numpy(), doc_ids=idx["docs_pairs"], pq=opq)
    if cfg == "spann":
        g = idx["cgraph"]
        ohn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], idx["cents"])
        return O, O.Spann(ohn, oivf)
    return O, oivf


SPANN_EF, SPANN_RATIO = 128, 1e9   # 64 explored centroids, no ratio pruning
HNSW_EF = 128


def oracle_search(args, osearch, Q, k, nprobe, nthreads=0):
    """-> (doc ids (B,k,2) u64, scores (B,k) f32, counts (B,) int) of the oracle for this config."""
    if args.quant == "hnsw":
        od, os_, oc, _ = osearch.search_batch(Q, k, HNSW_EF)
        return od, os_, oc
    if args.quant == "spann":
        return osearch.search_batch(Q, k, SPANN_EF, nprobe, SPANN_RATIO, nthreads) if nthreads else \
            osearch.search_batch(Q, k, SPANN_EF, nprobe, SPANN_RATIO)
    return osearch.search_batch(Q, k, nprobe, nthreads) if nthreads else osearch.search_batch(Q, k, nprobe)


def same_results(od, os_, oc, gi, gs, gc):
    """bit-identical doc ids and scores on the valid prefix of every query; equal counts"""
    oc = np.asarray(oc).astype(np.int64)
    if not np.array_equal(oc, np.asarray(gc).astype(np.int32).astype(np.int64)):
        return False
    for b in range(len(oc)):
        n = max(int(oc[b]), 0)
        if not (np.array_equal(np.asarray(od[b, :n], dtype=np.uint64), np.asarray(gi[b, :n]).view(np.uint64).reshape(-1, 2)) and
                np.array_equal(np.asarray(os_[b, :n], dtype=np.float32).view(np.uint32), np.asarray(gs[b, :n]).view(np.uint32))):
            return False
    return True


def main():
    args = parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, ncores)

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the search path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import muopdb_b200 as M
    from muopdb_b200 import _lib, sharding

    cfg = args.quant
    grow = cfg == "spann"                      # config 5: rows per rank fixed, the collection grows with N
    replicas = cfg == "hnsw" and world > 1     # config 4 does not shard in BASELINE.json: N independent replicas, no collective
    ctx = M.default_context(local_rank)
    col = make_collection(args, device, shard=0 if replicas else rank, nshards=1 if replicas else world, grow=grow)
    nlist = col["nlist"]
    nprobe = args.nprobe if (grow or replicas) else max(args.nprobe // world, 1)
    B = args.batch if (grow or replicas) else args.batch * world   # c3/c2: replicated global batch
    k = args.k
    m = args.dim // args.dsub

    # resident copies of the index: two consecutive steps' inputs must exceed L2
    row_bytes = {"pq": m + 4, "spann": m + 4, "flat": 4 * args.dim + 4, "hnsw": 4 * args.dim}[cfg]
    index_bytes = col["X"].shape[0] * row_bytes
    ncopies = int(min(24, max(1, -(-int(2.3 * L2_BYTES) // index_bytes))))
    idx = build_indices(args, col, ctx, M, ncopies)
    objs = idx["objs"]
    sharded = world > 1 and not replicas
    if sharded:
        sharding.init_comm(ctx)
        ctx.shard_overlap(True)   # batch i's result exchange runs next to batch i+1's kernels; ctx.sync() completes both
    params = M.SearchParams(k, SPANN_EF, False, nprobe, SPANN_RATIO) if cfg == "spann" else None

    Qall = col["Q"]  # (4 * B, dim) distinct query batches
    nbatches = Qall.shape[0] // B
    ext = torch.cuda.ExternalStream(ctx.stream, device=device)
    out_dev = (torch.zeros((B, k, 2), dtype=torch.int64, device=device), torch.zeros((B, k), dtype=torch.float32, device=device),
               torch.zeros((B,), dtype=torch.int32, device=device))
    stats_dev = {}

    def step_device(i, out=out_dev):
        Qb = Qall[(i % nbatches) * B:(i % nbatches + 1) * B]
        o = objs[i % ncopies]
        if cfg == "hnsw":
            r, st = o.ann_search_batch(Qb, k, HNSW_EF, with_stats=True)
            stats_dev["st"] = st
            return (r.doc_ids, r.scores, r.counts)
        if cfg == "spann":
            if sharded:
                o.shard_search_batch(Qb, params, out=out, shared_codebook=True)
            else:
                r = o.search_batch(Qb, params)
                return (r.doc_ids, r.scores, r.counts)
            return out
        if sharded:
            # one collective call: split query encode + all-gather of codes, local shard search, all-gather + merge
            o.shard_search_batch(Qb, k, nprobe, out=out, shared_codebook=True)
        else:
            o.search_batch(Qb, k, nprobe, out=out)
        return out

    def join():
        """main stream waits for the exchange stream (a sharded step's merge runs there)"""
        if sharded:
            _lib.check(ctx.lib.mgpu_stream_signal(ctx.h, ctx.stream), ctx.h)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (the clock sampler runs from here to the end of the e2e leg)
    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.cuda.stream(ext):
        # every resident index copy is searched at least once before the clock starts (an object's first search allocates its
        # per-index scratch, which synchronises the device)
        for i in range(max(args.warmup, 3, ncopies)):
            step_device(i)
    barrier()

    # ---- timed: device-resident queries, the K steps back to back (rotating over the index copies), one CUDA event pair on
    # the library stream around all of them.  Only the dominant kernel is additionally bracketed by events (for the roofline).
    dom_class = _lib.K_HNSW if cfg == "hnsw" else _lib.K_SCAN
    ctx.profile_reset()
    ctx.profile_enable(True, classes=[dom_class])
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(ext):
        e0.record(ext)
        th0 = time.perf_counter()
        for i in range(args.steps):
            step_device(i)
        host_enqueue_ms = (time.perf_counter() - th0) * 1e3 / args.steps
        join()
        e1.record(ext)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    ctx.profile_enable(False)
    dom_ms, dom_launches = ctx.profile_get(dom_class)
    # per-class breakdown: a second, untimed pass with every class bracketed
    ctx.profile_reset()
    ctx.profile_enable(True)
    nprof = min(args.steps, 10)
    with torch.cuda.stream(ext):
        for i in range(nprof):
            step_device(i)
    barrier()
    ctx.profile_enable(False)
    prof = {nme: ctx.profile_get(c)[0] / nprof for c, nme in enumerate(_lib.KERNEL_CLASS_NAMES)}
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- algorithmic bytes of the dominant kernel, from the last step
    with torch.cuda.stream(ext):
        res = step_device(0, out=tuple(torch.zeros_like(x) for x in out_dev))
    barrier()
    fallbacks = None
    if cfg == "hnsw":
        st = stats_dev["st"].cpu().numpy()
        deg = 32
        dom_bytes = int(st[:, 0].sum()) * args.dim * 4 + int(st[:, 1].sum()) * (16 + 4 * deg)
        rows_per_launch = int(st[:, 0].sum())
    else:
        lists = objs[0].posting_lists if cfg == "spann" else objs[0]
        dom_bytes = lists.last_scan_bytes()
        rows_per_launch = lists.last_scan_rows()
        fallbacks = lists.last_scan_fallbacks()

    # ---- e2e: host buffers through the C ABI (pinned queries H2D, results D2H inside the timed region), nothing subtracted
    Qh = [torch.empty((B, args.dim), dtype=torch.float32).pin_memory() for _ in range(nbatches)]
    for j in range(nbatches):
        Qh[j].copy_(Qall[j * B:(j + 1) * B])
    outs = [(torch.zeros((B, k, 2), dtype=torch.int64).pin_memory(), torch.zeros((B, k), dtype=torch.float32).pin_memory(),
             torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(2)]

    def submit(i):
        o = objs[i % ncopies]
        if cfg == "spann":
            if sharded:
                return o.shard_search_batch_submit(Qh[i % nbatches], params, outs[i & 1], shared_codebook=True)
            return o.search_batch_submit(Qh[i % nbatches], params, outs[i & 1])
        if cfg == "hnsw":
            return o.ann_search_batch_submit(Qh[i % nbatches], k, HNSW_EF, outs[i & 1])
        if sharded:
            return o.shard_search_batch_submit(Qh[i % nbatches], k, nprobe, outs[i & 1], shared_codebook=True)
        return o.search_batch_submit(Qh[i % nbatches], k, nprobe, outs[i & 1])

    def run_pipelined(n):
        prev = None
        for i in range(n):
            tk = submit(i)
            if prev is not None:
                objs[0].search_wait(prev)
            prev = tk
        if prev is not None:
            objs[0].search_wait(prev)

    def run_blocking(n):
        for i in range(n):
            o = objs[i % ncopies]
            if cfg == "hnsw":
                o.ann_search_batch(Qh[i % nbatches].numpy(), k, HNSW_EF)
            elif cfg == "spann":
                (o.shard_search_batch(Qh[i % nbatches], params, out=outs[0]) if sharded else o.search_batch(Qh[i % nbatches].numpy(), params))
            elif sharded:
                o.shard_search_batch(Qh[i % nbatches], k, nprobe, out=outs[0], shared_codebook=True)
            else:
                o.search_batch(Qh[i % nbatches], k, nprobe, out=outs[0])

    pipelined = True   # every config has a pipelined submit / wait form
    run_blocking(3)
    barrier()
    t0 = time.perf_counter()
    run_blocking(args.steps)
    e2e_blocking_s = time.perf_counter() - t0
    barrier()
    e2e_s, e2e_mode, e2e_same = e2e_blocking_s, "blocking host-buffer call per batch (H2D, kernels, D2H, synchronise)", None
    if pipelined:
        run_pipelined(4)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps)
        e2e_s = time.perf_counter() - t0
        barrier()
        e2e_mode = ("pipelined host-buffer calls (mgpu_%s_search_submit / mgpu_search_wait, 2 batches in flight): wall clock of %d "
                    "steps, every H2D/D2H copy inside, nothing subtracted" % (
                        ("shard_spann" if cfg == "spann" else "shard_ivf") if sharded else {"spann": "spann", "hnsw": "hnsw"}.get(cfg, "ivf"),
                        args.steps))
        # the pipelined results are the device call's results (same batch, same index copy)
        last = args.steps - 1
        with torch.cuda.stream(ext):
            ref = step_device(last, out=tuple(torch.zeros_like(x) for x in out_dev))
        barrier()
        e2e_same = bool(np.array_equal(ref[0].cpu().numpy(), outs[last & 1][0].numpy()))
    tt = torch.tensor([e2e_s, e2e_blocking_s], dtype=torch.float64, device=device)
    if world > 1:   # the job is as slow as its slowest rank
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s, e2e_blocking_s = float(tt[0].item()), float(tt[1].item())
    clocks = sampler.stop()

    # ---- recall@10 vs exact brute force (first batch), merged result
    with torch.cuda.stream(ext):
        res = step_device(0, out=tuple(torch.zeros_like(x) for x in out_dev))
    barrier()
    got = res[0][:, :, 0].clone()
    if sharded:
        # ground truth over the whole collection: gather the per-shard exact top-k ids, let the owning rank compute the exact
        # distance of every gathered id, reduce with MIN, take the k best
        gt_local = exact_topk(col["X"], col["docs"], Qall[:B], k)
        gl = [torch.zeros_like(gt_local) for _ in range(world)]
        dist.all_gather(gl, gt_local)
        owned = torch.cat(gl, 1)  # (B, world*k) candidate doc ids
        mask = (owned % world) == rank
        pos = torch.div(owned, world, rounding_mode="floor").clamp_(0, col["X"].shape[0] - 1)   # doc = row * world + rank
        dloc = torch.full(owned.shape, float("inf"), device=device)
        for b0 in range(0, B, 256):
            xb = col["X"][pos[b0:b0 + 256]]
            dloc[b0:b0 + 256] = ((xb - Qall[b0:b0 + 256, None, :]) ** 2).sum(-1)
        dloc[~mask] = float("inf")
        dist.all_reduce(dloc, op=dist.ReduceOp.MIN)
        gt = torch.gather(owned, 1, torch.topk(dloc, k, dim=1, largest=False).indices)
    else:
        gt = exact_topk(col["X"], col["docs"], Qall[:B], k)
    hits = (got[:, :, None] == gt[:, None, :]).any(-1).float().sum().item()
    recall = hits / (B * k)

    nq_job = B * (world if (replicas or (grow and False)) else 1)   # replicas answer different copies of the same batch
    value = nq_job * args.steps / (dev_ms / 1e3)
    e2e_value = nq_job * args.steps / e2e_s
    peak, peak_src = measured_peak_gbs()
    dom_avg_ms = dom_ms / max(args.steps, 1)   # per step: HNSW launches a register kernel + a (normally empty) generic one
    achieved = dom_bytes / (dom_avg_ms / 1e3) / 1e9 if dom_avg_ms > 0 else 0.0
    # the library reports which of a class's alternative kernels served the last call (mgpu_last_kernel): never a guess
    kernel_name = ctx.last_kernel(dom_class)
    shape_key = f"{args.config}:n{world}"
    if cfg == "pq":
        workload = f"IVF-PQ (m={m}, 8-bit) {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, batch={args.batch}/GPU, k={k}"
        shard_s = f"doc_id mod {world}; per shard nlist={nlist}, nprobe={nprobe}; global batch {B} replicated"
    elif cfg == "flat":
        workload = f"IVF flat-L2 {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, batch={args.batch}/GPU, k={k}"
        shard_s = f"doc_id mod {world}; per shard nlist={nlist}, nprobe={nprobe}; global batch {B} replicated"
    elif cfg == "hnsw":
        workload = f"HNSW M=32 ef_search={HNSW_EF} {args.n}x{args.dim}, batch={args.batch}, k={k} (kNN-built graph, {idx['graph']['num_layers']} layers)"
        shard_s = "not sharded (BASELINE.json config 4 is a 1-GPU config): %d independent replicas, no collective" % world
    else:
        workload = (f"SPANN: centroid HNSW (M=32, ef={SPANN_EF}) over {args.nlist} centroids + PQ (m={m}) lists, {args.n} rows per GPU "
                    f"({args.n * world} rows over {world} GPUs), {nprobe} explored centroids, no ratio pruning, batch={args.batch}, k={k}")
        shard_s = f"doc_id mod {world}: every rank holds one {args.n}-row shard with its own {nlist} lists; batch {B} replicated"

    out = {
        "metric": args.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3, ncopies),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"pq": "u8 codes, u32 fixed-point LUT sums, f32 exact re-rank", "spann": "f32 graph distances; u8 codes, u32 LUT sums, f32 re-rank",
                  "flat": "f32", "hnsw": "f32"}[cfg], "data": "synthetic",
        "config": {"workload": workload, "sharding": shard_s,
                   "l2": "no flush kernel: consecutive steps search different resident copies of the index "
                         f"({ncopies} copies x {index_bytes / 1e6:.0f} MB > 2 x 126 MB L2)" if ncopies > 1 else
                         f"no flush kernel: the {index_bytes / 1e6:.0f} MB index exceeds the 126 MB L2",
                   "recall_at_10": recall,
                   "data_distribution": "2048 Gaussian blobs in a 32-d latent space embedded in 768-d + isotropic noise 0.02, seed %d" % args.seed},
        "recall_at_10": recall,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * args.dim * 4, "d2h_bytes_per_step": B * k * 20 + B * 4,
                "ms_per_step": e2e_s * 1e3 / args.steps, "mode": e2e_mode,
                "pipelined_equals_device_call_on_last_batch": e2e_same,
                "blocking": {"value": nq_job * args.steps / e2e_blocking_s, "ms_per_step": e2e_blocking_s * 1e3 / args.steps}},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(shape_key),
                     "algorithmic_bytes_per_launch": dom_bytes, "rows_per_launch": rows_per_launch,
                     "launch_ms": dom_avg_ms, "launches": int(dom_launches),
                     "exact_fallback_queries_last_step": fallbacks},
        "kernel_ms_per_step": prof,
    }

    # ---- CPU baseline + parity.  N == 1: the oracle on a bounded sample of the same workload, timed, and the GPU result
    # compared bit for bit on that sample.  N > 1: every rank runs the oracle on ITS shard for the first `parity_queries`
    # queries, rank 0 merges the per-shard oracle results with the oracle's merge (snapshot.rs:49-63) and compares them with the
    # GPU's merged answer bit for bit.
    gi, gs, gc = res[0].cpu().numpy().view(np.uint64), res[1].cpu().numpy(), res[2].cpu().numpy()
    if not args.no_cpu_baseline and world == 1:
        O, osearch = build_oracle(args, col, idx)
        Qcpu = Qall.cpu().numpy()
        nthreads = ncores
        per_call = max(nthreads * 2, 1)
        while B % per_call:
            per_call -= 1  # slices must tile the first batch exactly (its results are compared with the GPU's)
        done, secs, rr = cpu_leg(lambda q: oracle_search(args, osearch, q, k, nprobe, nthreads), Qcpu, args.cpu_seconds, per_call, keep=B)
        ok, off = True, 0
        for od, os_, oc in rr:
            nqs = od.shape[0]
            ok = ok and same_results(od, os_, oc, gi[off:off + nqs], gs[off:off + nqs], gc[off:off + nqs])
            off += nqs
        out["cpu_baseline"] = {"value": done / secs, "unit": UNIT, "cores": nthreads, "kind": "port",
                               "sample": f"{done} queries (the {Qcpu.shape[0]} bench queries, cycled), {secs:.1f} s, C oracle (one query per "
                                         f"thread" + (", query re-quantized per probed list as in ivf/block_based/index.rs:193)" if cfg in ("pq", "spann") else ")"),
                               "parity_with_gpu_on_sample": bool(ok), "parity_queries": int(off)}
    elif not args.no_cpu_baseline and sharded:
        O, osearch = build_oracle(args, col, idx)
        ns = min(args.parity_queries, B)
        od, os_, oc = oracle_search(args, osearch, Qall[:ns].cpu().numpy(), k, nprobe)
        pack = [None] * world if rank == 0 else None
        dist.gather_object((np.asarray(od), np.asarray(os_), np.asarray(oc)), pack, dst=0)
        if rank == 0:
            ok = True
            for b in range(ns):
                live = [p for p in pack if int(p[2][b]) >= 0]     # a shard answering None contributes nothing
                if live:
                    d = np.concatenate([np.asarray(p[0][b, :int(p[2][b])]).reshape(-1, 2) for p in live])
                    sc = np.concatenate([np.asarray(p[1][b, :int(p[2][b])], dtype=np.float32) for p in live])
                    md, ms = O.merge_topk(d, sc, k)
                else:
                    md, ms = [], np.zeros(0, dtype=np.float32)
                nn = len(ms)
                gotd = [int(lo) | (int(hi) << 64) for lo, hi in gi[b, :nn].reshape(-1, 2)]
                ok = ok and int(np.int32(gc[b])) == (nn if live else -1) and gotd == [int(x) for x in md] and \
                    np.array_equal(gs[b, :nn].view(np.uint32), np.asarray(ms, dtype=np.float32).view(np.uint32))
            out["parity_with_oracle_on_sample"] = {"ok": bool(ok), "queries": int(ns), "shards": world,
                                                   "how": "per-shard C oracle on every rank + oracle merge_topk on rank 0 vs the GPU's "
                                                          "all-gathered + merged result, bit for bit"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args, ncores):
    """The reference's CPU algorithm (C oracle; the Rust crate cannot be built here: no rustc/cargo) on all host cores."""
    import torch
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    world = max(args.gpus, 1)
    col = make_collection(args, dev, shard=0, nshards=1, grow=args.quant == "spann")  # the CPU reference searches one whole collection
    idx = {"cents": col["centroids"].cpu().numpy(), "offs": col["offsets"].cpu().numpy().astype(np.uint64),
           "ids": col["list_ids"].cpu().numpy().astype(np.uint32), "docs_pairs": _pairs(col["docs"]), "graph": None, "cgraph": None}
    import oracle as O
    if args.quant in ("pq", "spann"):
        opq = O.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy())

        class _C:   # build_oracle reads idx["codes"].cpu().numpy()
            def __init__(self, a): self.a = a
            def cpu(self): return self
            def numpy(self): return self.a
        idx["codes"] = _C(opq.quantize(col["X"].cpu().numpy()))
    if args.quant == "hnsw":
        idx["graph"] = build_hnsw_arrays(col["X"])
    if args.quant == "spann":
        idx["cgraph"] = build_hnsw_arrays(col["centroids"], Msz=32, seed=3)
    O, osearch = build_oracle(args, col, idx)
    Qh = col["Q"].cpu().numpy()
    per_step = args.ref_queries_per_step or ncores * 4
    k, nprobe = args.k, args.nprobe
    for i in range(max(args.warmup, 1)):
        oracle_search(args, osearch, Qh[:min(per_step, 2 * ncores)], k, nprobe, ncores)
    t = 0.0
    for i in range(args.steps):
        lo = (i * per_step) % max(Qh.shape[0] - per_step, 1)
        t0 = time.perf_counter()
        oracle_search(args, osearch, Qh[lo:lo + per_step], k, nprobe, ncores)
        t += time.perf_counter() - t0
    qps = per_step * args.steps / t
    m = args.dim // args.dsub
    sample = f"{per_step} queries per step (bounded sample of the {args.batch}-query batch), {ncores} threads, one query per thread"
    workload = {"pq": f"IVF-PQ (m={m}, 8-bit) {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, batch={args.batch}/GPU, k={k}",
                "flat": f"IVF flat-L2 {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, batch={args.batch}/GPU, k={k}",
                "hnsw": f"HNSW M=32 ef_search={HNSW_EF} {args.n}x{args.dim}, batch={args.batch}, k={k}",
                "spann": f"SPANN: centroid HNSW + PQ (m={m}) lists, one {args.n}-row shard, {nprobe} explored centroids, batch={args.batch}, k={k}"}[args.quant]
    print(json.dumps({
        "impl": "reference", "metric": args.metric, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": sample},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())

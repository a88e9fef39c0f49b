#!/usr/bin/env python
"""bench.py -- QPS @ recall@10 of the batched IVF-PQ search path (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[2], "C3"): 1M x 768 f32 vectors, IVF nlist 4096, PQ m=96 (dsub 8) 8 bit, nprobe 64,
batch 1024, k 10; synthetic seeded Gaussian-blob data.  One "step" = one batch of 1024 queries per GPU
through coarse scoring -> query quantize -> PQ LUT scan -> exact re-rank -> doc-id remap.

  value     whole-job QPS, queries already resident in HBM, device time by CUDA events on the library's stream
  e2e       same through the host-buffer C-ABI call (mgpu_ivf_search with MGPU_HOST): pinned host queries in, results out
  roofline  posting-list scan kernel: algorithmic bytes (rows scanned x (m + 4) B) / event-timed kernel time vs measured HBM
  cpu_baseline  the C oracle (port of the reference's Rust path) on the host cores, bounded sample

N > 1 (torchrun): the 1M index is doc-sharded N ways (doc_id mod N), every rank scans its shard for the replicated
global batch of 1024*N queries with nprobe 64/N of its nlist 4096/N lists, then one NCCL all-gather + merge kernel gives
every rank the merged top-k ("weak": per-GPU work is constant in N).

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


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nlist", type=int, default=4096)
    p.add_argument("--nprobe", type=int, default=64)
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--dsub", type=int, default=8)
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU-baseline sample")
    p.add_argument("--ref-queries-per-step", type=int, default=0, help="--impl reference: queries per step (0 = 4 x threads)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


# ---- synthetic collection (setup, untimed; torch is only the generator/plumbing here) ---------------------------------------
def make_collection(args, device, shard=0, nshards=1):
    """Seeded blobs -> coarse centroids (Lloyd on a sample) -> posting lists -> PQ codebook.  Returns torch tensors on
    `device`.  Posting lists / codebooks are INPUTS of the search path, any valid index will do."""
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

    X = draw(N)
    # queries: fresh draws from the same mixture
    nq_total = args.batch * max(nshards, 1) * 4
    Q = draw(nq_total)
    doc_lo = torch.arange(N, device=device, dtype=torch.int64)
    # PQ codebook shared by all shards: per-subspace Lloyd on 10 000 sampled rows (reference default
    # product_quantization_num_training_rows, rs/config/src/collection.rs:190)
    m, K = D // args.dsub, 256
    samp = X[torch.randperm(N, generator=g, device=device)[:10000]].reshape(10000, m, args.dsub).permute(1, 0, 2).contiguous()
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
    if nshards > 1:
        keep = (doc_lo % nshards) == shard
        Xs, docs = X[keep].contiguous(), doc_lo[keep].contiguous()
    else:
        Xs, docs = X, doc_lo
    del X
    nlist = max(args.nlist // nshards, 1)
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
            time.sleep(0.001)

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


def ncu_traffic():
    """dram bytes per launch of the scan kernel from the committed ncu summary, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "scan_ncu_summary.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---- CPU reference arm --------------------------------------------------------------------------------------------------------
def build_oracle_index(col, args):
    import oracle as O
    X = col["X"]
    cb = col["codebook"].cpu().numpy()
    opq = O.ProductQuantizer(args.dim, args.dsub, 8, cb)
    return O, opq


def cpu_leg(O, oivf, Qh, k, nprobe, seconds, nthreads, per_call, keep):
    """Times the oracle on successive slices of Qh (cycling) until `seconds` of CPU wall time are used; the results of the
    first `keep` queries are returned for the parity check."""
    done, t_used, i = 0, 0.0, 0
    res = []
    oivf.search_batch(Qh[:min(per_call, Qh.shape[0])], k, nprobe, nthreads)  # warm-up (page faults, thread pool)
    while t_used < seconds:
        lo = i % Qh.shape[0]
        q = Qh[lo:lo + per_call]
        t0 = time.perf_counter()
        r = oivf.search_batch(q, k, nprobe, nthreads)
        t_used += time.perf_counter() - t0
        if i < keep:
            res.append(r)
        done += q.shape[0]
        i += per_call
    return done, t_used, res


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

    ctx = M.default_context(local_rank)
    col = make_collection(args, device, shard=rank, nshards=world)
    nlist, nprobe = col["nlist"], max(args.nprobe // world, 1)
    B = args.batch * world  # replicated global batch
    k = args.k
    m = args.dim // args.dsub

    pq = M.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy(), ctx=ctx)
    codes = pq.quantize(col["X"])  # device -> device, bit-exact with the reference's quantize
    ctx.sync()
    docs_pairs = np.zeros((col["docs"].shape[0], 2), dtype=np.uint64)
    docs_pairs[:, 0] = col["docs"].cpu().numpy().astype(np.uint64)
    ivf = M.BlockBasedIvf(col["centroids"].cpu().numpy(), col["offsets"].cpu().numpy().astype(np.uint64),
                          col["list_ids"].cpu().numpy().astype(np.uint32), codes, pq, doc_ids=docs_pairs, ctx=ctx)
    if world > 1:
        sharding.init_comm(ctx)

    Qall = col["Q"]  # (4 * B, dim) distinct query batches
    nbatches = Qall.shape[0] // B
    ext = torch.cuda.ExternalStream(ctx.stream, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    out_local = (torch.zeros((B, k, 2), dtype=torch.int64, device=device), torch.zeros((B, k), dtype=torch.float32, device=device),
                 torch.zeros((B,), dtype=torch.int32, device=device))
    out_merged = tuple(torch.zeros_like(t) for t in out_local)

    def step_device(i):
        Qb = Qall[(i % nbatches) * B:(i % nbatches + 1) * B]
        if world > 1:
            # one collective call: split query encode + all-gather of codes, local shard search, all-gather + merge
            ivf.shard_search_batch(Qb, k, nprobe, out=out_merged, shared_codebook=True)
            return out_merged
        ivf.search_batch(Qb, k, nprobe, out=out_local)
        return out_local

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (the clock sampler runs from here to the end of the e2e leg)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step_device(i)
    barrier()

    # ---- timed: device-resident queries, per-step CUDA events on the library stream, L2 flushed between steps.  Only the
    # scan kernel (the roofline kernel) is bracketed by events inside the timed region; the other launches run back to back
    # as they do in production.  The per-class breakdown comes from a second, untimed pass with every class bracketed.
    ctx.profile_reset()
    ctx.profile_enable(True, classes=[_lib.K_SCAN])
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scan_bytes = 0
    barrier()
    with torch.cuda.stream(ext):
        for i in range(args.steps):
            flush.zero_()
            ev[i][0].record(ext)
            step_device(i)
            ev[i][1].record(ext)
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ctx.launch_count() - launches0
    ctx.profile_enable(False)
    scan_ms, scan_launches = ctx.profile_get(_lib.K_SCAN)
    ctx.profile_reset()
    ctx.profile_enable(True)
    nprof = min(args.steps, 10)
    with torch.cuda.stream(ext):
        for i in range(nprof):
            flush.zero_()
            step_device(i)
    barrier()
    ctx.profile_enable(False)
    prof = {nme: ctx.profile_get(c)[0] * args.steps / nprof for c, nme in enumerate(_lib.KERNEL_CLASS_NAMES)}
    scan_bytes_per_launch = ivf.last_scan_bytes()  # every batch scans about the same number of rows; this is the last one
    rows_per_launch = ivf.last_scan_rows()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- e2e: host buffers through the C ABI (pinned queries H2D, results D2H inside the timed call)
    Qh = [torch.empty((B, args.dim), dtype=torch.float32).pin_memory() for _ in range(nbatches)]
    for j in range(nbatches):
        Qh[j].copy_(Qall[j * B:(j + 1) * B])
    h_ids = torch.zeros((B, k, 2), dtype=torch.int64).pin_memory()
    h_sc = torch.zeros((B, k), dtype=torch.float32).pin_memory()
    h_cn = torch.zeros((B,), dtype=torch.int32).pin_memory()

    def step_host(i):
        if world == 1:
            ivf.search_batch(Qh[i % nbatches], k, nprobe, out=(h_ids, h_sc, h_cn))
        else:
            # H2D, split encode, local search, exchange + merge, D2H: all inside the one host-buffer collective call
            ivf.shard_search_batch(Qh[i % nbatches], k, nprobe, out=(h_ids, h_sc, h_cn), shared_codebook=True)

    for i in range(3):
        step_host(i)
    barrier()
    e2e_s = 0.0
    for i in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()
        barrier()
        t0 = time.perf_counter()
        step_host(i)
        e2e_s += time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_blocking_s = float(t.item())
    e2e_s, e2e_mode = e2e_blocking_s, "blocking call per batch (mgpu_ivf_search, MGPU_HOST), L2 flushed + synchronised between calls"
    e2e_same, e2e_wall_ms, e2e_flush_ms = None, None, None

    if True:
        # pipelined form of the same host-buffer call (mgpu_ivf_search_submit / mgpu_search_wait): two batches in flight, so
        # batch i+1's H2D and batch i-1's D2H overlap batch i's kernels.  Every step still uploads its queries from pinned
        # memory and downloads its results.  The L2 flush runs on the library stream between batches; its event-timed
        # duration is subtracted from the wall clock (it is serial with the kernels).
        outs = [(h_ids, h_sc, h_cn), (torch.zeros_like(h_ids).pin_memory(), torch.zeros_like(h_sc).pin_memory(),
                                      torch.zeros_like(h_cn).pin_memory())]
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

        def run_pipelined(n, timed):
            prev = None
            for i in range(n):
                with torch.cuda.stream(ext):
                    if timed:
                        fev[i][0].record(ext)
                    flush.zero_()
                    if timed:
                        fev[i][1].record(ext)
                if world > 1:
                    tk = ivf.shard_search_batch_submit(Qh[i % nbatches], k, nprobe, outs[i & 1], shared_codebook=True)
                else:
                    tk = ivf.search_batch_submit(Qh[i % nbatches], k, nprobe, outs[i & 1])
                if prev is not None:
                    ivf.search_wait(prev)
                prev = tk
            ivf.search_wait(prev)

        run_pipelined(4, False)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps, True)
        wall = time.perf_counter() - t0
        barrier()
        flush_s = sum(a.elapsed_time(b) for a, b in fev) / 1e3
        e2e_s = wall - flush_s
        if world > 1:   # the job is as slow as its slowest rank
            t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_mode = ("pipelined host-buffer calls (mgpu_%sivf_search_submit/mgpu_search_wait, 2 batches in flight); wall clock of "
                    "%d steps minus the event-timed L2 flushes (%.3f ms/step) that run between batches"
                    % ("shard_" if world > 1 else "", args.steps, flush_s * 1e3 / args.steps))
        # the pipelined results are the blocking call's results
        last = args.steps - 1
        ref = ivf.shard_search_batch(Qh[last % nbatches], k, nprobe) if world > 1 else ivf.search_batch(Qh[last % nbatches], k, nprobe)
        got_ids = outs[last & 1][0].numpy()
        e2e_same = bool(np.array_equal(np.asarray(ref.doc_ids).view(np.int64).reshape(got_ids.shape), got_ids))
        e2e_wall_ms, e2e_flush_ms = wall * 1e3 / args.steps, flush_s * 1e3 / args.steps

    clocks = sampler.stop()

    # ---- recall@10 vs exact brute force (first batch), merged result
    res = step_device(0)
    barrier()
    got = res[0][:, :, 0].clone()
    if world > 1:
        # ground truth needs the whole collection: gather the per-shard exact top-k and merge on rank 0
        gt_local = exact_topk(col["X"], col["docs"], Qall[:B], k)
        gl = [torch.zeros_like(gt_local) for _ in range(world)]
        dist.all_gather(gl, gt_local)
        # distances are needed for an exact merge; recompute them per shard would need X: approximate by voting is
        # wrong, so compute exact distances of the gathered ids on the ranks that own them
        owned = torch.cat(gl, 1)  # (B, world*k) candidate doc ids
        mask = (owned % world) == rank
        # local position of an owned doc id: docs are arange filtered by modulo -> index = id // world
        pos = torch.div(owned, world, rounding_mode="floor").clamp_(0, col["X"].shape[0] - 1)
        dloc = torch.full(owned.shape, float("inf"), device=device)
        for b0 in range(0, B, 256):
            xb = col["X"][pos[b0:b0 + 256]]  # (256, world*k, dim)
            dloc[b0:b0 + 256] = ((xb - Qall[b0:b0 + 256, None, :]) ** 2).sum(-1)
        dloc[~mask] = float("inf")
        dist.all_reduce(dloc, op=dist.ReduceOp.MIN)
        gt = torch.gather(owned, 1, torch.topk(dloc, k, dim=1, largest=False).indices)
    else:
        gt = exact_topk(col["X"], col["docs"], Qall[:B], k)
    hits = (got[:, :, None] == gt[:, None, :]).any(-1).float().sum().item()
    recall = hits / (B * k)

    value = B * args.steps / (dev_ms / 1e3)
    e2e_value = B * args.steps / e2e_s
    peak, peak_src = measured_peak_gbs()
    scan_avg_ms = scan_ms / max(scan_launches, 1)
    achieved = scan_bytes_per_launch / (scan_avg_ms / 1e3) / 1e9 if scan_avg_ms > 0 else 0.0

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 codes, u32 fixed-point LUT sums, f32 exact re-rank", "data": "synthetic",
        "config": {"workload": f"IVF-PQ (m={m}, 8-bit) {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, "
                               f"batch={args.batch}/GPU, k={k}",
                   "sharding": f"doc_id mod {world}; per shard nlist={nlist}, nprobe={nprobe}; global batch {B} replicated",
                   "l2": "flushed between timed steps (256 MiB write)", "recall_at_10": recall,
                   "data_distribution": "2048 Gaussian blobs in a 32-d latent space embedded in 768-d + isotropic noise 0.02, seed %d" % args.seed},
        "recall_at_10": recall,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * args.dim * 4, "d2h_bytes_per_step": B * k * 20 + B * 4,
                "ms_per_step": e2e_s * 1e3 / args.steps, "mode": e2e_mode,
                "wall_ms_per_step_incl_flush": e2e_wall_ms, "flush_ms_per_step": e2e_flush_ms,
                "pipelined_equals_blocking_on_last_batch": e2e_same,
                "blocking": {"value": B * args.steps / e2e_blocking_s, "ms_per_step": e2e_blocking_s * 1e3 / args.steps}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_scan<PQ_FAST,3> (posting-list LUT scan)", "achieved": achieved, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                     "algorithmic_bytes_per_launch": scan_bytes_per_launch, "rows_per_launch": rows_per_launch,
                     "launch_ms": scan_avg_ms, "launches": int(scan_launches)},
        "kernel_ms_per_step": {nme: v / args.steps for nme, v in prof.items()},
    }

    # ---- CPU baseline (rank 0, N == 1 only): the oracle on a bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        opq = O.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy())
        oivf = O.Ivf(col["centroids"].cpu().numpy(), col["offsets"].cpu().numpy().astype(np.uint64),
                     col["list_ids"].cpu().numpy().astype(np.uint32), codes.cpu().numpy(), doc_ids=docs_pairs, pq=opq)
        Qcpu = Qall.cpu().numpy()
        nthreads = ncores
        per_call = max(nthreads * 2, 1)
        while B % per_call:
            per_call -= 1  # slices must tile the first batch exactly (its results are compared with the GPU's)
        done, secs, rr = cpu_leg(O, oivf, Qcpu, k, nprobe, args.cpu_seconds, nthreads, per_call=per_call, keep=B)
        cpu_qps = done / secs
        # parity on the sample: identical doc ids and bit-identical scores
        ok = True
        gi, gs = res[0].cpu().numpy().view(np.uint64), res[1].cpu().numpy()
        off = 0
        for od, os_, oc in rr:
            nqs = od.shape[0]
            ok = ok and np.array_equal(od, gi[off:off + nqs]) and np.array_equal(os_.view(np.uint32), gs[off:off + nqs].view(np.uint32))
            off += nqs
        out["cpu_baseline"] = {"value": cpu_qps, "unit": UNIT, "cores": nthreads, "kind": "port",
                               "sample": f"{done} queries (the {Qcpu.shape[0]} bench queries, cycled), {secs:.1f} s, C oracle (one query per thread, "
                                         f"query re-quantized per probed list as in ivf/block_based/index.rs:193)",
                               "parity_with_gpu_on_sample": bool(ok)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args, ncores):
    """The reference's CPU algorithm (C oracle; the Rust crate cannot be built here: no rustc/cargo) on all host cores."""
    import torch
    import oracle as O
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    world = max(args.gpus, 1)
    col = make_collection(args, dev, shard=0, nshards=1)  # the CPU reference searches the whole collection
    opq = O.ProductQuantizer(args.dim, args.dsub, 8, col["codebook"].cpu().numpy())
    Xh = col["X"].cpu().numpy()
    codes = opq.quantize(Xh)
    docs_pairs = np.zeros((Xh.shape[0], 2), dtype=np.uint64)
    docs_pairs[:, 0] = col["docs"].cpu().numpy().astype(np.uint64)
    oivf = O.Ivf(col["centroids"].cpu().numpy(), col["offsets"].cpu().numpy().astype(np.uint64),
                 col["list_ids"].cpu().numpy().astype(np.uint32), codes, doc_ids=docs_pairs, pq=opq)
    Qh = col["Q"].cpu().numpy()
    per_step = args.ref_queries_per_step or ncores * 4
    k, nprobe = args.k, args.nprobe
    for i in range(max(args.warmup, 1)):
        oivf.search_batch(Qh[:min(per_step, 2 * ncores)], k, nprobe, ncores)
    t = 0.0
    for i in range(args.steps):
        lo = (i * per_step) % max(Qh.shape[0] - per_step, 1)
        t0 = time.perf_counter()
        oivf.search_batch(Qh[lo:lo + per_step], k, nprobe, ncores)
        t += time.perf_counter() - t0
    qps = per_step * args.steps / t
    m = args.dim // args.dsub
    sample = f"{per_step} queries per step (bounded sample of the 1024-query batch), {ncores} threads, one query per thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IVF-PQ (m={m}, 8-bit) {args.n}x{args.dim}, nlist={args.nlist}, nprobe={args.nprobe}, "
                               f"batch={args.batch}/GPU, k={k}", "sample": sample},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())

// hnsw_spec.cu -- HNSW beam search, speculate-then-replay (flat rows; the fast path of mgpu_hnsw_search / Spann centroid search).
//   BlockBasedHnsw::ann_search   rs/index/src/hnsw/block_based/index.rs:159-210
//   BlockBasedHnsw::search_layer rs/index/src/hnsw/block_based/index.rs:212-287
//
// The reference's loop is strictly sequential: pop the nearest candidate, fetch its edges, test/mark visited, score the fresh
// neighbours, admit them one by one, pop again.  On a GPU every step of that chain is a memory round trip, and one expansion
// cost ~8 us in hnsw.cu's register kernel (edge row -> visited atomics -> row gathers -> admissions, three barriers).  But
// the chain is predictable: the entry behind the candidate just popped IS the next pop in ~90 % of the expansions (measured
// on the oracle).  So one CTA (8 warps) per query works in batches:
//
//   SPECULATE (all warps, parallel)  the first T entries of the candidate list are assumed to be the next T pops: their
//       edge rows are fetched together, every neighbour that is not in the visited set NOW is scored (rows prefetched to L2 in
//       one burst, half-warp per row, bit-exact 16-lane order), keys parked in shared memory;
//   REPLAY (warp 0, sequential, exact)  the reference's loop runs on registers and shared memory only: pop (with the strict
//       '>' stop rule); if the popped id is the next speculated one, its edges are walked in stored order, a neighbour is
//       fresh iff inserting it into the visited set succeeds (the set only grows, so fresh-now implies scored-at-speculation),
//       admissions happen one by one exactly like index.rs:255-281.  The first pop that was not speculated (a neighbour admitted
//       meanwhile jumped the queue) ends the batch -- nothing is lost but the unused scores.
//
// Same total orders as hnsw.cu (W ascending (key, id), C ascending (key, ~id)); here both are sorted arrays in shared memory and
// an expansion updates them ONCE with a warp-parallel merge (every element finds its final position by binary search) instead
// of one sorted insertion per admitted neighbour: a lone warp retires a dependent instruction every ~5 cycles, and the
// sequential admissions (about a thousand cycles each on the register lists) were 60 % of a batch.  The decisions that ARE
// order dependent -- which neighbours get in, given that the furthest key shrinks with every admission -- run on the 32
// largest entries of W, one per lane (about 20 instructions per admission).  Same counters (distance evaluations = fresh neighbours at replay, expansions = popped candidates with edges), so results AND
// out_stats equal the oracle's.  The visited set is a per-query open-addressing hash table in shared memory (no global
// bitmap, no atomics to L2); a query that outgrows it, or whose candidate list overflows its registers, raises err_flags[q] and
// is redone by k_hnsw_search.
#include "hnsw_device.cuh"

#define HS_TMAX 8         /* speculated candidates per batch: T <= warps, T <= HS_TMAX */
#define HS_HASH_LOG 13
#define HS_BITMAP_MAX_N 262144u   /* up to this many points the visited set is a plain bitmap in shared memory (32 KB) */
#define HS_HASH_CAP (1u << HS_HASH_LOG)
#define HS_EMPTY 0xFFFFFFFFu

#ifdef MGPU_SCAN_DBG
// experiment build only (make DBG=1): [0] batches [1] replayed expansions [2] rows scored [3] cycles speculate-1 [4] speculate-2
// [5] replay (thread 0 of every CTA)
__device__ unsigned long long g_hs_dbg[8];
#define HSD_T(v) const long long v = clock64()
#define HSD_ADD(i, x) do { if (tid == 0) atomicAdd(&g_hs_dbg[i], (unsigned long long)(x)); } while (0)
#else
#define HSD_T(v)
#define HSD_ADD(i, x)
#endif

__device__ __forceinline__ uint32_t hs_hash(uint32_t e) { return (e * 2654435761u) >> (32 - HS_HASH_LOG); }
__device__ __forceinline__ bool hs_contains(const uint32_t *hash, uint32_t e) {
  uint32_t h = hs_hash(e);
  for (;;) {
    const uint32_t v = hash[h];
    if (v == e) return true;
    if (v == HS_EMPTY) return false;
    h = (h + 1) & (HS_HASH_CAP - 1);
  }
}
// 1: newly inserted, 0: already present
__device__ __forceinline__ int hs_insert(uint32_t *hash, uint32_t e) {
  uint32_t h = hs_hash(e);
  for (;;) {
    const uint32_t old = atomicCAS(&hash[h], HS_EMPTY, e);
    if (old == HS_EMPTY) return 1;
    if (old == e) return 0;
    h = (h + 1) & (HS_HASH_CAP - 1);
  }
}

// visited set in shared memory: bitmap over the point ids when it fits, else the hash table
struct HsVisited {
  uint32_t *mem; bool bitmap;
  __device__ __forceinline__ bool contains(uint32_t e) const { return bitmap ? ((mem[e >> 5] >> (e & 31)) & 1u) != 0 : hs_contains(mem, e); }
  __device__ __forceinline__ int insert(uint32_t e) const {   // 1: newly inserted
    if (bitmap) { const uint32_t bit = 1u << (e & 31); return (atomicOr(&mem[e >> 5], bit) & bit) ? 0 : 1; }
    return hs_insert(mem, e);
  }
};

template <int METRIC, int EPL, int HS_WARPS, int HS_T>
__global__ void __launch_bounds__(HS_WARPS * 32, HS_WARPS == 16 ? 1 : 2) k_hnsw_spec(HnswDev g, HnswSearchArgs a, uint32_t *__restrict__ err_flags, uint32_t vis_words) {
  constexpr int HS_THREADS = HS_WARPS * 32;
  static_assert(HS_T <= HS_WARPS && HS_T <= HS_TMAX, "one warp fetches the edges of one speculated candidate");
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t ef = a.ef;
  constexpr int CAP = EPL * 32;
  float *sq = (float *)smem;                                   // dim floats
  uint32_t *hash = (uint32_t *)(sq + ((g.dim + 3) & ~3u));     // visited set: vis_words words (bitmap: ceil(n/32); hash: HS_HASH_CAP)
  const HsVisited vis{hash, g.n <= HS_BITMAP_MAX_N};
  uint32_t *sp_edge = hash + vis_words;                        // [T][32] neighbour ids (HS_EMPTY = no edge)
  uint32_t *sp_key = sp_edge + HS_T * 32;                      // [T][32] distance keys of the neighbours scored this batch
  uint32_t *work = sp_key + HS_T * 32;                         // (t << 5 | j) of the rows to score
  uint32_t *sp_id = work + HS_T * 32;                          // [T] speculated candidate ids
  uint32_t *sp_deg = sp_id + HS_T;                             // [T] stored edges of each
  int *st = (int *)(sp_deg + HS_T);                            // [0] nspec [1] stop [2] work count [3] next entry point [4] abort
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = blockIdx.x;
  for (uint32_t d = tid; d < g.dim; d += HS_THREADS) sq[d] = a.Q[(size_t)q * g.dim + d];
  for (uint32_t i = tid; i < vis_words; i += HS_THREADS) hash[i] = vis.bitmap ? 0u : HS_EMPTY;
  if (tid == 0) { st[2] = 0; st[4] = 0; }
  __syncthreads();

  unsigned long long n_dist = 0, n_expand = 0;  // warp 0 (uniform)
  uint32_t nvis = 0;

  // bit-exact NoQuantizer::distance by a half-warp: lane h owns lane-accumulator h of l2.rs:30-68 / dot_product.rs:38-71
  auto flat_distance_halfwarp = [&](uint32_t pid) -> float {
    const float *row = (const float *)g.rows + (size_t)pid * g.dim;
    const int h = lane & 15;
    const int n = (int)g.dim;
    float ret = 0.0f;
    int p = 0;
    const bool go16 = METRIC == MGPU_L2 ? (n / 16 > 0) : (n > 16);
    if (go16) {
      const int chunks = n / 16;
      float acc = 0.0f;
      for (int c0 = 0; c0 < chunks; c0 += 48) {
        float y[48];
#pragma unroll
        for (int i = 0; i < 48; i++) y[i] = (c0 + i < chunks) ? __ldg(row + (c0 + i) * 16 + h) : 0.0f;
#pragma unroll
        for (int i = 0; i < 48; i++) {
          if (c0 + i < chunks) {
            float x = sq[(c0 + i) * 16 + h];
            if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y[i]); acc = __fadd_rn(acc, __fmul_rn(d, d)); }
            else acc = __fadd_rn(acc, __fmul_rn(x, y[i]));
          }
        }
      }
      float s = -0.0f;
      const int basel = lane & 16;
#pragma unroll
      for (int l = 0; l < 16; l++) s = __fadd_rn(s, __shfl_sync(0xffffffffu, acc, basel + l));
      ret = __fadd_rn(ret, s);
      p = chunks * 16;
    }
    if (p < n) ret = ref_tail<METRIC>(PtrAcc{sq}, PtrAcc{row}, p, n, ret);
    if (METRIC == MGPU_L2) return sqrtf(ret);  // NoQuantizer::distance -> D::calculate (noq/mod.rs:44-51)
    return -ret;
  };

  RegList<EPL> W, Cd;  // only warp 0's copies are meaningful: W ascending (key, id), C ascending (key, ~id), lane-blocked registers
  int nW = 0, nC = 0;
  bool overflow = false;

  // warp 0: publish the next batch -- the first min(T, nC) entries of the candidate list
  auto plan = [&](int stop) {
    int ns = 0;
    if (!stop) {
      if (nC == 0) stop = 1;
      else {
        ns = nC < HS_T ? nC : HS_T;
#pragma unroll
        for (int t = 0; t < HS_T; t++) {
          const uint64_t c = Cd.get(t);
          if (lane == 0 && t < ns) sp_id[t] = ~(uint32_t)c;
        }
      }
    }
    if (lane == 0) { st[0] = ns; st[1] = stop; st[2] = 0; }
  };

  uint32_t ep = g.entry_point;
  for (int layer = (int)g.num_layers - 1; layer >= 0; layer--) {
    const uint64_t lvl_s = g.level_offsets[g.num_layers - 1 - layer];
    // entry: set_visited(ep); distance; push to both heaps (index.rs:219-233)
    if (warp == 0) {
      float ed = flat_distance_halfwarp(ep);
      ed = __shfl_sync(0xffffffffu, ed, 0);
      if (lane == 0 && ep < g.n) nvis += (uint32_t)vis.insert(ep);
      nvis = __shfl_sync(0xffffffffu, nvis, 0);
      const uint32_t kd = f2key(ed);
      W.init(); Cd.init();
      W.insert(((uint64_t)kd << 32) | ep);
      Cd.insert(((uint64_t)kd << 32) | (uint32_t)~ep);
      nW = 1; nC = 1;
      n_dist++;
      plan(0);
    }
    for (;;) {
      __syncthreads();                                  // (A) the batch is published
      if (st[1] || st[4]) break;
      const int nspec = st[0];
      HSD_T(d0);
      // ---- SPECULATE 1: edge rows of the speculated candidates, visited lookups, work list, L2 prefetch of the rows
      if (warp < nspec) {
        const uint32_t cur = sp_id[warp];
        uint32_t e = HS_EMPTY, deg = 0;
        if (layer == 0 && g.edges0 != nullptr) {
          if (cur < g.n && (uint32_t)lane < g.deg0) e = __ldg(g.edges0 + (size_t)cur * g.deg0 + lane);
          deg = (uint32_t)__popc(__ballot_sync(0xffffffffu, e != HS_EMPTY));   // stored order is dense from slot 0
        } else {
          long long idx = -1;
          if (layer == 0) idx = cur;
          else {
            const int32_t pos = cur < g.n ? g.upper_dense[(size_t)(layer - 1) * g.n + cur] : -1;
            if (pos >= 0) idx = (long long)pos - (long long)lvl_s;
          }
          uint64_t e_begin = 0, e_end = 0;
          if (idx >= 0 && lvl_s + (uint64_t)idx + 1 < g.n_edge_offsets) {
            e_begin = g.edge_offsets[lvl_s + idx];
            e_end = g.edge_offsets[lvl_s + idx + 1];
          }
          deg = (uint32_t)(e_end - e_begin);               // <= 32 (launcher)
          if ((uint32_t)lane < deg) e = g.edges[e_begin + lane];
        }
        const bool valid = e < g.n;                         // ids >= n are never visited nor scored (as in hnsw.cu)
        const bool unv = valid && !vis.contains(e);
        sp_edge[warp * 32 + lane] = valid ? e : HS_EMPTY;
        if (lane == 0) sp_deg[warp] = deg;
        const unsigned mk = __ballot_sync(0xffffffffu, unv);
        int base = 0;
        if (lane == 0 && mk) base = atomicAdd(&st[2], __popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (unv) {
          work[base + __popc(mk & ((1u << lane) - 1))] = (uint32_t)(warp * 32 + lane);
          const char *rowp = (const char *)g.rows + (size_t)e * g.dim * 4;
          const uint32_t lines = (g.dim * 4 + 127) / 128;
          for (uint32_t ln = 0; ln < lines; ln++) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + (size_t)ln * 128));
        }
      }
      __syncthreads();                                  // (B)
      HSD_T(d1);
      // ---- SPECULATE 2: distances of the unvisited neighbours, half-warp per row
      {
        const int nwork = st[2];
        const int rounds = (nwork + HS_WARPS * 2 - 1) / (HS_WARPS * 2);
        for (int r = 0; r < rounds; r++) {
          if (r * HS_WARPS * 2 + warp * 2 >= nwork) continue;   // warp-uniform
          const int i = r * HS_WARPS * 2 + warp * 2 + (lane >> 4);
          const uint32_t wi = work[i < nwork ? i : nwork - 1];
          const float d = flat_distance_halfwarp(sp_edge[wi]);
          if (i < nwork && (lane & 15) == 0) sp_key[wi] = f2key(d);
        }
      }
      __syncthreads();                                  // (C)
      HSD_T(d2);
      HSD_ADD(0, 1); HSD_ADD(2, st[2]); HSD_ADD(3, d1 - d0); HSD_ADD(4, d2 - d1);
      // ---- REPLAY: the reference's loop on the speculated data (index.rs:235-281)
      if (warp == 0) {
        int s = 0, stop = 0;
        for (;;) {
          if (s >= nspec) break;
          if (nC == 0) { stop = 1; break; }
          const uint64_t c = Cd.get(0);
          const uint32_t ck = (uint32_t)(c >> 32), cid = ~(uint32_t)c;
          uint32_t fk = (uint32_t)(W.get(nW - 1) >> 32);           // furthest key of W (nW >= 1), kept in a register
          if (ck > fk) { Cd.pop_front(); nC--; stop = 1; break; }   // strictly farther than the furthest: search_layer ends
          if (cid != sp_id[s]) break;                              // not speculated: next batch starts with it
          Cd.pop_front();
          nC--;
          if (sp_deg[s] != 0) {                                    // None => continue
            n_expand++;
            const uint32_t e = sp_edge[s * 32 + lane];
            const bool valid = e != HS_EMPTY;
            // the first occurrence of an id in the edge list is the one that can be fresh (index.rs:255-259)
            const unsigned peers = __match_any_sync(0xffffffffu, e);
            const bool leader = valid && (__ffs(peers) - 1 == lane);
            const int isnew = leader ? vis.insert(e) : 0;
            const unsigned fresh = __ballot_sync(0xffffffffu, isnew);
            const uint32_t kd_lane = sp_key[s * 32 + lane];
            nvis += (uint32_t)__popc(fresh);
            n_dist += (unsigned long long)__popc(fresh);
            // Admissions, sequential in edge order (index.rs:260-281): admit iff kd < furthest || |W| < ef, with the furthest
            // re-read before every neighbour.  While W is full its furthest key never grows, so a fresh neighbour that is not
            // below the furthest key of NOW can never be admitted later in this list: those are dropped in one ballot and
            // only the possible admissions are walked one by one.
            unsigned todo = __ballot_sync(0xffffffffu, isnew && (nW < (int)ef || kd_lane < fk));
            bool changed = false;
            while (todo) {
              const int j = __ffs(todo) - 1;
              todo &= todo - 1;
              const uint32_t kd = __shfl_sync(0xffffffffu, kd_lane, j), ee = __shfl_sync(0xffffffffu, e, j);
              if (kd < fk || nW < (int)ef) {
                if (nC == CAP) {
                  // the entry that falls off the end may only be lost if it can never be expanded
                  const uint32_t lastk = (uint32_t)(Cd.get(CAP - 1) >> 32);
                  if (!(nW == (int)ef && lastk > fk)) overflow = true;
                  nC--;
                }
                Cd.insert(((uint64_t)kd << 32) | (uint32_t)~ee);
                nC++;
                W.insert(((uint64_t)kd << 32) | ee);
                nW++;
                if (nW > (int)ef) nW--;  // pop the furthest (index.rs:277-279)
                fk = (uint32_t)(W.get(nW - 1) >> 32);
                changed = true;
              }
            }
            // candidates strictly farther than the furthest can never be expanded (they only trigger the `break`): truncate
            // the sorted tail once per expansion
            if (changed && nW == (int)ef) nC = Cd.count_le(nC, fk);
          }
          s++;
        }
        if ((!vis.bitmap && nvis > HS_HASH_CAP * 3 / 4) || overflow) { if (lane == 0) st[4] = 1; }
        plan(stop);
        HSD_T(d3);
        HSD_ADD(1, s); HSD_ADD(5, d3 - d2);
      }
    }
    if (st[4]) break;
    // ---- next layer's entry: min_by distance over the sorted working list == W[0] (index.rs:176-181)
    if (layer > 0) {
      if (warp == 0 && lane == 0) st[3] = nW > 0 ? (int)(uint32_t)W.v[0] : (int)ep;
      __syncthreads();
      ep = (uint32_t)st[3];
      __syncthreads();
    }
  }
  if (st[4]) {   // outgrew the visited set / the candidate registers: k_hnsw_search redoes this query
    if (tid == 0) err_flags[q] = 1;
    return;
  }
  // ---- results: working list is already sorted by (distance, point_id); truncate to k, map to doc ids (index.rs:185-204)
  if (warp == 0) {
    const uint32_t cnt = min((uint32_t)nW, a.k);
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const uint32_t i = lane * EPL + e;
      if (i < cnt) {
        const uint64_t w = W.v[e];
        const uint32_t pid = (uint32_t)w, kd = (uint32_t)(w >> 32);
        const uint32_t u = (kd & 0x80000000u) ? (kd ^ 0x80000000u) : ~kd;
        a.out_scores[(size_t)q * a.k + i] = __uint_as_float(u);
        if (a.out_pids) a.out_pids[(size_t)q * a.k + i] = pid;
        if (a.out_docs) {
          mgpu_u128 d;
          if (g.doc_ids) d = g.doc_ids[pid]; else { d.lo = pid; d.hi = 0; }
          a.out_docs[(size_t)q * a.k + i] = d;
        }
      }
    }
    if (lane == 0) {
      a.out_counts[q] = cnt;
      if (a.out_stats) { a.out_stats[2 * (size_t)q] = n_dist; a.out_stats[2 * (size_t)q + 1] = n_expand; }
    }
  }
}

// Applies when: flat rows, ef in 1..224 (register lists), every adjacency list holds <= 32 edges, and the upper layers have the
// dense position map.  *launched = false otherwise (the caller keeps its previous kernels).
int launch_hnsw_spec(mgpu_hnsw *h, const HnswDev &g, const HnswSearchArgs &a, uint32_t *err_flags, bool *launched) {
  mgpu_ctx *ctx = h->ctx;
  *launched = false;
  // MGPU_HNSW_SPEC=0: never; =1: whenever applicable; unset: when the rows do not fit L2 (measured: on a graph of a few
  // thousand L2-resident rows -- the SPANN centroid graph -- one expansion per round trip is already cheap and the replay
  // overhead makes this kernel slower than k_hnsw_search_reg)
  static const int mode = getenv("MGPU_HNSW_SPEC") ? (getenv("MGPU_HNSW_SPEC")[0] == '0' ? 0 : 1) : 2;
  if (mode == 2 && (size_t)h->n * h->dim * 4 <= (size_t)ctx->l2_bytes / 2) return MGPU_OK;
  if (mode == 0 || h->quant != MGPU_QUANT_NONE || a.ef == 0 || a.ef > 224 || h->max_degree > 32 || h->max_degree == 0) return MGPU_OK;
  if (h->num_layers > 1 && !g.upper_dense) return MGPU_OK;
  if (g.edges0 && g.deg0 > 32) return MGPU_OK;
  const uint32_t vis_words = h->n <= HS_BITMAP_MAX_N ? (uint32_t)((h->n + 31) / 32) + 1 : HS_HASH_CAP;
  // few queries (config 4: 256): 16 warps and 8 speculated candidates per CTA -- more rows scored per memory round trip; many
  // queries (the Spann centroid search: 1024): 8 warps / 4 candidates so that the whole batch is resident at once
  const bool wide = a.B <= (uint32_t)ctx->sm_count;   // one 16-warp CTA per SM (its registers allow no second one)
  static const bool t8 = getenv("MGPU_HNSW_T") && atoi(getenv("MGPU_HNSW_T")) == 8;
  const int T = (wide || t8) ? 8 : 4;
  const size_t smem = (size_t)((h->dim + 3) & ~3u) * 4 + (size_t)vis_words * 4 + (size_t)T * 32 * 4 * 3 + T * 8 + 64;
  if (smem > ctx->smem_optin) return MGPU_OK;
  const int epl = a.ef <= 48 ? 2 : (a.ef <= 128 ? 5 : 8);
#ifdef MGPU_SCAN_DBG
  {
    static int nl = 0;
    if (++nl == 8) {
      unsigned long long hd[8];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpyFromSymbol(hd, g_hs_dbg, sizeof(hd));
      fprintf(stderr, "[hnsw spec dbg] 7 launches: batches %llu replayed expansions %llu (%.2f per batch) rows scored %llu (%.1f per batch) | cycles "
                      "per batch: speculate-1 %.0f speculate-2 %.0f replay %.0f\n", hd[0], hd[1], (double)hd[1] / hd[0], hd[2], (double)hd[2] / hd[0],
              (double)hd[3] / hd[0], (double)hd[4] / hd[0], (double)hd[5] / hd[0]);
    }
  }
#endif
  LaunchScope ls(ctx, MGPU_K_HNSW);
#define HS_LAUNCH(MT, E, NWARP, TT)                                                                                   \
  do {                                                                                                                \
    cudaFuncSetAttribute(k_hnsw_spec<MT, E, NWARP, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    k_hnsw_spec<MT, E, NWARP, TT><<<a.B, NWARP * 32, smem, ctx->stream>>>(g, a, err_flags, vis_words);                \
  } while (0)
#define HS_LAUNCH_W(MT, E) do { if (wide) HS_LAUNCH(MT, E, 16, 8); else if (t8) HS_LAUNCH(MT, E, 8, 8); else HS_LAUNCH(MT, E, 8, 4); } while (0)
#define HS_LAUNCH_E(MT) do { if (epl == 2) HS_LAUNCH_W(MT, 2); else if (epl == 5) HS_LAUNCH_W(MT, 5); else HS_LAUNCH_W(MT, 8); } while (0)
  if (h->metric == MGPU_L2) HS_LAUNCH_E(MGPU_L2); else HS_LAUNCH_E(MGPU_DOT);
#undef HS_LAUNCH_E
#undef HS_LAUNCH_W
#undef HS_LAUNCH
  CUDA_TRY(ctx, cudaGetLastError());
  *launched = true;
  return MGPU_OK;
}

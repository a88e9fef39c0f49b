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
//       edge rows are fetched together (one warp each); a neighbour is "fresh if its candidate is replayed" when it is the
//       first occurrence in its own list, is not in the visited set NOW, and is not listed by an earlier candidate of the
//       batch -- one 32-bit mask per candidate.  Every such neighbour is scored (rows prefetched to L2 in one burst, half-warp
//       per row, all 48 loads of a row block in flight, bit-exact 16-lane order), keys parked in shared memory;
//   REPLAY (warp 0 decides, the CTA updates)  the reference's loop runs on shared memory only: pop (with the strict '>' stop
//       rule); if the popped id is the next speculated one, its fresh mask says which neighbours index.rs:255-281 would walk.
//       While W is full only neighbours below the furthest key can be admitted; those are walked one by one against the 32
//       largest entries of W held one per lane (an admission drops lane 0's entry and inserts the newcomer: ~20 instructions).
//       What the expansion admitted is then merged into the two sorted arrays by ALL warps -- one or two elements per thread,
//       each computing its final index (old index + newcomers below it; binary search for a newcomer) -- three barriers instead
//       of a lone warp looping over 128 + ~100 entries.  The first pop that was not speculated (a neighbour admitted meanwhile
//       jumped the queue) ends the batch -- nothing is lost but the unused scores;
//   MARK (next batch, all warps)  the fresh neighbours of the candidates that WERE replayed enter the visited set; the ones
//       behind the end of the replay leave no trace.
//
// Measured on config 4 (1M x 768, ef 128, 256 queries; cycles per batch of 3.41 replayed expansions): register-list replay
// 2.9k speculate-1 + 11.6k speculate-2 + 16-18k replay (kernel 2.39 ms); CTA merges + precomputed fresh masks: replay 8.7k
// (1.79 ms); 48 loads in flight per row block: speculate-2 ~4k (1.34 ms, 182k QPS; round 1's kernel: 3.24 ms).
//
// Same total orders as hnsw.cu (W ascending (key, id), C ascending (key, ~id)).  Same counters (distance evaluations = fresh
// neighbours at replay, expansions = popped candidates with edges), so results AND out_stats equal the oracle's.  The visited
// set is a bitmap (n <= 262144) or an open-addressing hash table in shared memory (no global bitmap, no atomics to L2); a
// query that outgrows it, or whose candidate array overflows, raises err_flags[q] and is redone by k_hnsw_search.
#include "hnsw_device.cuh"

#include <algorithm>

#define HS_TMAX 8         /* speculated candidates per batch: T <= warps, T <= HS_TMAX */
#define HS_HASH_LOG 13
#define HS_BITMAP_MAX_N 262144u   /* up to this many points the visited set is a plain bitmap in shared memory (32 KB) */
#define HS_HASH_CAP (1u << HS_HASH_LOG)
#define HS_EMPTY 0xFFFFFFFFu

#ifdef MGPU_SCAN_DBG
// experiment build only (make DBG=1): [0] batches [1] replayed expansions [2] rows scored [3] cycles speculate-1 [4] speculate-2
// [5] replay (thread 0 of every CTA)
__device__ unsigned long long g_hs_dbg[16];
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

// ---- W and C: sorted arrays of unique 64-bit composites in shared memory, updated once per expansion.
// number of entries of the ascending array A[0, n) that are < x
__device__ __forceinline__ int hs_lower_bound(const uint64_t *A, int n, uint64_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (A[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// In-place merge into the ascending array A[0, n): every lane whose bit is set in `mask` brings one composite `a` (all distinct,
// and distinct from A's); A[0, n + popc(mask)) is ascending afterwards.  Called by one full warp.  Chunks of 128 old elements
// are taken from the top down, loaded to registers before any of them is stored, so that moving up never overwrites an element
// that is still to be read.  Returns the number of entries of the merged array whose key (high word) is <= le_key.
__device__ __forceinline__ int hs_merge_inplace(uint64_t *A, int n, uint64_t a, unsigned mask, uint32_t le_key) {
  const int lane = threadIdx.x & 31;
  const bool mine = (mask >> lane) & 1u;
  int apos = 0;
  int nle = (mine && (uint32_t)(a >> 32) <= le_key) ? 1 : 0;
  for (unsigned m = mask; m; m &= m - 1) {
    const uint64_t x = shfl64(a, __ffs(m) - 1);
    apos += x < a ? 1 : 0;
  }
  for (int base = n > 0 ? ((n - 1) & ~127) : -1; base >= 0; base -= 128) {
    uint64_t o[4];
    int r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int idx = base + k * 32 + lane;
      o[k] = idx < n ? A[idx] : ~0ull;
      r[k] = 0;
      nle += (idx < n && (uint32_t)(o[k] >> 32) <= le_key) ? 1 : 0;
    }
    __syncwarp();
    for (unsigned m = mask; m; m &= m - 1) {
      const int t = __ffs(m) - 1;
      const uint64_t x = shfl64(a, t);
      int c = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool lt = o[k] < x;
        c += lt ? 1 : 0;
        r[k] += lt ? 0 : 1;
      }
      const int tot = __reduce_add_sync(0xffffffffu, c);
      if (lane == t) apos += tot;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int idx = base + k * 32 + lane;
      if (idx < n && r[k]) A[idx + r[k]] = o[k];
    }
    __syncwarp();
  }
  if (mine) A[apos] = a;
  __syncwarp();
  return __reduce_add_sync(0xffffffffu, nle);
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

template <int METRIC, int HS_WARPS, int HS_T>
__global__ void __launch_bounds__(HS_WARPS * 32, HS_WARPS == 16 ? 1 : (HS_WARPS == 4 ? 4 : (HS_WARPS == 2 ? 7 : 2))) k_hnsw_spec(HnswDev g, HnswSearchArgs a, uint32_t *__restrict__ err_flags, uint32_t vis_words, uint32_t vis_bitmap) {
  constexpr int HS_THREADS = HS_WARPS * 32;
  constexpr int HS_KPT = HS_WARPS >= 4 ? 2 : 4;       // elements per thread in the CTA merges: each half of the CTA holds
  constexpr int HS_MCAP = HS_THREADS / 2 * HS_KPT;    // up to HS_MCAP (>= 128) entries of W or of C
  static_assert(HS_T <= HS_WARPS && HS_T <= HS_TMAX, "one warp fetches the edges of one speculated candidate");
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t ef = a.ef;
  const int WCAP = (int)ef + 32, CCAP = 2 * (int)ef + 96;
  float *sq = (float *)smem;                                   // dim floats
  uint32_t *hash = (uint32_t *)(sq + ((g.dim + 3) & ~3u));     // visited set: vis_words words (bitmap: ceil(n/32); hash: HS_HASH_CAP)
  const HsVisited vis{hash, vis_bitmap != 0};
  uint32_t *sp_edge = hash + vis_words;                        // [T][32] neighbour ids (HS_EMPTY = no edge)
  uint32_t *sp_key = sp_edge + HS_T * 32;                      // [T][32] distance keys of the neighbours scored this batch
  uint32_t *work = sp_key + HS_T * 32;                         // (t << 5 | j) of the rows to score
  uint32_t *sp_id = work + HS_T * 32;                          // [T] speculated candidate ids
  uint32_t *sp_deg = sp_id + HS_T;                             // [T] stored edges of each
  uint32_t *sp_fresh = sp_deg + HS_T;                          // [T] lanes whose neighbour is fresh IF the candidates before it are replayed
  int *st = (int *)(sp_fresh + HS_T);                          // [0] nspec [1] stop [2] work count [3] next entry point [4] abort
                                                               // [5] candidates replayed by the last batch
                                                               // [7] merge request, [8..14] its arguments (see REPLAY)
  uint64_t *Wb = (uint64_t *)(((uintptr_t)(st + 16) + 15) & ~(uintptr_t)15);  // [WCAP] working list, ascending (key, id)
  uint64_t *Cb = Wb + WCAP;                                    // [CCAP] candidates, ascending (key, ~id), live part [head, nC)
  uint64_t *aw = Cb + CCAP;                                    // [32] newcomers that stay in W (W form)
  uint64_t *ac = aw + 32;                                      // [32] newcomers pushed to C (C form)

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
        if (c0 + 48 <= chunks) {
          // all 48 loads of the block are issued before the first is consumed (volatile keeps their order; left to itself the
          // compiler waits after every six -- eight L2 round trips per 3 KB row instead of one)
          const float *rp = row + c0 * 16 + h;
#pragma unroll
          for (int i = 0; i < 48; i++) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(y[i]) : "l"(rp + i * 16));
        } else {
#pragma unroll
          for (int i = 0; i < 48; i++) y[i] = (c0 + i < chunks) ? __ldg(row + (c0 + i) * 16 + h) : 0.0f;
        }
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

  // warp 0 (uniform registers): sizes; the live candidates are Cb[head, nC)
  int nW = 0, head = 0, nC = 0;
  bool overflow = false;

  // warp 0: publish the next batch -- the first min(T, live) entries of the candidate list
  auto plan = [&](int stop) {
    int ns = 0;
    if (!stop) {
      const int live = nC - head;
      if (live == 0) stop = 1;
      else {
        ns = live < HS_T ? live : HS_T;
        if (lane < ns) sp_id[lane] = ~(uint32_t)Cb[head + lane];
      }
    }
    if (lane == 0) { st[0] = ns; st[1] = stop; st[2] = 0; }
  };


  // warp 0: the composites xc of the lanes in `mask` join the candidates; with a full W, candidates strictly farther than its
  // furthest key fk can never be expanded (they only trigger the `break`) and are cut off the sorted tail
  // warp 0: room for na more candidates behind nC -- slide the live part back to the start of the array if need be (moving
  // down: chunks bottom up); false if it still does not fit
  auto candidate_room = [&](int na) -> bool {
    if (nC + na > CCAP && head > 0) {
      const int live = nC - head;
      for (int i0 = 0; i0 < live; i0 += 32) {
        const uint64_t v = i0 + lane < live ? Cb[head + i0 + lane] : 0ull;
        __syncwarp();
        if (i0 + lane < live) Cb[i0 + lane] = v;
        __syncwarp();
      }
      head = 0; nC = live;
    }
    return nC + na <= CCAP;
  };
  auto merge_candidates = [&](uint64_t xc, unsigned mask, uint32_t fk) {
    const int na = __popc(mask);
    if (!candidate_room(na)) { overflow = true; return; }
    const int nle = hs_merge_inplace(Cb + head, nC - head, xc, mask, fk);
    nC = nW == (int)ef ? head + nle : nC + na;
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
      head = 0;
      if (lane == 0) { Wb[0] = ((uint64_t)kd << 32) | ep; Cb[0] = ((uint64_t)kd << 32) | (uint32_t)~ep; }
      __syncwarp();
      nW = 1; nC = 1;
      n_dist++;
      plan(0);
      if (lane == 0) st[5] = 0;
    }
    __syncthreads();                                    // (A) the first batch is published
    for (;;) {
      // ---- the neighbours the last replay found fresh become visited now (set_visited, index.rs:255-259): the replay itself
      // only reads the masks computed below, and candidates it did not reach must leave no trace
      if (warp < st[5] && ((sp_fresh[warp] >> lane) & 1u)) vis.insert(sp_edge[warp * 32 + lane]);
      __syncthreads();                                  // (I)
      if (st[1] || st[4]) break;
      const int nspec = st[0];
      HSD_T(d0);
      // ---- SPECULATE 1: edge rows of the speculated candidates, visited lookups, work list, L2 prefetch of the rows
      uint32_t my_e = HS_EMPTY;
      bool fresh0 = false;
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
        if (!valid) e = HS_EMPTY;
        // the first occurrence of an id in the edge list is the one that can be fresh (index.rs:255-259)
        const unsigned peers = __match_any_sync(0xffffffffu, e);
        fresh0 = valid && (__ffs(peers) - 1 == lane) && !vis.contains(e);
        my_e = e;
        sp_edge[warp * 32 + lane] = e;
        if (lane == 0) sp_deg[warp] = deg;
        const unsigned mk = __ballot_sync(0xffffffffu, fresh0);
        int base = 0;
        if (lane == 0 && mk) base = atomicAdd(&st[2], __popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (fresh0) {
          work[base + __popc(mk & ((1u << lane) - 1))] = (uint32_t)(warp * 32 + lane);
          const char *rowp = (const char *)g.rows + (size_t)e * g.dim * 4;
          const uint32_t lines = (g.dim * 4 + 127) / 128;
          for (uint32_t ln = 0; ln < lines; ln++) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + (size_t)ln * 128));
        }
      }
      __syncthreads();                                  // (B)
      HSD_T(d1);
      // a neighbour that an EARLIER candidate of this batch also lists is visited by the time this candidate is expanded
      if (warp < nspec) {
        bool dup = false;
        for (int t = 0; t < warp; t++) {
#pragma unroll
          for (int l = 0; l < 32; l++) dup |= sp_edge[t * 32 + l] == my_e;
        }
        const unsigned fm = __ballot_sync(0xffffffffu, fresh0 && !dup);
        if (lane == 0) sp_fresh[warp] = fm;
      }
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
      // ---- REPLAY: the reference's loop on the speculated data (index.rs:235-281).  Warp 0 walks the speculated candidates;
      // the other warps wait at (S1).  An expansion that admitted neighbours leaves the two sorted arrays to be updated: that is
      // handed to the whole CTA (st[7] = 1) -- one or two elements per thread, every element computing its own final position
      // -- because a lone warp looping over 128 + 100 entries was 60 % of the batch.
      int s = 0, stop = 0, cadd_le = 0, resume = 0;
      uint32_t e = HS_EMPTY, kd = 0;                    // warp 0: the expansion in progress
      unsigned fresh = 0;
      for (;;) {
        if (warp == 0) {
          int cmd = 0;
          for (;;) {
            if (!resume) {
              if (s >= nspec) break;
              if (nC == head) { stop = 1; break; }
              HSD_T(ta);
              const uint64_t c = Cb[head];
              const uint32_t ck = (uint32_t)(c >> 32), cid = ~(uint32_t)c;
              const uint32_t fk0 = (uint32_t)(Wb[nW - 1] >> 32);     // furthest key of W (nW >= 1)
              if (ck > fk0) { head++; stop = 1; break; }             // strictly farther than the furthest: search_layer ends
              if (cid != sp_id[s]) break;                            // not speculated: next batch starts with it
              head++;                                                // pop
              if (sp_deg[s] == 0) { s++; continue; }                 // None => continue
              n_expand++;
              fresh = sp_fresh[s];
              e = sp_edge[s * 32 + lane];
              kd = sp_key[s * 32 + lane];
              nvis += (uint32_t)__popc(fresh);
              n_dist += (unsigned long long)__popc(fresh);
              HSD_T(tb);
              HSD_ADD(9, tb - ta);
              // ---- admissions (index.rs:260-281): in edge order, admit iff kd < furthest || |W| < ef, furthest re-read every
              // time.  (1) W not full yet: the first ef - nW fresh neighbours are admitted unconditionally, nothing is evicted
              if (nW < (int)ef && fresh) {
                const uint64_t xw = ((uint64_t)kd << 32) | e, xc = ((uint64_t)kd << 32) | (uint32_t)~e;
                const int room = (int)ef - nW;
                const int rank = __popc(fresh & ((1u << lane) - 1));
                const bool take = ((fresh >> lane) & 1u) && rank < room;
                const unsigned admitted = __ballot_sync(0xffffffffu, take);
                const int na = __popc(admitted);
                fresh &= ~admitted;
                if (!candidate_room(na)) overflow = true;
                else if (nW <= HS_MCAP && nC - head <= HS_MCAP) {
                  if (take) { aw[rank] = xw; ac[rank] = xc; }
                  cadd_le = na;
                  if (lane == 0) {
                    st[8] = na; st[9] = nW; st[10] = na; st[11] = head; st[12] = nC - head; st[13] = -1 /* keep every candidate */; st[14] = 0;
                  }
                  nW += na;
                  resume = 1;
                  cmd = 1;
                  break;
                } else {
                  hs_merge_inplace(Wb, nW, xw, admitted, 0u);
                  nW += na;
                  merge_candidates(xc, admitted, (uint32_t)(Wb[nW - 1] >> 32));
                }
              }
            }
            resume = 0;
            // (2) W full: its furthest key never grows, so a fresh neighbour that is not below the furthest key of NOW can never
            // be admitted later in this list -- those are dropped with one ballot.  The rest is walked one by one against the
            // 32 largest entries of W held one per lane (descending): an admission removes lane 0's entry (the furthest) and
            // inserts the newcomer in place.
            if (nW == (int)ef && fresh) {
              uint32_t fk = (uint32_t)(Wb[nW - 1] >> 32);
              unsigned todo = __ballot_sync(0xffffffffu, ((fresh >> lane) & 1u) && kd < fk);
              if (todo) {
                HSD_T(r0);
                HSD_ADD(7, __popc(todo));
                unsigned adm2 = 0;
                uint64_t tk = lane < nW ? Wb[nW - 1 - lane] : 0ull;  // 0 = no entry (no real composite is 0)
                while (todo) {
                  const int j = __ffs(todo) - 1;
                  todo &= todo - 1;
                  const uint32_t kj = __shfl_sync(0xffffffffu, kd, j), ej = __shfl_sync(0xffffffffu, e, j);
                  if (kj < fk) {
                    const uint64_t x = ((uint64_t)kj << 32) | ej;
                    const uint64_t nx = shfl64(tk, (lane + 1) & 31);
                    const uint64_t below = lane == 31 ? 0ull : nx;          // tail after dropping lane 0's entry
                    tk = below > x ? below : ((lane == 0 || tk > x) ? x : tk);
                    fk = (uint32_t)(shfl64(tk, 0) >> 32);
                    adm2 |= 1u << j;
                  }
                }
                HSD_T(r1);
                HSD_ADD(8, r1 - r0);
                if (adm2) {
                  HSD_ADD(6, __popc(adm2)); HSD_ADD(11, 1);
                  // Every admission evicted the furthest entry of the moment, so what was evicted is: the newcomers above the
                  // final furthest entry, and as many of the LARGEST old entries as newcomers stay.  The final furthest entry is
                  // lane 0's, unless 32 admissions emptied the lanes of old entries and the 33rd largest old entry is above it.
                  const uint64_t xw = ((uint64_t)kd << 32) | e, xc = ((uint64_t)kd << 32) | (uint32_t)~e;
                  uint64_t top = shfl64(tk, 0);
                  if (nW > 32) { const uint64_t w33 = Wb[nW - 33]; top = w33 > top ? w33 : top; }
                  fk = (uint32_t)(top >> 32);
                  const unsigned stay = __ballot_sync(0xffffffffu, ((adm2 >> lane) & 1u) && xw <= top);
                  HSD_ADD(12, __popc(stay));
                  // the newcomers evicted again were pushed to C as well (and are cut off unless tied with the furthest key)
                  const int naW = __popc(stay), naC = __popc(adm2);
                  if (!candidate_room(naC)) overflow = true;
                  else if (nW - naW <= HS_MCAP && nC - head <= HS_MCAP) {
                    const unsigned below = (1u << lane) - 1;
                    if ((stay >> lane) & 1u) aw[__popc(stay & below)] = xw;
                    if ((adm2 >> lane) & 1u) ac[__popc(adm2 & below)] = xc;
                    cadd_le = __popc(__ballot_sync(0xffffffffu, ((adm2 >> lane) & 1u) && kd <= fk));
                    if (lane == 0) {
                      st[8] = naW; st[9] = nW - naW; st[10] = naC; st[11] = head; st[12] = nC - head; st[13] = (int)fk; st[14] = 0;
                    }
                    cmd = 1;
                    s++;
                    break;
                  } else {
                    hs_merge_inplace(Wb, nW - naW, xw, stay, 0u);
                    merge_candidates(xc, adm2, fk);
                  }
                }
              }
            }
            s++;
          }
          if (!cmd) {
            if ((!vis.bitmap && nvis > HS_HASH_CAP * 3 / 4) || overflow) { if (lane == 0) st[4] = 1; }
            HSD_T(tp);
            plan(stop);
            if (lane == 0) st[5] = s;
            HSD_T(d3);
            HSD_ADD(1, s); HSD_ADD(5, d3 - d2); HSD_ADD(10, d3 - tp);
          }
          if (lane == 0) st[7] = cmd;
        }
        __syncthreads();                                // (S1) next batch published, or a merge request
        if (!st[7]) break;
        HSD_T(tm0);
        {
          // threads [0, HALF): W <- Wb[0, st[9]) + aw[0, st[8]);  threads [HALF, 2 HALF): C <- Cb[head, head + st[12]) + ac[0, st[10])
          constexpr int HALF = HS_THREADS / 2;
          const bool isC = tid >= HALF;
          const int i = isC ? tid - HALF : tid;
          uint64_t *A = isC ? Cb + st[11] : Wb;
          const uint64_t *add = isC ? ac : aw;
          const int n = isC ? st[12] : st[9], na = isC ? st[10] : st[8];
          const uint32_t fkk = (uint32_t)st[13];
          uint64_t x[HS_KPT];
          int r[HS_KPT];
#pragma unroll
          for (int k = 0; k < HS_KPT; k++) { const int idx = i + k * HALF; x[k] = idx < n ? A[idx] : ~0ull; r[k] = 0; }
#pragma unroll 4
          for (int j = 0; j < na; j++) {
            const uint64_t v = add[j];
#pragma unroll
            for (int k = 0; k < HS_KPT; k++) r[k] += v < x[k] ? 1 : 0;
          }
          int apos = -1;
          uint64_t amine = 0;
          if (i >= HALF - na) {                         // a newcomer: elements below it, old and new
            amine = add[HALF - 1 - i];
            apos = hs_lower_bound(A, n, amine);
#pragma unroll 4
            for (int j = 0; j < na; j++) apos += add[j] < amine ? 1 : 0;
          }
          if (isC) {                                    // candidates strictly farther than the furthest key are cut off
#pragma unroll
            for (int k = 0; k < HS_KPT; k++) {
              const int idx = i + k * HALF;
              if (idx < n && (uint32_t)(x[k] >> 32) <= fkk && (idx + 1 == n || (uint32_t)(A[idx + 1] >> 32) > fkk)) st[14] = idx + 1;
            }
          }
          __syncthreads();                              // (S2) every element is in a register
#pragma unroll
          for (int k = 0; k < HS_KPT; k++) { const int idx = i + k * HALF; if (idx < n && r[k]) A[idx + r[k]] = x[k]; }
          if (apos >= 0) A[apos] = amine;
          __syncthreads();                              // (S3)
          if (warp == 0) nC = head + st[14] + cadd_le;
          HSD_T(tm1);
          HSD_ADD(13, tm1 - tm0);
        }
      }
    }
    if (st[4]) break;
    // ---- next layer's entry: min_by distance over the sorted working list == W[0] (index.rs:176-181)
    if (layer > 0) {
      if (warp == 0 && lane == 0) st[3] = nW > 0 ? (int)(uint32_t)Wb[0] : (int)ep;
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
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint64_t w = Wb[i];
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
    if (lane == 0) {
      a.out_counts[q] = cnt;
      if (a.out_stats) { a.out_stats[2 * (size_t)q] = n_dist; a.out_stats[2 * (size_t)q + 1] = n_expand; }
    }
  }
}

// Applies when: flat rows, ef in 1..2048, every adjacency list holds <= 32 edges, and the upper layers have the
// dense position map.  *launched = false otherwise (the caller keeps its previous kernels).
int launch_hnsw_spec(mgpu_hnsw *h, const HnswDev &g, const HnswSearchArgs &a, uint32_t *err_flags, bool *launched) {
  mgpu_ctx *ctx = h->ctx;
  *launched = false;
  // MGPU_HNSW_SPEC=0: never (hnsw.cu's kernels); unset / 1: whenever applicable
  static const int mode = getenv("MGPU_HNSW_SPEC") ? (getenv("MGPU_HNSW_SPEC")[0] == '0' ? 0 : 1) : 1;
  if (mode == 0 || h->quant != MGPU_QUANT_NONE || a.ef == 0 || a.ef > 2048 || h->max_degree > 32 || h->max_degree == 0) return MGPU_OK;
  if (h->num_layers > 1 && !g.upper_dense) return MGPU_OK;
  if (g.edges0 && g.deg0 > 32) return MGPU_OK;
  // MGPU_HNSW_BITMAP_MAX=<n>: test knob -- graphs above n points use the hash table (default 262144)
  static const uint32_t bitmap_max = getenv("MGPU_HNSW_BITMAP_MAX") ? (uint32_t)atoll(getenv("MGPU_HNSW_BITMAP_MAX")) : HS_BITMAP_MAX_N;
  const uint32_t vis_bitmap = h->n <= std::min(bitmap_max, HS_BITMAP_MAX_N) ? 1u : 0u;
  const uint32_t vis_words = vis_bitmap ? (uint32_t)((h->n + 31) / 32) + 1 : HS_HASH_CAP;
  const bool wide = a.B <= (uint32_t)ctx->sm_count;   // one 16-warp CTA per SM (its registers allow no second one)
  // Three shapes (measured): few queries (<= one per SM): 16 warps, 8 candidates per batch; a graph whose
  // rows stay in L2 searched by many queries (the Spann centroid graph: 4096 x 768, 1024 queries): 2 warps, 2 candidates --
  // seven CTAs per SM keep the whole batch resident (the replay is one warp's latency chain: what counts is how many of
  // them an SM interleaves) and the L2-resident rows make a batch's scoring short, so speculating further only wastes it.
  // Kernel time per 1024 queries: 1.15 ms, vs 1.49 with 4 warps (MGPU_HNSW_SMALL=1), 2.1 with 8 warps / 4 candidates and
  // 1.54 for hnsw.cu's register-list kernel.  Otherwise 8 warps, 4 candidates.
  static const int small_env = getenv("MGPU_HNSW_SMALL") ? atoi(getenv("MGPU_HNSW_SMALL")) : -1;
  const bool l2_resident = (size_t)h->n * h->dim * 4 <= (size_t)ctx->l2_bytes / 2;
  const bool small = !wide && (small_env >= 0 ? small_env != 0 : (l2_resident && a.B >= 2u * (uint32_t)ctx->sm_count));
  const int T = wide ? 8 : (small ? 2 : 4);
  const size_t smem = (size_t)((h->dim + 3) & ~3u) * 4 + (size_t)vis_words * 4 + (size_t)T * 32 * 4 * 3 + T * 12 + 64 + 16 +
                      (size_t)((a.ef + 32) + (2 * a.ef + 96) + 64) * 8;
  if (smem > ctx->smem_optin) return MGPU_OK;
#ifdef MGPU_SCAN_DBG
  {
    static int nl = 0;
    if (++nl == 8) {
      unsigned long long hd[16];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpyFromSymbol(hd, g_hs_dbg, sizeof(hd));
      fprintf(stderr, "[hnsw spec dbg] 7 launches: batches %llu replayed expansions %llu (%.2f per batch) rows scored %llu (%.1f per batch) | cycles "
                      "per batch: speculate-1 %.0f speculate-2 %.0f replay %.0f\n", hd[0], hd[1], (double)hd[1] / hd[0], hd[2], (double)hd[2] / hd[0],
              (double)hd[3] / hd[0], (double)hd[4] / hd[0], (double)hd[5] / hd[0]);
      fprintf(stderr, "[hnsw spec dbg] per replayed expansion: walked %.2f admitted %.2f stay %.2f; expansions with admissions %.2f; cycles per such expansion: "
                      "walk %.0f | fixed part per expansion %.0f, plan per batch %.0f, CTA merge %.0f\n", (double)hd[7] / hd[1], (double)hd[6] / hd[1], (double)hd[12] / hd[1],
              (double)hd[11] / hd[1], (double)hd[8] / (hd[11] ? hd[11] : 1), (double)hd[9] / hd[1], (double)hd[10] / hd[0],
              (double)hd[13] / (hd[11] ? hd[11] : 1));
    }
  }
#endif
  LaunchScope ls(ctx, MGPU_K_HNSW, nullptr, wide ? "k_hnsw_spec<16 warps,T=8> (hnsw_spec.cu)" : (small ? "k_hnsw_spec<2 warps,T=2> (hnsw_spec.cu)" : "k_hnsw_spec<8 warps,T=4> (hnsw_spec.cu)"));
#define HS_LAUNCH(MT, NWARP, TT)                                                                                   \
  do {                                                                                                             \
    cudaFuncSetAttribute(k_hnsw_spec<MT, NWARP, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    k_hnsw_spec<MT, NWARP, TT><<<a.B, NWARP * 32, smem, ctx->stream>>>(g, a, err_flags, vis_words, vis_bitmap);                \
  } while (0)
#define HS_LAUNCH_W(MT) do { if (wide) HS_LAUNCH(MT, 16, 8); else if (small && small_env == 1) HS_LAUNCH(MT, 4, 2); else if (small) HS_LAUNCH(MT, 2, 2); else HS_LAUNCH(MT, 8, 4); } while (0)
  if (h->metric == MGPU_L2) HS_LAUNCH_W(MGPU_L2); else HS_LAUNCH_W(MGPU_DOT);
#undef HS_LAUNCH_W
#undef HS_LAUNCH
  CUDA_TRY(ctx, cudaGetLastError());
  *launched = true;
  return MGPU_OK;
}

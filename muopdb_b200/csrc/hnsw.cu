// hnsw.cu -- batched HNSW beam search, strict reference semantics.
//   BlockBasedHnsw::ann_search   rs/index/src/hnsw/block_based/index.rs:159-210
//   BlockBasedHnsw::search_layer rs/index/src/hnsw/block_based/index.rs:212-287
//   get_edges_for_point / entry  rs/index/src/hnsw/block_based/graph_storage.rs:459-554
//
// One CTA (4 warps) per query.  The beam state lives in shared memory as two SORTED arrays of 64-bit composites:
//   W  working list, ascending (distance key, point id)       -- max-heap `working_list` of the reference
//   C  candidates,   ascending (distance key, ~point id)      -- min-heap `candidates` (pop = nearest, ties: larger id)
// Any priority structure that realises the same total order reproduces the reference's pops/evictions exactly.
// Candidates farther than the current furthest of a full W can never be expanded (they only trigger the `break`), so C is
// truncated to distance <= furthest -- it stays bounded by ef + ties.
// Per expansion: the neighbour list is filtered against the per-query visited bitmap (atomicOr, edge order preserved),
// the unvisited rows are gathered from HBM and scored bit-exactly in the reference's 16-lane order (two rows per warp:
// lane h of a half-warp owns lane-accumulator h, 64-byte coalesced row reads), then warp 0 applies the admissions
// sequentially in edge order, exactly like index.rs:255-281.
#include "internal.cuh"
#include "pq_device.cuh"
#include "hnsw_device.cuh"

#define HN_THREADS 128
#define HN_WARPS 4

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// warp-collective insert of `key` into the ascending array A[lo, n) (shared memory); returns the new n
__device__ __forceinline__ int warp_sorted_insert(uint64_t *A, int lo, int n, uint64_t key) {
  const int lane = lane_id();
  uint32_t cnt = 0;
  for (int i = lo + lane; i < n; i += 32) cnt += (A[i] <= key) ? 1u : 0u;
  const int pos = lo + (int)warp_sum_u32(cnt);
  for (int base = n - 1; base >= pos; base -= 32) {
    int i = base - lane;
    uint64_t v = 0;
    if (i >= pos) v = A[i];
    __syncwarp();
    if (i >= pos) A[i + 1] = v;
    __syncwarp();
  }
  if (lane == 0) A[pos] = key;
  __syncwarp();
  return n + 1;
}

template <int QUANT, int METRIC>
__global__ void __launch_bounds__(HN_THREADS) k_hnsw_search(HnswDev g, HnswSearchArgs a, uint32_t *__restrict__ visited_all,
                                                             uint32_t vis_words, const uint8_t *__restrict__ qcodes_all,
                                                             uint32_t capC, uint32_t *__restrict__ err_flags, int only_flagged) {
  extern __shared__ __align__(16) uint8_t smem[];
  if (only_flagged) {
    // second pass after k_hnsw_search_reg: redo the queries whose candidate list overflowed its registers
    if (!err_flags[blockIdx.x]) return;
    uint32_t *vis = visited_all + (size_t)blockIdx.x * vis_words;
    for (uint32_t i = threadIdx.x; i < vis_words; i += HN_THREADS) vis[i] = 0;
    __syncthreads();
  }
  const uint32_t ef = a.ef;
  uint64_t *W = (uint64_t *)smem;                 // ef + 1
  uint64_t *C = W + (ef + 2);                     // capC + 1
  float *sq = (float *)(C + capC + 2);            // dim floats (flat) / unused
  uint32_t *nb = (uint32_t *)(sq + ((g.dim + 3) & ~3u));  // 32 neighbour ids
  uint32_t *nbk = nb + 32;                        // 32 distance keys
  int *st = (int *)(nbk + 32);                    // state: [0]=nW [1]=headC [2]=nC [3]=cur point [4]=stop [5]=nu

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = blockIdx.x;
  uint32_t *visited = visited_all + (size_t)q * vis_words;
  const uint8_t *qc = QUANT == MGPU_QUANT_PQ ? qcodes_all + (size_t)q * g.m : nullptr;
  if (QUANT == MGPU_QUANT_NONE)
    for (uint32_t d = tid; d < g.dim; d += HN_THREADS) sq[d] = a.Q[(size_t)q * g.dim + d];
  __syncthreads();

  unsigned long long n_dist = 0, n_expand = 0;  // tracked by thread 0

  // distance of the query to point `pid`; called by a full half-warp for flat rows (returns on all 16 lanes) or by a
  // single thread for PQ rows
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
      // 48 independent 64-byte row segments in flight per half-warp (one round trip for a 768-d row), consumed in order
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
    if (p < n) ret = ref_tail<METRIC>(PtrAcc{sq}, PtrAcc{row}, p, n, ret);  // every lane redundantly; tiny
    if (METRIC == MGPU_L2) return sqrtf(ret);  // NoQuantizer::distance -> D::calculate (noq/mod.rs:44-51)
    return -ret;
  };
  auto pq_distance_thread = [&](uint32_t pid) -> float {
    return pq_distance_streaming<METRIC>(g.cb, g.m, g.K, g.dsub, RowMajorCode{qc},
                                         RowMajorCode{(const uint8_t *)g.rows + (size_t)pid * g.m});
  };

  // pop the nearest candidate (thread 0 only): `continue` while the working list is empty, `break` when the candidate is
  // strictly farther than the furthest of the working list (index.rs:235-247)
  auto pop_next = [&]() {
    int head = st[1], nC = st[2], nW = st[0];
    int stop = 0;
    uint32_t cur = 0;
    for (;;) {
      if (head >= nC) { stop = 1; break; }
      uint64_t c = C[head++];
      if (nW == 0) continue;
      uint32_t ck = (uint32_t)(c >> 32), fk = (uint32_t)(W[nW - 1] >> 32);
      if (ck > fk) { stop = 1; break; }
      cur = ~(uint32_t)c;
      break;
    }
    st[1] = head; st[3] = (int)cur; st[4] = stop;
  };

  uint32_t ep = g.entry_point;
  for (int layer = (int)g.num_layers - 1; layer >= 0; layer--) {
    // ---- search_layer(ep, ef, layer) -------------------------------------------------------------------------------
    const uint64_t lvl_s = g.level_offsets[g.num_layers - 1 - layer];
    const uint64_t lvl_e = g.level_offsets[g.num_layers - layer];
    // entry: set_visited(ep); distance; push to both heaps (index.rs:219-233)
    float ed = 0.0f;
    if (QUANT == MGPU_QUANT_NONE) { if (warp == 0) ed = flat_distance_halfwarp(ep); }
    else if (tid == 0) ed = pq_distance_thread(ep);
    if (tid == 0) {
      if (ep < g.n) atomicOr(&visited[ep >> 5], 1u << (ep & 31));
      uint32_t kd = f2key(ed);
      W[0] = ((uint64_t)kd << 32) | ep;
      C[0] = ((uint64_t)kd << 32) | (uint32_t)~ep;
      st[0] = 1; st[1] = 0; st[2] = 1; st[4] = 0;
      n_dist++;
      pop_next();
    }
    for (;;) {
      // the nearest candidate was popped by thread 0 at the end of the previous admission phase (index.rs:235-247)
      __syncthreads();
      if (st[4]) break;
      const uint32_t cur = (uint32_t)st[3];
      // ---- edges of cur at this layer (graph_storage.rs:459-521)
      long long idx = -1;
      if (layer == 0) idx = cur;
      else {
        // first position of `cur` inside points[lvl_s, lvl_e): one load from the dense per-layer map when the index has one
        // (create builds it while (num_layers-1) * n * 4 bytes stays small), else a binary search in the sorted copy
        if (g.upper_dense) {
          int32_t pos = cur < g.n ? g.upper_dense[(size_t)(layer - 1) * g.n + cur] : -1;
          if (pos >= 0) idx = (long long)pos - (long long)lvl_s;
        } else {
          long long lo = (long long)lvl_s, hi = (long long)lvl_e - 1;
          while (lo <= hi) {
            long long mid = (lo + hi) >> 1;
            uint32_t v = g.upper_pid[mid];
            if (v < cur) lo = mid + 1; else hi = mid - 1;
          }
          if (lo < (long long)lvl_e && g.upper_pid[lo] == cur) idx = (long long)g.upper_pos[lo] - (long long)lvl_s;
        }
      }
      uint64_t e_begin = 0, e_end = 0;
      if (idx >= 0 && lvl_s + (uint64_t)idx + 1 < g.n_edge_offsets) {
        e_begin = g.edge_offsets[lvl_s + idx];
        e_end = g.edge_offsets[lvl_s + idx + 1];
      }
      if (e_begin == e_end) {  // None => continue (uniform across the CTA)
        __syncthreads();
        if (tid == 0) pop_next();
        continue;
      }
      if (tid == 0) n_expand++;
      for (uint64_t eb = e_begin; eb < e_end; eb += 32) {
        // ---- visited filter, edge order preserved (index.rs:255-259)
        if (warp == 0) {
          uint64_t ei = eb + lane;
          uint32_t e = 0;
          bool fresh = false;
          if (ei < e_end) {
            e = g.edges[ei];
            if (e < g.n) {
              uint32_t bit = 1u << (e & 31);
              uint32_t old = atomicOr(&visited[e >> 5], bit);
              fresh = !(old & bit);
            }
          }
          unsigned mk = __ballot_sync(0xffffffffu, fresh);
          if (fresh) nb[__popc(mk & ((1u << lane) - 1))] = e;
          if (lane == 0) st[5] = __popc(mk);
        }
        __syncthreads();
        const int nu = st[5];
        // ---- distances (index.rs:264 -> :289-298)
        if (QUANT == MGPU_QUANT_NONE) {
          const int rounds = (nu + HN_WARPS * 2 - 1) / (HN_WARPS * 2);
          for (int r = 0; r < rounds; r++) {
            // both half-warps stay in the loop together (the lane reduction shuffles are full-warp);
            // an out-of-range half re-scores the last neighbour and discards the result
            int j = r * HN_WARPS * 2 + warp * 2 + (lane >> 4);
            int jj = j < nu ? j : (nu - 1);
            float d = flat_distance_halfwarp(nb[jj]);
            if (j < nu && (lane & 15) == 0) nbk[j] = f2key(d);
          }
        } else {
          if (tid < nu) nbk[tid] = f2key(pq_distance_thread(nb[tid]));
        }
        __syncthreads();
        // ---- admissions, sequential in edge order (index.rs:260-281)
        if (warp == 0) {
          int nW = st[0], head = st[1], nC = st[2];
          for (int j = 0; j < nu; j++) {
            if (nW == 0) continue;  // peek() == None => continue (only when ef == 0)
            uint32_t kd = nbk[j], e = nb[j];
            uint32_t fk = (uint32_t)(W[nW - 1] >> 32);
            if (kd < fk || nW < (int)ef) {
              if (nC + 1 > (int)capC) {
                // compact C (drop the consumed prefix)
                int live = nC - head;
                for (int base = 0; base < live; base += 32) {
                  int i = base + lane;
                  uint64_t v = 0;
                  if (i < live) v = C[head + i];
                  __syncwarp();
                  if (i < live) C[i] = v;
                  __syncwarp();
                }
                head = 0; nC = live;
                if (nC + 1 > (int)capC) { if (lane == 0) err_flags[q] = 1; nC = capC - 1; }
              }
              nC = warp_sorted_insert(C, head, nC, ((uint64_t)kd << 32) | (uint32_t)~e);
              nW = warp_sorted_insert(W, 0, nW, ((uint64_t)kd << 32) | e);
              if (nW > (int)ef) nW--;  // pop the furthest (index.rs:277-279)
              if (nW == (int)ef && nW > 0) {
                // candidates strictly farther than the furthest can never be expanded: truncate the sorted tail
                uint32_t fk2 = (uint32_t)(W[nW - 1] >> 32);
                while (nC > head && (uint32_t)(C[nC - 1] >> 32) > fk2) nC--;
              }
            }
          }
          if (lane == 0) {
            st[0] = nW; st[1] = head; st[2] = nC;
            n_dist += nu;
            if (eb + 32 >= e_end) pop_next();  // last edge batch: pop the next candidate right away (saves a barrier)
          }
        }
        // no barrier here: warps 1..3 only touch nb/nbk again after the next filter barrier, which warp 0 reaches last
      }
    }
    // ---- next layer's entry: min_by distance over the sorted working list == W[0] (index.rs:176-181)
    if (layer > 0) {
      if (st[0] > 0) ep = (uint32_t)W[0];
      __syncthreads();
    }
  }
  // ---- results: working list is already sorted by (distance, point_id); truncate to k, map to doc ids (index.rs:185-204)
  const int nW = st[0];
  const uint32_t cnt = min((uint32_t)nW, a.k);
  for (uint32_t i = tid; i < cnt; i += HN_THREADS) {
    uint64_t w = W[i];
    uint32_t pid = (uint32_t)w, kd = (uint32_t)(w >> 32);
    uint32_t u = (kd & 0x80000000u) ? (kd ^ 0x80000000u) : ~kd;
    a.out_scores[(size_t)q * a.k + i] = __uint_as_float(u);
    if (a.out_pids) a.out_pids[(size_t)q * a.k + i] = pid;
    if (a.out_docs) {
      mgpu_u128 d;
      if (g.doc_ids) d = g.doc_ids[pid]; else { d.lo = pid; d.hi = 0; }
      a.out_docs[(size_t)q * a.k + i] = d;
    }
  }
  if (tid == 0) {
    a.out_counts[q] = cnt;
    if (a.out_stats) { a.out_stats[2 * (size_t)q] = n_dist; a.out_stats[2 * (size_t)q + 1] = n_expand; }
  }
}

// ---- register-resident beam (the fast path) ---------------------------------------------------------------------------------
// Same algorithm and the same total orders as k_hnsw_search, but the two priority structures live in the REGISTERS of
// warp 0 as sorted, lane-blocked arrays (lane l holds entries EPL*l .. EPL*l + EPL-1):
//   W  ascending (distance key, point id);  a full W drops its furthest by decrementing nW (stale tail entries are larger
//      than anything admitted later, so they never move in front of a live entry and fall off the end);
//   C  ascending (distance key, ~point id); pop = shift the whole array left by one entry.
// A sorted insert is ~EPL compares + one ballot + three shuffles instead of a shared-memory scan-and-shift loop, which was
// the serial bottleneck of every expansion (ncu: warp 0 executed ~2.5 k instructions per expansion, the other warps sat at
// the barrier).  Capacity: 32*EPL entries each; a candidate that would fall off the end of C while still expandable
// raises err_flags[q] and the launcher re-runs that query with k_hnsw_search (never observed short of mass exact ties).
template <int QUANT, int METRIC, int EPL>
__global__ void __launch_bounds__(HN_THREADS) k_hnsw_search_reg(HnswDev g, HnswSearchArgs a, uint32_t *__restrict__ visited_all,
                                                                 uint32_t vis_words, const uint8_t *__restrict__ qcodes_all,
                                                                 uint32_t *__restrict__ err_flags) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t ef = a.ef;
  constexpr int CAP = EPL * 32;
  float *sq = (float *)smem;                                  // dim floats (flat) / unused
  uint32_t *nb = (uint32_t *)(sq + ((g.dim + 3) & ~3u));      // 32 neighbour ids
  uint32_t *nbk = nb + 32;                                    // 32 distance keys
  int *st = (int *)(nbk + 32);                                // [3]=cur point [4]=stop [5]=nu [6]=next entry point

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = blockIdx.x;
  uint32_t *visited = visited_all + (size_t)q * vis_words;
  const uint8_t *qc = QUANT == MGPU_QUANT_PQ ? qcodes_all + (size_t)q * g.m : nullptr;
  if (QUANT == MGPU_QUANT_NONE)
    for (uint32_t d = tid; d < g.dim; d += HN_THREADS) sq[d] = a.Q[(size_t)q * g.dim + d];
  __syncthreads();

  unsigned long long n_dist = 0, n_expand = 0;  // tracked by warp 0 (uniform)

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
  auto pq_distance_thread = [&](uint32_t pid) -> float {
    return pq_distance_streaming<METRIC>(g.cb, g.m, g.K, g.dsub, RowMajorCode{qc},
                                         RowMajorCode{(const uint8_t *)g.rows + (size_t)pid * g.m});
  };

  RegList<EPL> W, Cd;  // only warp 0's copies are meaningful
  int nW = 0, nC = 0;
  bool overflow = false;

  // warp 0: pop the nearest candidate; `continue` while the working list is empty, `break` when the candidate is strictly
  // farther than the furthest of the working list (index.rs:235-247).  Publishes cur / stop through shared memory.
  auto pop_next = [&]() {
    int stop = 0;
    uint32_t cur = 0;
    for (;;) {
      if (nC == 0) { stop = 1; break; }
      const uint64_t c = Cd.get(0);
      Cd.pop_front();
      nC--;
      if (nW == 0) continue;
      const uint32_t ck = (uint32_t)(c >> 32), fk = (uint32_t)(W.get(nW - 1) >> 32);
      if (ck > fk) { stop = 1; break; }
      cur = ~(uint32_t)c;
      break;
    }
    if (lane == 0) { st[3] = (int)cur; st[4] = stop; }
  };

  uint32_t ep = g.entry_point;
  for (int layer = (int)g.num_layers - 1; layer >= 0; layer--) {
    const uint64_t lvl_s = g.level_offsets[g.num_layers - 1 - layer];
    const uint64_t lvl_e = g.level_offsets[g.num_layers - layer];
    // entry: set_visited(ep); distance; push to both heaps (index.rs:219-233)
    if (warp == 0) {
      float ed = 0.0f;
      if (QUANT == MGPU_QUANT_NONE) ed = flat_distance_halfwarp(ep);
      else { if (lane == 0) ed = pq_distance_thread(ep); }
      ed = __shfl_sync(0xffffffffu, ed, 0);
      if (lane == 0 && ep < g.n) atomicOr(&visited[ep >> 5], 1u << (ep & 31));
      const uint32_t kd = f2key(ed);
      W.init(); Cd.init();
      W.insert(((uint64_t)kd << 32) | ep);
      Cd.insert(((uint64_t)kd << 32) | (uint32_t)~ep);
      nW = 1; nC = 1;
      n_dist++;
      pop_next();
    }
    for (;;) {
      __syncthreads();
      if (st[4]) break;
      const uint32_t cur = (uint32_t)st[3];
      // ---- edges of cur at this layer (graph_storage.rs:459-521)
      const bool inline0 = layer == 0 && g.edges0 != nullptr;
      uint64_t e_begin = 0, e_end = 0;
      uint32_t my_edge = 0xFFFFFFFFu;   // inline0: lane l of every warp holds edge l (the rows are prefetched by all warps)
      if (inline0) {
        if (cur < g.n && (uint32_t)lane < g.deg0) my_edge = __ldg(g.edges0 + (size_t)cur * g.deg0 + lane);
        uint32_t second = 0xFFFFFFFFu;
        if (g.deg0 > 32 && cur < g.n && (uint32_t)lane + 32 < g.deg0) second = __ldg(g.edges0 + (size_t)cur * g.deg0 + 32 + lane);
        const unsigned m0 = __ballot_sync(0xffffffffu, my_edge != 0xFFFFFFFFu), m1 = __ballot_sync(0xffffffffu, second != 0xFFFFFFFFu);
        e_end = (uint64_t)(__popc(m0) + __popc(m1));   // stored order is dense from slot 0
        if (g.prefetch_rows && my_edge < g.n) {
          // pull the neighbour rows towards L2 while the visited filter's atomics are in flight: the rows of the fresh
          // neighbours are then L2 hits.  Warp w prefetches its quarter of every row's 128-byte lines.
          const char *rowp = (const char *)g.rows + (size_t)my_edge * g.dim * 4;
          const uint32_t lines = (g.dim * 4 + 127) / 128;
          for (uint32_t ln = warp; ln < lines; ln += HN_WARPS) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + (size_t)ln * 128));
        }
        if (e_end > 32) my_edge = 0xFFFFFFFFu;  // rare wide rows: the generic batch loop below re-reads them
      } else {
        long long idx = -1;
        if (layer == 0) idx = cur;
        else if (g.upper_dense) {
          int32_t pos = cur < g.n ? g.upper_dense[(size_t)(layer - 1) * g.n + cur] : -1;
          if (pos >= 0) idx = (long long)pos - (long long)lvl_s;
        } else {
          long long lo = (long long)lvl_s, hi = (long long)lvl_e - 1;
          while (lo <= hi) {
            long long mid = (lo + hi) >> 1;
            uint32_t v = g.upper_pid[mid];
            if (v < cur) lo = mid + 1; else hi = mid - 1;
          }
          if (lo < (long long)lvl_e && g.upper_pid[lo] == cur) idx = (long long)g.upper_pos[lo] - (long long)lvl_s;
        }
        if (idx >= 0 && lvl_s + (uint64_t)idx + 1 < g.n_edge_offsets) {
          e_begin = g.edge_offsets[lvl_s + idx];
          e_end = g.edge_offsets[lvl_s + idx + 1];
        }
      }
      if (e_begin == e_end) {  // None => continue (uniform across the CTA)
        __syncthreads();
        if (warp == 0) pop_next();
        continue;
      }
      if (warp == 0) n_expand++;
      for (uint64_t eb = e_begin; eb < e_end; eb += 32) {
        // ---- visited filter, edge order preserved (index.rs:255-259)
        if (warp == 0) {
          uint64_t ei = eb + lane;
          uint32_t e = 0;
          bool fresh = false;
          if (ei < e_end) {
            if (inline0) e = (e_end <= 32) ? my_edge : __ldg(g.edges0 + (size_t)cur * g.deg0 + ei);
            else e = g.edges[ei];
            if (e < g.n) {
              uint32_t bit = 1u << (e & 31);
              uint32_t old = atomicOr(&visited[e >> 5], bit);
              fresh = !(old & bit);
            }
          }
          unsigned mk = __ballot_sync(0xffffffffu, fresh);
          if (fresh) nb[__popc(mk & ((1u << lane) - 1))] = e;
          if (lane == 0) st[5] = __popc(mk);
        }
        __syncthreads();
        const int nu = st[5];
        // ---- distances (index.rs:264 -> :289-298)
        if (QUANT == MGPU_QUANT_NONE) {
          const int rounds = (nu + HN_WARPS * 2 - 1) / (HN_WARPS * 2);
          for (int r = 0; r < rounds; r++) {
            if (r * HN_WARPS * 2 + warp * 2 >= nu) continue;  // warp-uniform: no row for this warp in this round
            int j = r * HN_WARPS * 2 + warp * 2 + (lane >> 4);
            int jj = j < nu ? j : (nu - 1);
            float d = flat_distance_halfwarp(nb[jj]);
            if (j < nu && (lane & 15) == 0) nbk[j] = f2key(d);
          }
        } else {
          if (tid < nu) nbk[tid] = f2key(pq_distance_thread(nb[tid]));
        }
        __syncthreads();
        // ---- admissions, sequential in edge order (index.rs:260-281)
        if (warp == 0) {
          for (int j = 0; j < nu; j++) {
            if (nW == 0) continue;  // peek() == None => continue (only when ef == 0)
            const uint32_t kd = nbk[j], e = nb[j];
            const uint32_t fk = (uint32_t)(W.get(nW - 1) >> 32);
            if (kd < fk || nW < (int)ef) {
              if (nC == CAP) {
                // the entry that falls off the end may only be lost if it can never be expanded
                const uint32_t lastk = (uint32_t)(Cd.get(CAP - 1) >> 32);
                if (!(nW == (int)ef && lastk > fk)) overflow = true;
                nC--;
              }
              Cd.insert(((uint64_t)kd << 32) | (uint32_t)~e);
              nC++;
              W.insert(((uint64_t)kd << 32) | e);
              nW++;
              if (nW > (int)ef) nW--;  // pop the furthest (index.rs:277-279)
              if (nW == (int)ef && nW > 0) {
                // candidates strictly farther than the furthest can never be expanded: truncate the sorted tail
                const uint32_t fk2 = (uint32_t)(W.get(nW - 1) >> 32);
                nC = Cd.count_le(nC, fk2);
              }
            }
          }
          n_dist += nu;
          if (eb + 32 >= e_end) pop_next();  // last edge batch: pop the next candidate right away (saves a barrier)
        }
        // no barrier here: warps 1..3 only touch nb/nbk again after the next filter barrier, which warp 0 reaches last
      }
    }
    // ---- next layer's entry: min_by distance over the sorted working list == W[0] (index.rs:176-181)
    if (layer > 0) {
      if (warp == 0 && lane == 0) st[6] = nW > 0 ? (int)(uint32_t)W.v[0] : (int)ep;
      __syncthreads();
      ep = (uint32_t)st[6];
      __syncthreads();
    }
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
      if (overflow) err_flags[q] = 1;
    }
  }
}

int launch_hnsw_search(mgpu_hnsw *h, const HnswSearchArgs &a) {
  mgpu_ctx *ctx = h->ctx;
  if (a.B == 0) return MGPU_OK;
  const uint32_t ef = a.ef;
  const uint32_t capC = 2 * ef + 256;
  const uint32_t vis_words = (uint32_t)((h->n + 31) / 32) + 1;
  size_t smem = (size_t)(ef + 2) * 8 + (size_t)(capC + 2) * 8 + (size_t)((h->dim + 3) & ~3u) * 4 + 64 * 4 + 16 * 4;
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "hnsw search: ef=%u needs %zu bytes of shared memory", ef, smem);
  // workspace: visited bitmaps, query codes, error flags
  size_t need = 0;
  need = ws_need(need, (size_t)a.B * vis_words * 4);
  need = ws_need(need, (size_t)a.B * (h->pq ? h->pq->m : 0));
  need = ws_need(need, (size_t)a.B * 4);
  // NOTE: the caller's staged buffers live at the front of the workspace; this launcher uses a private allocation
  uint8_t *priv = nullptr;
  CUDA_TRY(ctx, cudaMallocAsync((void **)&priv, need + 256, ctx->stream));
  WsAlloc w(priv, need + 256);
  uint32_t *visited = w.get<uint32_t>((size_t)a.B * vis_words);
  uint8_t *qcodes = w.get<uint8_t>((size_t)a.B * (h->pq ? h->pq->m : 0));
  uint32_t *err = w.get<uint32_t>(a.B);
  CUDA_TRY(ctx, cudaMemsetAsync(err, 0, (size_t)a.B * 4, ctx->stream));
  HnswDev g;
  g.edges = h->d_edges; g.points = h->d_points; g.upper_pid = h->d_upper_sorted_pid; g.upper_pos = h->d_upper_sorted_pos;
  static const bool use_dense = !(getenv("MGPU_HNSW_DENSE") && getenv("MGPU_HNSW_DENSE")[0] == '0');
  g.upper_dense = use_dense ? h->d_upper_dense : nullptr;
  static const bool use_edges0 = !(getenv("MGPU_HNSW_EDGES0") && getenv("MGPU_HNSW_EDGES0")[0] == '0');
  static const bool use_prefetch = !(getenv("MGPU_HNSW_PREFETCH") && getenv("MGPU_HNSW_PREFETCH")[0] == '0');
  g.edges0 = use_edges0 ? h->d_edges0 : nullptr; g.deg0 = h->deg0;
  g.prefetch_rows = use_prefetch && h->quant == MGPU_QUANT_NONE ? 1u : 0u;
  g.edge_offsets = h->d_edge_offsets; g.level_offsets = h->d_level_offsets; g.rows = h->d_rows; g.doc_ids = h->d_doc_ids;
  g.cb = h->pq ? h->pq->d_cb : nullptr; g.dim = h->dim; g.qdim = h->qdim; g.num_layers = h->num_layers;
  g.entry_point = h->entry_point; g.m = h->pq ? h->pq->m : 0; g.K = h->pq ? h->pq->K : 0; g.dsub = h->pq ? h->pq->dsub : 0;
  g.n = h->n; g.n_edge_offsets = h->n_edge_offsets;
  int s = MGPU_OK;
  if (h->quant == MGPU_QUANT_PQ) s = launch_pq_quantize(h->pq, a.Q, a.B, qcodes);  // index.rs:168
  static const bool no_reg = getenv("MGPU_HNSW_REG") && getenv("MGPU_HNSW_REG")[0] == '0';
  int epl = no_reg ? 0 : (ef <= 48 ? 2 : (ef <= 128 ? 5 : (ef <= 224 ? 8 : 0)));
  const size_t smem_reg = (size_t)((h->dim + 3) & ~3u) * 4 + 64 * 4 + 16 * 4;
  // fast path: speculate-then-replay kernel (hnsw_spec.cu); the queries it flags are redone by the generic kernel below
  bool spec = false;
  if (s == MGPU_OK) s = launch_hnsw_spec(h, g, a, err, &spec);
  if (spec) epl = -1;
  // the global visited bitmaps (B x n bits: 32 MB for 256 queries on 1M points) serve the kernels below; after the batch kernel
  // only flagged queries are redone, and those clear their own bitmap
  if (s == MGPU_OK && !spec && cudaMemsetAsync(visited, 0, (size_t)a.B * vis_words * 4, ctx->stream) != cudaSuccess)
    s = mgpu_fail(ctx, MGPU_ERR_CUDA, "hnsw search: memset failed");
  if (s == MGPU_OK && epl > 0) {
    LaunchScope ls(ctx, MGPU_K_HNSW, nullptr, "k_hnsw_search_reg (hnsw.cu)");
#define HN_LAUNCH_R(QT, MT, E)                                                                                          \
  do {                                                                                                                  \
    cudaFuncSetAttribute(k_hnsw_search_reg<QT, MT, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_reg);     \
    k_hnsw_search_reg<QT, MT, E><<<a.B, HN_THREADS, smem_reg, ctx->stream>>>(g, a, visited, vis_words, qcodes, err);    \
  } while (0)
#define HN_LAUNCH_RE(QT, MT)                                                                  \
  do {                                                                                        \
    if (epl == 2) HN_LAUNCH_R(QT, MT, 2); else if (epl == 5) HN_LAUNCH_R(QT, MT, 5); else HN_LAUNCH_R(QT, MT, 8); \
  } while (0)
    if (h->quant == MGPU_QUANT_NONE) { if (h->metric == MGPU_L2) HN_LAUNCH_RE(MGPU_QUANT_NONE, MGPU_L2); else HN_LAUNCH_RE(MGPU_QUANT_NONE, MGPU_DOT); }
    else { if (h->metric == MGPU_L2) HN_LAUNCH_RE(MGPU_QUANT_PQ, MGPU_L2); else HN_LAUNCH_RE(MGPU_QUANT_PQ, MGPU_DOT); }
#undef HN_LAUNCH_RE
#undef HN_LAUNCH_R
    if (cudaGetLastError() != cudaSuccess) s = mgpu_fail(ctx, MGPU_ERR_CUDA, "hnsw search launch failed");
  }
  if (s == MGPU_OK) {
    // generic kernel: the whole batch when the register kernel does not apply, otherwise only the flagged queries
    const int only_flagged = epl ? 1 : 0;
    LaunchScope ls(ctx, MGPU_K_HNSW, nullptr, only_flagged ? nullptr : "k_hnsw_search (hnsw.cu)");
#define HN_LAUNCH(QT, MT)                                                                                       \
  do {                                                                                                          \
    cudaFuncSetAttribute(k_hnsw_search<QT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    k_hnsw_search<QT, MT><<<a.B, HN_THREADS, smem, ctx->stream>>>(g, a, visited, vis_words, qcodes, capC, err, only_flagged); \
  } while (0)
    if (h->quant == MGPU_QUANT_NONE) { if (h->metric == MGPU_L2) HN_LAUNCH(MGPU_QUANT_NONE, MGPU_L2); else HN_LAUNCH(MGPU_QUANT_NONE, MGPU_DOT); }
    else { if (h->metric == MGPU_L2) HN_LAUNCH(MGPU_QUANT_PQ, MGPU_L2); else HN_LAUNCH(MGPU_QUANT_PQ, MGPU_DOT); }
#undef HN_LAUNCH
    if (cudaGetLastError() != cudaSuccess) s = mgpu_fail(ctx, MGPU_ERR_CUDA, "hnsw search launch failed");
  }
  cudaFreeAsync(priv, ctx->stream);
  return s;
}

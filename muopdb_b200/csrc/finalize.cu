// finalize.cu -- per-query epilogue of a scan: exact re-score of the PQ survivors, the reference's result ordering,
// point-id -> doc-id remap, and the cross-shard merge.
//   search_with_centroids           rs/index/src/ivf/block_based/index.rs:250-285  (order by (distance, point_id))
//   search_with_centroids_and_remap rs/index/src/ivf/block_based/index.rs:298-332  (order by (score, doc_id))
//   Snapshot merge                  rs/index/src/collection/snapshot.rs:49-63,79-108
#include "internal.cuh"
#include "pq_device.cuh"
#include "finalize_device.cuh"

// generic path: one warp per query, one candidate per lane (any dsub; flat rows arrive with exact keys already)
template <int METRIC>
__global__ void __launch_bounds__(128) k_finalize(FinalizeArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.B) return;
  if (a.qstate && a.qstate[q]) return;   // deferred by the scan: the exact fallback answers this query
  uint64_t ckey = a.cand_key[(size_t)q * MGPU_NCAND + lane];
  uint32_t slot = a.cand_slot[(size_t)q * MGPU_NCAND + lane];
  bool valid = slot != MGPU_EMPTY_SLOT;
  const uint32_t nscan = __popc(__ballot_sync(0xffffffffu, valid));
  const uint32_t pid = (uint32_t)ckey;
  uint32_t skey = (uint32_t)(ckey >> 32);
  const uint32_t fkey = skey;
  if (a.cb != nullptr && a.prune) valid = prune_keep(a.k, valid, skey, lane, a.key16 != 0, a.m, a.cert_slack);
  if (a.cb != nullptr && valid) {
    // exact Quantizer::distance(quantized_query, row, StreamingSIMD) (index.rs:203-207, pq/mod.rs:231-266)
    float d;
    RowMajorCode qc{a.qcodes + (size_t)q * a.m};
    if (a.pq_fast) d = pq_distance_streaming<METRIC>(a.cb, a.m, a.K, a.dsub, qc, FastLayoutCode{a.codes, slot, a.ng});
    else d = pq_distance_streaming<METRIC>(a.cb, a.m, a.K, a.dsub, qc, RowMajorCode{a.codes + (size_t)slot * a.m});
    skey = f2key(d);
  }
  finalize_tail(a, q, lane, valid, skey, pid, slot, fkey, nscan);
}

// dsub == 8 (the reference default): one CTA of 128 threads per query.  Staging: the query's own centroids (m x 8 floats) and
// the 32 candidates' code words go to shared memory with 16-byte loads.  Scoring: 2 threads per candidate, thread (c, h)
// owns lane accumulators 4h..4h+3 of ProductQuantizer::distance's shared 8-lane sum (pq/mod.rs:231-266) and walks the m
// subspaces in order -- the additions happen in exactly the reference's order while the m codebook gathers of a candidate
// are independent 16-byte loads (no dependent code -> centroid -> accumulate latency chain per subspace).
#define FIN8_THREADS 128
#define FIN8_MLP 24
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
template <int METRIC>
__global__ void __launch_bounds__(FIN8_THREADS) k_finalize_pq8(FinalizeArgs a) {
  extern __shared__ __align__(16) uint8_t fsm[];
  const uint32_t m = a.m, mp = ((m + 15) & ~15u) + 4;  // +4: odd word stride between candidates (bank spread)
  float4 *qv4 = (float4 *)fsm;                     // m x 2: centroid of the query code per subspace
  uint8_t *cc = (uint8_t *)(qv4 + (size_t)m * 2);  // 32 x mp: candidate codes
  uint32_t *skeys = (uint32_t *)(cc + (size_t)MGPU_NCAND * mp);  // 32 exact score keys
  uint32_t *sslot = skeys + MGPU_NCAND;            // 32 slots
  float4 *ring = (float4 *)(sslot + MGPU_NCAND);   // FIN8_MLP x 64 gathered centroid halves (16-byte aligned: all sizes above are)
  const uint32_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  if (a.qstate && a.qstate[q]) return;   // deferred by the scan: the exact fallback answers this query
  uint32_t nscan = 0, fkey = 0;
  if (tid < MGPU_NCAND) {
    uint32_t slot = a.cand_slot[(size_t)q * MGPU_NCAND + tid];
    fkey = (uint32_t)(a.cand_key[(size_t)q * MGPU_NCAND + tid] >> 32);
    nscan = __popc(__ballot_sync(0xffffffffu, slot != MGPU_EMPTY_SLOT));
    if (a.prune && !prune_keep(a.k, slot != MGPU_EMPTY_SLOT, fkey, lane, a.key16 != 0, a.m, a.cert_slack))
      slot = MGPU_EMPTY_SLOT;  // cannot reach the exact top k: treated like an empty candidate from here on
    sslot[tid] = slot;
  }
  const uint8_t *qc = a.qcodes + (size_t)q * m;
  for (uint32_t i = tid; i < m * 2; i += FIN8_THREADS) {
    const uint32_t s = i >> 1, h = i & 1;
    qv4[i] = __ldg((const float4 *)(a.cb + ((size_t)s * a.K + qc[s]) * 8) + h);
  }
  __syncthreads();
  if (a.pq_fast) {
    // fast layout: a row's codes are 2*ng segments of 16 bytes; byte i of segment (g, u) is subspace 32g + (l ^ (16u + i))
    for (uint32_t i = tid; i < MGPU_NCAND * 2 * a.ng; i += FIN8_THREADS) {
      const uint32_t c = i & 31, seg = i >> 5, g = seg >> 1, u = seg & 1;
      const uint32_t slot = sslot[c];
      const uint32_t l = slot & 31;
      // an empty candidate gets code 0 everywhere: it is scored like any other (keeps the warps converged) and ignored later
      uint4 v = make_uint4(0, 0, 0, 0);
      if (slot != MGPU_EMPTY_SLOT) v = __ldg((const uint4 *)(a.codes + (size_t)(slot >> 5) * pq_fast_chunk_bytes(a.ng) + ((size_t)g * 2 + u) * 512 + l * 16));
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
      uint8_t *dst = cc + c * mp + 32 * g;
#pragma unroll
      for (int b = 0; b < 16; b++) dst[l ^ (16 * u + b)] = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
    }
  } else {
    for (uint32_t i = tid; i < MGPU_NCAND * m; i += FIN8_THREADS) {
      const uint32_t c = i / m, s = i - c * m;
      const uint32_t slot = sslot[c];
      cc[c * mp + s] = slot != MGPU_EMPTY_SLOT ? a.codes[(size_t)slot * m + s] : (uint8_t)0;
    }
  }
  __syncthreads();
  if (tid < 2 * MGPU_NCAND) {
    const uint32_t c = tid >> 1, h = tid & 1;
    const uint8_t *code = cc + c * mp;
    const float4 *cb4 = (const float4 *)a.cb + h;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    // The m codebook gathers of a candidate are independent, but a register-destination load is scheduled next to its use
    // (2-3 in flight), which makes this loop a chain of L2 round trips.  cp.async has no register destination: FIN8_MLP
    // 16-byte gathers are issued back to back into a thread-private (bank-interleaved) ring, waited for once, then consumed
    // in subspace order.
    // empty and pruned candidates issue no gathers (their key is never read)
    const bool live = sslot[c] != MGPU_EMPTY_SLOT;
    for (uint32_t s0 = 0; live && s0 < m; s0 += FIN8_MLP) {
      const uint32_t nb = min((uint32_t)FIN8_MLP, m - s0);
      for (uint32_t j = 0; j < nb; j++)
        cp_async16(&ring[j * (2 * MGPU_NCAND) + tid], cb4 + ((size_t)(s0 + j) * a.K + code[s0 + j]) * 2);
      cp_async_commit_wait_all();
      for (uint32_t j = 0; j < nb; j++) {
        const float4 x = qv4[(s0 + j) * 2 + h], y = ring[j * (2 * MGPU_NCAND) + tid];
        if (METRIC == MGPU_L2) {
          float d;
          d = __fsub_rn(x.x, y.x); a0 = __fadd_rn(a0, __fmul_rn(d, d));
          d = __fsub_rn(x.y, y.y); a1 = __fadd_rn(a1, __fmul_rn(d, d));
          d = __fsub_rn(x.z, y.z); a2 = __fadd_rn(a2, __fmul_rn(d, d));
          d = __fsub_rn(x.w, y.w); a3 = __fadd_rn(a3, __fmul_rn(d, d));
        } else {
          a0 = __fadd_rn(a0, __fmul_rn(x.x, y.x)); a1 = __fadd_rn(a1, __fmul_rn(x.y, y.y));
          a2 = __fadd_rn(a2, __fmul_rn(x.z, y.z)); a3 = __fadd_rn(a3, __fmul_rn(x.w, y.w));
        }
      }
    }
    // sum_16.reduce_sum() + sum_8.reduce_sum() + sum_4.reduce_sum() + sum_1 with empty 16/4/1 phases (pq/mod.rs:263-265)
    float s8 = -0.0f;
    const int base = lane & ~1;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      s8 = __fadd_rn(s8, __shfl_sync(0xffffffffu, a0, base + j));
      s8 = __fadd_rn(s8, __shfl_sync(0xffffffffu, a1, base + j));
      s8 = __fadd_rn(s8, __shfl_sync(0xffffffffu, a2, base + j));
      s8 = __fadd_rn(s8, __shfl_sync(0xffffffffu, a3, base + j));
    }
    float r = __fadd_rn(__fadd_rn(-0.0f, 0.0f), s8);   // reduce(16 zero lanes) = +0.0
    r = __fadd_rn(r, __fadd_rn(-0.0f, 0.0f));
    r = __fadd_rn(r, 0.0f);
    if (h == 0) skeys[c] = f2key(METRIC == MGPU_L2 ? r : -r);
  }
  __syncthreads();
  if (tid < 32) {
    const uint32_t slot = sslot[lane];
    const bool valid = slot != MGPU_EMPTY_SLOT;
    const uint32_t pid = (uint32_t)a.cand_key[(size_t)q * MGPU_NCAND + lane];
    finalize_tail(a, q, lane, valid, skeys[lane], pid, slot, fkey, nscan);
  }
}

int launch_finalize(mgpu_ctx *ctx, const FinalizeArgs &a) {
  if (a.B == 0) return MGPU_OK;
  if (a.k > MGPU_NCAND) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported by the scan kernels yet", a.k, MGPU_NCAND);
  LaunchScope ls(ctx, MGPU_K_FINALIZE);
  const size_t smem8 = (size_t)a.m * 32 + (size_t)MGPU_NCAND * (((a.m + 15) & ~15u) + 4) + MGPU_NCAND * 8 + (size_t)FIN8_MLP * 2 * MGPU_NCAND * 16;
  static const bool no8 = getenv("MGPU_FINALIZE8") && getenv("MGPU_FINALIZE8")[0] == '0';
  if (a.cb != nullptr && a.dsub == 8 && smem8 <= 96 * 1024 && !no8) {
    if (a.metric == MGPU_L2) {
      CUDA_TRY(ctx, cudaFuncSetAttribute(k_finalize_pq8<MGPU_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
      k_finalize_pq8<MGPU_L2><<<a.B, FIN8_THREADS, smem8, ctx->stream>>>(a);
    } else {
      CUDA_TRY(ctx, cudaFuncSetAttribute(k_finalize_pq8<MGPU_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
      k_finalize_pq8<MGPU_DOT><<<a.B, FIN8_THREADS, smem8, ctx->stream>>>(a);
    }
  } else {
    unsigned grid = (a.B + 3) / 4;
    if (a.metric == MGPU_L2) k_finalize<MGPU_L2><<<grid, 128, 0, ctx->stream>>>(a);
    else k_finalize<MGPU_DOT><<<grid, 128, 0, ctx->stream>>>(a);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- merge of S partial top-k lists per query ------------------------------------------------------------------------
// Concatenate, order by IdWithScore::cmp, truncate to k.  One block per query, rank-by-counting in shared memory.
// A partial count of UINT32_MAX is Spann::search's `None` (spann/index.rs:229-231): it contributes nothing, and the merged
// count is UINT32_MAX only when every partial result is None.
__global__ void k_merge_topk(const mgpu_u128 *__restrict__ docs, const float *__restrict__ scores,
                             const uint32_t *__restrict__ counts, uint32_t S, uint32_t B, uint32_t k,
                             mgpu_u128 *__restrict__ out_docs, float *__restrict__ out_scores,
                             uint32_t *__restrict__ out_counts) {
  extern __shared__ __align__(16) uint8_t sm[];
  const uint32_t n = S * k;
  mgpu_u128 *sd = (mgpu_u128 *)sm;
  uint32_t *sk = (uint32_t *)(sd + n);
  uint32_t *sv = sk + n;  // valid flag
  const uint32_t q = blockIdx.x;
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    uint32_t s = e / k, i = e % k;
    const uint32_t cs = counts[(size_t)s * B + q];
    bool valid = cs != 0xFFFFFFFFu && i < cs;
    size_t src = ((size_t)s * B + q) * k + i;
    sd[e] = valid ? docs[src] : mgpu_u128{0, 0};
    sk[e] = valid ? f2key(scores[src]) : 0xFFFFFFFFu;
    sv[e] = valid;
  }
  __syncthreads();
  uint32_t total = 0;
  for (uint32_t e = 0; e < n; e++) total += sv[e];
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    if (!sv[e]) continue;
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) {
      if (!sv[j]) continue;
      bool less = doc_less(sk[j], sd[j], sk[e], sd[e]);
      bool same = sk[j] == sk[e] && sd[j].lo == sd[e].lo && sd[j].hi == sd[e].hi;
      rank += (less || (same && j < e)) ? 1u : 0u;
    }
    if (rank < k) {
      out_docs[(size_t)q * k + rank] = sd[e];
      out_scores[(size_t)q * k + rank] = key2f(sk[e]);
    }
  }
  if (threadIdx.x == 0) {
    bool all_none = true;
    for (uint32_t s = 0; s < S; s++) all_none = all_none && counts[(size_t)s * B + q] == 0xFFFFFFFFu;
    out_counts[q] = all_none ? 0xFFFFFFFFu : (total < k ? total : k);
  }
}

int launch_merge_topk(mgpu_ctx *ctx, const mgpu_u128 *docs, const float *scores, const uint32_t *counts, uint32_t S,
                      uint32_t B, uint32_t k, mgpu_u128 *out_docs, float *out_scores, uint32_t *out_counts, cudaStream_t st) {
  if (B == 0 || k == 0) return MGPU_OK;
  size_t n = (size_t)S * k;
  size_t smem = n * (sizeof(mgpu_u128) + 8);
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "merge of %zu candidates per query exceeds shared memory", n);
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = n >= 256 ? 256 : (int)((n + 31) / 32 * 32);
  if (!st) st = ctx->stream;
  LaunchScope ls(ctx, MGPU_K_MERGE, st);
  k_merge_topk<<<B, threads, smem, st>>>(docs, scores, counts, S, B, k, out_docs, out_scores, out_counts);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- multi-round top-k (k > 32) -----------------------------------------------------------------------------------------
// After one scan round the query's 32 candidates are sorted by composite (ranking key << 32 | point id).  A full round
// (32 valid entries) only REPORTS the entries strictly below its last composite c32; the group equal to c32 (duplicates of one
// point in several probed lists share a composite) is left to the next round, whose scan ignores composites < c32.  A short
// round has seen every remaining row: everything is reported and later rounds find nothing.
__global__ void k_round_prepare(uint64_t *__restrict__ cand_key, uint32_t *__restrict__ cand_slot, uint32_t B,
                                uint64_t *__restrict__ lower_bound, uint32_t *__restrict__ reported, uint32_t want,
                                uint32_t *__restrict__ unfinished) {
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= B) return;
  const uint64_t key = cand_key[(size_t)q * MGPU_NCAND + lane];
  const uint32_t slot = cand_slot[(size_t)q * MGPU_NCAND + lane];
  const unsigned valid = __ballot_sync(0xffffffffu, slot != MGPU_EMPTY_SLOT);
  if (valid == 0xffffffffu) {
    const uint64_t c32 = shfl64(key, 31);
    if (key == c32) cand_slot[(size_t)q * MGPU_NCAND + lane] = MGPU_EMPTY_SLOT;
    const uint32_t nrep = __popc(__ballot_sync(0xffffffffu, key != c32));
    if (lane == 0) {
      lower_bound[q] = c32;
      // per-query count of reported candidates: the host keeps adding rounds until every query that has not seen a short
      // round holds `want` (= k + spare) of them -- a round that ends inside a group of equal composites reports < 31
      const uint32_t tot = reported[q] + nrep;
      reported[q] = tot;
      if (tot < want) atomicAdd(unfinished, 1u);
    }
  } else if (lane == 0) {
    lower_bound[q] = MGPU_EMPTY_KEY;
  }
}

int launch_round_prepare(mgpu_ctx *ctx, uint64_t *cand_key, uint32_t *cand_slot, uint32_t B, uint64_t *lower_bound,
                         uint32_t *reported, uint32_t want, uint32_t *unfinished) {
  if (B == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_FINALIZE);
  k_round_prepare<<<(B + 3) / 4, 128, 0, ctx->stream>>>(cand_key, cand_slot, B, lower_bound, reported, want, unfinished);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// Merge of the rounds' exact results: the reference keeps the k smallest by (distance, point_id) (bounded heap,
// index.rs:265-274; utils.rs:71-76) and only then maps to doc ids and orders by (score, doc_id) (index.rs:311-326).  Both
// orders by counting in shared memory; one block per query.  pids/scores: [R][B][32], counts: [R][B].
__global__ void k_merge_rounds(const uint32_t *__restrict__ pids, const float *__restrict__ scores,
                               const uint32_t *__restrict__ counts, uint32_t R, uint32_t B, uint32_t k,
                               const mgpu_u128 *__restrict__ doc_ids, int remap, uint32_t *__restrict__ out_pids,
                               mgpu_u128 *__restrict__ out_docs, float *__restrict__ out_scores,
                               uint32_t *__restrict__ out_counts) {
  extern __shared__ __align__(16) uint8_t sm[];
  const uint32_t n = R * MGPU_NCAND;
  mgpu_u128 *sd = (mgpu_u128 *)sm;      // doc id of a selected entry
  uint32_t *sk = (uint32_t *)(sd + n);  // score key
  uint32_t *sp = sk + n;                // point id
  uint32_t *sf = sp + n;                // 0 absent, 1 present, 2 selected (among the k smallest by (distance, point_id))
  const uint32_t q = blockIdx.x;
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    const uint32_t r = e / MGPU_NCAND, i = e % MGPU_NCAND;
    const bool valid = i < counts[(size_t)r * B + q];
    const size_t src = ((size_t)r * B + q) * MGPU_NCAND + i;
    sk[e] = valid ? f2key(scores[src]) : 0xFFFFFFFFu;
    sp[e] = valid ? pids[src] : 0xFFFFFFFFu;
    sf[e] = valid ? 1u : 0u;
  }
  __syncthreads();
  uint32_t total = 0;
  for (uint32_t e = 0; e < n; e++) total += sf[e] ? 1u : 0u;
  const uint32_t cnt = total < k ? total : k;
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    if (!sf[e]) continue;
    uint32_t rank = 0;
    const uint32_t ke = sk[e], pe = sp[e];
    for (uint32_t j = 0; j < n; j++) {
      const uint32_t kj = sk[j], pj = sp[j];   // absent entries carry the maximal (key, pid): never counted
      rank += (kj < ke || (kj == ke && (pj < pe || (pj == pe && j < e)))) ? 1u : 0u;
    }
    if (rank < k) {
      if (!remap) {
        out_pids[(size_t)q * k + rank] = pe;
        out_scores[(size_t)q * k + rank] = key2f(ke);
      } else {
        mgpu_u128 d;
        if (doc_ids) d = doc_ids[pe]; else { d.lo = pe; d.hi = 0; }
        sd[e] = d;
      }
    }
    // publish the selection after the ranks of this pass are final for everyone (flags are only read as != 0 above)
    if (rank < k && remap) atomicExch(&sf[e], 2u);
  }
  __syncthreads();
  if (remap) {
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
      if (sf[e] != 2u) continue;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n; j++) {
        if (sf[j] != 2u) continue;
        const bool less = doc_less(sk[j], sd[j], sk[e], sd[e]);
        const bool same = sk[j] == sk[e] && sd[j].lo == sd[e].lo && sd[j].hi == sd[e].hi;
        rank += (less || (same && j < e)) ? 1u : 0u;
      }
      out_docs[(size_t)q * k + rank] = sd[e];
      out_scores[(size_t)q * k + rank] = key2f(sk[e]);
    }
  }
  if (threadIdx.x == 0) out_counts[q] = cnt;
}

int launch_merge_rounds(mgpu_ctx *ctx, const uint32_t *pids, const float *scores, const uint32_t *counts, uint32_t R, uint32_t B,
                        uint32_t k, const mgpu_u128 *doc_ids, uint32_t *out_pids, mgpu_u128 *out_docs, float *out_scores,
                        uint32_t *out_counts) {
  if (B == 0 || k == 0) return MGPU_OK;
  const size_t n = (size_t)R * MGPU_NCAND;
  const size_t smem = n * (sizeof(mgpu_u128) + 12);
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "merge of %zu candidates per query exceeds shared memory", n);
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_merge_rounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(ctx, MGPU_K_FINALIZE);
  k_merge_rounds<<<B, 256, smem, ctx->stream>>>(pids, scores, counts, R, B, k, doc_ids, out_docs != nullptr ? 1 : 0, out_pids, out_docs,
                                                 out_scores, out_counts);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}


// finalize.cu -- per-query epilogue of a scan: exact re-score of the PQ survivors, the reference's result ordering,
// point-id -> doc-id remap, and the cross-shard merge.
//   search_with_centroids           rs/index/src/ivf/block_based/index.rs:250-285  (order by (distance, point_id))
//   search_with_centroids_and_remap rs/index/src/ivf/block_based/index.rs:298-332  (order by (score, doc_id))
//   Snapshot merge                  rs/index/src/collection/snapshot.rs:49-63,79-108
#include "internal.cuh"
#include "pq_device.cuh"

__device__ __forceinline__ float key2f(uint32_t key) {
  uint32_t u = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
  return __uint_as_float(u);
}

struct FastLayoutCode {
  const uint8_t *codes; uint32_t slot, ng;
  __device__ __forceinline__ uint32_t operator()(uint32_t s) const {
    uint32_t g = s >> 5, j = s & 31, l = slot & 31, t = j ^ l;
    size_t chunk = slot >> 5;
    return codes[((chunk * ng + g) * 2 + (t >> 4)) * 512 + l * 16 + (t & 15)];
  }
};

__device__ __forceinline__ bool doc_less(uint32_t ka, mgpu_u128 a, uint32_t kb, mgpu_u128 b) {
  if (ka != kb) return ka < kb;
  if (a.hi != b.hi) return a.hi < b.hi;
  return a.lo < b.lo;
}

template <int METRIC>
__global__ void __launch_bounds__(128) k_finalize(FinalizeArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.B) return;
  uint64_t ckey = a.cand_key[(size_t)q * MGPU_NCAND + lane];
  uint32_t slot = a.cand_slot[(size_t)q * MGPU_NCAND + lane];
  const bool valid = slot != MGPU_EMPTY_SLOT;
  const uint32_t pid = (uint32_t)ckey;
  uint32_t skey = (uint32_t)(ckey >> 32);
  if (a.cb != nullptr && valid) {
    // exact Quantizer::distance(quantized_query, row, StreamingSIMD) (index.rs:203-207, pq/mod.rs:231-266)
    float d;
    RowMajorCode qc{a.qcodes + (size_t)q * a.m};
    if (a.pq_fast) d = pq_distance_streaming<METRIC>(a.cb, a.m, a.K, a.dsub, qc, FastLayoutCode{a.codes, slot, a.ng});
    else d = pq_distance_streaming<METRIC>(a.cb, a.m, a.K, a.dsub, qc, RowMajorCode{a.codes + (size_t)slot * a.m});
    skey = f2key(d);
  }
  WarpTop32 w;
  w.key = valid ? (((uint64_t)skey << 32) | pid) : MGPU_EMPTY_KEY;
  w.pay = slot;
  w.sort();  // ascending (distance, point_id): PointAndDistance::cmp (rs/index/src/utils.rs:71-76)
  const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, w.key != MGPU_EMPTY_KEY));
  const uint32_t count = min(a.k, nvalid);
  const uint32_t rk = (uint32_t)(w.key >> 32), rpid = (uint32_t)w.key;
  const float score = key2f(rk);
  if (a.out_pids && (uint32_t)lane < count) {
    a.out_pids[(size_t)q * a.k + lane] = rpid;
    if (!a.out_docs) a.out_scores[(size_t)q * a.k + lane] = score;
  }
  if (a.out_docs) {
    mgpu_u128 doc;
    if ((uint32_t)lane < count && a.doc_ids) doc = a.doc_ids[rpid];
    else { doc.lo = rpid; doc.hi = 0; }
    // IdWithScore::cmp (utils.rs:95-114): (score, doc_id); rank by counting among the kept lanes
    uint32_t rank = 0;
    for (uint32_t j = 0; j < count; j++) {
      uint32_t kj = __shfl_sync(0xffffffffu, rk, j);
      mgpu_u128 dj;
      dj.lo = shfl64(doc.lo, j); dj.hi = shfl64(doc.hi, j);
      bool less = doc_less(kj, dj, rk, doc);
      bool same = (kj == rk) && dj.lo == doc.lo && dj.hi == doc.hi;
      rank += (less || (same && j < (uint32_t)lane)) ? 1u : 0u;
    }
    if ((uint32_t)lane < count) {
      a.out_docs[(size_t)q * a.k + rank] = doc;
      a.out_scores[(size_t)q * a.k + rank] = score;
    }
  }
  if (lane == 0 && a.out_counts) a.out_counts[q] = count;
}

int launch_finalize(mgpu_ctx *ctx, const FinalizeArgs &a) {
  if (a.B == 0) return MGPU_OK;
  if (a.k > MGPU_NCAND) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported by the scan kernels yet", a.k, MGPU_NCAND);
  unsigned grid = (a.B + 3) / 4;
  LaunchScope ls(ctx, MGPU_K_FINALIZE);
  if (a.metric == MGPU_L2) k_finalize<MGPU_L2><<<grid, 128, 0, ctx->stream>>>(a);
  else k_finalize<MGPU_DOT><<<grid, 128, 0, ctx->stream>>>(a);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- merge of S partial top-k lists per query ------------------------------------------------------------------------
// Concatenate, order by IdWithScore::cmp, truncate to k.  One block per query, rank-by-counting in shared memory.
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
    bool valid = i < counts[(size_t)s * B + q];
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
  if (threadIdx.x == 0) out_counts[q] = total < k ? total : k;
}

int launch_merge_topk(mgpu_ctx *ctx, const mgpu_u128 *docs, const float *scores, const uint32_t *counts, uint32_t S,
                      uint32_t B, uint32_t k, mgpu_u128 *out_docs, float *out_scores, uint32_t *out_counts) {
  if (B == 0 || k == 0) return MGPU_OK;
  size_t n = (size_t)S * k;
  size_t smem = n * (sizeof(mgpu_u128) + 8);
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "merge of %zu candidates per query exceeds shared memory", n);
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_merge_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = n >= 256 ? 256 : (int)((n + 31) / 32 * 32);
  LaunchScope ls(ctx, MGPU_K_MERGE);
  k_merge_topk<<<B, threads, smem, ctx->stream>>>(docs, scores, counts, S, B, k, out_docs, out_scores, out_counts);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

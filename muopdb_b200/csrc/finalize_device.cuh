// finalize_device.cuh -- device helpers of the per-query epilogue shared by finalize.cu and the exact fallback scan (scan.cu).
#pragma once
#include "internal.cuh"

__device__ __forceinline__ float key2f(uint32_t key) {
  uint32_t u = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
  return __uint_as_float(u);
}

struct FastLayoutCode {
  const uint8_t *codes; uint32_t slot, ng;
  __device__ __forceinline__ uint32_t operator()(uint32_t s) const {
    return codes[pq_fast_code_offset(slot >> 5, ng, slot & 31, s)];
  }
};

__device__ __forceinline__ bool doc_less(uint32_t ka, mgpu_u128 a, uint32_t kb, mgpu_u128 b) {
  if (ka != kb) return ka < kb;
  if (a.hi != b.hi) return a.hi < b.hi;
  return a.lo < b.lo;
}

// Which of a query's 32 candidates can still reach the exact top k (FinalizeArgs::prune)?  Called by one full warp; lane =
// candidate.  Returns this lane's verdict.  err16: keys of the 16-bit scan (bound = k-th key + 2 E); otherwise the 32-bit
// per-query-scale keys (relative 2^-12 + 8192 units).
__device__ __forceinline__ bool prune_keep(uint32_t k, bool valid, uint32_t key, int lane, bool key16 = false, uint32_t m = 0,
                                           float slack = 1.0f) {
  const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
  if (nvalid <= k) return valid;
  // k-th smallest key among the valid candidates, by counting (keys may repeat)
  uint32_t rank = 0;
#pragma unroll 8
  for (int j = 0; j < 32; j++) {
    const uint32_t kj = __shfl_sync(0xffffffffu, key, j);
    const bool vj = __shfl_sync(0xffffffffu, (int)valid, j);
    rank += (vj && (kj < key || (kj == key && j < lane))) ? 1u : 0u;
  }
  const unsigned who = __ballot_sync(0xffffffffu, valid && rank == k - 1);
  const uint32_t kth = __shfl_sync(0xffffffffu, key, __ffs(who) - 1);
  const uint64_t bound = key16 ? (uint64_t)kth + 2ull * key16_err(m, key, slack) : (uint64_t)kth + (kth >> 12) + 8192u;
  return valid && (uint64_t)key <= bound;
}

// Ordering + remap tail shared by both finalize kernels; called by one full warp per query.  `skey` is the exact score
// key of this lane's candidate, `valid` whether the lane holds one.
// fkey / nscan (16-bit scan only): this lane's fixed-point ranking key and the number of candidates the scan produced.
// Returns false when the answer could not be certified (the caller has appended the query to the fallback list).
__device__ __forceinline__ bool finalize_tail(const FinalizeArgs &a, uint32_t q, int lane, bool valid, uint32_t skey,
                                              uint32_t pid, uint32_t slot, uint32_t fkey = 0, uint32_t nscan = 0) {
  WarpTop32 w;
  w.key = valid ? (((uint64_t)skey << 32) | pid) : MGPU_EMPTY_KEY;
  w.pay = (uint32_t)lane;
  w.sort();  // ascending (distance, point_id): PointAndDistance::cmp (rs/index/src/utils.rs:71-76)
  const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, w.key != MGPU_EMPTY_KEY));
  const uint32_t count = min(a.k, nvalid);
  if (a.key16 && nscan == MGPU_NCAND) {
    // Certification.  A row outside the candidate list has key >= K32 (the largest candidate key), and
    //   gscale * exact score >= key - E(key) >= K32 - E(K32)        (E: entry rounding + fp32 summation-order slack);
    // the k-th best candidate's EXACT score s_k is known by now, so the answer is exact whenever gscale * s_k < K32 - E(K32).
    const uint32_t k32 = __reduce_max_sync(0xffffffffu, fkey);
    const uint32_t sk_kth = __shfl_sync(0xffffffffu, (uint32_t)(w.key >> 32), (int)(count ? count - 1 : 0));
    const double lhs = (double)a.gscale * (double)key2f(sk_kth) * (1.0 + 1.0e-6) + 1.0 + (double)key16_err(a.m, k32, a.cert_slack);
    if (count < a.k || !(lhs < (double)k32)) {
      if (lane == 0) a.uncert_list[atomicAdd(a.uncert_count, 1u)] = q;
      return false;
    }
  }
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
  return true;
}


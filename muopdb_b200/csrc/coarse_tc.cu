// coarse_tc.cu -- IVF coarse scoring on the 5th-gen tensor cores (tcgen05 + TMEM, operands fed by TMA), followed by an
// exact fp32 re-score of a provably sufficient candidate set.
//
// Replaces the arithmetic of BlockBasedIvf::find_nearest_centroids (rs/index/src/ivf/block_based/index.rs:147-163): the
// reference computes sqrt(sum (q-c)^2) for every centroid in fp32 and keeps the nprobe smallest.  Here
//   1. k_split_bf16      x -> (hi, lo) bf16 pair, x ~= hi + lo; rows laid out [hi | hi | lo] (queries) / [hi | lo | hi]
//                        (centroids) so that ONE K = 3*dim bf16 GEMM yields q.c with ~2^-17 relative error;
//   2. k_coarse_gemm     D~[b][c] = |q|^2 + |c|^2 - 2 q.c  : 128x256 tile per CTA, 4-stage TMA->smem ring (SWIZZLE_128B),
//                        one elected thread issues tcgen05.mma (M128 N256 K16, fp32 accumulators in 256 TMEM columns),
//                        4 epilogue warps read TMEM with tcgen05.ld and fuse the norm terms;
//   3. k_coarse_select   per query: radix-select tau = nprobe-th smallest D~, candidates = {c : D~ <= tau + 2 eps}, exact
//                        bit-faithful sqrt-L2 for the candidates (16-lane order), final (distance, index) ordering.
// With |D~ - d_ref| <= eps for every centroid, the candidate set contains the reference's nprobe nearest (DESIGN.md
// "coarse scoring"), so the probe lists are IDENTICAL to the exact path's, including ties.
#include <cuda.h>
#include <cuda_bf16.h>

#include "internal.cuh"
#include "scan_common.cuh"

#define TC_BM 128
#define TC_BN 256
#define TC_BK 64
#define TC_STAGES 4
#define TC_THREADS 192
#define TC_ERR_REL 3.0e-4f

// Diagnostic for the margin rule's silent cliff: [0] centroids in the uncertain band (re-scored exactly),
// [1] queries, summed over every selection since the last reset.  On centred data the band holds a handful of centroids per
// query; on un-centred data (|x|^2 >> distances) it grows towards all C -- still exact, but the selection then costs a full
// exact scoring.  Read by mgpu_coarse_band_stats.
__device__ unsigned long long g_coarse_band[2];


// ---- split ---------------------------------------------------------------------------------------------------------------
__global__ void k_split_bf16(const float *__restrict__ X, uint64_t n, uint32_t dim, uint32_t Kp, int is_centroid,
                             __nv_bfloat16 *__restrict__ out, float *__restrict__ norms) {
  uint64_t row = blockIdx.x;
  const float *x = X + row * dim;
  __nv_bfloat16 *o = out + row * Kp;
  float part = 0.0f;
  if ((dim & 3u) == 0 && ((uintptr_t)x & 15u) == 0) {
    // four elements per thread: one 16-byte load, three 8-byte stores (rows are 16-byte aligned: dim % 4 == 0, Kp % 64 == 0)
    const float4 *x4 = (const float4 *)x;
    for (uint32_t d4 = threadIdx.x; d4 < dim / 4; d4 += blockDim.x) {
      const float4 v = x4[d4];
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
      uint2 H, L;
      H.x = *(const uint32_t *)&h01; H.y = *(const uint32_t *)&h23;
      L.x = *(const uint32_t *)&l01; L.y = *(const uint32_t *)&l23;
      *(uint2 *)(o + 4 * d4) = H;
      *(uint2 *)(o + dim + 4 * d4) = is_centroid ? L : H;
      *(uint2 *)(o + 2 * dim + 4 * d4) = is_centroid ? H : L;
      part += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
      float v = x[d];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      o[d] = h;
      o[dim + d] = is_centroid ? l : h;
      o[2 * dim + d] = is_centroid ? h : l;
      part += v * v;
    }
  }
  for (uint32_t d = 3 * dim + threadIdx.x; d < Kp; d += blockDim.x) o[d] = __float2bfloat16_rn(0.0f);
  __shared__ float sred[32];
  for (int o2 = 16; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; w++) s += sred[w];
    norms[row] = s;
  }
}

// ---- tcgen05 / TMA wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (cute/arch/mma_sm100_desc.hpp)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

struct TcSmem {
  // stage buffers first (1024-byte aligned for SWIZZLE_128B)
  uint8_t a[TC_STAGES][TC_BM * TC_BK * 2];
  uint8_t b[TC_STAGES][TC_BN * TC_BK * 2];
  uint64_t full[TC_STAGES], empty[TC_STAGES], accum_full;
  uint32_t tmem_base;
};

// grid = (ceil(C / 256), ceil(B / 128)); one output tile per CTA.
__global__ void __launch_bounds__(TC_THREADS, 1)
k_coarse_gemm(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, const float *__restrict__ qn,
              const float *__restrict__ cn, uint32_t B, uint32_t C, uint32_t Kp, float *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem &s = *reinterpret_cast<TcSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
  const uint32_t nkb = Kp / TC_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; i++) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 256 fp32 accumulator columns; the allocating warp also frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem_base;

  if (warp == 0) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      for (uint32_t kb = 0; kb < nkb; kb++) {
        const uint32_t st = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
        mbar_wait(&s.empty[st], ph ^ 1);
        mbar_expect_tx(&s.full[st], (TC_BM + TC_BN) * TC_BK * 2);
        tma_load_2d(s.a[st], &map_q, &s.full[st], (int)(kb * TC_BK), (int)m0);
        tma_load_2d(s.b[st], &map_c, &s.full[st], (int)(kb * TC_BK), (int)n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      // instruction descriptor (kind::f16): c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (uint32_t kb = 0; kb < nkb; kb++) {
        const uint32_t st = kb % TC_STAGES, ph = (kb / TC_STAGES) & 1;
        mbar_wait(&s.full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = umma_desc_sw128(smem_u32(s.a[st])), bd = umma_desc_sw128(smem_u32(s.b[st]));
#pragma unroll
        for (int k = 0; k < TC_BK / 16; k++)  // +32 bytes (encoded +2) per K=16 step inside the 128-byte swizzle atom
          umma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0u);
        umma_commit(&s.empty[st]);  // frees the smem stage when these MMAs retire
      }
      umma_commit(&s.accum_full);   // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const uint32_t quarter = warp & 3;
    mbar_wait(&s.accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t row = m0 + quarter * 32 + lane;
    const float qnr = row < B ? qn[row] : 0.0f;
#pragma unroll 1
    for (int cchunk = 0; cchunk < TC_BN / 32; cchunk++) {
      uint32_t r[32];
      tmem_ld32(tmem + ((quarter * 32u) << 16) + (uint32_t)(cchunk * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const uint32_t col0 = n0 + cchunk * 32;
      if (row < B) {
        float *dst = out + (size_t)row * C + col0;
        if (col0 + 32 <= C && (C & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float4 v;
            v.x = qnr + cn[col0 + 4 * j + 0] - 2.0f * __uint_as_float(r[4 * j + 0]);
            v.y = qnr + cn[col0 + 4 * j + 1] - 2.0f * __uint_as_float(r[4 * j + 1]);
            v.z = qnr + cn[col0 + 4 * j + 2] - 2.0f * __uint_as_float(r[4 * j + 2]);
            v.w = qnr + cn[col0 + 4 * j + 3] - 2.0f * __uint_as_float(r[4 * j + 3]);
            *(float4 *)(dst + 4 * j) = v;
          }
        } else {
          for (int j = 0; j < 32; j++)
            if (col0 + j < C) dst[j] = qnr + cn[col0 + j] - 2.0f * __uint_as_float(r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// ---- selection with a provable margin + exact re-score -------------------------------------------------------------------
// One CTA per query.  Let d_c be the reference's fp32 squared distance, D~_c the tensor-core value, |D~_c - d_c| <= eps,
// tau the nprobe-th smallest D~ and [lo, hi] a key interval known to contain tau (a 22-bit radix bin).  Then
//   * D~_c < lo - 2 eps  =>  d_c < tau - eps <= d_(nprobe): c is in EVERY valid answer ("sure");
//   * every member of the answer has D~_c <= tau + 2 eps <= hi + 2 eps ("band" = the rest of those).
// need_order != 0 (mgpu_ivf_coarse: nearest-first ids + distances): every centroid with D~ <= hi + 2 eps is re-scored with
// the bit-faithful sqrt-L2 and ordered by (distance, index) -- identical to the exact path.
// need_order == 0 (search: the result of scanning a probe SET does not depend on its order, index.rs:265-274): sure
// centroids are emitted directly, only the band is re-scored and its (nprobe - #sure) nearest complete the set.
// smem: ckey[cap] (u64; its first C words double as the radix keys until the band is known), cand[cap], query, hist.
#define SEL_THREADS 512
#define SEL_BINS 2048
__device__ __forceinline__ float sel_key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

__global__ void __launch_bounds__(SEL_THREADS, 3)
k_coarse_select(const float *__restrict__ Dt, const float *__restrict__ Q, const float *__restrict__ centroids,
                const float *__restrict__ qn, float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, uint32_t cand_cap,
                int need_order, uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ overflow,
                const uint32_t *__restrict__ only_flagged) {
  if (only_flagged && !only_flagged[blockIdx.x]) return;  // fallback launch: queries the warp kernel could not finish
  extern __shared__ __align__(16) uint8_t sm[];
  uint64_t *ckey = (uint64_t *)sm;                 // cand_cap (power of two >= C)
  uint32_t *keys = (uint32_t *)sm;                 // C, aliases ckey: dead before the first ckey store
  uint32_t *cand = (uint32_t *)(ckey + cand_cap);  // cand_cap
  float *sq = (float *)(cand + cand_cap);          // dim
  uint32_t *hist = (uint32_t *)(sq + ((dim + 3) & ~3u));  // SEL_BINS
  uint32_t *misc = hist + SEL_BINS;                // [0] bin, [1] remaining rank, [2] band count, [3] sure count, [4] kmin, [5] kmax
  const uint32_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *row = Dt + (size_t)q * C;
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  for (uint32_t i = tid; i < C / 4; i += SEL_THREADS) {  // C % 4 == 0 (launcher)
    const float4 v = __ldg((const float4 *)row + i);
    uint4 k4 = make_uint4(f2key(v.x), f2key(v.y), f2key(v.z), f2key(v.w));
    ((uint4 *)keys)[i] = k4;
    kmin = min(min(kmin, k4.x), min(k4.y, min(k4.z, k4.w)));
    kmax = max(max(kmax, k4.x), max(k4.y, max(k4.z, k4.w)));
  }
  for (uint32_t d = tid; d < dim; d += SEL_THREADS) sq[d] = Q[(size_t)q * dim + d];
  for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
  if (tid == 0) { misc[0] = 0; misc[1] = nprobe - 1; misc[2] = 0; misc[3] = 0; misc[4] = 0xFFFFFFFFu; misc[5] = 0u; }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if (lane == 0) { atomicMin(&misc[4], kmin); atomicMax(&misc[5], kmax); }
  __syncthreads();
  // two-level histogram over the query's OWN key range [kmin, kmax] (keys are monotone in the distance): level 1 splits
  // the range into <= 2048 equal key intervals, level 2 splits the interval holding rank nprobe-1 again.  The result is a
  // key interval [lo_key, hi_key] of width range / 2^22 (or a single key) that contains tau.
  uint32_t base = misc[4];
  uint32_t width_log = 32 - __clz((misc[5] - base) | 1u);          // bits needed for (k - base)
  uint32_t lo_key = base, hi_key = misc[5];
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t shift = width_log > 11 ? width_log - 11 : 0;      // bin = (k - base) >> shift  in [0, 2048)
    for (uint32_t i = tid; i < C; i += SEL_THREADS) {
      const uint32_t k = keys[i];
      if (k >= lo_key && k <= hi_key) atomicAdd(&hist[(k - base) >> shift], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      constexpr int PER = SEL_BINS / 32;
      uint32_t sum = 0;
      for (int j = 0; j < PER; j++) sum += hist[lane * PER + j];
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t rank = misc[1];
      const uint32_t excl = incl - sum;
      if (rank >= excl && rank < incl) {  // exactly one lane
        uint32_t r = rank - excl;
        int bb = 0;
        for (; bb < PER - 1; bb++) { uint32_t h = hist[lane * PER + bb]; if (r < h) break; r -= h; }
        misc[0] = (uint32_t)(lane * PER + bb);
        misc[1] = r;
      }
    }
    __syncthreads();
    const uint32_t bsel = misc[0];
    lo_key = base + (bsel << shift);
    hi_key = shift ? lo_key + ((1u << shift) - 1u) : lo_key;
    if (shift == 0) break;
    base = lo_key;
    width_log = shift;
    if (pass == 0) {
      for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
    }
  }
  const float eps = TC_ERR_REL * (qn[q] + cn_max);
  const uint32_t lim_hi = f2key(sel_key2f(hi_key) + 2.0f * eps);
  const uint32_t lim_lo = need_order ? 0u : f2key(sel_key2f(lo_key) - 2.0f * eps);
  // classification: sure (emitted right away, set-only mode) / band (to be re-scored)
  for (uint32_t i0 = 0; i0 < C; i0 += SEL_THREADS) {
    const uint32_t i = i0 + tid;
    const uint32_t k = i < C ? keys[i] : 0xFFFFFFFFu;
    const bool sure = i < C && k < lim_lo;
    const bool band = i < C && !sure && k <= lim_hi;
    const unsigned ms = __ballot_sync(0xffffffffu, sure), mb = __ballot_sync(0xffffffffu, band);
    uint32_t bs = 0, bb = 0;
    if (lane == 0) {
      if (ms) bs = atomicAdd(&misc[3], (uint32_t)__popc(ms));
      if (mb) bb = atomicAdd(&misc[2], (uint32_t)__popc(mb));
    }
    bs = __shfl_sync(0xffffffffu, bs, 0);
    bb = __shfl_sync(0xffffffffu, bb, 0);
    const unsigned below = (1u << lane) - 1;
    if (sure) { uint32_t pos = bs + __popc(ms & below); if (pos < nprobe) out_ids[(size_t)q * nprobe + pos] = i; }
    if (band) { uint32_t pos = bb + __popc(mb & below); if (pos < cand_cap) cand[pos] = i; }
  }
  __syncthreads();
  uint32_t ncand = misc[2];
  const uint32_t nsure = min(misc[3], nprobe);
  const uint32_t need = nprobe - nsure;
  if (ncand > cand_cap) { if (tid == 0) atomicAdd(overflow, 1u); ncand = cand_cap; }
  if (tid == 0) { atomicAdd(&g_coarse_band[0], (unsigned long long)ncand); atomicAdd(&g_coarse_band[1], 1ull); }
  if (!need_order && ncand <= need) {
    // the whole band belongs to the answer: nothing to decide, nothing to re-score
    for (uint32_t i = tid; i < ncand; i += SEL_THREADS) out_ids[(size_t)q * nprobe + nsure + i] = cand[i];
    return;
  }
  // exact sqrt-L2 (l2.rs:30-74) for every band centroid
  const int n = (int)dim;
  if ((dim & 3u) == 0) {
    // 4 lanes per pair: lane r of a group owns lane-accumulators 4r..4r+3 of the reference's 16 (float4 loads)
    const int r = lane & 3, grp = lane >> 2;
    const uint32_t per_round = (SEL_THREADS / 32) * 8;
    for (uint32_t base = 0; base < ncand; base += per_round) {
      if (base + warp * 8 >= ncand) continue;  // warp-uniform: nothing for this warp in this round
      const uint32_t j = base + warp * 8 + grp;
      const uint32_t jj = j < ncand ? j : ncand - 1;
      const uint32_t cidx = cand[jj];
      const float *crow = centroids + (size_t)cidx * dim;
      const int chunks = n / 16;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
      const float4 *c4 = (const float4 *)crow + r;
      const float4 *q4 = (const float4 *)sq + r;
      int c = 0;
      for (; c + 4 <= chunks; c += 4) {
        float4 y[4];
#pragma unroll
        for (int u = 0; u < 4; u++) y[u] = __ldg(c4 + (c + u) * 4);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const float4 x = q4[(c + u) * 4];
          float d;
          d = __fsub_rn(x.x, y[u].x); a0 = __fadd_rn(a0, __fmul_rn(d, d));
          d = __fsub_rn(x.y, y[u].y); a1 = __fadd_rn(a1, __fmul_rn(d, d));
          d = __fsub_rn(x.z, y[u].z); a2 = __fadd_rn(a2, __fmul_rn(d, d));
          d = __fsub_rn(x.w, y[u].w); a3 = __fadd_rn(a3, __fmul_rn(d, d));
        }
      }
      for (; c < chunks; c++) {
        const float4 y = __ldg(c4 + c * 4), x = q4[c * 4];
        float d;
        d = __fsub_rn(x.x, y.x); a0 = __fadd_rn(a0, __fmul_rn(d, d));
        d = __fsub_rn(x.y, y.y); a1 = __fadd_rn(a1, __fmul_rn(d, d));
        d = __fsub_rn(x.z, y.z); a2 = __fadd_rn(a2, __fmul_rn(d, d));
        d = __fsub_rn(x.w, y.w); a3 = __fadd_rn(a3, __fmul_rn(d, d));
      }
      float s2 = -0.0f;  // ordered lane reduction 0..15
      const int gl = lane & ~3;
#pragma unroll
      for (int l = 0; l < 4; l++) {
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a0, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a1, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a2, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a3, gl + l));
      }
      float ret = __fadd_rn(0.0f, s2);
      const int p = chunks * 16;
      if (p < n) ret = ref_tail<MGPU_L2>(PtrAcc{sq}, PtrAcc{crow}, p, n, ret);
      const float dist = sqrtf(ret);
      if (j < ncand && r == 0) ckey[j] = ((uint64_t)f2key(dist) << 32) | cidx;
    }
  } else {
    // half-warp per pair, lane h owns lane-accumulator h
    const int h = lane & 15, half = lane >> 4;
    const uint32_t per_round = (SEL_THREADS / 32) * 2;
    for (uint32_t base = 0; base < ncand; base += per_round) {
      if (base + warp * 2 >= ncand) continue;
      const uint32_t j = base + warp * 2 + half;
      const uint32_t jj = j < ncand ? j : ncand - 1;
      const uint32_t cidx = cand[jj];
      const float *crow = centroids + (size_t)cidx * dim;
      const int chunks = n / 16;
      float acc = 0.0f;
#pragma unroll 8
      for (int c = 0; c < chunks; c++) {
        float d = __fsub_rn(sq[c * 16 + h], __ldg(crow + c * 16 + h));
        acc = __fadd_rn(acc, __fmul_rn(d, d));
      }
      float s2 = -0.0f;
      const int basel = lane & 16;
#pragma unroll
      for (int l = 0; l < 16; l++) s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, acc, basel + l));
      float ret = __fadd_rn(0.0f, s2);
      const int p = chunks * 16;
      if (p < n) ret = ref_tail<MGPU_L2>(PtrAcc{sq}, PtrAcc{crow}, p, n, ret);
      const float dist = sqrtf(ret);
      if (j < ncand && h == 0) ckey[j] = ((uint64_t)f2key(dist) << 32) | cidx;
    }
  }
  __syncthreads();
  uint32_t *oi = out_ids + (size_t)q * nprobe + nsure;
  float *od = out_dist ? out_dist + (size_t)q * nprobe + nsure : nullptr;
  if (ncand <= 2048) {
    // order by (distance total order, centroid index): rank by counting (keys are distinct: the index is part of the key)
    for (uint32_t i = tid; i < ncand; i += SEL_THREADS) {
      const uint64_t kk = ckey[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < ncand; j++) rank += ckey[j] < kk ? 1u : 0u;
      if (rank < need) {
        oi[rank] = (uint32_t)kk;
        if (od) od[rank] = sel_key2f((uint32_t)(kk >> 32));
      }
    }
    return;
  }
  for (uint32_t i = ncand + tid; i < cand_cap; i += SEL_THREADS) ckey[i] = MGPU_EMPTY_KEY;
  __syncthreads();
  uint32_t P2 = 1;
  while (P2 < ncand) P2 <<= 1;
  for (uint32_t k = 2; k <= P2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < P2; i += SEL_THREADS) {
        uint32_t ixj = i ^ j;
        if (ixj > i) {
          uint64_t x = ckey[i], y = ckey[ixj];
          bool up = (i & k) == 0;
          if ((x > y) == up) { ckey[i] = y; ckey[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = tid; i < need; i += SEL_THREADS) {
    uint64_t kk = ckey[i];
    oi[i] = (uint32_t)kk;
    if (od) od[i] = sel_key2f((uint32_t)(kk >> 32));
  }
}

// ---- small-CTA selection (the fast path) -------------------------------------------------------------------------------------
// Same rule as k_coarse_select, one CTA of 4 warps per query, built for latency: the D~ row (C <= 4096 values) is pulled
// into shared memory with ONE burst of cp.async (every 16-byte piece in flight at once), converted to ordered keys in
// place, and all later passes (range, two histogram levels, classification) run at shared-memory speed.  Sure centroids
// are emitted in ascending D~ order (the scan prunes best when the nearest lists come first), the band is re-scored
// exactly and appended.  Queries whose band does not fit the candidate lists are flagged and finished by
// k_coarse_select.  Optionally emits work[q] = number of 32-row chunks in the chosen lists (the scan's LPT schedule).
#define SELQ_THREADS 128
#define SELQ_WARPS (SELQ_THREADS / 32)
#define SELW_CAP 256
#define SELW_MAXC 4096
#define SELW_BINS 1024
#define SELW_BYTES (SELW_MAXC * 4 + SELW_BINS * 4 + SELW_CAP * 8 + SELW_CAP * 8 + SELW_CAP * 4 + 64)
#define SELW_MLP 8   /* 4 warps x 8 x 32 lanes x 16 B = the 16 KB key space */
__device__ __forceinline__ void selw_cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void selw_cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(SELQ_THREADS)
k_coarse_select_q(const float *__restrict__ Dt, const float *__restrict__ Q, const float *__restrict__ centroids,
                  const float *__restrict__ qn, float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, int need_order,
                  uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ flags,
                  const uint32_t *__restrict__ chunk_start, uint32_t *__restrict__ work_out) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t q = blockIdx.x;
  uint32_t *keys = (uint32_t *)sm;                       // C ordered keys (first the raw floats)
  uint32_t *hist = keys + SELW_MAXC;                     // SELW_BINS
  uint64_t *skey = (uint64_t *)(hist + SELW_BINS);       // SELW_CAP: (approximate key, index) of the sure centroids
  uint64_t *ckey = skey + SELW_CAP;                      // SELW_CAP: (exact key, index) of the band
  uint32_t *cand = (uint32_t *)(ckey + SELW_CAP);        // SELW_CAP: band indices
  uint32_t *misc = cand + SELW_CAP;                      // [0] sure [1] band [2] bin [3] rank [4..7] warp sums [8] kmin [9] kmax [10] work
  const float4 *row4 = (const float4 *)(Dt + (size_t)q * C);
  const uint32_t C4 = C / 4;
  // ---- the row, one burst; every thread converts the pieces it fetched itself (no barrier needed in between)
  for (uint32_t i = tid; i < C4; i += SELQ_THREADS) selw_cp_async16((uint4 *)keys + i, row4 + i);
  for (int i = tid; i < SELW_BINS / 4; i += SELQ_THREADS) ((uint4 *)hist)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { misc[0] = 0; misc[1] = 0; misc[3] = nprobe - 1; misc[8] = 0xFFFFFFFFu; misc[9] = 0u; misc[10] = 0u; misc[11] = 0u; }
  selw_cp_async_wait();
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll 4
  for (uint32_t i = tid; i < C4; i += SELQ_THREADS) {
    const float4 v = ((const float4 *)keys)[i];
    const uint4 k4 = make_uint4(f2key(v.x), f2key(v.y), f2key(v.z), f2key(v.w));
    ((uint4 *)keys)[i] = k4;
    kmin = min(min(kmin, k4.x), min(k4.y, min(k4.z, k4.w)));
    kmax = max(max(kmax, k4.x), max(k4.y, max(k4.z, k4.w)));
  }
  // Pre-filter: every warp sorts its 32 per-thread minima (shuffle bitonic network) and takes the ceil(nprobe/4)-th smallest;
  // the maximum of the four is an upper bound T of tau (4 x ceil(nprobe/4) >= nprobe distinct keys are <= it), and only a few
  // percent of the keys are <= T.  The histogram passes below then touch ~100 keys instead of all C with shared-memory
  // atomics (2 cycles per lane each).
  uint32_t tsort = kmin;
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const uint32_t other = __shfl_xor_sync(0xffffffffu, tsort, j);
      const bool up = (lane & k2) == 0, lower = (lane & j) == 0;
      tsort = (lower == up) ? min(tsort, other) : max(tsort, other);
    }
  }
  const uint32_t twarp = __shfl_sync(0xffffffffu, tsort, (int)min((nprobe + 3) / 4, 32u) - 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __syncthreads();  // misc / hist initialised
  if (lane == 0) { atomicMin(&misc[8], kmin); atomicMax(&misc[9], kmax); atomicMax(&misc[11], twarp); }
  __syncthreads();
  // ---- two histogram levels over [kmin, T]: the key interval [lo_key, hi_key] of width range / 2^20 holding tau
  uint32_t base = misc[8];
  const uint32_t top_key = nprobe <= SELQ_THREADS ? min(misc[11], misc[9]) : misc[9];
  uint32_t width_log = 32 - __clz((top_key - base) | 1u);
  uint32_t lo_key = base, hi_key = top_key;
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t shift = width_log > 10 ? width_log - 10 : 0;
#pragma unroll 4
    for (uint32_t i = tid; i < C4; i += SELQ_THREADS) {
      const uint4 k4 = ((const uint4 *)keys)[i];
      if (k4.x >= lo_key && k4.x <= hi_key) atomicAdd(&hist[(k4.x - base) >> shift], 1u);
      if (k4.y >= lo_key && k4.y <= hi_key) atomicAdd(&hist[(k4.y - base) >> shift], 1u);
      if (k4.z >= lo_key && k4.z <= hi_key) atomicAdd(&hist[(k4.z - base) >> shift], 1u);
      if (k4.w >= lo_key && k4.w <= hi_key) atomicAdd(&hist[(k4.w - base) >> shift], 1u);
    }
    __syncthreads();
    // thread t owns bins [8 t, 8 t + 8); rotated read order spreads a warp over the banks
    constexpr int PER = SELW_BINS / SELQ_THREADS;
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { loc[j] = hist[tid * PER + ((j + lane) & (PER - 1))]; sum += loc[j]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) misc[4 + w] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int j = 0; j < SELQ_WARPS; j++) before += j < w ? misc[4 + j] : 0u;
    const uint32_t rank = misc[3];
    const uint32_t excl = before + incl - sum;
    if (rank >= excl && rank < excl + sum) {  // exactly one thread
      uint32_t rr = rank - excl;
      int bb = 0;
      for (; bb < PER - 1; bb++) { uint32_t h = hist[tid * PER + bb]; if (rr < h) break; rr -= h; }
      misc[2] = (uint32_t)(tid * PER + bb);
      misc[3] = rr;
    }
    __syncthreads();
    const uint32_t bsel = misc[2];
    lo_key = base + (bsel << shift);
    hi_key = shift ? lo_key + ((1u << shift) - 1u) : lo_key;
    if (shift == 0) break;
    base = lo_key;
    width_log = shift;
    if (pass == 0) {
      for (int i = tid; i < SELW_BINS / 4; i += SELQ_THREADS) ((uint4 *)hist)[i] = make_uint4(0, 0, 0, 0);
      __syncthreads();
    }
  }
  const float eps = TC_ERR_REL * (qn[q] + cn_max);
  const uint32_t lim_hi = f2key(sel_key2f(hi_key) + 2.0f * eps);
  const uint32_t lim_lo = need_order ? 0u : f2key(sel_key2f(lo_key) - 2.0f * eps);
  // ---- classification (hits are rare: plain shared-memory counters)
#pragma unroll 4
  for (uint32_t i = tid; i < C4; i += SELQ_THREADS) {
    const uint4 k4 = ((const uint4 *)keys)[i];
    const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      if (kk[c] < lim_lo) {
        const uint32_t pos = atomicAdd(&misc[0], 1u);
        if (pos < SELW_CAP) skey[pos] = ((uint64_t)kk[c] << 32) | (4 * i + c);
      } else if (kk[c] <= lim_hi) {
        const uint32_t pos = atomicAdd(&misc[1], 1u);
        if (pos < SELW_CAP) cand[pos] = 4 * i + c;
      }
    }
  }
  __syncthreads();
  const uint32_t nsure = min(misc[0], nprobe);   // <= nprobe - 1 by construction
  const uint32_t ncand = misc[1];
  if (ncand > SELW_CAP || nsure > SELW_CAP) {   // does not fit the candidate lists: k_coarse_select finishes this query
    if (tid == 0) { flags[q] = 1u; if (work_out) work_out[q] = 0; }
    return;
  }
  if (tid == 0) { atomicAdd(&g_coarse_band[0], (unsigned long long)ncand); atomicAdd(&g_coarse_band[1], 1ull); }
  const uint32_t need = nprobe - nsure;
  uint32_t *oi = out_ids + (size_t)q * nprobe;
  uint32_t mywork = 0;
  // sure centroids, ascending approximate distance (rank by counting; keys are distinct)
  for (uint32_t e = tid; e < nsure; e += SELQ_THREADS) {
    const uint64_t kk = skey[e];
    uint32_t r = 0;
    for (uint32_t j = 0; j < nsure; j++) r += skey[j] < kk ? 1u : 0u;
    oi[r] = (uint32_t)kk;
    if (chunk_start) mywork += chunk_start[(uint32_t)kk + 1] - chunk_start[(uint32_t)kk];
  }
  if (!need_order && ncand <= need) {  // the whole band belongs to the answer
    for (uint32_t e = tid; e < ncand; e += SELQ_THREADS) {
      oi[nsure + e] = cand[e];
      if (chunk_start) mywork += chunk_start[cand[e] + 1] - chunk_start[cand[e]];
    }
  } else {
    // ---- exact sqrt-L2 (l2.rs:30-74) of the band: 4 lanes per pair, lane r owns lane-accumulators 4r..4r+3.
    // cp.async (no register destination) keeps SELW_MLP 16-byte centroid loads in flight per lane -- a plain load is
    // scheduled next to its use, which turns the loop into a chain of L2 round trips.  The ring re-uses the key space.
    const int n = (int)dim, chunks = n / 16;
    const int r4 = lane & 3, grp = lane >> 2, gl = lane & ~3;
    const float4 *q4 = (const float4 *)(Q + (size_t)q * dim) + r4;
    float4 *ring = (float4 *)keys + w * (SELW_MLP * 32);
    // small band (the usual case): ONE burst brings every band row and the query into shared memory (key + histogram
    // space, both dead by now), so the exact arithmetic runs from shared memory after a single L2 round trip
    const uint32_t dim4 = dim / 4;
    const bool burst = (size_t)ncand * dim <= SELW_MAXC && dim <= SELW_BINS;
    float4 *rowbuf = (float4 *)keys, *qbuf = (float4 *)hist;
    if (burst) {
      __syncthreads();  // every thread is done reading keys / hist
      for (uint32_t i = tid; i < ncand * dim4; i += SELQ_THREADS) {
        const uint32_t j = i / dim4, piece = i - j * dim4;
        selw_cp_async16(rowbuf + i, (const float4 *)(centroids + (size_t)cand[j] * dim) + piece);
      }
      for (uint32_t i = tid; i < dim4; i += SELQ_THREADS) selw_cp_async16(qbuf + i, (const float4 *)(Q + (size_t)q * dim) + i);
      selw_cp_async_wait();
      __syncthreads();
    }
    for (uint32_t b0 = w * 8; b0 < ncand; b0 += SELQ_WARPS * 8) {
      const uint32_t j = b0 + grp;
      const uint32_t jc = j < ncand ? j : ncand - 1;
      const uint32_t cidx = cand[jc];
      const float *crow = centroids + (size_t)cidx * dim;
      const float4 *c4 = (const float4 *)crow + r4;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
      if (burst) {
        const float4 *rb = rowbuf + (size_t)jc * dim4 + r4, *qb = qbuf + r4;
#pragma unroll 4
        for (int c = 0; c < chunks; c++) {
          const float4 x = qb[c * 4], y = rb[c * 4];
          float d;
          d = __fsub_rn(x.x, y.x); a0 = __fadd_rn(a0, __fmul_rn(d, d));
          d = __fsub_rn(x.y, y.y); a1 = __fadd_rn(a1, __fmul_rn(d, d));
          d = __fsub_rn(x.z, y.z); a2 = __fadd_rn(a2, __fmul_rn(d, d));
          d = __fsub_rn(x.w, y.w); a3 = __fadd_rn(a3, __fmul_rn(d, d));
        }
      } else {
        for (int c = 0; c < chunks; c += SELW_MLP) {
          const int nb = min(SELW_MLP, chunks - c);
          for (int u = 0; u < nb; u++) selw_cp_async16(&ring[u * 32 + lane], c4 + (c + u) * 4);
          selw_cp_async_wait();
          for (int u = 0; u < nb; u++) {
            const float4 x = __ldg(q4 + (c + u) * 4), y = ring[u * 32 + lane];
            float d;
            d = __fsub_rn(x.x, y.x); a0 = __fadd_rn(a0, __fmul_rn(d, d));
            d = __fsub_rn(x.y, y.y); a1 = __fadd_rn(a1, __fmul_rn(d, d));
            d = __fsub_rn(x.z, y.z); a2 = __fadd_rn(a2, __fmul_rn(d, d));
            d = __fsub_rn(x.w, y.w); a3 = __fadd_rn(a3, __fmul_rn(d, d));
          }
        }
      }
      float s2 = -0.0f;  // ordered lane reduction 0..15
#pragma unroll
      for (int l = 0; l < 4; l++) {
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a0, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a1, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a2, gl + l));
        s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, a3, gl + l));
      }
      float ret = __fadd_rn(0.0f, s2);
      const int p = chunks * 16;
      if (p < n) ret = ref_tail<MGPU_L2>(PtrAcc{Q + (size_t)q * dim}, PtrAcc{crow}, p, n, ret);
      const float dist = sqrtf(ret);
      if (j < ncand && r4 == 0) ckey[j] = ((uint64_t)f2key(dist) << 32) | cidx;
    }
    __syncthreads();
    float *od = out_dist ? out_dist + (size_t)q * nprobe + nsure : nullptr;
    for (uint32_t e = tid; e < ncand; e += SELQ_THREADS) {
      const uint64_t kk = ckey[e];
      uint32_t r = 0;
      for (uint32_t j = 0; j < ncand; j++) r += ckey[j] < kk ? 1u : 0u;
      if (r < need) {
        oi[nsure + r] = (uint32_t)kk;
        if (od) od[r] = sel_key2f((uint32_t)(kk >> 32));
        if (chunk_start) mywork += chunk_start[(uint32_t)kk + 1] - chunk_start[(uint32_t)kk];
      }
    }
  }
  if (work_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mywork += __shfl_xor_sync(0xffffffffu, mywork, o);
    if (lane == 0 && mywork) atomicAdd(&misc[10], mywork);
    __syncthreads();
    if (tid == 0) work_out[q] = misc[10];
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static int make_map(mgpu_ctx *ctx, CUtensorMap *map, const void *base, uint64_t rows, uint32_t Kp, uint32_t box_rows) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return mgpu_fail(ctx, MGPU_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {Kp, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mgpu_fail(ctx, MGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MGPU_OK;
}

uint32_t coarse_tc_kp(uint32_t dim) { return ((3 * dim + TC_BK - 1) / TC_BK) * TC_BK; }

bool coarse_tc_applicable(mgpu_ctx *ctx, uint32_t dim, uint32_t C, uint32_t nprobe) {
  static const int mode = getenv("MGPU_COARSE") ? (getenv("MGPU_COARSE")[0] == 'e' ? 0 : 2) : 1;  // exact | auto | tc
  if (mode == 0) return false;
  if (C % 4 != 0 || dim < 16) return false;
  // the candidate buffers are sized for ALL centroids, so the margin rule can never overflow them
  size_t cap = 1;
  while (cap < (size_t)C) cap <<= 1;
  size_t smem = ((dim + 3) & ~3u) * 4 + cap * 12 + SEL_BINS * 4 + 64;
  if (smem > ctx->smem_optin || nprobe > C) return false;
  if (mode == 2) return true;
  return C >= 256 && dim >= 64;
}

int launch_split_bf16(mgpu_ctx *ctx, const float *dX, uint64_t n, uint32_t dim, int is_centroid, void *d_out, float *d_norms) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_COARSE);
  k_split_bf16<<<(unsigned)n, 128, 0, ctx->stream>>>(dX, n, dim, coarse_tc_kp(dim), is_centroid, (__nv_bfloat16 *)d_out, d_norms);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// Tensor-core squared-L2 estimates D~ (B x C) of B rows against the pre-split centroids: split the rows, then the K = 3*dim
// bf16 GEMM with the norm terms fused in the epilogue.  d_xsplit: B x Kp bf16; d_xn: B floats (|x|^2).
int launch_tc_distances(mgpu_ctx *ctx, const float *dX, uint32_t B, const void *d_csplit, const float *d_cn, uint32_t C, uint32_t dim,
                        void *d_xsplit, float *d_xn, float *d_Dt) {
  const uint32_t Kp = coarse_tc_kp(dim);
  MGPU_TRY(launch_split_bf16(ctx, dX, B, dim, 0, d_xsplit, d_xn));
  CUtensorMap mq, mc;
  MGPU_TRY(make_map(ctx, &mq, d_xsplit, B, Kp, TC_BM));
  MGPU_TRY(make_map(ctx, &mc, d_csplit, C, Kp, TC_BN));
  size_t smem = sizeof(TcSmem) + 1024;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_coarse_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((C + TC_BN - 1) / TC_BN, (B + TC_BM - 1) / TC_BM);
  LaunchScope ls(ctx, MGPU_K_COARSE);
  k_coarse_gemm<<<grid, TC_THREADS, smem, ctx->stream>>>(mq, mc, d_xn, d_cn, B, C, Kp, d_Dt);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// d_Dt: B x C floats of workspace; d_qsplit: B x Kp bf16; d_qn: B floats
int launch_coarse_tc(mgpu_ctx *ctx, const float *dQ, uint32_t B, const float *d_centroids, const void *d_csplit, const float *d_cn,
                     float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, void *d_qsplit, float *d_qn, float *d_Dt,
                     uint32_t *d_overflow, uint32_t *d_flags, int need_order, uint32_t *out_ids, float *out_dist,
                     const uint32_t *chunk_start, uint32_t *d_work, bool *work_done, cudaEvent_t after_gemm) {
  // MGPU_FORK=0: side-stream work (the query encode) starts before the GEMM; 1 (default): after it
  static const int fork_at = getenv("MGPU_FORK") ? atoi(getenv("MGPU_FORK")) : 1;
  if (after_gemm && fork_at == 0) CUDA_TRY(ctx, cudaEventRecord(after_gemm, ctx->stream));
  MGPU_TRY(launch_tc_distances(ctx, dQ, B, d_csplit, d_cn, C, dim, d_qsplit, d_qn, d_Dt));
  if (after_gemm && fork_at != 0) CUDA_TRY(ctx, cudaEventRecord(after_gemm, ctx->stream));  // fork point for work that may overlap the selection
  uint32_t cap = 1;
  while (cap < C) cap <<= 1;
  size_t ssel = ((dim + 3) & ~3u) * 4 + (size_t)cap * 12 + SEL_BINS * 4 + 64;
  static const bool force_order = getenv("MGPU_COARSE_ORDER") && getenv("MGPU_COARSE_ORDER")[0] == '1';
  if (force_order || out_dist) need_order = 1;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_coarse_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssel));
  static const bool no_warp = getenv("MGPU_SELECT_FAST") && getenv("MGPU_SELECT_FAST")[0] == '0';
  const uint32_t *only_flagged = nullptr;
  if (work_done) *work_done = false;
  if (!no_warp && (dim & 3u) == 0 && nprobe <= SELW_CAP && C <= SELW_MAXC && d_flags) {
    CUDA_TRY(ctx, cudaMemsetAsync(d_flags, 0, (size_t)B * 4, ctx->stream));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_coarse_select_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_BYTES));
    LaunchScope ls(ctx, MGPU_K_SELECT);
    k_coarse_select_q<<<B, SELQ_THREADS, SELW_BYTES, ctx->stream>>>(d_Dt, dQ, d_centroids, d_qn, cn_max, C, dim, nprobe, need_order,
                                                                     out_ids, out_dist, d_flags, d_work ? chunk_start : nullptr, d_work);
    CUDA_TRY(ctx, cudaGetLastError());
    only_flagged = d_flags;
    if (work_done && d_work) *work_done = true;
  }
  {
    LaunchScope ls(ctx, MGPU_K_SELECT);
    k_coarse_select<<<B, SEL_THREADS, ssel, ctx->stream>>>(d_Dt, dQ, d_centroids, d_qn, cn_max, C, dim, nprobe, cap, need_order, out_ids,
                                                            out_dist, d_overflow, only_flagged);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return MGPU_OK;
}

int coarse_band_stats(mgpu_ctx *ctx, uint64_t out[2], int reset) {
  unsigned long long h[2] = {0, 0};
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyFromSymbol(h, g_coarse_band, sizeof(h)));
  out[0] = h[0]; out[1] = h[1];
  if (reset) { h[0] = h[1] = 0; CUDA_TRY(ctx, cudaMemcpyToSymbol(g_coarse_band, h, sizeof(h))); }
  return MGPU_OK;
}

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

// ---- split ---------------------------------------------------------------------------------------------------------------
__global__ void k_split_bf16(const float *__restrict__ X, uint64_t n, uint32_t dim, uint32_t Kp, int is_centroid,
                             __nv_bfloat16 *__restrict__ out, float *__restrict__ norms) {
  uint64_t row = blockIdx.x;
  const float *x = X + row * dim;
  __nv_bfloat16 *o = out + row * Kp;
  float part = 0.0f;
  for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
    float v = x[d];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    o[d] = h;
    o[dim + d] = is_centroid ? l : h;
    o[2 * dim + d] = is_centroid ? h : l;
    part += v * v;
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
// One CTA per query.  smem: keys[C] (ordered keys of D~), hist[256], cand[<= C] indices, exact key64 list.
#define SEL_THREADS 512
__global__ void __launch_bounds__(SEL_THREADS)
k_coarse_select(const float *__restrict__ Dt, const float *__restrict__ Q, const float *__restrict__ centroids,
                const float *__restrict__ qn, float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, uint32_t cand_cap,
                uint32_t *__restrict__ out_ids, float *__restrict__ out_dist, uint32_t *__restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint32_t *keys = (uint32_t *)sm;                 // C
  float *sq = (float *)(keys + C);                 // dim (16-byte aligned: C % 4 == 0 is required by the launcher)
  uint64_t *ckey = (uint64_t *)(sq + ((dim + 3) & ~3u));  // cand_cap (power of two)
  uint32_t *cand = (uint32_t *)(ckey + cand_cap);  // cand_cap
  uint32_t *hist = cand + cand_cap;                // 256
  uint32_t *misc = hist + 256;                     // [0] prefix, [1] remaining rank, [2] candidate count
  const uint32_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *row = Dt + (size_t)q * C;
  for (uint32_t i = tid; i < C; i += SEL_THREADS) keys[i] = f2key(row[i]);
  for (uint32_t d = tid; d < dim; d += SEL_THREADS) sq[d] = Q[(size_t)q * dim + d];
  if (tid == 0) { misc[0] = 0; misc[1] = nprobe - 1; misc[2] = 0; }
  __syncthreads();
  // radix select (MSB first, 8 bits per pass) of the key with rank nprobe-1.  Keys cluster in a few bins, so increments are
  // aggregated per warp with match_any (one shared-memory atomic per distinct bin per warp instruction) and the bin scan is
  // a warp-parallel prefix sum.
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = misc[0];
    const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (uint32_t i0 = 0; i0 < C; i0 += SEL_THREADS) {
      const uint32_t i = i0 + tid;
      uint32_t bin = 0xFFFFu;  // "not a candidate of this pass"
      if (i < C) {
        uint32_t k = keys[i];
        if ((k & mask) == prefix) bin = (k >> shift) & 255u;
      }
      const unsigned grp = __match_any_sync(0xffffffffu, bin);
      if (bin != 0xFFFFu && lane == __ffs(grp) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(grp));
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) { loc[j] = hist[lane * 8 + j]; sum += loc[j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t rank = misc[1];
      const uint32_t excl = incl - sum;
      const bool mine = rank >= excl && rank < incl;  // exactly one lane
      if (mine) {
        uint32_t r = rank - excl, b = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { if (r >= loc[j] && b == (uint32_t)j) { r -= loc[j]; b = j + 1; } }
        misc[0] = prefix | ((uint32_t)(lane * 8 + b) << shift);
        misc[1] = r;
      }
    }
    __syncthreads();
  }
  const uint32_t tau_key = misc[0];
  // candidates: D~ <= tau + 2 eps   (eps bounds |D~ - reference fp32 squared distance|)
  const uint32_t tk = (tau_key & 0x80000000u) ? (tau_key ^ 0x80000000u) : ~tau_key;
  const float tau = __uint_as_float(tk);
  const float eps = TC_ERR_REL * (qn[q] + cn_max);
  const uint32_t lim_key = f2key(tau + 2.0f * eps);
  for (uint32_t i = tid; i < C; i += SEL_THREADS) {
    if (keys[i] <= lim_key) {
      uint32_t pos = atomicAdd(&misc[2], 1u);
      if (pos < cand_cap) cand[pos] = i;
    }
  }
  __syncthreads();
  uint32_t ncand = misc[2];
  if (ncand > cand_cap) { if (tid == 0) atomicAdd(overflow, 1u); ncand = cand_cap; }
  // exact sqrt-L2 (l2.rs:30-74) for every candidate: half-warp per pair, lane h owns lane-accumulator h
  const int h = lane & 15, half = lane >> 4;
  const uint32_t npairs_round = (SEL_THREADS / 32) * 2;
  for (uint32_t base = 0; base < ncand; base += npairs_round) {
    uint32_t j = base + warp * 2 + half;
    uint32_t jj = j < ncand ? j : ncand - 1;
    const uint32_t cidx = cand[jj];
    const float *crow = centroids + (size_t)cidx * dim;
    const int n = (int)dim;
    float ret = 0.0f;
    int p = 0;
    if (n / 16 > 0) {
      const int chunks = n / 16;
      float acc = 0.0f;
#pragma unroll 16
      for (int c = 0; c < chunks; c++) {  // 16 independent 64-byte row segments in flight per half-warp
        float d = __fsub_rn(sq[c * 16 + h], __ldg(crow + c * 16 + h));
        acc = __fadd_rn(acc, __fmul_rn(d, d));
      }
      float s2 = -0.0f;
      const int basel = lane & 16;
#pragma unroll
      for (int l = 0; l < 16; l++) s2 = __fadd_rn(s2, __shfl_sync(0xffffffffu, acc, basel + l));
      ret = __fadd_rn(ret, s2);
      p = chunks * 16;
    }
    if (p < n) ret = ref_tail<MGPU_L2>(PtrAcc{sq}, PtrAcc{crow}, p, n, ret);
    const float dist = sqrtf(ret);
    if (j < ncand && h == 0) ckey[j] = ((uint64_t)f2key(dist) << 32) | cidx;
  }
  for (uint32_t i = ncand + tid; i < cand_cap; i += SEL_THREADS) ckey[i] = MGPU_EMPTY_KEY;
  __syncthreads();
  // order by (distance total order, centroid index): bitonic sort over the smallest power of two >= ncand
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
  for (uint32_t i = tid; i < nprobe; i += SEL_THREADS) {
    uint64_t kk = ckey[i];
    out_ids[(size_t)q * nprobe + i] = (uint32_t)kk;
    if (out_dist) {
      uint32_t kd = (uint32_t)(kk >> 32);
      out_dist[(size_t)q * nprobe + i] = __uint_as_float((kd & 0x80000000u) ? (kd ^ 0x80000000u) : ~kd);
    }
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
  size_t smem = (size_t)C * 4 + ((dim + 3) & ~3u) * 4 + cap * 12 + 1024 + 64;
  if (smem > ctx->smem_optin || nprobe > C) return false;
  if (mode == 2) return true;
  return C >= 1024 && dim >= 64;
}

int launch_split_bf16(mgpu_ctx *ctx, const float *dX, uint64_t n, uint32_t dim, int is_centroid, void *d_out, float *d_norms) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_COARSE);
  k_split_bf16<<<(unsigned)n, 128, 0, ctx->stream>>>(dX, n, dim, coarse_tc_kp(dim), is_centroid, (__nv_bfloat16 *)d_out, d_norms);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// d_Dt: B x C floats of workspace; d_qsplit: B x Kp bf16; d_qn: B floats
int launch_coarse_tc(mgpu_ctx *ctx, const float *dQ, uint32_t B, const float *d_centroids, const void *d_csplit, const float *d_cn,
                     float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, void *d_qsplit, float *d_qn, float *d_Dt,
                     uint32_t *d_overflow, uint32_t *out_ids, float *out_dist) {
  const uint32_t Kp = coarse_tc_kp(dim);
  MGPU_TRY(launch_split_bf16(ctx, dQ, B, dim, 0, d_qsplit, d_qn));
  CUtensorMap mq, mc;
  MGPU_TRY(make_map(ctx, &mq, d_qsplit, B, Kp, TC_BM));
  MGPU_TRY(make_map(ctx, &mc, d_csplit, C, Kp, TC_BN));
  size_t smem = sizeof(TcSmem) + 1024;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_coarse_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    dim3 grid((C + TC_BN - 1) / TC_BN, (B + TC_BM - 1) / TC_BM);
    LaunchScope ls(ctx, MGPU_K_COARSE);
    k_coarse_gemm<<<grid, TC_THREADS, smem, ctx->stream>>>(mq, mc, d_qn, d_cn, B, C, Kp, d_Dt);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  uint32_t cap = 1;
  while (cap < C) cap <<= 1;
  size_t ssel = (size_t)C * 4 + ((dim + 3) & ~3u) * 4 + (size_t)cap * 12 + 1024 + 64;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_coarse_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssel));
  {
    LaunchScope ls(ctx, MGPU_K_SELECT);
    k_coarse_select<<<B, SEL_THREADS, ssel, ctx->stream>>>(d_Dt, dQ, d_centroids, d_qn, cn_max, C, dim, nprobe, cap, out_ids, out_dist,
                                                            d_overflow);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return MGPU_OK;
}

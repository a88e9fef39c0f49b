// coarse.cu -- all-pairs distance matrix + per-row k-smallest selection.
//   BlockBasedIvf::find_nearest_centroids  rs/index/src/ivf/block_based/index.rs:147-163  (sqrt L2, query x centroid)
//   IvfBuilder::find_nearest_centroids     rs/index/src/ivf/builder.rs:268-282            (squared L2, vector x centroid)
//   DistanceCalculator::calculate          rs/utils/src/lib.rs:17-36                       (generic batch form)
//
// Exact path: every (row, column) pair is accumulated in the reference's 16-lane order by ONE thread (16 lane
// accumulators per pair, 2x2 pairs per thread), so distances are bit-identical to the CPU and the selection reproduces
// the reference's ordering including exact ties.  Tiles are staged through shared memory: 32x32 pairs per CTA, 64 dims
// per stage.
#include "internal.cuh"

#define DM_KC 64
#define DM_PITCH (DM_KC + 4)

template <int METRIC>
__global__ void __launch_bounds__(256) k_distance_tile(const float *__restrict__ A, uint64_t nA,
                                                        const float *__restrict__ Bm, uint64_t nB, uint32_t dim,
                                                        int do_sqrt, float *__restrict__ out) {
  __shared__ __align__(16) float As[32 * DM_PITCH];
  __shared__ __align__(16) float Bs[32 * DM_PITCH];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const uint64_t a0 = (uint64_t)blockIdx.y * 32, b0 = (uint64_t)blockIdx.x * 32;
  float acc[2][2][16];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int l = 0; l < 16; l++) acc[i][j][l] = 0.0f;

  for (uint32_t k0 = 0; k0 < dim; k0 += DM_KC) {
    const uint32_t kc = min((uint32_t)DM_KC, dim - k0);  // multiple of 16
    // stage: 32 rows x kc floats per matrix, float4 granularity
    for (uint32_t i = threadIdx.x; i < 32 * (DM_KC / 4); i += 256) {
      uint32_t r = i / (DM_KC / 4), c4 = i % (DM_KC / 4);
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (c4 * 4 < kc) {
        if (a0 + r < nA) va = *(const float4 *)(A + (a0 + r) * dim + k0 + c4 * 4);
        if (b0 + r < nB) vb = *(const float4 *)(Bm + (b0 + r) * dim + k0 + c4 * 4);
      }
      *(float4 *)(As + r * DM_PITCH + c4 * 4) = va;
      *(float4 *)(Bs + r * DM_PITCH + c4 * 4) = vb;
    }
    __syncthreads();
    const uint32_t chunks = kc / 16;
    for (uint32_t c = 0; c < chunks; c++) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float4 av[2], bv[2];
        av[0] = *(const float4 *)(As + ty * DM_PITCH + c * 16 + j * 4);
        av[1] = *(const float4 *)(As + (ty + 16) * DM_PITCH + c * 16 + j * 4);
        bv[0] = *(const float4 *)(Bs + tx * DM_PITCH + c * 16 + j * 4);
        bv[1] = *(const float4 *)(Bs + (tx + 16) * DM_PITCH + c * 16 + j * 4);
#pragma unroll
        for (int ia = 0; ia < 2; ia++)
#pragma unroll
          for (int ib = 0; ib < 2; ib++) {
            const float x[4] = {av[ia].x, av[ia].y, av[ia].z, av[ia].w};
            const float y[4] = {bv[ib].x, bv[ib].y, bv[ib].z, bv[ib].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              float &s = acc[ia][ib][j * 4 + e];
              if (METRIC == MGPU_L2) { float d = __fsub_rn(x[e], y[e]); s = __fadd_rn(s, __fmul_rn(d, d)); }
              else s = __fadd_rn(s, __fmul_rn(x[e], y[e]));
            }
          }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ia = 0; ia < 2; ia++)
#pragma unroll
    for (int ib = 0; ib < 2; ib++) {
      uint64_t r = a0 + ty + ia * 16, c = b0 + tx + ib * 16;
      if (r < nA && c < nB) {
        float v = __fadd_rn(0.0f, ordered_reduce(acc[ia][ib], 16));
        if (METRIC == MGPU_L2) { if (do_sqrt) v = sqrtf(v); }
        else v = -v;
        out[r * nB + c] = v;
      }
    }
}

// any dim: one thread per pair, operands straight from global/L2 (small problems and odd dims only)
template <int METRIC>
__global__ void k_distance_generic(const float *__restrict__ A, uint64_t nA, const float *__restrict__ Bm, uint64_t nB,
                                   uint32_t dim, int do_sqrt, float *__restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA * nB) return;
  uint64_t r = i / nB, c = i % nB;
  float v = ref_distance<METRIC>(PtrAcc{A + r * dim}, PtrAcc{Bm + c * dim}, (int)dim);
  if (METRIC == MGPU_L2 && do_sqrt) v = sqrtf(v);
  out[i] = v;
}

int launch_distance_matrix(mgpu_ctx *ctx, const float *dA, uint64_t nA, const float *dB, uint64_t nB, uint32_t dim,
                           int metric, int mode, float *dout, int kernel_class) {
  if (nA == 0 || nB == 0) return MGPU_OK;
  bool tiled = dim % 16 == 0 && dim >= 16 && !(metric == MGPU_DOT && dim == 16) &&
               ((uintptr_t)dA % 16 == 0) && ((uintptr_t)dB % 16 == 0);
  LaunchScope ls(ctx, kernel_class);
  if (tiled) {
    dim3 grid((unsigned)((nB + 31) / 32), (unsigned)((nA + 31) / 32));
    if (grid.y > 65535) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "distance matrix: too many rows per call (%llu)", (unsigned long long)nA);
    if (metric == MGPU_L2) k_distance_tile<MGPU_L2><<<grid, 256, 0, ctx->stream>>>(dA, nA, dB, nB, dim, mode, dout);
    else k_distance_tile<MGPU_DOT><<<grid, 256, 0, ctx->stream>>>(dA, nA, dB, nB, dim, mode, dout);
  } else {
    uint64_t total = nA * nB;
    unsigned grid = (unsigned)((total + 127) / 128);
    if (metric == MGPU_L2) k_distance_generic<MGPU_L2><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, mode, dout);
    else k_distance_generic<MGPU_DOT><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, mode, dout);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- per-row selection of the nsel smallest by (value total order, column index) -----------------------------------------
// select_nth_unstable_by + sort_by(total_cmp) (index.rs:158-161); the reference leaves the order of exactly equal
// distances unspecified -- we break ties by column index.  Shared-memory bitonic sort.
__global__ void __launch_bounds__(512) k_select_smallest(const float *__restrict__ D, uint32_t C, uint32_t P2,
                                                          uint32_t nsel, uint32_t *__restrict__ out_ids,
                                                          float *__restrict__ out_vals) {
  extern __shared__ __align__(16) uint64_t keys[];
  const uint32_t b = blockIdx.x;
  const float *row = D + (size_t)b * C;
  for (uint32_t i = threadIdx.x; i < P2; i += blockDim.x)
    keys[i] = i < C ? (((uint64_t)f2key(row[i]) << 32) | i) : MGPU_EMPTY_KEY;
  __syncthreads();
  for (uint32_t k = 2; k <= P2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < P2; i += blockDim.x) {
        uint32_t ixj = i ^ j;
        if (ixj > i) {
          uint64_t x = keys[i], y = keys[ixj];
          bool up = (i & k) == 0;
          if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < nsel; i += blockDim.x) {
    uint64_t kk = keys[i];
    uint32_t idx = (uint32_t)kk;
    out_ids[(size_t)b * nsel + i] = idx;
    if (out_vals) out_vals[(size_t)b * nsel + i] = row[idx];
  }
}

int launch_select_smallest(mgpu_ctx *ctx, const float *dD, uint32_t B, uint32_t C, uint32_t nsel, uint32_t *out_ids,
                           float *out_vals) {
  if (B == 0) return MGPU_OK;
  uint32_t P2 = 1;
  while (P2 < C) P2 <<= 1;
  size_t smem = (size_t)P2 * 8;
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "selection over %u columns exceeds shared memory", C);
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_select_smallest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(ctx, MGPU_K_SELECT);
  int threads = P2 / 2 >= 512 ? 512 : (P2 / 2 >= 32 ? (int)(P2 / 2) : 32);
  k_select_smallest<<<B, threads, smem, ctx->stream>>>(dD, C, P2, nsel, out_ids, out_vals);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// acceptance rule of IvfBuilder::build_posting_lists (ivf/builder.rs:312-327)
__global__ void k_assign_filter(const uint32_t *__restrict__ sel_ids, const float *__restrict__ sel_vals, uint64_t n,
                                uint32_t r, float threshold, uint32_t *__restrict__ out_cids,
                                uint32_t *__restrict__ out_counts) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dmin = sel_vals[i * r];
  uint32_t cnt = 0;
  for (uint32_t j = 0; j < r; j++) out_cids[i * r + j] = 0xFFFFFFFFu;
  for (uint32_t j = 0; j < r; j++) {
    float d = sel_vals[i * r + j];
    if (fabsf(__fsub_rn(d, dmin)) <= __fmul_rn(dmin, threshold)) out_cids[i * r + cnt++] = sel_ids[i * r + j];
  }
  out_counts[i] = cnt;
}

int launch_assign_filter(mgpu_ctx *ctx, const uint32_t *sel_ids, const float *sel_vals, uint64_t n, uint32_t r,
                         float threshold, uint32_t *out_cids, uint32_t *out_counts) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_OTHER);
  k_assign_filter<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(sel_ids, sel_vals, n, r, threshold, out_cids, out_counts);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

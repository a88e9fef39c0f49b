// pq.cu -- ProductQuantizer kernels: code-pair table, batch quantize, batch SDC distance.
// Reference: rs/quantization/src/pq/mod.rs:152-177 (quantize), :202-278 (distance).
#include "internal.cuh"
#include "pq_device.cuh"

// table[s][a][b] = score contribution of subspace s when the query code is a and the row code is b.
//   L2 : ||cb[s][a] - cb[s][b]||^2,  dot: -<cb[s][a], cb[s][b]>
// restricted to what ProductQuantizer::distance(StreamingSIMD) actually accumulates (pq/mod.rs:231-266): the 16/8/4-lane
// phases of every subspace, but the scalar tail (dsub % 4 dims) of the LAST subspace only -- `sum_1 = ...` is an
// assignment, so earlier tails are overwritten (pq/mod.rs:259-261).
// Used by the scan to build the per-query LUT as a 1 KB row gather instead of re-reading the whole codebook; the scan
// only ranks with it (fixed point), exact scores come from pq_distance_streaming.
// One block per (s, a); also reduces rowmin/rowmax for the fixed-point LUT scaling.
template <int METRIC>
__global__ void k_pq_build_table(const float *__restrict__ cb, uint32_t dsub, uint32_t K, float *__restrict__ table,
                                 float *__restrict__ rowmin, float *__restrict__ rowmax) {
  uint32_t s = blockIdx.y, a = blockIdx.x;
  const float *ca = cb + ((size_t)s * K + a) * dsub;
  float mn = 3.4e38f, mx = -3.4e38f;
  for (uint32_t b = threadIdx.x; b < K; b += blockDim.x) {
    const float *cbv = cb + ((size_t)s * K + b) * dsub;
    const uint32_t lanes_end = dsub - (dsub % 4);
    const uint32_t end = (s + 1 == gridDim.y) ? dsub : lanes_end;
    float v = 0.0f;
    for (uint32_t d = 0; d < end; d++) {
      float x = ca[d], y = cbv[d];
      if (METRIC == MGPU_L2) { float df = __fsub_rn(x, y); v = __fadd_rn(v, __fmul_rn(df, df)); }
      else v = __fsub_rn(v, __fmul_rn(x, y));
    }
    table[((size_t)s * K + a) * K + b] = v;
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ float smn[32], smx[32];
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane_id() == 0) { smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; i++) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    rowmin[(size_t)s * K + a] = mn; rowmax[(size_t)s * K + a] = mx;
  }
}

// maximum row range of the table (one block) and the 16-bit image
__global__ void k_pq_max_range(const float *__restrict__ rowmin, const float *__restrict__ rowmax, uint32_t n, float *__restrict__ out) {
  __shared__ float sm[32];
  float mx = 0.0f;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, rowmax[i] - rowmin[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t w = 1; w < (blockDim.x + 31) / 32; w++) mx = fmaxf(mx, sm[w]);
    out[0] = mx;
  }
}
__global__ void k_pq_table16(const float *__restrict__ table, const float *__restrict__ rowmin, uint64_t n, uint32_t K, float gscale,
                             uint16_t *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __fmul_rn(__fsub_rn(table[i], rowmin[i / K]), gscale);
  out[i] = (uint16_t)min(__float2uint_rn(v), 65535u);
}

// d_table16 / gscale (needs the table, rowmin, rowmax); synchronises once to read the range back
int launch_pq_build_table16(mgpu_pq *pq) {
  mgpu_ctx *ctx = pq->ctx;
  if (pq->K != 256) return MGPU_OK;
  const uint64_t n = (uint64_t)pq->m * pq->K * pq->K;
  float *d_mx = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&d_mx, 16));
  {
    LaunchScope ls(ctx, MGPU_K_OTHER);
    k_pq_max_range<<<1, 1024, 0, ctx->stream>>>(pq->d_rowmin, pq->d_rowmax, pq->m * pq->K, d_mx);
  }
  float mx = 0.0f;
  cudaError_t e = cudaMemcpyAsync(&mx, d_mx, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_mx);
  CUDA_TRY(ctx, e);
  if (!(mx > 0.0f) || !(mx < 3.0e38f)) { pq->gscale = 0.0f; return MGPU_OK; }   // degenerate codebook: the 32-bit scan handles it
  pq->gscale = 65535.0f / mx;
  CUDA_TRY(ctx, cudaMalloc((void **)&pq->d_table16, n * 2));
  LaunchScope ls(ctx, MGPU_K_OTHER);
  k_pq_table16<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(pq->d_table, pq->d_rowmin, n, pq->K, pq->gscale, pq->d_table16);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

int launch_pq_build_table(mgpu_pq *pq) {
  mgpu_ctx *ctx = pq->ctx;
  dim3 grid(pq->K, pq->m);
  int threads = pq->K >= 256 ? 256 : (pq->K >= 32 ? (int)pq->K : 32);
  LaunchScope ls(ctx, MGPU_K_OTHER);
  if (pq->metric == MGPU_L2)
    k_pq_build_table<MGPU_L2><<<grid, threads, 0, ctx->stream>>>(pq->d_cb, pq->dsub, pq->K, pq->d_table, pq->d_rowmin, pq->d_rowmax);
  else
    k_pq_build_table<MGPU_DOT><<<grid, threads, 0, ctx->stream>>>(pq->d_cb, pq->dsub, pq->K, pq->d_table, pq->d_rowmin, pq->d_rowmax);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// Quantizer::quantize for a batch (pq/mod.rs:152-177): grid = (row tiles, subspaces).  The subspace's centroids sit in
// shared memory; QT_SPLIT threads share one input row, thread j scoring centroids j, j + QT_SPLIT, ... with the bit-exact
// calculate_squared and keeping its first minimum; the QT_SPLIT partial minima are combined by (distance, index), which
// is the reference's strict-`<` first-minimum rule.  ALWAYS the L2 calculator, whatever D is (pq/mod.rs:168).
#define QT_THREADS 256
#define QT_SPLIT 4
#define QT_ROWS (QT_THREADS / QT_SPLIT)
__global__ void __launch_bounds__(QT_THREADS) k_pq_quantize(const float *__restrict__ X, uint64_t n, uint32_t dim,
                                                             const float *__restrict__ cb, uint32_t dsub, uint32_t K,
                                                             uint32_t m, uint8_t *__restrict__ codes) {
  extern __shared__ __align__(16) float sm[];
  float *scb = sm;                       // K * dsub
  float *sx = sm + (size_t)K * dsub;     // QT_ROWS * (dsub + 1)
  uint32_t s = blockIdx.y;
  uint64_t row0 = (uint64_t)blockIdx.x * QT_ROWS;
  const float *gcb = cb + (size_t)s * K * dsub;
  if (((K * dsub) & 3u) == 0) {
    for (uint32_t i = threadIdx.x; i < K * dsub / 4; i += blockDim.x) ((float4 *)scb)[i] = __ldg((const float4 *)gcb + i);
  } else {
    for (uint32_t i = threadIdx.x; i < K * dsub; i += blockDim.x) scb[i] = gcb[i];
  }
  uint32_t pitch = dsub + 1;
  for (uint32_t i = threadIdx.x; i < QT_ROWS * dsub; i += blockDim.x) {
    uint32_t r = i / dsub, d = i % dsub;
    uint64_t row = row0 + r;
    sx[r * pitch + d] = row < n ? X[row * dim + (size_t)s * dsub + d] : 0.0f;
  }
  __syncthreads();
  const uint32_t r = threadIdx.x / QT_SPLIT, j = threadIdx.x % QT_SPLIT;
  uint64_t row = row0 + r;
  const float *x = sx + r * pitch;
  uint32_t best = 0;
  float best_d = 3.40282347e+38f;
  if (dsub == 8) {
    // common case: the row's subvector lives in registers, every centroid is two 16-byte shared loads.  calculate_squared
    // on 8 values is one 8-lane chunk: lane accumulators 0 + d*d (== d*d: a square is never -0), ordered reduce starting
    // from -0.0 (-0.0 + p == p for p >= +0), then 0 + s (== s).  The identities hold for NaN as well, so only the 8
    // subtractions, 8 products and 7 ordered additions are executed.
    float xv[8];
#pragma unroll
    for (int l = 0; l < 8; l++) xv[l] = x[l];
#pragma unroll 4
    for (uint32_t c = j; c < K; c += QT_SPLIT) {
      const float4 c0 = *(const float4 *)(scb + c * 8), c1 = *(const float4 *)(scb + c * 8 + 4);
      const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      float d0 = __fsub_rn(xv[0], cv[0]);
      float s2 = __fmul_rn(d0, d0);
#pragma unroll
      for (int l = 1; l < 8; l++) { float d = __fsub_rn(xv[l], cv[l]); s2 = __fadd_rn(s2, __fmul_rn(d, d)); }
      if (s2 < best_d) { best_d = s2; best = c; }
    }
  } else {
    for (uint32_t c = j; c < K; c += QT_SPLIT) {
      float d = ref_distance<MGPU_L2>(PtrAcc{x}, PtrAcc{scb + (size_t)c * dsub}, (int)dsub);
      if (d < best_d) { best_d = d; best = c; }
    }
  }
  // combine the QT_SPLIT partial first-minima: smaller distance, then smaller index
#pragma unroll
  for (int o = 1; o < QT_SPLIT; o <<= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best_d, o);
    const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (od < best_d || (od == best_d && ob < best)) { best_d = od; best = ob; }
  }
  if (j == 0 && row < n) codes[row * m + s] = (uint8_t)best;
}

int launch_pq_quantize(mgpu_pq *pq, const float *dX, uint64_t n, uint8_t *dcodes, cudaStream_t st) {
  mgpu_ctx *ctx = pq->ctx;
  if (!st) st = ctx->stream;
  if (n == 0) return MGPU_OK;
  size_t smem = ((size_t)pq->K * pq->dsub + (size_t)QT_ROWS * (pq->dsub + 1)) * sizeof(float);
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "pq quantize: codebook slice of %zu bytes exceeds shared memory", smem);
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_pq_quantize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n + QT_ROWS - 1) / QT_ROWS), pq->m);
  LaunchScope ls(ctx, MGPU_K_QUANTIZE, st);
  k_pq_quantize<<<grid, QT_THREADS, smem, st>>>(dX, n, pq->dim, pq->d_cb, pq->dsub, pq->K, pq->m, dcodes);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

template <int METRIC>
__global__ void k_pq_distance_pairs(const float *__restrict__ cb, uint32_t m, uint32_t K, uint32_t dsub,
                                    const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, uint64_t n,
                                    float *__restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = pq_distance_streaming<METRIC>(cb, m, K, dsub, RowMajorCode{a + i * m}, RowMajorCode{b + i * m});
}

int launch_pq_distance_pairs(mgpu_pq *pq, const uint8_t *da, const uint8_t *db, uint64_t n, float *dout) {
  mgpu_ctx *ctx = pq->ctx;
  if (n == 0) return MGPU_OK;
  unsigned grid = (unsigned)((n + 127) / 128);
  LaunchScope ls(ctx, MGPU_K_OTHER);
  if (pq->metric == MGPU_L2) k_pq_distance_pairs<MGPU_L2><<<grid, 128, 0, ctx->stream>>>(pq->d_cb, pq->m, pq->K, pq->dsub, da, db, n, dout);
  else k_pq_distance_pairs<MGPU_DOT><<<grid, 128, 0, ctx->stream>>>(pq->d_cb, pq->m, pq->K, pq->dsub, da, db, n, dout);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

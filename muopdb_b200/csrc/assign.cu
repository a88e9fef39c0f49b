// assign.cu -- build-time arithmetic next to the query kernels:
//   LaneConformingDistanceCalculator<LANES, D>::calculate_squared   rs/utils/src/distance/lane_conforming.rs:16-27
//   assignment step of KMeansBuilder::run_lloyd                     rs/utils/src/kmeans_builder/kmeans_builder.rs:199-221
//     (calculator chosen by the dimension, kmeans_builder.rs:126-136)
//   ProductQuantizer::original_vector                               rs/quantization/src/pq/mod.rs:184-200
//
// k-means assignment: argmin_c (T::calculate_squared(x, c) + penalty_c), folded from (0, f32::MAX) with a strict '<' (first
// minimum wins, NaN never wins).  Large problems (L2, dim % 16 == 0) use the tensor-core distance estimate of coarse_tc.cu:
// with |D~_c - d_c| <= eps for every centroid, the true winner c* satisfies D~_c* + p_c* <= min_c (D~_c + p_c) + 2 eps, so
// re-scoring the centroids inside that band with the reference's arithmetic and taking their first minimum reproduces the
// reference exactly (same argument as DESIGN.md "coarse scoring").  Everything else runs the exact SIMT distance matrix.
#include <algorithm>

#include "internal.cuh"

// ---- lane-conforming all-pairs -----------------------------------------------------------------------------------------------
// one thread per pair; LANES accumulators over the whole vector (dim % LANES == 0), ordered reduce, outermost_op
template <int METRIC, int LANES>
__device__ __forceinline__ float lane_conforming_pair(const float *__restrict__ a, const float *__restrict__ b, uint32_t dim) {
  float acc[LANES];
#pragma unroll
  for (int l = 0; l < LANES; l++) acc[l] = 0.0f;
  for (uint32_t c = 0; c < dim; c += LANES) {
#pragma unroll
    for (int l = 0; l < LANES; l++) {
      const float x = a[c + l], y = b[c + l];
      if (METRIC == MGPU_L2) { const float d = __fsub_rn(x, y); acc[l] = __fadd_rn(acc[l], __fmul_rn(d, d)); }
      else acc[l] = __fadd_rn(acc[l], __fmul_rn(x, y));
    }
  }
  float s = -0.0f;  // reduce_sum: ordered, identity -0.0
#pragma unroll
  for (int l = 0; l < LANES; l++) s = __fadd_rn(s, acc[l]);
  return METRIC == MGPU_L2 ? s : -s;  // D::outermost_op
}

template <int METRIC, int LANES>
__global__ void k_distance_lanes(const float *__restrict__ A, uint64_t nA, const float *__restrict__ Bm, uint64_t nB, uint32_t dim,
                                 float *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA * nB) return;
  const uint64_t r = i / nB, c = i % nB;
  out[i] = lane_conforming_pair<METRIC, LANES>(A + r * dim, Bm + c * dim, dim);
}

template <int METRIC>
static int launch_lanes_m(mgpu_ctx *ctx, const float *dA, uint64_t nA, const float *dB, uint64_t nB, uint32_t dim, int lanes, float *dout) {
  const uint64_t total = nA * nB;
  const unsigned grid = (unsigned)((total + 127) / 128);
  switch (lanes) {
    case 1: k_distance_lanes<METRIC, 1><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, dout); break;
    case 2: k_distance_lanes<METRIC, 2><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, dout); break;
    case 4: k_distance_lanes<METRIC, 4><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, dout); break;
    case 8: k_distance_lanes<METRIC, 8><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, dout); break;
    case 16: k_distance_lanes<METRIC, 16><<<grid, 128, 0, ctx->stream>>>(dA, nA, dB, nB, dim, dout); break;
    default: return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "lane count %d is not one of 1, 2, 4, 8, 16", lanes);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

int launch_distance_lanes(mgpu_ctx *ctx, const float *dA, uint64_t nA, const float *dB, uint64_t nB, uint32_t dim, int metric,
                          int lanes, float *dout, int kernel_class) {
  if (nA == 0 || nB == 0) return MGPU_OK;
  if (nA * nB > 0xFFFFFFFFull * 64) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "lane-conforming distance matrix too large for one call");
  // LANES = 16 is the first phase of the generic cascade with no remainder: the tiled kernel computes exactly that
  if (lanes == 16 && dim % 16 == 0 && !(metric == MGPU_DOT && dim == 16))
    return launch_distance_matrix(ctx, dA, nA, dB, nB, dim, metric, 0, dout, kernel_class);
  LaunchScope ls(ctx, kernel_class);
  return metric == MGPU_L2 ? launch_lanes_m<MGPU_L2>(ctx, dA, nA, dB, nB, dim, lanes, dout)
                           : launch_lanes_m<MGPU_DOT>(ctx, dA, nA, dB, nB, dim, lanes, dout);
}

// ---- k-means assignment: exact matrix -> first minimum of (d + penalty) ---------------------------------------------------------
// one warp per row; lanes stride over the centroids, each keeps its first minimum, then the warp combines by (cost, index)
__global__ void __launch_bounds__(256) k_kmeans_argmin(const float *__restrict__ D, uint64_t n, uint32_t C,
                                                        const float *__restrict__ pen, uint32_t *__restrict__ labels,
                                                        float *__restrict__ costs) {
  const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float *d = D + row * C;
  float best = 3.40282347e+38f;
  uint32_t bi = 0xFFFFFFFFu;
  for (uint32_t c = lane; c < C; c += 32) {
    const float v = __fadd_rn(d[c], pen ? pen[c] : 0.0f);
    if (v < best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    labels[row] = bi == 0xFFFFFFFFu ? 0u : bi;   // nothing below f32::MAX: the fold keeps (0, f32::MAX)
    if (costs) costs[row] = best;
  }
}

// ---- k-means assignment: tensor-core estimate -> band -> exact re-score --------------------------------------------------------
// one warp per row.  Dt = D~ (n x C) from k_coarse_gemm.  eps as in coarse_tc.cu (TC_ERR_REL * (|x|^2 + max |c|^2)) plus the
// rounding of the penalty addition.
#define KM_ERR_REL 3.0e-4f
__global__ void __launch_bounds__(256) k_kmeans_pick_tc(const float *__restrict__ Dt, const float *__restrict__ X,
                                                         const float *__restrict__ centroids, const float *__restrict__ xn,
                                                         const float *__restrict__ cn_max_p, uint64_t n, uint32_t C, uint32_t dim,
                                                         const float *__restrict__ pen, uint32_t *__restrict__ labels,
                                                         float *__restrict__ costs) {
  const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float *d = Dt + row * C;
  float vmin = 3.40282347e+38f;
  for (uint32_t c = lane; c < C; c += 32) vmin = fminf(vmin, d[c] + (pen ? pen[c] : 0.0f));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  // + the rounding of the penalty additions (one ulp of a value next to the minimum)
  const float eps = KM_ERR_REL * (xn[row] + cn_max_p[0]) + 1.0e-6f * fabsf(vmin);
  const float lim = vmin + 2.0f * eps;
  const float *x = X + row * dim;
  float best = 3.40282347e+38f;
  uint32_t bi = 0xFFFFFFFFu;
  const int h = lane & 15, half = lane >> 4;
  for (uint32_t c0 = 0; c0 < C; c0 += 32) {
    const uint32_t c = c0 + lane;
    const bool inband = c < C && !(d[c] + (pen ? pen[c] : 0.0f) > lim);   // NaN estimates stay in the band
    unsigned mask = __ballot_sync(0xffffffffu, inband);
    while (mask) {
      // two band members at a time, one per half-warp: lane h owns lane-accumulator h of LaneConforming<16> (== l2.rs:30-68
      // for dim % 16 == 0)
      const int l0 = __ffs(mask) - 1;
      mask &= mask - 1;
      int l1 = -1;
      if (mask) { l1 = __ffs(mask) - 1; mask &= mask - 1; }
      const int mine = half == 0 ? l0 : l1;
      const uint32_t cidx = c0 + (uint32_t)(mine < 0 ? l0 : mine);
      const float *crow = centroids + (size_t)cidx * dim;
      float acc = 0.0f;
      for (uint32_t k = 0; k < dim; k += 16) {
        const float df = __fsub_rn(x[k + h], __ldg(crow + k + h));
        acc = __fadd_rn(acc, __fmul_rn(df, df));
      }
      float s = -0.0f;
      const int basel = lane & 16;
#pragma unroll
      for (int l = 0; l < 16; l++) s = __fadd_rn(s, __shfl_sync(0xffffffffu, acc, basel + l));
      const float v = __fadd_rn(s, pen ? pen[cidx] : 0.0f);
      if (mine >= 0 && h == 0 && (v < best || (v == best && cidx < bi))) { best = v; bi = cidx; }
    }
  }
  // combine lanes 0 and 16
  const float ob = __shfl_sync(0xffffffffu, best, 16);
  const uint32_t oi = __shfl_sync(0xffffffffu, bi, 16);
  if (lane == 0) {
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    labels[row] = bi == 0xFFFFFFFFu ? 0u : bi;
    if (costs) costs[row] = best;
  }
}

int launch_kmeans_argmin(mgpu_ctx *ctx, const float *dD, uint64_t n, uint32_t C, const float *d_pen, uint32_t *d_labels, float *d_costs) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_SELECT);
  k_kmeans_argmin<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(dD, n, C, d_pen, d_labels, d_costs);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

__global__ void k_max_f32(const float *__restrict__ v, uint32_t n, float *__restrict__ out) {
  __shared__ float sm[32];
  float mx = 0.0f;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, v[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t w = 1; w < (blockDim.x + 31) / 32; w++) mx = fmaxf(mx, sm[w]);
    out[0] = mx * 1.000001f;
  }
}

int launch_max_f32(mgpu_ctx *ctx, const float *d_v, uint32_t n, float *d_out) {
  LaunchScope ls(ctx, MGPU_K_OTHER);
  k_max_f32<<<1, 1024, 0, ctx->stream>>>(d_v, n, d_out);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

int launch_kmeans_pick_tc(mgpu_ctx *ctx, const float *d_Dt, const float *dX, const float *d_centroids, const float *d_xn,
                          const float *d_cn_max, uint64_t n, uint32_t C, uint32_t dim, const float *d_pen, uint32_t *d_labels,
                          float *d_costs) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_SELECT);
  k_kmeans_pick_tc<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(d_Dt, dX, d_centroids, d_xn, d_cn_max, n, C, dim, d_pen, d_labels,
                                                                      d_costs);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- ProductQuantizer::original_vector (pq/mod.rs:184-200): a codebook gather ----------------------------------------------------
__global__ void k_pq_original(const float *__restrict__ cb, uint32_t m, uint32_t K, uint32_t dsub, const uint8_t *__restrict__ codes,
                              uint64_t n, float *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t dim = (uint64_t)m * dsub;
  if (i >= n * dim) return;
  const uint64_t row = i / dim;
  const uint32_t d = (uint32_t)(i % dim), s = d / dsub, e = d % dsub;
  out[i] = cb[((size_t)s * K + codes[row * m + s]) * dsub + e];
}

int launch_pq_original(mgpu_pq *pq, const uint8_t *d_codes, uint64_t n, float *d_out) {
  mgpu_ctx *ctx = pq->ctx;
  if (n == 0) return MGPU_OK;
  const uint64_t total = n * pq->dim;
  LaunchScope ls(ctx, MGPU_K_OTHER);
  k_pq_original<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(pq->d_cb, pq->m, pq->K, pq->dsub, d_codes, n, d_out);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// ---- gathers behind the BlockBasedIvf accessors (index.rs:350-384) --------------------------------------------------------------
__global__ void k_gather_docs(const mgpu_u128 *__restrict__ doc_ids, const uint32_t *__restrict__ pids, uint32_t n, uint64_t nvec,
                              mgpu_u128 *__restrict__ out, uint32_t *__restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = pids[i];
  if (p >= nvec) { atomicAdd(bad, 1u); out[i] = mgpu_u128{0, 0}; return; }
  out[i] = doc_ids ? doc_ids[p] : mgpu_u128{p, 0};
}

int launch_gather_docs(mgpu_ctx *ctx, const mgpu_u128 *d_doc_ids, const uint32_t *d_pids, uint32_t n, uint64_t nvec, mgpu_u128 *d_out,
                       uint32_t *d_bad) {
  if (n == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_OTHER);
  k_gather_docs<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_doc_ids, d_pids, n, nvec, d_out, d_bad);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// point id -> one slot holding its row.  A point living in several lists has the same row in each of them; the smallest
// slot is kept so that the map does not depend on thread scheduling.
__global__ void k_pid_slot(const uint32_t *__restrict__ slot_pid, uint64_t nslots, uint32_t *__restrict__ pid_slot) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const uint32_t p = slot_pid[s];
  if (p != MGPU_EMPTY_SLOT) atomicMin(&pid_slot[p], (uint32_t)s);
}

int ivf_ensure_pid_slot(mgpu_ivf *ivf) {
  mgpu_ctx *ctx = ivf->ctx;
  if (ivf->d_pid_slot) return MGPU_OK;
  const uint64_t nslots = ivf->total_chunks * 32;
  CUDA_TRY(ctx, cudaMalloc((void **)&ivf->d_pid_slot, std::max<uint64_t>(ivf->n, 1) * 4));
  CUDA_TRY(ctx, cudaMemsetAsync(ivf->d_pid_slot, 0xFF, std::max<uint64_t>(ivf->n, 1) * 4, ctx->stream));
  if (nslots) {
    LaunchScope ls(ctx, MGPU_K_OTHER);
    k_pid_slot<<<(unsigned)((nslots + 255) / 256), 256, 0, ctx->stream>>>(ivf->d_slot_pid, nslots, ivf->d_pid_slot);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return MGPU_OK;
}

// BlockBasedIvf::get_vector (index.rs:372-384): the quantized row of a point, read back out of the scan layouts
// (internal.cuh): PQ fast = XOR-permuted 16-byte units, PQ generic = row-major by slot, flat = [dim4][32 lanes] float4.
__global__ void k_gather_rows(const uint32_t *__restrict__ pid_slot, const uint32_t *__restrict__ pids, uint32_t n, uint64_t nvec,
                              const uint8_t *__restrict__ codes, const float *__restrict__ rows, uint32_t m, uint32_t ng, int pq_fast,
                              uint32_t dim, uint32_t dim4, uint8_t *__restrict__ out_codes, float *__restrict__ out_rows,
                              uint32_t *__restrict__ bad) {
  const uint32_t i = blockIdx.x;
  const uint32_t p = pids[i];
  uint32_t slot = p < nvec ? pid_slot[p] : MGPU_EMPTY_SLOT;
  if (slot == MGPU_EMPTY_SLOT) { if (threadIdx.x == 0) atomicAdd(bad, 1u); return; }
  const uint32_t l = slot & 31;
  const size_t chunk = slot >> 5;
  if (out_codes) {
    for (uint32_t s = threadIdx.x; s < m; s += blockDim.x) {
      uint8_t c;
      if (pq_fast) c = codes[pq_fast_code_offset(chunk, ng, l, s)];
      else c = codes[(size_t)slot * m + s];
      out_codes[(size_t)i * m + s] = c;
    }
  } else {
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x)
      out_rows[(size_t)i * dim + d] = rows[((chunk * dim4 + (d >> 2)) * 32 + l) * 4 + (d & 3)];
  }
}

int launch_gather_rows(mgpu_ivf *ivf, const uint32_t *d_pids, uint32_t n, void *d_out, uint32_t *d_bad) {
  mgpu_ctx *ctx = ivf->ctx;
  if (n == 0) return MGPU_OK;
  MGPU_TRY(ivf_ensure_pid_slot(ivf));
  LaunchScope ls(ctx, MGPU_K_OTHER);
  const bool pq = ivf->quant == MGPU_QUANT_PQ;
  k_gather_rows<<<n, 128, 0, ctx->stream>>>(ivf->d_pid_slot, d_pids, n, ivf->n, ivf->d_codes, ivf->d_rows, pq ? ivf->pq->m : 0, ivf->ng,
                                             ivf->pq_fast ? 1 : 0, ivf->dim, ivf->dim4, pq ? (uint8_t *)d_out : nullptr,
                                             pq ? nullptr : (float *)d_out, d_bad);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// api_build.cu -- extern "C" entry points next to the search calls: lane-conforming distances, the k-means assignment step,
// ProductQuantizer::original_vector and the BlockBasedIvf accessors (doc ids / point ids / stored rows).
#include <unordered_map>

#include "internal.cuh"

namespace {
struct U128Hash {
  size_t operator()(const std::pair<uint64_t, uint64_t> &k) const {
    uint64_t h = k.first * 0x9E3779B97F4A7C15ull;
    h ^= (k.second + 0x7F4A7C15ull + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
typedef std::unordered_map<std::pair<uint64_t, uint64_t>, uint32_t, U128Hash> DocMap;
}  // namespace

void ivf_free_doc_map(mgpu_ivf *ivf) {
  delete (DocMap *)ivf->doc_map;
  ivf->doc_map = nullptr;
}

// doc_id_to_point_id (index.rs:67-73): built from the doc-id table in point-id order, a repeated doc id keeps its LAST point
static int ivf_doc_map(mgpu_ivf *ivf, DocMap **out) {
  mgpu_ctx *ctx = ivf->ctx;
  if (!ivf->doc_map) {
    DocMap *mp = new DocMap();
    mp->reserve(ivf->n);
    if (ivf->d_doc_ids) {
      std::vector<mgpu_u128> h(ivf->n);
      if (ivf->n) {
        cudaError_t e = cudaMemcpyAsync(h.data(), ivf->d_doc_ids, ivf->n * sizeof(mgpu_u128), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { delete mp; return mgpu_fail(ctx, MGPU_ERR_CUDA, "reading the doc-id table failed: %s", cudaGetErrorString(e)); }
      }
      for (uint64_t i = 0; i < ivf->n; i++) (*mp)[{h[i].lo, h[i].hi}] = (uint32_t)i;
    } else {
      for (uint64_t i = 0; i < ivf->n; i++) (*mp)[{i, 0}] = (uint32_t)i;
    }
    ivf->doc_map = mp;
  }
  *out = (DocMap *)ivf->doc_map;
  return MGPU_OK;
}

extern "C" {

/* LaneConformingDistanceCalculator<LANES, D>::calculate_squared (lane_conforming.rs:16-27), all pairs. */
int mgpu_distance_batch_lanes(mgpu_ctx *ctx, const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim, int metric,
                              int lanes, float *out, int mem) {
  if (!ctx || !A || !B || !out || dim == 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch_lanes: null/zero argument");
  if (metric != MGPU_L2 && metric != MGPU_DOT) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch_lanes: bad metric");
  if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch_lanes: LANES must be 1, 2, 4, 8 or 16");
  // chunks_exact(LANES) silently drops a remainder in the reference; the calculator is only defined for conforming dims
  if (dim % (uint32_t)lanes != 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch_lanes: dim %u is not a multiple of LANES %d", dim, lanes);
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  const size_t bA = nA * dim * 4, bB = nB * dim * 4, bO = nA * nB * 4;
  const float *dA = A, *dB = B;
  float *dO = out;
  if (mem == MGPU_HOST) {
    MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(ws_need(0, bA), bB), bO)));
    WsAlloc w(ctx->ws, ctx->ws_bytes);
    float *a = w.get<float>(nA * dim), *b = w.get<float>(nB * dim);
    dO = w.get<float>(nA * nB);
    CUDA_TRY(ctx, cudaMemcpyAsync(a, A, bA, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(b, B, bB, cudaMemcpyHostToDevice, ctx->stream));
    dA = a; dB = b;
  }
  MGPU_TRY(launch_distance_lanes(ctx, dA, nA, dB, nB, dim, metric, lanes, dO, MGPU_K_OTHER));
  if (mem == MGPU_HOST) {
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dO, bO, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MGPU_OK;
}

/* Assignment step of KMeansBuilder::run_lloyd (kmeans_builder.rs:199-221). */
int mgpu_kmeans_assign(mgpu_ctx *ctx, const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim, int metric,
                       const float *penalties, uint32_t *out_labels, float *out_costs, int mem) {
  if (!ctx || (n && (!X || !out_labels)) || !centroids || nlist == 0 || dim == 0)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "kmeans_assign: null/zero argument");
  if (metric != MGPU_L2 && metric != MGPU_DOT) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "kmeans_assign: bad metric");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  // calculator by dimension (kmeans_builder.rs:126-136); 0 = D::calculate_squared itself
  const int lanes = dim % 16 == 0 ? 16 : (dim % 8 == 0 ? 8 : (dim % 4 == 0 ? 4 : 0));
  static const bool no_tc = getenv("MGPU_KMEANS_TC") && getenv("MGPU_KMEANS_TC")[0] == '0';
  const bool tc = !no_tc && metric == MGPU_L2 && lanes == 16 && coarse_tc_applicable(ctx, dim, nlist, 1);
  const uint32_t Kp = coarse_tc_kp(dim);
  uint64_t slab = std::max<uint64_t>(128, std::min<uint64_t>(n, (512ull << 20) / ((uint64_t)nlist * 4)));
  slab = std::min<uint64_t>(slab, 65535ull * 32);
  size_t need = 0;
  need = ws_need(need, mem == MGPU_HOST ? (size_t)nlist * dim * 4 : 0);
  need = ws_need(need, mem == MGPU_HOST && penalties ? (size_t)nlist * 4 : 0);
  need = ws_need(need, mem == MGPU_HOST ? slab * dim * 4 : 0);
  need = ws_need(need, slab * nlist * 4);
  need = ws_need(need, mem == MGPU_HOST ? slab * 4 : 0);
  need = ws_need(need, mem == MGPU_HOST ? slab * 4 : 0);
  need = ws_need(need, tc ? (size_t)nlist * Kp * 2 : 0);
  need = ws_need(need, tc ? (size_t)nlist * 4 + 16 : 0);
  need = ws_need(need, tc ? slab * Kp * 2 : 0);
  need = ws_need(need, tc ? slab * 4 : 0);
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  float *sC = w.get<float>(mem == MGPU_HOST ? (size_t)nlist * dim : 0);
  float *sP = w.get<float>(mem == MGPU_HOST && penalties ? nlist : 0);
  float *sX = w.get<float>(mem == MGPU_HOST ? slab * dim : 0);
  float *dD = w.get<float>(slab * nlist);
  uint32_t *sL = w.get<uint32_t>(mem == MGPU_HOST ? slab : 0);
  float *sV = w.get<float>(mem == MGPU_HOST ? slab : 0);
  uint16_t *csplit = w.get<uint16_t>(tc ? (size_t)nlist * Kp : 0);
  float *cn = w.get<float>(tc ? nlist + 4 : 0);
  uint16_t *xsplit = w.get<uint16_t>(tc ? slab * Kp : 0);
  float *xn = w.get<float>(tc ? slab : 0);
  const float *dC = centroids, *dP = penalties;
  if (mem == MGPU_HOST) {
    CUDA_TRY(ctx, cudaMemcpyAsync(sC, centroids, (size_t)nlist * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
    dC = sC;
    if (penalties) { CUDA_TRY(ctx, cudaMemcpyAsync(sP, penalties, (size_t)nlist * 4, cudaMemcpyHostToDevice, ctx->stream)); dP = sP; }
  }
  if (tc) {
    MGPU_TRY(launch_split_bf16(ctx, dC, nlist, dim, 1, csplit, cn));
    MGPU_TRY(launch_max_f32(ctx, cn, nlist, cn + nlist));
  }
  for (uint64_t i = 0; i < n; i += slab) {
    const uint64_t cnt = std::min(slab, n - i);
    const float *dX = X + i * dim;
    if (mem == MGPU_HOST) { CUDA_TRY(ctx, cudaMemcpyAsync(sX, X + i * dim, cnt * dim * 4, cudaMemcpyHostToDevice, ctx->stream)); dX = sX; }
    uint32_t *oL = mem == MGPU_DEVICE ? out_labels + i : sL;
    float *oV = mem == MGPU_DEVICE ? (out_costs ? out_costs + i : nullptr) : (out_costs ? sV : nullptr);
    if (tc) {
      MGPU_TRY(launch_tc_distances(ctx, dX, (uint32_t)cnt, csplit, cn, nlist, dim, xsplit, xn, dD));
      MGPU_TRY(launch_kmeans_pick_tc(ctx, dD, dX, dC, xn, cn + nlist, cnt, nlist, dim, dP, oL, oV));
    } else {
      if (lanes) MGPU_TRY(launch_distance_lanes(ctx, dX, cnt, dC, nlist, dim, metric, lanes, dD, MGPU_K_COARSE));
      else MGPU_TRY(launch_distance_matrix(ctx, dX, cnt, dC, nlist, dim, metric, 0, dD, MGPU_K_COARSE));
      MGPU_TRY(launch_kmeans_argmin(ctx, dD, cnt, nlist, dP, oL, oV));
    }
    if (mem == MGPU_HOST) {
      CUDA_TRY(ctx, cudaMemcpyAsync(out_labels + i, oL, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
      if (out_costs) CUDA_TRY(ctx, cudaMemcpyAsync(out_costs + i, oV, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return MGPU_OK;
}

/* ProductQuantizer::original_vector (pq/mod.rs:184-200) for n code words: out is n x dim. */
int mgpu_pq_original_vector(mgpu_pq *pq, const uint8_t *codes, uint64_t n, float *out, int mem) {
  if (!pq || (n && (!codes || !out))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = pq->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  if (mem == MGPU_DEVICE) return launch_pq_original(pq, codes, n, out);
  const uint64_t slab = std::max<uint64_t>(1, (128ull << 20) / ((uint64_t)pq->dim * 4));
  MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(0, slab * pq->m), slab * pq->dim * 4)));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  uint8_t *dc = w.get<uint8_t>(slab * pq->m);
  float *dout = w.get<float>(slab * pq->dim);
  for (uint64_t i = 0; i < n; i += slab) {
    const uint64_t cnt = std::min(slab, n - i);
    CUDA_TRY(ctx, cudaMemcpyAsync(dc, codes + i * pq->m, cnt * pq->m, cudaMemcpyHostToDevice, ctx->stream));
    MGPU_TRY(launch_pq_original(pq, dc, cnt, dout));
    CUDA_TRY(ctx, cudaMemcpyAsync(out + i * pq->dim, dout, cnt * pq->dim * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MGPU_OK;
}

/* BlockBasedIvf::get_doc_id / get_doc_ids (index.rs:350-366): doc ids of n point ids, same order.  HOST buffers. */
int mgpu_ivf_get_doc_ids(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n, mgpu_u128 *out_doc_ids) {
  if (!ivf || (n && (!point_ids || !out_doc_ids))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(ws_need(0, (size_t)n * 4), (size_t)n * 16), 16)));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  uint32_t *dp = w.get<uint32_t>(n);
  mgpu_u128 *dd = w.get<mgpu_u128>(n);
  uint32_t *bad = w.get<uint32_t>(4);
  CUDA_TRY(ctx, cudaMemcpyAsync(dp, point_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
  MGPU_TRY(launch_gather_docs(ctx, ivf->d_doc_ids, dp, n, ivf->n, dd, bad));
  uint32_t hbad = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_doc_ids, dd, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (hbad) return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "get_doc_ids: %u point ids out of range (the reference returns an error)", hbad);
  return MGPU_OK;
}

/* BlockBasedIvf::get_point_id (index.rs:469-471): *found = 0 encodes None. */
int mgpu_ivf_get_point_id(mgpu_ivf *ivf, const mgpu_u128 *doc_id, int *found, uint32_t *point_id) {
  if (!ivf || !doc_id || !found || !point_id) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  DocMap *mp;
  MGPU_TRY(ivf_doc_map(ivf, &mp));
  auto it = mp->find({doc_id->lo, doc_id->hi});
  *found = it != mp->end();
  *point_id = *found ? it->second : 0;
  return MGPU_OK;
}

/* BlockBasedIvf::get_vector (index.rs:372-384) for n point ids: the stored (quantized) rows, n x quantized_dimension of u8
 * (PQ) or f32 (NoQuantizer).  HOST buffers. */
int mgpu_ivf_get_vectors(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n, void *out_rows) {
  if (!ivf || (n && (!point_ids || !out_rows))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  const size_t rb = ivf->quant == MGPU_QUANT_PQ ? ivf->pq->m : (size_t)ivf->dim * 4;
  MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(ws_need(0, (size_t)n * 4), (size_t)n * rb), 16)));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  uint32_t *dp = w.get<uint32_t>(n);
  uint8_t *dr = w.get<uint8_t>((size_t)n * rb);
  uint32_t *bad = w.get<uint32_t>(4);
  CUDA_TRY(ctx, cudaMemcpyAsync(dp, point_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
  MGPU_TRY(launch_gather_rows(ivf, dp, n, dr, bad));
  uint32_t hbad = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_rows, dr, (size_t)n * rb, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (hbad) return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "get_vectors: %u point ids out of range or stored in no posting list", hbad);
  return MGPU_OK;
}

/* BlockBasedIvf::invalidate_batch(doc_ids) (index.rs:417-452): doc id -> point id (None => not invalidated), then
 * DashSet::insert, whose return value (newly inserted?) decides whether the doc id counts as invalidated.  out_ok (may be
 * NULL): n flags; *out_num_ok (may be NULL): how many were newly invalidated. */
int mgpu_ivf_invalidate_docs(mgpu_ivf *ivf, const mgpu_u128 *doc_ids, uint32_t n, uint8_t *out_ok, uint32_t *out_num_ok) {
  if (!ivf || (n && !doc_ids)) return MGPU_ERR_INVALID_ARG;
  if (out_num_ok) *out_num_ok = 0;
  std::vector<uint32_t> pids;
  std::vector<uint32_t> idx;
  {
    mgpu_ctx *ctx = ivf->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    DocMap *mp;
    MGPU_TRY(ivf_doc_map(ivf, &mp));
    for (uint32_t i = 0; i < n; i++) {
      if (out_ok) out_ok[i] = 0;
      auto it = mp->find({doc_ids[i].lo, doc_ids[i].hi});
      if (it != mp->end()) { pids.push_back(it->second); idx.push_back(i); }
    }
  }
  uint32_t num = 0;
  for (size_t j = 0; j < pids.size(); j++) {   // sequential, like the reference's loop: a repeated doc id succeeds once
    int was = 0;
    MGPU_TRY(mgpu_ivf_is_invalidated(ivf, pids[j], &was));
    if (!was) {
      MGPU_TRY(mgpu_ivf_invalidate(ivf, &pids[j], 1));
      if (out_ok) out_ok[idx[j]] = 1;
      num++;
    }
  }
  if (out_num_ok) *out_num_ok = num;
  return MGPU_OK;
}

/* BlockBasedIvf::is_invalidated(doc_id) (index.rs:454-459): unknown doc ids are not invalidated. */
int mgpu_ivf_is_doc_invalidated(mgpu_ivf *ivf, const mgpu_u128 *doc_id, int *out) {
  if (!ivf || !doc_id || !out) return MGPU_ERR_INVALID_ARG;
  int found = 0;
  uint32_t pid = 0;
  MGPU_TRY(mgpu_ivf_get_point_id(ivf, doc_id, &found, &pid));
  *out = 0;
  if (!found) return MGPU_OK;
  return mgpu_ivf_is_invalidated(ivf, pid, out);
}

}  // extern "C"

// api.cu -- the extern "C" boundary declared in include/muopdb_gpu.h: contexts, resident index handles and the
// per-call orchestration (H2D staging -> kernels on the ctx stream -> D2H).  No CPU compute fallback exists:
// every entry point that computes needs a CUDA device.
#include <chrono>
#include <dlfcn.h>
#include <stdarg.h>

#include <algorithm>

#include "internal.cuh"

int mgpu_fail(mgpu_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int mgpu_ws_reserve(mgpu_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return MGPU_OK;
  if (ctx->ws) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_bytes = 0; }
  size_t want = bytes + bytes / 4;
  CUDA_TRY(ctx, cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return MGPU_OK;
}

int mgpu_pinned_reserve(mgpu_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->pinned_bytes) return MGPU_OK;
  if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
  size_t want = bytes + bytes / 4;
  CUDA_TRY(ctx, cudaMallocHost(&ctx->pinned, want));
  ctx->pinned_bytes = want;
  return MGPU_OK;
}

template <class T>
static int dev_alloc_copy(mgpu_ctx *ctx, T **dst, const T *src, size_t n, bool src_is_device = false) {
  *dst = nullptr;
  if (n == 0) n = 1;
  CUDA_TRY(ctx, cudaMalloc((void **)dst, n * sizeof(T)));
  if (src) CUDA_TRY(ctx, cudaMemcpyAsync(*dst, src, n * sizeof(T), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  return MGPU_OK;
}

extern "C" {

const char *mgpu_version(void) { return "muopdb_b200 0.1 (sm_100a)"; }

int mgpu_init(int device, mgpu_ctx **out) {
  if (!out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return MGPU_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = new mgpu_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MGPU_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return MGPU_ERR_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MGPU_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return MGPU_ERR_CUDA; }
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  cudaEventCreate(&ctx->t0);
  cudaEventCreate(&ctx->t1);
  *out = ctx;
  return MGPU_OK;
}

void mgpu_destroy(mgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mgpu_comm_destroy(ctx);
  mgpu_profile_reset(ctx);
  for (auto &pp : ctx->pipe) {
    if (pp.ev_out && pp.pending) cudaEventSynchronize(pp.ev_out);
    if (pp.q) cudaFree(pp.q);
    if (pp.out) cudaFree(pp.out);
    if (pp.ev_h2d) cudaEventDestroy(pp.ev_h2d);
    if (pp.ev_done) cudaEventDestroy(pp.ev_done);
    if (pp.ev_out) cudaEventDestroy(pp.ev_out);
  }
  if (ctx->shard_ws) cudaFree(ctx->shard_ws);
  for (auto &sl : ctx->shard_slot) {
    if (sl.buf) cudaFree(sl.buf);
    if (sl.ev_local) cudaEventDestroy(sl.ev_local);
    if (sl.ev_xdone) cudaEventDestroy(sl.ev_xdone);
  }
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  cudaEventDestroy(ctx->t0);
  cudaEventDestroy(ctx->t1);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_join);
  cudaStreamSynchronize(ctx->aux_stream);
  cudaStreamDestroy(ctx->aux_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *mgpu_last_error(mgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int mgpu_sync(mgpu_ctx *ctx) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));   // a sharded call's exchange may still be in flight
  return MGPU_OK;
}
void *mgpu_stream(mgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// Ordering against a caller-owned stream (the ctx stream is non-blocking: it does not synchronise with the legacy stream or
// with any other stream implicitly).  wait: the ctx stream waits for everything enqueued on `other` so far (inputs produced
// there, e.g. a zero-fill of the output buffers); signal: `other` waits for everything enqueued on the ctx streams so far
// (results).  The events are created per call: two calls may be in flight from different threads of one ctx user.
int mgpu_stream_wait(mgpu_ctx *ctx, void *other) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  if ((cudaStream_t)other == ctx->stream) return MGPU_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaEvent_t e;
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaError_t r = cudaEventRecord(e, (cudaStream_t)other);
  if (r == cudaSuccess) r = cudaStreamWaitEvent(ctx->stream, e, 0);
  cudaEventDestroy(e);   // released once the wait has consumed it
  CUDA_TRY(ctx, r);
  return MGPU_OK;
}
int mgpu_stream_signal(mgpu_ctx *ctx, void *other) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t srcs[2] = {ctx->stream, ctx->comm_stream};
  for (cudaStream_t st : srcs) {
    if (!st || st == (cudaStream_t)other) continue;
    cudaEvent_t e;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t r = cudaEventRecord(e, st);
    if (r == cudaSuccess) r = cudaStreamWaitEvent((cudaStream_t)other, e, 0);
    cudaEventDestroy(e);
    CUDA_TRY(ctx, r);
  }
  return MGPU_OK;
}
int mgpu_device_sm_count(mgpu_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int mgpu_timer_start(mgpu_ctx *ctx) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  CUDA_TRY(ctx, cudaEventRecord(ctx->t0, ctx->stream));
  return MGPU_OK;
}
int mgpu_timer_stop(mgpu_ctx *ctx, float *ms) {
  if (!ctx || !ms) return MGPU_ERR_INVALID_ARG;
  CUDA_TRY(ctx, cudaEventRecord(ctx->t1, ctx->stream));
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->t1));
  CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
  return MGPU_OK;
}

int mgpu_profile_enable(mgpu_ctx *ctx, int on) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  // on > 0: every class; on < 0: -on is a bit mask of classes (e.g. -(1 << MGPU_K_SCAN): only the scan kernel, so that the
  // other launches of a timed region run back to back without event records between them); 0: off
  ctx->profiling = on > 0 ? 0xFFFFFFFFu : (on < 0 ? (uint32_t)(-on) : 0u);
  return MGPU_OK;
}
int mgpu_profile_reset(mgpu_ctx *ctx) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  cudaStreamSynchronize(ctx->stream);
  for (auto &p : ctx->prof) {
    for (auto e : p.ev) cudaEventDestroy(e);
    p.ev.clear(); p.launches = 0; p.done_ms = 0.f;
  }
  return MGPU_OK;
}
int mgpu_coarse_band_stats(mgpu_ctx *ctx, uint64_t out[2], int reset) {
  if (!ctx || !out) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  return coarse_band_stats(ctx, out, reset);
}

const char *mgpu_last_kernel(mgpu_ctx *ctx, int cls) {
  if (!ctx || cls < 0 || cls >= MGPU_K_COUNT) return "";
  std::lock_guard<std::mutex> g(ctx->mu);
  return ctx->last_kernel[cls] ? ctx->last_kernel[cls] : "";
}

int mgpu_profile_get(mgpu_ctx *ctx, int cls, float *total_ms, uint64_t *launches) {
  if (!ctx || cls < 0 || cls >= MGPU_K_COUNT) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  auto &p = ctx->prof[cls];
  for (size_t i = 0; i + 1 < p.ev.size(); i += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
    p.done_ms += ms;
  }
  for (auto e : p.ev) cudaEventDestroy(e);
  p.ev.clear();
  if (total_ms) *total_ms = p.done_ms;
  if (launches) *launches = p.launches;
  return MGPU_OK;
}
uint64_t mgpu_launch_count(mgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- helpers: stage caller buffers ------------------------------------------------------------------------------------
// in: returns a device pointer holding `bytes` of src (copying through the workspace when src is host memory)
static int stage_in(mgpu_ctx *ctx, const void *src, size_t bytes, int mem, void *ws_slot, const void **dptr) {
  if (mem == MGPU_DEVICE) { *dptr = src; return MGPU_OK; }
  CUDA_TRY(ctx, cudaMemcpyAsync(ws_slot, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dptr = ws_slot;
  return MGPU_OK;
}
static int stage_out(mgpu_ctx *ctx, void *dst, const void *dsrc, size_t bytes, int mem) {
  if (mem == MGPU_DEVICE || dst == nullptr) return MGPU_OK;
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MGPU_OK;
}

// ---- distances ------------------------------------------------------------------------------------------------------------
int mgpu_distance_batch(mgpu_ctx *ctx, const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim,
                        int metric, int squared, float *out, int mem) {
  if (!ctx || !A || !B || !out || dim == 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch: null/zero argument");
  if (metric != MGPU_L2 && metric != MGPU_DOT) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "distance_batch: bad metric");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  size_t bA = nA * dim * 4, bB = nB * dim * 4, bO = nA * nB * 4;
  const float *dA = A, *dB = B;
  float *dO = out;
  if (mem == MGPU_HOST) {
    size_t need = ws_need(ws_need(ws_need(0, bA), bB), bO);
    MGPU_TRY(mgpu_ws_reserve(ctx, need));
    WsAlloc w(ctx->ws, ctx->ws_bytes);
    float *a = w.get<float>(nA * dim), *b = w.get<float>(nB * dim);
    dO = w.get<float>(nA * nB);
    CUDA_TRY(ctx, cudaMemcpyAsync(a, A, bA, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(b, B, bB, cudaMemcpyHostToDevice, ctx->stream));
    dA = a; dB = b;
  }
  MGPU_TRY(launch_distance_matrix(ctx, dA, nA, dB, nB, dim, metric, squared ? 0 : 1, dO, MGPU_K_OTHER));
  if (mem == MGPU_HOST) {
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dO, bO, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MGPU_OK;
}

// ---- product quantizer ----------------------------------------------------------------------------------------------------
int mgpu_pq_create(mgpu_ctx *ctx, uint32_t dim, uint32_t dsub, uint32_t nbits, const float *codebook, int metric,
                   mgpu_pq **out) {
  if (!ctx || !out || !codebook) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_create: null argument");
  *out = nullptr;
  // ProductQuantizerConfig::validate (pq/mod.rs:41-46)
  if (dsub == 0 || dim == 0 || dim % dsub != 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_create: Dimensions are not valid");
  if (nbits == 0 || nbits > 8) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "pq_create: num_bits must be in 1..8 (codes are u8)");
  if (metric != MGPU_L2 && metric != MGPU_DOT) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_create: bad metric");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  mgpu_pq *pq = new mgpu_pq();
  pq->ctx = ctx; pq->dim = dim; pq->dsub = dsub; pq->nbits = nbits; pq->m = dim / dsub; pq->K = 1u << nbits; pq->metric = metric;
  size_t ncb = (size_t)pq->m * pq->K * dsub;
  int s = dev_alloc_copy(ctx, &pq->d_cb, codebook, ncb);
  if (s == MGPU_OK) s = dev_alloc_copy<float>(ctx, &pq->d_table, nullptr, (size_t)pq->m * pq->K * pq->K);
  if (s == MGPU_OK) s = dev_alloc_copy<float>(ctx, &pq->d_rowmin, nullptr, (size_t)pq->m * pq->K);
  if (s == MGPU_OK) s = dev_alloc_copy<float>(ctx, &pq->d_rowmax, nullptr, (size_t)pq->m * pq->K);
  if (s == MGPU_OK) s = launch_pq_build_table(pq);
  if (s == MGPU_OK) s = launch_pq_build_table16(pq);
  if (s == MGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) s = mgpu_fail(ctx, MGPU_ERR_CUDA, "pq_create: table build failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (s != MGPU_OK) { mgpu_pq_destroy(pq); return s; }
  *out = pq;
  return MGPU_OK;
}

void mgpu_pq_destroy(mgpu_pq *pq) {
  if (!pq) return;
  cudaSetDevice(pq->ctx->device);
  cudaFree(pq->d_cb); cudaFree(pq->d_table); cudaFree(pq->d_rowmin); cudaFree(pq->d_rowmax); cudaFree(pq->d_table16);
  delete pq;
}

int mgpu_pq_quantize_batch(mgpu_pq *pq, const float *X, uint64_t n, uint8_t *codes, int mem) {
  if (!pq || (n && (!X || !codes))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = pq->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  const float *dX = X;
  uint8_t *dC = codes;
  if (mem == MGPU_HOST) {
    // bounded staging: encode in slabs so that huge inputs do not need a huge workspace
    const uint64_t slab = std::max<uint64_t>(1, (256ull << 20) / ((uint64_t)pq->dim * 4));
    MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(0, slab * pq->dim * 4), slab * pq->m)));
    WsAlloc w(ctx->ws, ctx->ws_bytes);
    float *dx = w.get<float>(slab * pq->dim);
    uint8_t *dc = w.get<uint8_t>(slab * pq->m);
    for (uint64_t i = 0; i < n; i += slab) {
      uint64_t cnt = std::min(slab, n - i);
      CUDA_TRY(ctx, cudaMemcpyAsync(dx, X + i * pq->dim, cnt * pq->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
      MGPU_TRY(launch_pq_quantize(pq, dx, cnt, dc));
      CUDA_TRY(ctx, cudaMemcpyAsync(codes + i * pq->m, dc, cnt * pq->m, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MGPU_OK;
  }
  return launch_pq_quantize(pq, dX, n, dC);
}

int mgpu_pq_distance_batch(mgpu_pq *pq, const uint8_t *a, const uint8_t *b, uint64_t n, float *out, int mem) {
  if (!pq || (n && (!a || !b || !out))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = pq->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  if (mem == MGPU_DEVICE) return launch_pq_distance_pairs(pq, a, b, n, out);
  size_t bc = n * pq->m;
  MGPU_TRY(mgpu_ws_reserve(ctx, ws_need(ws_need(ws_need(0, bc), bc), n * 4)));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  uint8_t *da = w.get<uint8_t>(bc), *db = w.get<uint8_t>(bc);
  float *dout = w.get<float>(n);
  CUDA_TRY(ctx, cudaMemcpyAsync(da, a, bc, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(db, b, bc, cudaMemcpyHostToDevice, ctx->stream));
  MGPU_TRY(launch_pq_distance_pairs(pq, da, db, n, dout));
  CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

// ---- IVF ------------------------------------------------------------------------------------------------------------------
int mgpu_ivf_create(mgpu_ctx *ctx, uint32_t dim, uint32_t nlist, const float *centroids, const uint64_t *list_offsets,
                    const uint32_t *list_point_ids, int quant, int metric, mgpu_pq *pq, const void *rows, int rows_mem,
                    uint64_t n, const mgpu_u128 *doc_ids, mgpu_ivf **out) {
  if (!ctx || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  if (!centroids || !list_offsets || dim == 0 || nlist == 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: null/zero argument");
  if (quant == MGPU_QUANT_PQ && (!pq || pq->dim != dim)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: PQ quantizer missing or of the wrong dimension");
  if (quant != MGPU_QUANT_PQ && quant != MGPU_QUANT_NONE) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: bad quantizer kind");
  if (n >= 0xFFFFFFFFull) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "ivf_create: point ids are u32");
  uint64_t total_ids = list_offsets[nlist];
  if (total_ids && (!list_point_ids || !rows)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: null rows/ids");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);

  mgpu_ivf *ivf = new mgpu_ivf();
  ivf->ctx = ctx; ivf->dim = dim; ivf->nlist = nlist; ivf->n = n; ivf->quant = quant;
  ivf->metric = quant == MGPU_QUANT_PQ ? pq->metric : metric;
  ivf->pq = quant == MGPU_QUANT_PQ ? pq : nullptr;
  ivf->dim4 = (dim + 3) / 4;
  // chunked slot table on the host
  std::vector<uint32_t> chunk_start(nlist + 1), list_len(nlist);
  uint64_t chunks = 0;
  for (uint32_t c = 0; c < nlist; c++) {
    if (list_offsets[c + 1] < list_offsets[c]) { delete ivf; return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: list_offsets not monotone"); }
    uint64_t len = list_offsets[c + 1] - list_offsets[c];
    chunk_start[c] = (uint32_t)chunks;
    list_len[c] = (uint32_t)len;
    chunks += (len + 31) / 32;
  }
  chunk_start[nlist] = (uint32_t)chunks;
  if (chunks * 32 >= 0xFFFFFFFFull) { delete ivf; return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "ivf_create: more than 2^32 padded slots"); }
  ivf->total_chunks = chunks;
  std::vector<uint32_t> slot_pid(std::max<uint64_t>(chunks * 32, 1), MGPU_EMPTY_SLOT);
  for (uint32_t c = 0; c < nlist; c++) {
    uint64_t b = list_offsets[c], len = list_len[c];
    for (uint64_t i = 0; i < len; i++) {
      uint32_t pid = list_point_ids[b + i];
      if (pid >= n) { delete ivf; return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_create: point id %u out of range", pid); }
      slot_pid[(uint64_t)chunk_start[c] * 32 + i] = pid;
    }
  }
  ivf->h_list_len = list_len;
  int s = dev_alloc_copy(ctx, &ivf->d_centroids, centroids, (size_t)nlist * dim);
  if (s == MGPU_OK && dim >= 16 && nlist % 4 == 0) {
    // tensor-core coarse scoring operands: bf16 (hi, lo) split of the centroids + their squared norms
    double mx = 0.0;
    for (uint32_t c = 0; c < nlist; c++) {
      double a = 0.0;
      for (uint32_t d = 0; d < dim; d++) a += (double)centroids[(size_t)c * dim + d] * centroids[(size_t)c * dim + d];
      mx = std::max(mx, a);
    }
    ivf->cn_max = (float)(mx * 1.000001);
    s = dev_alloc_copy<uint16_t>(ctx, (uint16_t **)&ivf->d_csplit, nullptr, (size_t)nlist * coarse_tc_kp(dim));
    if (s == MGPU_OK) s = dev_alloc_copy<float>(ctx, &ivf->d_cn, nullptr, nlist);
    if (s == MGPU_OK) s = launch_split_bf16(ctx, ivf->d_centroids, nlist, dim, 1, ivf->d_csplit, ivf->d_cn);
  }
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &ivf->d_chunk_start, chunk_start.data(), nlist + 1);
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &ivf->d_list_len, list_len.data(), nlist);
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &ivf->d_slot_pid, slot_pid.data(), slot_pid.size());
  if (s == MGPU_OK && doc_ids) s = dev_alloc_copy(ctx, &ivf->d_doc_ids, doc_ids, n);
  if (s == MGPU_OK) {
    s = dev_alloc_copy<uint32_t>(ctx, &ivf->d_invalid, nullptr, (n + 31) / 32 + 1);
    if (s == MGPU_OK && cudaMemsetAsync(ivf->d_invalid, 0, ((n + 31) / 32 + 1) * 4, ctx->stream) != cudaSuccess) s = MGPU_ERR_CUDA;
  }
  if (s == MGPU_OK) {
    s = dev_alloc_copy<unsigned long long>(ctx, &ivf->d_scan_rows, nullptr, 2);
    if (s == MGPU_OK) cudaMemsetAsync(ivf->d_scan_rows, 0, 16, ctx->stream);
  }
  // rows -> chunked layout
  void *d_src = nullptr;
  bool own_src = false;
  size_t row_bytes = quant == MGPU_QUANT_PQ ? pq->m : (size_t)dim * 4;
  if (s == MGPU_OK && chunks) {
    if (rows_mem == MGPU_DEVICE) d_src = const_cast<void *>(rows);
    else {
      if (cudaMalloc(&d_src, std::max<size_t>(n * row_bytes, 16)) != cudaSuccess) s = mgpu_fail(ctx, MGPU_ERR_OOM, "ivf_create: staging %zu bytes of rows failed", n * row_bytes);
      else { own_src = true; if (cudaMemcpyAsync(d_src, rows, n * row_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) s = MGPU_ERR_CUDA; }
    }
  }
  if (s == MGPU_OK) {
    size_t nslots = std::max<uint64_t>(chunks * 32, 1);
    if (quant == MGPU_QUANT_PQ) {
      ivf->pq_fast = pq->nbits == 8 && pq->m % 32 == 0 && pq->m / 32 <= 4;
      ivf->ng = ivf->pq_fast ? pq->m / 32 : 0;
      ivf->bytes_per_row = pq->m + 4;
      s = dev_alloc_copy<uint8_t>(ctx, &ivf->d_codes, nullptr,
                                  ivf->pq_fast ? std::max<uint64_t>(chunks, 1) * pq_fast_chunk_bytes(ivf->ng) : nslots * pq->m);
    } else {
      ivf->bytes_per_row = (uint64_t)dim * 4 + 4;
      s = dev_alloc_copy<float>(ctx, &ivf->d_rows, nullptr, nslots * ivf->dim4 * 4);
    }
  }
  if (s == MGPU_OK && chunks) s = launch_build_layout(ivf, d_src);
  if (s == MGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    s = mgpu_fail(ctx, MGPU_ERR_CUDA, "ivf_create: layout build failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (own_src) cudaFree(d_src);
  if (s != MGPU_OK) { mgpu_ivf_destroy(ivf); return s; }
  *out = ivf;
  return MGPU_OK;
}

void mgpu_ivf_destroy(mgpu_ivf *ivf) {
  if (!ivf) return;
  cudaSetDevice(ivf->ctx->device);
  cudaStreamSynchronize(ivf->ctx->stream);
  cudaFree(ivf->d_centroids); cudaFree(ivf->d_csplit); cudaFree(ivf->d_cn); cudaFree(ivf->d_chunk_start); cudaFree(ivf->d_list_len); cudaFree(ivf->d_slot_pid);
  cudaFree(ivf->d_codes); cudaFree(ivf->d_rows); cudaFree(ivf->d_doc_ids); cudaFree(ivf->d_invalid); cudaFree(ivf->d_scan_rows);
  cudaFree(ivf->d_scan_overflow); cudaFree(ivf->d_pid_slot); cudaFree(ivf->d_qstate);
  ivf_free_doc_map(ivf);
  delete ivf;
}

uint64_t mgpu_ivf_num_vectors(mgpu_ivf *ivf) { return ivf ? ivf->n : 0; }
uint32_t mgpu_ivf_num_clusters(mgpu_ivf *ivf) { return ivf ? ivf->nlist : 0; }

__global__ void k_set_bits(uint32_t *bitmap, const uint32_t *ids, uint32_t n, uint64_t limit) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ids[i] < limit) atomicOr(&bitmap[ids[i] >> 5], 1u << (ids[i] & 31));
}

int mgpu_ivf_invalidate(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n) {
  if (!ivf || (n && !point_ids)) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  MGPU_TRY(mgpu_ws_reserve(ctx, (size_t)n * 4));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ws, point_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  {
    LaunchScope ls(ctx, MGPU_K_OTHER);
    k_set_bits<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ivf->d_invalid, (const uint32_t *)ctx->ws, n, ivf->n);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ivf->n_invalid += n;
  return MGPU_OK;
}

int mgpu_ivf_is_invalidated(mgpu_ivf *ivf, uint32_t point_id, int *out) {
  if (!ivf || !out) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  if (point_id >= ivf->n) { *out = 0; return MGPU_OK; }
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  uint32_t w = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&w, ivf->d_invalid + (point_id >> 5), 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = (w >> (point_id & 31)) & 1;
  return MGPU_OK;
}

uint64_t mgpu_ivf_last_scan_rows(mgpu_ivf *ivf) {
  if (!ivf) return 0;
  std::lock_guard<std::mutex> g(ivf->ctx->mu);
  cudaSetDevice(ivf->ctx->device);
  unsigned long long v = 0;
  cudaMemcpyAsync(&v, ivf->d_scan_rows, 8, cudaMemcpyDeviceToHost, ivf->ctx->stream);
  cudaStreamSynchronize(ivf->ctx->stream);
  return v;
}
uint64_t mgpu_ivf_last_scan_fallbacks(mgpu_ivf *ivf) {
  if (!ivf || !ivf->d_scan_overflow || !ivf->last_scan_was16) return 0;
  std::lock_guard<std::mutex> g(ivf->ctx->mu);
  cudaSetDevice(ivf->ctx->device);
  uint32_t v = 0;
  cudaMemcpyAsync(&v, ivf->d_scan_overflow, 4, cudaMemcpyDeviceToHost, ivf->ctx->stream);
  cudaStreamSynchronize(ivf->ctx->stream);
  return v;
}
uint64_t mgpu_ivf_last_scan_bytes(mgpu_ivf *ivf) { return ivf ? mgpu_ivf_last_scan_rows(ivf) * ivf->bytes_per_row : 0; }

// coarse scoring on device buffers: dQ (B x dim) -> d_ids (B x nprobe), d_dist (optional)
static size_t ivf_coarse_extra_ws(mgpu_ivf *ivf, uint32_t B) {  // query split + norms + overflow counter
  return ws_need(ws_need(ws_need(ws_need(0, (size_t)B * coarse_tc_kp(ivf->dim) * 2), (size_t)B * 4), 16), (size_t)B * 4) + 256;
}

// need_order: the caller wants the probes nearest-first (mgpu_ivf_coarse); a search only needs the probe SET
// fork_ev (optional): recorded at the point after which work that does not depend on the probes may run concurrently
// d_work / work_done (optional): per-query chunk counts of the chosen lists, when the selection kernel can emit them
static int ivf_coarse_dev(mgpu_ivf *ivf, const float *dQ, uint32_t B, uint32_t nprobe, float *dD, uint32_t *d_ids, float *d_dist,
                          void *extra_ws, int need_order, cudaEvent_t fork_ev = nullptr, uint32_t *d_work = nullptr,
                          bool *work_done = nullptr) {
  mgpu_ctx *ctx = ivf->ctx;
  if (work_done) *work_done = false;
  if (ivf->d_csplit && extra_ws && coarse_tc_applicable(ctx, ivf->dim, ivf->nlist, nprobe)) {
    // tensor-core pass + exact re-score of a provably sufficient candidate set: identical probes to the exact path
    WsAlloc w(extra_ws, ivf_coarse_extra_ws(ivf, B));
    void *qsplit = w.get<uint16_t>((size_t)B * coarse_tc_kp(ivf->dim));
    float *qn = w.get<float>(B);
    uint32_t *ovf = w.get<uint32_t>(4);
    uint32_t *flags = w.get<uint32_t>(B);
    return launch_coarse_tc(ctx, dQ, B, ivf->d_centroids, ivf->d_csplit, ivf->d_cn, ivf->cn_max, ivf->nlist, ivf->dim, nprobe, qsplit,
                            qn, dD, ovf, flags, need_order, d_ids, d_dist, ivf->d_chunk_start, d_work, work_done, fork_ev);
  }
  if (fork_ev) CUDA_TRY(ctx, cudaEventRecord(fork_ev, ctx->stream));
  // always the L2 calculator with sqrt (index.rs:155), whatever the quantizer's metric
  MGPU_TRY(launch_distance_matrix(ctx, dQ, B, ivf->d_centroids, ivf->nlist, ivf->dim, MGPU_L2, 1, dD, MGPU_K_COARSE));
  return launch_select_smallest(ctx, dD, B, ivf->nlist, nprobe, d_ids, d_dist);
}

int mgpu_ivf_coarse(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t nprobe, uint32_t *out_ids, float *out_dist, int mem) {
  if (!ivf || (B && (!Q || !out_ids))) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = ivf->ctx;
  if (nprobe == 0 || nprobe > ivf->nlist) return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "num_probes %u out of range 1..%u (the reference panics)", nprobe, ivf->nlist);
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  size_t bQ = (size_t)B * ivf->dim * 4, bD = (size_t)B * ivf->nlist * 4, bI = (size_t)B * nprobe * 4;
  size_t need = ws_need(ws_need(ws_need(ws_need(ws_need(0, bQ), bD), bI), bI), ivf_coarse_extra_ws(ivf, B));
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  float *sQ = w.get<float>((size_t)B * ivf->dim);
  float *dD = w.get<float>((size_t)B * ivf->nlist);
  uint32_t *sI = w.get<uint32_t>((size_t)B * nprobe);
  float *sV = w.get<float>((size_t)B * nprobe);
  uint8_t *xws = w.get<uint8_t>(ivf_coarse_extra_ws(ivf, B));
  const void *dQ;
  MGPU_TRY(stage_in(ctx, Q, bQ, mem, sQ, &dQ));
  uint32_t *dI = mem == MGPU_DEVICE ? out_ids : sI;
  float *dV = mem == MGPU_DEVICE ? out_dist : (out_dist ? sV : nullptr);
  MGPU_TRY(ivf_coarse_dev(ivf, (const float *)dQ, B, nprobe, dD, dI, dV, xws, 1));
  MGPU_TRY(stage_out(ctx, out_ids, dI, bI, mem));
  MGPU_TRY(stage_out(ctx, out_dist, dV, bI, mem));
  if (mem == MGPU_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

// scan + finalize on device buffers
static int ivf_scan_dev(mgpu_ivf *ivf, const float *dQ, uint32_t B, const uint32_t *d_probes, uint32_t max_probes,
                        const uint32_t *d_counts, uint32_t k, uint8_t *d_qcodes, uint64_t *d_ckey, uint32_t *d_cslot,
                        uint32_t *d_order, uint32_t *d_out_pids, mgpu_u128 *d_out_docs, float *d_out_scores, uint32_t *d_out_counts,
                        bool qcodes_on_aux = false, const uint32_t *d_filter = nullptr, uint64_t filter_stride = 0,
                        bool have_work = false, bool have_codes = false) {
  mgpu_ctx *ctx = ivf->ctx;
  CUDA_TRY(ctx, cudaMemsetAsync(ivf->d_scan_rows, 0, 8, ctx->stream));
  ScanArgs a;
  memset(&a, 0, sizeof(a));
  a.chunk_start = ivf->d_chunk_start; a.list_len = ivf->d_list_len; a.slot_pid = ivf->d_slot_pid;
  a.invalid = ivf->n_invalid ? ivf->d_invalid : nullptr;
  a.codes = ivf->d_codes; a.rows = ivf->d_rows; a.dim = ivf->dim; a.dim4 = ivf->dim4; a.ng = ivf->ng;
  a.Q = dQ; a.B = B; a.probes = d_probes; a.max_probes = max_probes; a.probe_counts = d_counts;
  a.cand_key = d_ckey; a.cand_slot = d_cslot; a.rows_scanned = ivf->d_scan_rows; a.next_query = (unsigned int *)(ivf->d_scan_rows + 1);
  a.metric = ivf->metric;
  a.filter = d_filter; a.filter_stride = filter_stride;
  if (ivf->quant == MGPU_QUANT_PQ) {
    mgpu_pq *pq = ivf->pq;
    a.m = pq->m; a.K = pq->K; a.table = pq->d_table; a.rowmin = pq->d_rowmin; a.rowmax = pq->d_rowmax;
    // the query is quantized with the same codebook (index.rs:193) -- once per query here, not once per list
    if (have_codes) {   // the caller already holds the query codes (sharded search: encoded once across the ranks) ...
      if (ctx->ext_codes_on_aux) {   // ... or is still gathering them on the side stream
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        ctx->ext_codes_on_aux = false;
      }
    }
    else if (qcodes_on_aux) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // encoded on the side stream
    else MGPU_TRY(launch_pq_quantize(pq, dQ, B, d_qcodes));
    a.qcodes = d_qcodes;
  }
  // longest-first query schedule for the persistent scan CTAs (MGPU_PLAN=0 disables it)
  static const bool use_plan = !(getenv("MGPU_PLAN") && getenv("MGPU_PLAN")[0] == '0');
  if (use_plan && d_order && B > 2 * (uint32_t)ctx->sm_count) {
    MGPU_TRY(launch_plan_queries(ivf, d_probes, max_probes, d_counts, B, d_order, d_order + B, have_work));
    a.order = d_order;
  }
  FinalizeArgs f;
  memset(&f, 0, sizeof(f));
  f.cand_key = d_ckey; f.cand_slot = d_cslot; f.B = B; f.k = k; f.slot_pid = ivf->d_slot_pid; f.metric = ivf->metric;
  // PQ ranks by a fixed-point key and re-scores exactly, so it wants spare candidates beyond k: one round gives 32 - k of
  // them; from k > 16 on the multi-round path keeps the same MGPU_ROUND_SPARE margin as for k > 32
  if (k > MGPU_NCAND || (ivf->quant == MGPU_QUANT_PQ && k > MGPU_NCAND - MGPU_ROUND_SPARE)) {
    // ---- multi-round top-k.  Every round is the ordinary scan restricted to rows whose composite (ranking key, point id)
    // is not below the bound left by the previous round, followed by the exact re-score of its (up to) 31 newly reported
    // candidates; the rounds' exact results are merged with the reference's ordering.  Rounds run until every query has
    // either seen a short round (all of its rows reported) or holds k + MGPU_ROUND_SPARE reported candidates: a round that
    // ends inside a group of equal composites (a point living in several probed lists) reports fewer than 31, so the count
    // is tracked per query on the device instead of assuming 31 per round.  The spare candidates cover swaps between the
    // fixed-point ranking and the exact scores at the k boundary.
    if (k > MGPU_MAX_K) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", k, MGPU_MAX_K);
    const uint32_t want = k + MGPU_ROUND_SPARE;
    const uint32_t R0 = (want + 30) / 31;
    // hard cap: twice the nominal count (covers every point sitting in up to 16 probed lists), bounded by what the merge
    // kernel can hold in shared memory
    const uint32_t Rcap = std::min<uint32_t>(2 * R0 + 2, (uint32_t)(ctx->smem_optin / (MGPU_NCAND * 28)));
    const size_t per = (size_t)B * MGPU_NCAND;
    size_t need = 0;
    need = ws_need(need, (size_t)B * 8); need = ws_need(need, Rcap * per * 4); need = ws_need(need, Rcap * per * 4);
    need = ws_need(need, (size_t)Rcap * B * 4); need = ws_need(need, (size_t)B * 4); need = ws_need(need, 16);
    uint8_t *priv = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync((void **)&priv, need + 256, ctx->stream));
    WsAlloc pw(priv, need + 256);
    uint64_t *lb = pw.get<uint64_t>(B);
    uint32_t *rP = pw.get<uint32_t>(Rcap * per); float *rS = pw.get<float>(Rcap * per); uint32_t *rC = pw.get<uint32_t>((size_t)Rcap * B);
    uint32_t *reported = pw.get<uint32_t>(B); uint32_t *unfinished = pw.get<uint32_t>(4);
    int st = MGPU_OK;
    if (cudaMemsetAsync(lb, 0, (size_t)B * 8, ctx->stream) != cudaSuccess) st = mgpu_fail(ctx, MGPU_ERR_CUDA, "memset failed");
    if (st == MGPU_OK && cudaMemsetAsync(reported, 0, (size_t)B * 4, ctx->stream) != cudaSuccess) st = mgpu_fail(ctx, MGPU_ERR_CUDA, "memset failed");
    a.lower_bound = lb;
    if (ivf->quant == MGPU_QUANT_PQ) {
      f.cb = ivf->pq->d_cb; f.codes = ivf->d_codes; f.qcodes = d_qcodes; f.m = ivf->pq->m; f.K = ivf->pq->K;
      f.dsub = ivf->pq->dsub; f.ng = ivf->ng; f.pq_fast = ivf->pq_fast;
    }
    f.doc_ids = nullptr; f.k = MGPU_NCAND; f.prune = false;
    uint32_t r = 0, target = R0;
    while (st == MGPU_OK) {
      for (; r < target && st == MGPU_OK; r++) {
        st = launch_scan(ivf, a);
        if (st == MGPU_OK && cudaMemsetAsync(unfinished, 0, 4, ctx->stream) != cudaSuccess) st = mgpu_fail(ctx, MGPU_ERR_CUDA, "memset failed");
        if (st == MGPU_OK) st = launch_round_prepare(ctx, d_ckey, d_cslot, B, lb, reported, want, unfinished);
        f.out_docs = nullptr; f.out_pids = rP + r * per; f.out_scores = rS + r * per; f.out_counts = rC + (size_t)r * B;
        if (st == MGPU_OK) st = launch_finalize(ctx, f);
      }
      if (st != MGPU_OK) break;
      uint32_t left = 0;   // queries whose last round was full and that still hold fewer than `want` reported candidates
      if (cudaMemcpyAsync(&left, unfinished, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess) { st = mgpu_fail(ctx, MGPU_ERR_CUDA, "multi-round top-k: reading the round status failed"); break; }
      if (left == 0) break;
      if (r >= Rcap) {
        st = mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "multi-round top-k: %u rounds did not yield k + %d candidates for %u queries "
                                                   "(a point living in more than 16 probed lists?)", r, MGPU_ROUND_SPARE, left);
        break;
      }
      target = std::min(Rcap, r + std::max<uint32_t>(1, R0 / 8));
    }
    // the k smallest by (distance, point_id) over all rounds, then (remap variants) doc ids ordered by (score, doc_id)
    if (st == MGPU_OK) st = launch_merge_rounds(ctx, rP, rS, rC, r, B, k, ivf->d_doc_ids, d_out_pids, d_out_docs, d_out_scores, d_out_counts);
    cudaFreeAsync(priv, ctx->stream);
    return st;
  }
  static const bool no_prune = getenv("MGPU_FINALIZE_PRUNE") && getenv("MGPU_FINALIZE_PRUNE")[0] == '0';
  if (ivf->quant == MGPU_QUANT_PQ) {
    f.cb = ivf->pq->d_cb; f.codes = ivf->d_codes; f.qcodes = d_qcodes; f.m = ivf->pq->m; f.K = ivf->pq->K;
    f.dsub = ivf->pq->dsub; f.ng = ivf->ng; f.pq_fast = ivf->pq_fast;
    f.prune = ivf->metric == MGPU_L2 && !no_prune;
  }
  f.doc_ids = ivf->d_doc_ids;
  f.out_pids = d_out_pids; f.out_docs = d_out_docs; f.out_scores = d_out_scores; f.out_counts = d_out_counts;
  if (ivf->quant == MGPU_QUANT_PQ && scan_pq16_applicable(ivf, a)) {
    // ---- 16-bit-LUT scan (scan_pq16.cu) -> exact re-rank + certification -> exact fallback for whatever is left.  The
    // fallback launch is a no-op when every query was certified (the usual case); its list is filled on the device.
    if (!ivf->d_scan_overflow || ivf->scan_overflow_cap < B || !ivf->d_qstate) {
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ivf->d_scan_overflow); cudaFree(ivf->d_qstate);
      ivf->d_scan_overflow = nullptr; ivf->d_qstate = nullptr; ivf->scan_overflow_cap = 0;
      CUDA_TRY(ctx, cudaMalloc((void **)&ivf->d_scan_overflow, ((size_t)B + 1) * 4));
      CUDA_TRY(ctx, cudaMalloc((void **)&ivf->d_qstate, (size_t)B * 4));
      ivf->scan_overflow_cap = B;
    }
    a.overflow_count = ivf->d_scan_overflow;
    a.overflow_list = ivf->d_scan_overflow + 1;
    CUDA_TRY(ctx, cudaMemsetAsync(a.overflow_count, 0, 4, ctx->stream));
    ivf->last_scan_was16 = true;
    MGPU_TRY(launch_scan_pq16(ivf, a, ivf->d_qstate));
    static const float cert_slack = getenv("MGPU_CERT_SLACK") ? (float)atof(getenv("MGPU_CERT_SLACK")) : 1.0f;
    f.key16 = 1; f.gscale = ivf->pq->gscale; f.cert_slack = cert_slack; f.qstate = ivf->d_qstate;
    f.uncert_count = a.overflow_count; f.uncert_list = a.overflow_list;
    MGPU_TRY(launch_finalize(ctx, f));
    return launch_scan_pq_exact_list(ivf, a, f);
  }
  ivf->last_scan_was16 = false;
  MGPU_TRY(launch_scan(ivf, a));
  return launch_finalize(ctx, f);
}

// slot of a pipelined search: lazily created streams/events, staging grown on demand; a slot whose previous ticket was never
// waited for is drained first (its results are already on their way to the caller's buffers)
static int pipe_prepare(mgpu_ctx *ctx, mgpu_ctx::Pipe *pp, size_t q_bytes, size_t out_bytes) {
  if (!ctx->h2d_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  if (!pp->ev_h2d) {
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&pp->ev_h2d, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&pp->ev_done, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&pp->ev_out, cudaEventDisableTiming));
  }
  if (pp->pending) { CUDA_TRY(ctx, cudaEventSynchronize(pp->ev_out)); pp->pending = false; }
  if (pp->q_bytes < q_bytes) {
    if (pp->q) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(pp->q); pp->q = nullptr; pp->q_bytes = 0; }
    CUDA_TRY(ctx, cudaMalloc(&pp->q, q_bytes));
    pp->q_bytes = q_bytes;
  }
  if (pp->out_bytes < out_bytes) {
    if (pp->out) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(pp->out); pp->out = nullptr; pp->out_bytes = 0; }
    CUDA_TRY(ctx, cudaMalloc(&pp->out, out_bytes));
    pp->out_bytes = out_bytes;
  }
  return MGPU_OK;
}

// first half of a pipelined host-buffer call: claims the slot, uploads Q on the H2D stream, makes the main stream wait for it
// copy_off / copy_bytes: the part of Q this call uploads (the sharded calls upload one rank's slice and all-gather the rest)
static int pipe_begin(mgpu_ctx *ctx, const void *Q, size_t q_bytes, size_t out_bytes, mgpu_ctx::Pipe **out_pp, size_t copy_off = 0,
                      size_t copy_bytes = (size_t)-1) {
  mgpu_ctx::Pipe *pp = &ctx->pipe[ctx->pipe_seq & 1];
  MGPU_TRY(pipe_prepare(ctx, pp, q_bytes, out_bytes));
  if (copy_bytes == (size_t)-1) copy_bytes = q_bytes;
  if (pp->used) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->h2d_stream, pp->ev_done, 0));  // the previous user of the buffer has read it
  if (copy_bytes) CUDA_TRY(ctx, cudaMemcpyAsync((char *)pp->q + copy_off, (const char *)Q + copy_off, copy_bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
  CUDA_TRY(ctx, cudaEventRecord(pp->ev_h2d, ctx->h2d_stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, pp->ev_h2d, 0));
  *out_pp = pp;
  return MGPU_OK;
}
// second half: results leave on the D2H stream; the ticket completes when they have landed in the caller's buffers
// results_ready (optional): the event after which the results are complete when they are produced off the main stream (the
// sharded calls' exchange stream); ev_done still marks the point after which the slot's query staging may be overwritten
static int pipe_end(mgpu_ctx *ctx, mgpu_ctx::Pipe *pp, uint32_t B, uint32_t k, const mgpu_u128 *dD, const float *dS, const uint32_t *dC,
                    mgpu_u128 *out_docs, float *out_scores, uint32_t *out_counts, uint64_t *ticket, cudaEvent_t results_ready = nullptr) {
  CUDA_TRY(ctx, cudaEventRecord(pp->ev_done, ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, results_ready ? results_ready : pp->ev_done, 0));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_docs, dD, (size_t)B * k * 16, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_scores, dS, (size_t)B * k * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_counts, dC, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  CUDA_TRY(ctx, cudaEventRecord(pp->ev_out, ctx->d2h_stream));
  pp->seq = ++ctx->pipe_seq; pp->pending = true; pp->used = true;
  *ticket = pp->seq;
  return MGPU_OK;
}

// shared implementation of scan / scan_remap / search
static int ivf_search_impl(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                           const uint32_t *probe_counts, uint32_t nprobe_coarse, uint32_t k, uint32_t *out_pids,
                           mgpu_u128 *out_docs, float *out_scores, uint32_t *out_counts, int mem,
                           const uint32_t *filter_bits = nullptr, uint64_t filter_stride = 0, uint64_t *ticket = nullptr,
                           const uint8_t *d_qcodes_ext = nullptr, bool caller_holds_lock = false) {
  mgpu_ctx *ctx = ivf->ctx;
  const bool do_coarse = probe_ids == nullptr;
  if (ticket) *ticket = 0;
  if (ticket && (mem != MGPU_HOST || !out_docs || out_pids)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "search_submit: host buffers and doc-id output only");
  if (do_coarse) {
    if (nprobe_coarse == 0 || nprobe_coarse > ivf->nlist)
      return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "num_probes %u out of range 1..%u (the reference panics)", nprobe_coarse, ivf->nlist);
    max_probes = nprobe_coarse;
  }
  if (k > MGPU_MAX_K) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", k, MGPU_MAX_K);
  if (B && (!Q || !out_scores || !out_counts || (!out_pids && !out_docs))) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "search: null buffer");
  std::unique_lock<std::mutex> g(ctx->mu, std::defer_lock);
  if (!caller_holds_lock) g.lock();
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  if (k == 0) {  // heap of capacity 0 keeps nothing (index.rs:265-274)
    if (mem == MGPU_HOST) memset(out_counts, 0, (size_t)B * 4);
    else CUDA_TRY(ctx, cudaMemsetAsync(out_counts, 0, (size_t)B * 4, ctx->stream));
    return MGPU_OK;
  }
  if (max_probes == 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "search: max_probes == 0");
  const uint64_t fwords = (ivf->n + 31) / 32;
  if (filter_bits && filter_stride != 0 && filter_stride < fwords)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "search: filter stride %llu < %llu words per bitmap", (unsigned long long)filter_stride, (unsigned long long)fwords);
  const size_t bF = filter_bits && mem == MGPU_HOST ? (filter_stride ? (size_t)B * filter_stride : (size_t)fwords) * 4 : 0;
  const uint32_t m = ivf->pq ? ivf->pq->m : 0;
  size_t bQ = (size_t)B * ivf->dim * 4, bP = (size_t)B * max_probes * 4;
  size_t need = 0;
  need = ws_need(need, bQ);                                   // Q
  need = ws_need(need, bP);                                   // probes
  need = ws_need(need, (size_t)B * 4);                        // probe counts
  need = ws_need(need, do_coarse ? (size_t)B * ivf->nlist * 4 : 0);  // distance matrix
  need = ws_need(need, (size_t)B * m);                        // query codes
  need = ws_need(need, (size_t)B * MGPU_NCAND * 8);           // cand keys
  need = ws_need(need, (size_t)B * MGPU_NCAND * 4);           // cand slots
  need = ws_need(need, (size_t)B * 8);                        // query schedule + work
  need = ws_need(need, (size_t)B * k * 16);                   // out docs
  need = ws_need(need, (size_t)B * k * 4);                    // out pids
  need = ws_need(need, (size_t)B * k * 4);                    // out scores
  need = ws_need(need, (size_t)B * 4);                        // out counts
  need = ws_need(need, do_coarse ? ivf_coarse_extra_ws(ivf, B) : 0);
  need = ws_need(need, bF);                                   // planner filter bitmaps
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  float *sQ = w.get<float>((size_t)B * ivf->dim);
  uint32_t *sP = w.get<uint32_t>((size_t)B * max_probes);
  uint32_t *sPC = w.get<uint32_t>(B);
  float *dD = w.get<float>(do_coarse ? (size_t)B * ivf->nlist : 0);
  uint8_t *dQC = w.get<uint8_t>((size_t)B * m);
  uint64_t *dCK = w.get<uint64_t>((size_t)B * MGPU_NCAND);
  uint32_t *dCS = w.get<uint32_t>((size_t)B * MGPU_NCAND);
  uint32_t *dOrd = w.get<uint32_t>((size_t)B * 2);
  mgpu_u128 *sDocs = w.get<mgpu_u128>((size_t)B * k);
  uint32_t *sPids = w.get<uint32_t>((size_t)B * k);
  float *sScores = w.get<float>((size_t)B * k);
  uint32_t *sCounts = w.get<uint32_t>(B);
  uint8_t *xws = w.get<uint8_t>(do_coarse ? ivf_coarse_extra_ws(ivf, B) : 0);
  uint32_t *sF = w.get<uint32_t>(bF / 4);

  const void *dQ;
  mgpu_ctx::Pipe *pp = nullptr;
  if (ticket) {
    // pipelined call: queries go through this slot's own staging buffer on the H2D stream
    MGPU_TRY(pipe_begin(ctx, Q, bQ, (size_t)B * k * 24 + (size_t)B * 4 + 768, &pp));
    dQ = pp->q;
  } else {
    MGPU_TRY(stage_in(ctx, Q, bQ, mem, sQ, &dQ));
  }
  const uint32_t *dF = filter_bits;
  if (bF) { const void *t; MGPU_TRY(stage_in(ctx, filter_bits, bF, mem, sF, &t)); dF = (const uint32_t *)t; }
  const uint32_t *dP, *dPC = nullptr;
  bool qcodes_on_aux = false, have_work = false;
  if (do_coarse) {
    // the query encode (index.rs:193) does not depend on the probes: it runs on the side stream next to the (latency
    // bound) selection kernel, forked right after the coarse GEMM -- the GEMM itself wants every SM's shared memory
    static const bool use_aux = !(getenv("MGPU_AUX_STREAM") && getenv("MGPU_AUX_STREAM")[0] == '0');
    const bool fork = use_aux && ivf->quant == MGPU_QUANT_PQ && !d_qcodes_ext;
    MGPU_TRY(ivf_coarse_dev(ivf, (const float *)dQ, B, nprobe_coarse, dD, sP, nullptr, xws, 0, fork ? ctx->ev_fork : nullptr,
                            dOrd + B, &have_work));
    if (fork) {
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
      MGPU_TRY(launch_pq_quantize(ivf->pq, (const float *)dQ, B, dQC, ctx->aux_stream));
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
      qcodes_on_aux = true;
    }
    dP = sP;
  } else {
    const void *t;
    MGPU_TRY(stage_in(ctx, probe_ids, bP, mem, sP, &t));
    dP = (const uint32_t *)t;
    if (probe_counts) { MGPU_TRY(stage_in(ctx, probe_counts, (size_t)B * 4, mem, sPC, &t)); dPC = (const uint32_t *)t; }
  }
  uint32_t *oP = out_pids ? (mem == MGPU_DEVICE ? out_pids : sPids) : nullptr;
  mgpu_u128 *oD = out_docs ? (mem == MGPU_DEVICE ? out_docs : sDocs) : nullptr;
  float *oS = mem == MGPU_DEVICE ? out_scores : sScores;
  uint32_t *oC = mem == MGPU_DEVICE ? out_counts : sCounts;
  if (pp) {
    WsAlloc wo(pp->out, pp->out_bytes);
    oD = wo.get<mgpu_u128>((size_t)B * k); oS = wo.get<float>((size_t)B * k); oC = wo.get<uint32_t>(B);
  }
  MGPU_TRY(ivf_scan_dev(ivf, (const float *)dQ, B, dP, max_probes, dPC, k, d_qcodes_ext ? (uint8_t *)d_qcodes_ext : dQC, dCK, dCS, dOrd, oP, oD,
                        oS, oC, qcodes_on_aux, dF, filter_stride, have_work, d_qcodes_ext != nullptr));
  if (pp) {
    return pipe_end(ctx, pp, B, k, oD, oS, oC, out_docs, out_scores, out_counts, ticket);
  }
  if (mem == MGPU_HOST) {
    MGPU_TRY(stage_out(ctx, out_pids, oP, (size_t)B * k * 4, mem));
    MGPU_TRY(stage_out(ctx, out_docs, oD, (size_t)B * k * 16, mem));
    MGPU_TRY(stage_out(ctx, out_scores, oS, (size_t)B * k * 4, mem));
    MGPU_TRY(stage_out(ctx, out_counts, oC, (size_t)B * 4, mem));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MGPU_OK;
}

static int check_probes_host(mgpu_ivf *ivf, const uint32_t *probe_ids, uint32_t B, uint32_t max_probes,
                             const uint32_t *probe_counts, int mem) {
  if (mem != MGPU_HOST || !probe_ids) return MGPU_OK;
  for (uint32_t b = 0; b < B; b++) {
    uint32_t np = probe_counts ? std::min(probe_counts[b], max_probes) : max_probes;
    for (uint32_t i = 0; i < np; i++)
      if (probe_ids[(size_t)b * max_probes + i] >= ivf->nlist)
        return mgpu_fail(ivf->ctx, MGPU_ERR_OUT_OF_RANGE, "centroid id %u out of range", probe_ids[(size_t)b * max_probes + i]);
  }
  return MGPU_OK;
}

int mgpu_ivf_scan(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                  const uint32_t *probe_counts, uint32_t k, uint32_t *out_point_ids, float *out_scores,
                  uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  if (B && !probe_ids) return mgpu_fail(ivf->ctx, MGPU_ERR_INVALID_ARG, "ivf_scan: null probe list");
  MGPU_TRY(check_probes_host(ivf, probe_ids, B, max_probes, probe_counts, mem));
  return ivf_search_impl(ivf, Q, B, probe_ids, max_probes, probe_counts, 0, k, out_point_ids, nullptr, out_scores, out_counts, mem);
}

int mgpu_ivf_scan_remap(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                        const uint32_t *probe_counts, uint32_t k, mgpu_u128 *out_doc_ids, float *out_scores,
                        uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  if (B && !probe_ids) return mgpu_fail(ivf->ctx, MGPU_ERR_INVALID_ARG, "ivf_scan_remap: null probe list");
  MGPU_TRY(check_probes_host(ivf, probe_ids, B, max_probes, probe_counts, mem));
  return ivf_search_impl(ivf, Q, B, probe_ids, max_probes, probe_counts, 0, k, nullptr, out_doc_ids, out_scores, out_counts, mem);
}

int mgpu_ivf_search(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, mgpu_u128 *out_doc_ids,
                    float *out_scores, uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  return ivf_search_impl(ivf, Q, B, nullptr, 0, nullptr, nprobe, k, nullptr, out_doc_ids, out_scores, out_counts, mem);
}

/* Pipelined BlockBasedIvf::search over host buffers: enqueue and return; mgpu_search_wait(ticket) completes it. */
int mgpu_ivf_search_submit(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, mgpu_u128 *out_doc_ids,
                           float *out_scores, uint32_t *out_counts, uint64_t *ticket) {
  if (!ivf || !ticket) return MGPU_ERR_INVALID_ARG;
  if (B == 0 || k == 0) {  // nothing to put in flight
    *ticket = 0;
    return ivf_search_impl(ivf, Q, B, nullptr, 0, nullptr, nprobe, k, nullptr, out_doc_ids, out_scores, out_counts, MGPU_HOST);
  }
  return ivf_search_impl(ivf, Q, B, nullptr, 0, nullptr, nprobe, k, nullptr, out_doc_ids, out_scores, out_counts, MGPU_HOST, nullptr, 0,
                         ticket);
}

int mgpu_search_wait(mgpu_ctx *ctx, uint64_t ticket) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  if (ticket == 0) return MGPU_OK;
  cudaEvent_t ev = nullptr;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ticket > ctx->pipe_seq) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "search_wait: unknown ticket %llu", (unsigned long long)ticket);
    mgpu_ctx::Pipe &pp = ctx->pipe[(ticket - 1) & 1];
    if (pp.seq != ticket || !pp.pending) return MGPU_OK;  // already drained (waited for, or its slot was reused)
    ev = pp.ev_out;
  }
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaEventSynchronize(ev));  // not under the lock: other threads keep submitting
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    mgpu_ctx::Pipe &pp = ctx->pipe[(ticket - 1) & 1];
    if (pp.seq == ticket) pp.pending = false;
  }
  return MGPU_OK;
}

/* Planner filter hook (index.rs:212-226): BlockBasedIvf::search with Some(planner). */
int mgpu_ivf_search_filtered(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, const uint32_t *filter_bits,
                             uint64_t filter_stride_words, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  return ivf_search_impl(ivf, Q, B, nullptr, 0, nullptr, nprobe, k, nullptr, out_doc_ids, out_scores, out_counts, mem, filter_bits,
                         filter_stride_words);
}

int mgpu_ivf_scan_remap_filtered(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                                 const uint32_t *probe_counts, uint32_t k, const uint32_t *filter_bits, uint64_t filter_stride_words,
                                 mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  if (B && !probe_ids) return mgpu_fail(ivf->ctx, MGPU_ERR_INVALID_ARG, "ivf_scan_remap_filtered: null probe list");
  MGPU_TRY(check_probes_host(ivf, probe_ids, B, max_probes, probe_counts, mem));
  return ivf_search_impl(ivf, Q, B, probe_ids, max_probes, probe_counts, 0, k, nullptr, out_doc_ids, out_scores, out_counts, mem,
                         filter_bits, filter_stride_words);
}

// ---- build-time assignment -----------------------------------------------------------------------------------------------
int mgpu_ivf_assign(mgpu_ctx *ctx, const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim,
                    uint32_t max_clusters, float threshold, uint32_t *out_cids, uint32_t *out_counts, int mem) {
  if (!ctx || (n && (!X || !centroids || !out_cids || !out_counts))) return MGPU_ERR_INVALID_ARG;
  if (max_clusters == 0 || max_clusters > nlist) return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "max_clusters_per_vector %u out of range 1..%u (the reference panics)", max_clusters, nlist);
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (n == 0) return MGPU_OK;
  const uint32_t r = max_clusters;
  // slabs of rows so the n x nlist matrix stays bounded
  uint64_t slab = std::max<uint64_t>(32, std::min<uint64_t>(n, (512ull << 20) / ((uint64_t)nlist * 4)));
  slab = std::min<uint64_t>(slab, 65535ull * 32);
  size_t need = 0;
  need = ws_need(need, (size_t)nlist * dim * 4);
  need = ws_need(need, slab * dim * 4);
  need = ws_need(need, slab * nlist * 4);
  need = ws_need(need, slab * r * 4);
  need = ws_need(need, slab * r * 4);
  need = ws_need(need, slab * r * 4);
  need = ws_need(need, slab * 4);
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  float *sC = w.get<float>((size_t)nlist * dim);
  float *sX = w.get<float>(slab * dim);
  float *dD = w.get<float>(slab * nlist);
  uint32_t *dSelI = w.get<uint32_t>(slab * r);
  float *dSelV = w.get<float>(slab * r);
  uint32_t *sCid = w.get<uint32_t>(slab * r);
  uint32_t *sCnt = w.get<uint32_t>(slab);
  const void *dC;
  MGPU_TRY(stage_in(ctx, centroids, (size_t)nlist * dim * 4, mem, sC, &dC));
  for (uint64_t i = 0; i < n; i += slab) {
    uint64_t cnt = std::min(slab, n - i);
    const void *dX;
    MGPU_TRY(stage_in(ctx, X + i * dim, cnt * dim * 4, mem, sX, &dX));
    MGPU_TRY(launch_distance_matrix(ctx, (const float *)dX, cnt, (const float *)dC, nlist, dim, MGPU_L2, 0, dD, MGPU_K_COARSE));
    MGPU_TRY(launch_select_smallest(ctx, dD, (uint32_t)cnt, nlist, r, dSelI, dSelV));
    uint32_t *oI = mem == MGPU_DEVICE ? out_cids + i * r : sCid;
    uint32_t *oN = mem == MGPU_DEVICE ? out_counts + i : sCnt;
    MGPU_TRY(launch_assign_filter(ctx, dSelI, dSelV, cnt, r, threshold, oI, oN));
    if (mem == MGPU_HOST) {
      CUDA_TRY(ctx, cudaMemcpyAsync(out_cids + i * r, oI, cnt * r * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaMemcpyAsync(out_counts + i, oN, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return MGPU_OK;
}

// ---- merge ----------------------------------------------------------------------------------------------------------------
int mgpu_merge_topk(mgpu_ctx *ctx, const mgpu_u128 *doc_ids, const float *scores, const uint32_t *counts, uint32_t S,
                    uint32_t B, uint32_t k, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  if (B && S && k && (!doc_ids || !scores || !counts || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "merge_topk: null buffer");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  if (k == 0 || S == 0) {
    if (mem == MGPU_HOST) memset(out_counts, 0, (size_t)B * 4);
    else CUDA_TRY(ctx, cudaMemsetAsync(out_counts, 0, (size_t)B * 4, ctx->stream));
    return MGPU_OK;
  }
  if (mem == MGPU_DEVICE) return launch_merge_topk(ctx, doc_ids, scores, counts, S, B, k, out_doc_ids, out_scores, out_counts);
  size_t nin = (size_t)S * B * k, nout = (size_t)B * k;
  size_t need = 0;
  need = ws_need(need, nin * 16); need = ws_need(need, nin * 4); need = ws_need(need, (size_t)S * B * 4);
  need = ws_need(need, nout * 16); need = ws_need(need, nout * 4); need = ws_need(need, (size_t)B * 4);
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  mgpu_u128 *dD = w.get<mgpu_u128>(nin); float *dS = w.get<float>(nin); uint32_t *dC = w.get<uint32_t>((size_t)S * B);
  mgpu_u128 *oD = w.get<mgpu_u128>(nout); float *oS = w.get<float>(nout); uint32_t *oC = w.get<uint32_t>(B);
  CUDA_TRY(ctx, cudaMemcpyAsync(dD, doc_ids, nin * 16, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(dS, scores, nin * 4, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(dC, counts, (size_t)S * B * 4, cudaMemcpyHostToDevice, ctx->stream));
  MGPU_TRY(launch_merge_topk(ctx, dD, dS, dC, S, B, k, oD, oS, oC));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_doc_ids, oD, nout * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_scores, oS, nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_counts, oC, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

// ---- NCCL (dlopen'ed; the library has no link-time dependency on it) ------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_ncclGetUniqueId)(nccl_uid_t *);
typedef int (*fn_ncclCommInitRank)(void **, int, nccl_uid_t, int);
typedef int (*fn_ncclCommDestroy)(void *);
typedef int (*fn_ncclAllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_ncclGetErrorString)(int);
typedef int (*fn_ncclGroup)(void);

static void *g_nccl = nullptr;
static void *nccl_sym(const char *name) {
  if (!g_nccl) {
    // a process that already imported torch has its bundled libnccl mapped: prefer that copy
    const char *cands[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; cands[i] && !g_nccl; i++) g_nccl = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl) {
      const char *env = getenv("MGPU_NCCL_LIB");
      if (env) g_nccl = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!g_nccl) return nullptr;
  }
  return dlsym(g_nccl, name);
}

int mgpu_comm_unique_id(uint8_t out_id[128]) {
  auto f = (fn_ncclGetUniqueId)nccl_sym("ncclGetUniqueId");
  if (!f) return MGPU_ERR_NCCL;
  nccl_uid_t id;
  if (f(&id) != 0) return MGPU_ERR_NCCL;
  memcpy(out_id, id.internal, 128);
  return MGPU_OK;
}

typedef int (*fn_ncclCommSplit)(void *, int, int, void **, void *);

int mgpu_comm_init(mgpu_ctx *ctx, int nranks, int rank, const uint8_t id[128]) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return MGPU_ERR_INVALID_ARG;
  auto f = (fn_ncclCommInitRank)nccl_sym("ncclCommInitRank");
  if (!f) return mgpu_fail(ctx, MGPU_ERR_NCCL, "libnccl not found (set MGPU_NCCL_LIB)");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  nccl_uid_t uid;
  memcpy(uid.internal, id, 128);
  int r = f(&ctx->nccl_comm, nranks, uid, rank);
  if (r != 0) {
    auto es = (fn_ncclGetErrorString)nccl_sym("ncclGetErrorString");
    return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclCommInitRank failed: %s", es ? es(r) : "?");
  }
  ctx->nranks = nranks; ctx->rank = rank;
  // a second communicator for the result exchange, so that it can run on its own stream next to the query-code all-gather
  // of the following batch (collectives of ONE communicator are ordered, whatever stream they are given)
  static const bool no_split = getenv("MGPU_COMM_SPLIT") && getenv("MGPU_COMM_SPLIT")[0] == '0';
  auto split = (fn_ncclCommSplit)nccl_sym("ncclCommSplit");
  ctx->nccl_comm_x = nullptr;
  if (split && !no_split && split(ctx->nccl_comm, 0, rank, &ctx->nccl_comm_x, nullptr) != 0) ctx->nccl_comm_x = nullptr;
  if (!ctx->comm_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  return MGPU_OK;
}

int mgpu_comm_destroy(mgpu_ctx *ctx) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  auto f = (fn_ncclCommDestroy)nccl_sym("ncclCommDestroy");
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  if (ctx->nccl_comm_x) { if (f) f(ctx->nccl_comm_x); ctx->nccl_comm_x = nullptr; }
  if (ctx->nccl_comm) { if (f) f(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  return MGPU_OK;
}

int mgpu_shard_overlap(mgpu_ctx *ctx, int on) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->shard_overlap = on != 0;
  return MGPU_OK;
}

// ---- result exchange --------------------------------------------------------------------------------------------------------
// One slot of exchange buffers: this rank's local top-k (docs | scores | counts), the gathered lists of all ranks, and (host
// callers) the merged result before its D2H copy.
struct ShardBufs {
  mgpu_u128 *locD; float *locS; uint32_t *locC;
  mgpu_u128 *gD; float *gS; uint32_t *gC;
  mgpu_u128 *outD; float *outS; uint32_t *outC;
};

// Claims the next slot.  The main stream waits until the slot's previous exchange has consumed its buffers.
static int shard_slot_begin(mgpu_ctx *ctx, uint32_t B, uint32_t k, mgpu_ctx::ShardSlot **out_slot, ShardBufs *b) {
  if (!ctx->comm_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  mgpu_ctx::ShardSlot *sl = &ctx->shard_slot[ctx->shard_seq++ & 1];
  if (!sl->ev_local) {
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl->ev_local, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl->ev_xdone, cudaEventDisableTiming));
  }
  const uint32_t S = (uint32_t)ctx->nranks;
  const size_t nloc = (size_t)B * k;
  size_t need = 0;
  need = ws_need(need, nloc * 16); need = ws_need(need, nloc * 4); need = ws_need(need, (size_t)B * 4);
  need = ws_need(need, S * nloc * 16); need = ws_need(need, S * nloc * 4); need = ws_need(need, (size_t)S * B * 4);
  need = ws_need(need, nloc * 16); need = ws_need(need, nloc * 4); need = ws_need(need, (size_t)B * 4);
  need += 256;
  if (sl->used) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl->ev_xdone, 0));
  if (sl->bytes < need) {
    if (sl->buf) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream)); CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(sl->buf); sl->buf = nullptr; sl->bytes = 0; }
    CUDA_TRY(ctx, cudaMalloc(&sl->buf, need + need / 4));
    sl->bytes = need + need / 4;
  }
  WsAlloc w(sl->buf, sl->bytes);
  b->locD = w.get<mgpu_u128>(nloc); b->locS = w.get<float>(nloc); b->locC = w.get<uint32_t>(B);
  b->gD = w.get<mgpu_u128>(S * nloc); b->gS = w.get<float>(S * nloc); b->gC = w.get<uint32_t>((size_t)S * B);
  b->outD = w.get<mgpu_u128>(nloc); b->outS = w.get<float>(nloc); b->outC = w.get<uint32_t>(B);
  *out_slot = sl;
  return MGPU_OK;
}

// All-gather of the per-shard top-k over NVLink + merge (snapshot.rs:60-61,105-106), on the exchange stream: it starts when
// the local search of this batch is done and does not hold up the main stream.
// MGPU_HOST_PROF=1: host microseconds spent inside the NCCL enqueue calls and inside the whole sharded call, printed every 64
// calls (where does the host's time per step go at N = 8?)
struct HostProf {
  double nccl_us = 0, call_us = 0, pre_us = 0, local_us = 0, xchg_us = 0; uint64_t calls = 0;
  static bool on() { static const bool v = getenv("MGPU_HOST_PROF") && getenv("MGPU_HOST_PROF")[0] == '1'; return v; }
  static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};
static HostProf g_hostprof;

static int shard_exchange(mgpu_ctx *ctx, mgpu_ctx::ShardSlot *sl, const ShardBufs &b, const mgpu_u128 *locD, const float *locS,
                          const uint32_t *locC, uint32_t B, uint32_t k, mgpu_u128 *oD, float *oS, uint32_t *oC) {
  auto ag = (fn_ncclAllGather)nccl_sym("ncclAllGather");
  if (!ag) return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclAllGather not found");
  void *comm = ctx->nccl_comm_x ? ctx->nccl_comm_x : ctx->nccl_comm;
  const uint32_t S = (uint32_t)ctx->nranks;
  const size_t nloc = (size_t)B * k;
  CUDA_TRY(ctx, cudaEventRecord(sl->ev_local, ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, sl->ev_local, 0));
  // ncclChar = 0; three small all-gathers (B*k*20 + B*4 bytes per rank), grouped so that NCCL issues them as one launch (the
  // exchange is latency bound: three separate collectives cost three launch + handshake latencies)
  auto gs = (fn_ncclGroup)nccl_sym("ncclGroupStart");
  auto ge = (fn_ncclGroup)nccl_sym("ncclGroupEnd");
  const double hp0 = HostProf::on() ? HostProf::now() : 0;
  const bool grouped = gs && ge && gs() == 0;
  int r = ag(locD, b.gD, nloc * 16, 0, comm, ctx->comm_stream);
  if (r == 0) r = ag(locS, b.gS, nloc * 4, 0, comm, ctx->comm_stream);
  if (r == 0) r = ag(locC, b.gC, (size_t)B * 4, 0, comm, ctx->comm_stream);
  if (grouped) { const int r2 = ge(); if (r == 0) r = r2; }
  if (HostProf::on()) g_hostprof.nccl_us += HostProf::now() - hp0;
  if (r != 0) return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclAllGather failed (%d)", r);
  ctx->launches += 3;
  MGPU_TRY(launch_merge_topk(ctx, b.gD, b.gS, b.gC, S, B, k, oD, oS, oC, ctx->comm_stream));
  CUDA_TRY(ctx, cudaEventRecord(sl->ev_xdone, ctx->comm_stream));
  sl->used = true;
  return MGPU_OK;
}

int mgpu_shard_allgather_merge(mgpu_ctx *ctx, const mgpu_u128 *local_doc_ids, const float *local_scores,
                               const uint32_t *local_counts, uint32_t B, uint32_t k, mgpu_u128 *out_doc_ids,
                               float *out_scores, uint32_t *out_counts) {
  if (!ctx) return MGPU_ERR_INVALID_ARG;
  if (!ctx->nccl_comm) return mgpu_fail(ctx, MGPU_ERR_NCCL, "shard_allgather_merge: communicator not initialised");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0 || k == 0) return MGPU_OK;
  mgpu_ctx::ShardSlot *sl;
  ShardBufs b;
  MGPU_TRY(shard_slot_begin(ctx, B, k, &sl, &b));
  MGPU_TRY(shard_exchange(ctx, sl, b, local_doc_ids, local_scores, local_counts, B, k, out_doc_ids, out_scores, out_counts));
  if (!ctx->shard_overlap) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl->ev_xdone, 0));
  return MGPU_OK;
}

static int spann_search_impl(mgpu_spann *sp, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                             uint32_t num_explored_centroids, float ratio, const uint32_t *filter_bits, uint64_t filter_stride,
                             mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem, bool caller_holds_lock = false,
                             const uint8_t *d_qcodes_ext = nullptr);

// The whole sharded query path in one call: every rank passes the SAME replicated batch and searches its own
// doc-shard -- BlockBasedIvf::search (ivf != null) or Spann::search (sp != null, config 5); the per-shard top-k lists are
// all-gathered and merged with the leaf ordering (snapshot.rs:60-61,105-106), so every rank returns the merged result.  With
// a codebook shared by all shards (shared_codebook != 0) the query encode (index.rs:193) -- per-query work that does not
// shrink with the shard -- is split across the ranks: rank r encodes queries [r*ceil(B/N), (r+1)*ceil(B/N)) and one
// all-gather of B x m code bytes replaces N-1 redundant encodes per query.
static int shard_search_impl(mgpu_ivf *ivf, mgpu_spann *sp, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, uint32_t ef,
                             uint32_t num_explored, float ratio, int shared_codebook, mgpu_u128 *out_doc_ids, float *out_scores,
                             uint32_t *out_counts, int mem, uint64_t *ticket) {
  if (!ivf && !sp) return MGPU_ERR_INVALID_ARG;
  if (sp) ivf = sp->lists;
  mgpu_ctx *ctx = ivf->ctx;
  if (ticket) *ticket = 0;
  if (!ctx->nccl_comm) return mgpu_fail(ctx, MGPU_ERR_NCCL, "shard search: communicator not initialised (mgpu_comm_init)");
  if (B && (!Q || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "shard search: null buffer");
  if (!sp && (nprobe == 0 || nprobe > ivf->nlist)) return mgpu_fail(ctx, MGPU_ERR_OUT_OF_RANGE, "num_probes %u out of range 1..%u (the reference panics)", nprobe, ivf->nlist);
  if (k > MGPU_MAX_K) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", k, MGPU_MAX_K);
  auto ag = (fn_ncclAllGather)nccl_sym("ncclAllGather");
  if (!ag) return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclAllGather not found");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  struct CallTimer {
    double t0 = HostProf::on() ? HostProf::now() : 0;
    ~CallTimer() {
      if (!HostProf::on()) return;
      g_hostprof.call_us += HostProf::now() - t0;
      static const int win = getenv("MGPU_HOST_PROF_N") ? std::max(1, atoi(getenv("MGPU_HOST_PROF_N"))) : 64;
      if (++g_hostprof.calls % win == 0) {
        const double n = (double)win;
        fprintf(stderr, "[host prof] last window of sharded calls: %.1f us per call on the host (before the local search %.1f, local search %.1f, "
                        "exchange %.1f), %.1f us inside NCCL enqueues\n", g_hostprof.call_us / n, g_hostprof.pre_us / n, g_hostprof.local_us / n,
                g_hostprof.xchg_us / n, g_hostprof.nccl_us / n);
        g_hostprof.call_us = g_hostprof.pre_us = g_hostprof.local_us = g_hostprof.xchg_us = g_hostprof.nccl_us = 0;
      }
    }
  } call_timer;
  if (k == 0 && !sp) {
    if (mem == MGPU_HOST) memset(out_counts, 0, (size_t)B * 4);
    else CUDA_TRY(ctx, cudaMemsetAsync(out_counts, 0, (size_t)B * 4, ctx->stream));
    return MGPU_OK;
  }
  const uint32_t kk = std::max<uint32_t>(k, 1);   // Spann with top_k = 0 still reports None / empty per query
  const uint32_t N = (uint32_t)ctx->nranks, r = (uint32_t)ctx->rank;
  const bool split_encode = shared_codebook && ivf->quant == MGPU_QUANT_PQ && N > 1;
  const uint32_t m = ivf->pq ? ivf->pq->m : 0;
  const uint32_t slice = (B + N - 1) / N;
  size_t need = 0;
  // HOST buffers: every rank holds the same replicated batch, so rank r uploads only rows [r * slice, (r + 1) * slice) and
  // one all-gather over NVLink completes the batch on every GPU (N x less PCIe traffic per rank; at N = 8 the full upload of
  // 8192 x 768 floats per rank was what bounded the end-to-end rate)
  static const bool no_split_upload = getenv("MGPU_SPLIT_UPLOAD") && getenv("MGPU_SPLIT_UPLOAD")[0] == '0';
  const bool split_upload = mem == MGPU_HOST && N > 1 && !no_split_upload;
  const size_t row_bytes = (size_t)ivf->dim * 4, slice_bytes = (size_t)slice * row_bytes;
  const size_t q_alloc = split_upload ? (size_t)N * slice_bytes : (size_t)B * row_bytes;
  const uint32_t up_lo = std::min(B, r * slice), up_cnt = std::min(B, up_lo + slice) - up_lo;
  need = ws_need(need, mem == MGPU_HOST ? q_alloc : 0);
  need = ws_need(need, split_encode ? (size_t)N * slice * m : 0);
  need = ws_need(need, split_encode ? (size_t)slice * m : 0);
  if (ctx->shard_ws_bytes < need) {
    if (ctx->shard_ws) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->shard_ws); ctx->shard_ws = nullptr; ctx->shard_ws_bytes = 0; }
    CUDA_TRY(ctx, cudaMalloc(&ctx->shard_ws, need + need / 4 + 256));
    ctx->shard_ws_bytes = need + need / 4 + 256;
  }
  WsAlloc w(ctx->shard_ws, ctx->shard_ws_bytes);
  float *sQ = w.get<float>(mem == MGPU_HOST ? q_alloc / 4 : 0);
  uint8_t *codes_all = w.get<uint8_t>(split_encode ? (size_t)N * slice * m : 0);
  uint8_t *codes_mine = w.get<uint8_t>(split_encode ? (size_t)slice * m : 0);
  mgpu_ctx::ShardSlot *sl;
  ShardBufs b;
  MGPU_TRY(shard_slot_begin(ctx, B, kk, &sl, &b));
  const void *dQv;
  mgpu_ctx::Pipe *pp = nullptr;
  if (ticket) {   // pipelined: this slot's staging for the queries and the merged results, copies on the side streams
    if (split_upload) MGPU_TRY(pipe_begin(ctx, Q, q_alloc, (size_t)B * kk * 24 + (size_t)B * 4 + 768, &pp, (size_t)up_lo * row_bytes, (size_t)up_cnt * row_bytes));
    else MGPU_TRY(pipe_begin(ctx, Q, (size_t)B * row_bytes, (size_t)B * kk * 24 + (size_t)B * 4 + 768, &pp));
    dQv = pp->q;
  } else if (split_upload) {
    if (up_cnt) CUDA_TRY(ctx, cudaMemcpyAsync((char *)sQ + (size_t)up_lo * row_bytes, (const char *)Q + (size_t)up_lo * row_bytes, (size_t)up_cnt * row_bytes,
                                              cudaMemcpyHostToDevice, ctx->stream));
    dQv = sQ;
  } else {
    MGPU_TRY(stage_in(ctx, Q, (size_t)B * row_bytes, mem, sQ, &dQv));
  }
  const float *dQ = (const float *)dQv;
  if (split_upload) {   // in-place all-gather of the query slices (rank r's slice sits at r * slice_bytes)
    int rc = ag((const char *)dQv + (size_t)r * slice_bytes, (void *)dQv, slice_bytes, 0, ctx->nccl_comm, ctx->stream);
    if (rc != 0) return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclAllGather (query slices) failed (%d)", rc);
    ctx->launches += 1;
  }
  const uint8_t *ext = nullptr;
  if (split_encode) {
    // on the side stream: neither the coarse GEMM nor the selection needs the codes, only the scan does (it waits for ev_join)
    static const bool aux_off = getenv("MGPU_AUX_STREAM") && getenv("MGPU_AUX_STREAM")[0] == '0';
    cudaStream_t es = aux_off ? ctx->stream : ctx->aux_stream;
    if (!aux_off) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));          // the batch is on the device
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    }
    const uint32_t lo = std::min(B, r * slice), cnt = std::min(B, lo + slice) - lo;
    if (cnt) MGPU_TRY(launch_pq_quantize(ivf->pq, dQ + (size_t)lo * ivf->dim, cnt, codes_mine, es));
    const double hp0 = HostProf::on() ? HostProf::now() : 0;
    int rc = ag(codes_mine, codes_all, (size_t)slice * m, 0, ctx->nccl_comm, es);
    if (!aux_off && rc == 0) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
      ctx->ext_codes_on_aux = true;
    }
    if (HostProf::on()) g_hostprof.nccl_us += HostProf::now() - hp0;
    if (rc != 0) return mgpu_fail(ctx, MGPU_ERR_NCCL, "ncclAllGather (query codes) failed (%d)", rc);
    ctx->launches += 1;
    ext = codes_all;  // rank j's slice starts at j*slice*m = (first query of the slice)*m: query q sits at q*m
  }
  const double hp1 = HostProf::on() ? HostProf::now() : 0;
  if (HostProf::on()) g_hostprof.pre_us += hp1 - call_timer.t0;
  if (sp) MGPU_TRY(spann_search_impl(sp, dQ, B, k, ef, num_explored, ratio, nullptr, 0, b.locD, b.locS, b.locC, MGPU_DEVICE, true, ext));
  else MGPU_TRY(ivf_search_impl(ivf, dQ, B, nullptr, 0, nullptr, nprobe, k, nullptr, b.locD, b.locS, b.locC, MGPU_DEVICE, nullptr, 0, nullptr, ext, true));
  const double hp2 = HostProf::on() ? HostProf::now() : 0;
  if (HostProf::on()) g_hostprof.local_us += hp2 - hp1;
  mgpu_u128 *oD = mem == MGPU_DEVICE ? out_doc_ids : b.outD;
  float *oS = mem == MGPU_DEVICE ? out_scores : b.outS;
  uint32_t *oC = mem == MGPU_DEVICE ? out_counts : b.outC;
  if (pp) {
    WsAlloc wo(pp->out, pp->out_bytes);
    oD = wo.get<mgpu_u128>((size_t)B * kk); oS = wo.get<float>((size_t)B * kk); oC = wo.get<uint32_t>(B);
  }
  MGPU_TRY(shard_exchange(ctx, sl, b, b.locD, b.locS, b.locC, B, kk, oD, oS, oC));
  if (HostProf::on()) g_hostprof.xchg_us += HostProf::now() - hp2;
  if (pp) return pipe_end(ctx, pp, B, k, oD, oS, oC, out_doc_ids, out_scores, out_counts, ticket, sl->ev_xdone);
  if (mem == MGPU_DEVICE) {
    // stream-ordered contract: later work on the ctx stream sees the merged result -- unless the caller opted into overlap
    // (mgpu_shard_overlap), in which case the next batch's kernels run next to this exchange and mgpu_sync completes both
    if (!ctx->shard_overlap) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl->ev_xdone, 0));
    return MGPU_OK;
  }
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl->ev_xdone, 0));
  MGPU_TRY(stage_out(ctx, out_doc_ids, oD, (size_t)B * k * 16, mem));
  MGPU_TRY(stage_out(ctx, out_scores, oS, (size_t)B * k * 4, mem));
  MGPU_TRY(stage_out(ctx, out_counts, oC, (size_t)B * 4, mem));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

int mgpu_shard_ivf_search(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, int shared_codebook,
                          mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem) {
  if (!ivf) return MGPU_ERR_INVALID_ARG;
  return shard_search_impl(ivf, nullptr, Q, B, k, nprobe, 0, 0, 0.f, shared_codebook, out_doc_ids, out_scores, out_counts, mem, nullptr);
}

/* Pipelined form for page-locked host buffers (mgpu_search_wait completes it); collective like the blocking call. */
int mgpu_shard_ivf_search_submit(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, int shared_codebook,
                                 mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, uint64_t *ticket) {
  if (!ivf || !ticket) return MGPU_ERR_INVALID_ARG;
  const bool inflight = B != 0 && k != 0;
  if (!inflight) *ticket = 0;
  return shard_search_impl(ivf, nullptr, Q, B, k, nprobe, 0, 0, 0.f, shared_codebook, out_doc_ids, out_scores, out_counts, MGPU_HOST,
                           inflight ? ticket : nullptr);
}

/* Config 5: Spann::search per doc-shard + exchange. */
int mgpu_shard_spann_search(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids,
                            float centroid_distance_ratio, int shared_codebook, mgpu_u128 *out_doc_ids, float *out_scores,
                            uint32_t *out_counts, int mem) {
  if (!s) return MGPU_ERR_INVALID_ARG;
  return shard_search_impl(nullptr, s, Q, B, top_k, 0, ef, num_explored_centroids, centroid_distance_ratio, shared_codebook, out_doc_ids,
                           out_scores, out_counts, mem, nullptr);
}

int mgpu_shard_spann_search_submit(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                                   uint32_t num_explored_centroids, float centroid_distance_ratio, int shared_codebook,
                                   mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, uint64_t *ticket) {
  if (!s || !ticket) return MGPU_ERR_INVALID_ARG;
  const bool inflight = B != 0 && top_k != 0;
  if (!inflight) *ticket = 0;
  return shard_search_impl(nullptr, s, Q, B, top_k, 0, ef, num_explored_centroids, centroid_distance_ratio, shared_codebook, out_doc_ids,
                           out_scores, out_counts, MGPU_HOST, inflight ? ticket : nullptr);
}

}  // extern "C"

// ---- HNSW / SPANN ------------------------------------------------------------------------------------------------------------
extern "C" {

int mgpu_hnsw_create(mgpu_ctx *ctx, uint32_t dim, uint32_t num_layers, const uint32_t *edges, uint64_t n_edges,
                     const uint32_t *points, uint64_t n_points, const uint64_t *edge_offsets, uint64_t n_edge_offsets,
                     const uint64_t *level_offsets, int quant, int metric, mgpu_pq *pq, const void *rows, int rows_mem,
                     uint64_t n, const mgpu_u128 *doc_ids, mgpu_hnsw **out) {
  if (!ctx || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  if (dim == 0 || num_layers == 0 || !edge_offsets || !level_offsets || n_edge_offsets == 0 || (n && !rows))
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: null/zero argument");
  if (quant == MGPU_QUANT_PQ && (!pq || pq->dim != dim)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: PQ quantizer missing or of the wrong dimension");
  if (n_edges && !edges) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: null edges");
  // edge_offsets = one entry per upper-layer node, then n layer-0 entries, then one terminal entry (hnsw/writer.rs:100-140);
  // level_offsets[num_layers] is n_upper + n in our builders and n_upper + n + 1 in files written by the reference
  if (level_offsets[num_layers - 1] > n_points || level_offsets[num_layers - 1] >= n_edge_offsets ||
      n > n_edge_offsets - 1 - level_offsets[num_layers - 1])
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: level_offsets inconsistent with points/edge_offsets");
  for (uint32_t li = 0; li + 1 < num_layers; li++)
    if (level_offsets[li + 1] < level_offsets[li]) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: level_offsets not monotone");
  // the kernels index `edges` with these: every offset inside the edge array, ascending
  for (uint64_t i = 0; i < n_edge_offsets; i++)
    if (edge_offsets[i] > n_edges || (i && edge_offsets[i] < edge_offsets[i - 1]))
      return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: edge_offsets[%llu] = %llu is out of order or beyond the %llu edges",
                       (unsigned long long)i, (unsigned long long)edge_offsets[i], (unsigned long long)n_edges);
  for (uint64_t i = 0; i < n_edges; i++)
    if (edges[i] >= n) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_create: edge %llu points at %u >= %llu points", (unsigned long long)i, edges[i], (unsigned long long)n);
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  mgpu_hnsw *h = new mgpu_hnsw();
  h->ctx = ctx; h->dim = dim; h->num_layers = num_layers; h->n = n; h->n_edges = n_edges; h->n_points = n_points;
  h->n_edge_offsets = n_edge_offsets; h->quant = quant; h->metric = quant == MGPU_QUANT_PQ ? pq->metric : metric;
  h->pq = quant == MGPU_QUANT_PQ ? pq : nullptr;
  h->qdim = quant == MGPU_QUANT_PQ ? pq->m : dim;
  h->h_level_offsets.assign(level_offsets, level_offsets + num_layers + 1);
  // entry point (graph_storage.rs:527-554)
  h->entry_point = 0;
  if (num_layers == 1) {
    for (uint64_t i = 0; i + 1 < n_edge_offsets; i++) if (edge_offsets[i + 1] > edge_offsets[i]) { h->entry_point = (uint32_t)i; break; }
  } else if (n_points > level_offsets[0]) h->entry_point = points[level_offsets[0]];
  // per-layer sorted (point id -> first position) tables replace the reference's linear scan (graph_storage.rs:423-449)
  uint64_t n_upper = level_offsets[num_layers - 1];
  std::vector<uint32_t> spid(std::max<uint64_t>(n_upper, 1)), spos(std::max<uint64_t>(n_upper, 1));
  for (uint32_t li = 0; li + 1 < num_layers; li++) {
    uint64_t s = level_offsets[li], e = level_offsets[li + 1];
    std::vector<std::pair<uint32_t, uint32_t>> v;
    v.reserve(e - s);
    for (uint64_t i = s; i < e; i++) v.push_back({points[i], (uint32_t)i});
    std::stable_sort(v.begin(), v.end(), [](const std::pair<uint32_t, uint32_t> &x, const std::pair<uint32_t, uint32_t> &y) { return x.first < y.first; });
    for (uint64_t i = s; i < e; i++) { spid[i] = v[i - s].first; spos[i] = v[i - s].second; }
  }
  size_t row_bytes = quant == MGPU_QUANT_PQ ? pq->m : (size_t)dim * 4;
  // dense (layer, point) -> first position map: one load instead of a binary search per expansion on the upper layers,
  // which matter because the reference runs the full `ef` beam on every layer (index.rs:172-183)
  std::vector<int32_t> dense;
  if (num_layers > 1 && (uint64_t)(num_layers - 1) * n * 4 <= (1ull << 30)) {
    dense.assign((size_t)(num_layers - 1) * n, -1);
    for (uint32_t li = 0; li + 1 < num_layers; li++) {
      uint32_t layer = num_layers - 1 - li;
      for (uint64_t i = level_offsets[li]; i < level_offsets[li + 1]; i++) {
        uint32_t pid = points[i];
        if (pid < n && dense[(size_t)(layer - 1) * n + pid] < 0) dense[(size_t)(layer - 1) * n + pid] = (int32_t)i;
      }
    }
  }
  // fixed-stride copy of the layer-0 adjacency lists (layer 0 is addressed by point id, graph_storage.rs:465-478)
  std::vector<uint32_t> edges0;
  {
    const uint64_t l0 = level_offsets[num_layers - 1];
    uint64_t maxdeg = 0;
    bool ok = true;
    for (uint64_t p = 0; p < n && ok; p++) {
      if (edge_offsets[l0 + p + 1] < edge_offsets[l0 + p] || edge_offsets[l0 + p + 1] > n_edges) ok = false;
      else maxdeg = std::max<uint64_t>(maxdeg, edge_offsets[l0 + p + 1] - edge_offsets[l0 + p]);
    }
    const uint64_t stride = (maxdeg + 7) & ~7ull;
    if (ok && maxdeg > 0 && stride <= 64 && n * stride * 4 <= (4ull << 30)) {
      h->deg0 = (uint32_t)stride;
      edges0.assign((size_t)n * stride, 0xFFFFFFFFu);
      for (uint64_t p = 0; p < n; p++) {
        const uint64_t b = edge_offsets[l0 + p], e = edge_offsets[l0 + p + 1];
        for (uint64_t i = b; i < e; i++) edges0[(size_t)p * stride + (i - b)] = edges[i];
      }
    }
  }
  for (uint64_t i = 0; i + 1 < n_edge_offsets; i++) h->max_degree = std::max<uint32_t>(h->max_degree, (uint32_t)std::min<uint64_t>(edge_offsets[i + 1] - edge_offsets[i], 0xFFFFFFFFull));
  int s = dev_alloc_copy(ctx, &h->d_edges, edges, n_edges);
  if (s == MGPU_OK && !edges0.empty()) s = dev_alloc_copy(ctx, &h->d_edges0, edges0.data(), edges0.size());
  if (s == MGPU_OK && !dense.empty()) s = dev_alloc_copy(ctx, &h->d_upper_dense, dense.data(), dense.size());
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &h->d_points, points, n_points);
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &h->d_edge_offsets, edge_offsets, n_edge_offsets);
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &h->d_level_offsets, level_offsets, num_layers + 1);
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &h->d_upper_sorted_pid, spid.data(), spid.size());
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, &h->d_upper_sorted_pos, spos.data(), spos.size());
  if (s == MGPU_OK) s = dev_alloc_copy(ctx, (uint8_t **)&h->d_rows, (const uint8_t *)rows, n * row_bytes, rows_mem == MGPU_DEVICE);
  if (s == MGPU_OK && doc_ids) s = dev_alloc_copy(ctx, &h->d_doc_ids, doc_ids, n);
  if (s == MGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) s = mgpu_fail(ctx, MGPU_ERR_CUDA, "hnsw_create: upload failed");
  if (s != MGPU_OK) { mgpu_hnsw_destroy(h); return s; }
  *out = h;
  return MGPU_OK;
}

void mgpu_hnsw_destroy(mgpu_hnsw *h) {
  if (!h) return;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  cudaFree(h->d_edges); cudaFree(h->d_points); cudaFree(h->d_edge_offsets); cudaFree(h->d_level_offsets);
  cudaFree(h->d_upper_sorted_pid); cudaFree(h->d_upper_sorted_pos); cudaFree(h->d_upper_dense); cudaFree(h->d_edges0); cudaFree(h->d_rows); cudaFree(h->d_doc_ids);
  delete h;
}

int mgpu_hnsw_search(mgpu_hnsw *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef, mgpu_u128 *out_doc_ids,
                     float *out_scores, uint32_t *out_counts, uint64_t *out_stats, int mem) {
  if (!h) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = h->ctx;
  if (B && (!Q || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_search: null buffer");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  if (h->n == 0) {
    if (mem == MGPU_HOST) memset(out_counts, 0, (size_t)B * 4); else CUDA_TRY(ctx, cudaMemsetAsync(out_counts, 0, (size_t)B * 4, ctx->stream));
    return MGPU_OK;
  }
  HnswSearchArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.k = k; a.ef = ef;
  if (mem == MGPU_DEVICE) {
    a.Q = Q; a.out_docs = out_doc_ids; a.out_scores = out_scores; a.out_counts = out_counts; a.out_stats = out_stats;
    return launch_hnsw_search(h, a);
  }
  size_t kk = std::max<uint32_t>(k, 1);
  size_t need = 0;
  need = ws_need(need, (size_t)B * h->dim * 4); need = ws_need(need, (size_t)B * kk * 16); need = ws_need(need, (size_t)B * kk * 4);
  need = ws_need(need, (size_t)B * 4); need = ws_need(need, (size_t)B * 16);
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  float *dQ = w.get<float>((size_t)B * h->dim);
  mgpu_u128 *dD = w.get<mgpu_u128>((size_t)B * kk); float *dS = w.get<float>((size_t)B * kk);
  uint32_t *dC = w.get<uint32_t>(B); uint64_t *dSt = w.get<uint64_t>((size_t)B * 2);
  CUDA_TRY(ctx, cudaMemcpyAsync(dQ, Q, (size_t)B * h->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  a.Q = dQ; a.out_docs = dD; a.out_scores = dS; a.out_counts = dC; a.out_stats = out_stats ? dSt : nullptr;
  MGPU_TRY(launch_hnsw_search(h, a));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_doc_ids, dD, (size_t)B * k * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_scores, dS, (size_t)B * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_counts, dC, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_stats) CUDA_TRY(ctx, cudaMemcpyAsync(out_stats, dSt, (size_t)B * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

/* Pipelined form of mgpu_hnsw_search for page-locked host buffers (see mgpu_ivf_search_submit). */
int mgpu_hnsw_search_submit(mgpu_hnsw *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef, mgpu_u128 *out_doc_ids,
                            float *out_scores, uint32_t *out_counts, uint64_t *ticket) {
  if (!h || !ticket) return MGPU_ERR_INVALID_ARG;
  *ticket = 0;
  mgpu_ctx *ctx = h->ctx;
  if (B && (!Q || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_search_submit: null buffer");
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  if (h->n == 0 || k == 0) { memset(out_counts, 0, (size_t)B * 4); return MGPU_OK; }   // nothing in flight: ticket 0
  mgpu_ctx::Pipe *pp = nullptr;
  MGPU_TRY(pipe_begin(ctx, Q, (size_t)B * h->dim * 4, (size_t)B * k * 24 + (size_t)B * 4 + 768, &pp));
  WsAlloc wo(pp->out, pp->out_bytes);
  mgpu_u128 *dD = wo.get<mgpu_u128>((size_t)B * k); float *dS = wo.get<float>((size_t)B * k); uint32_t *dC = wo.get<uint32_t>(B);
  HnswSearchArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.k = k; a.ef = ef;
  a.Q = (const float *)pp->q; a.out_docs = dD; a.out_scores = dS; a.out_counts = dC; a.out_stats = nullptr;
  MGPU_TRY(launch_hnsw_search(h, a));
  return pipe_end(ctx, pp, B, k, dD, dS, dC, out_doc_ids, out_scores, out_counts, ticket);
}

int mgpu_spann_create(mgpu_ctx *ctx, mgpu_hnsw *centroids, mgpu_ivf *posting_lists, mgpu_spann **out) {
  if (!ctx || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  if (!centroids || !posting_lists) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_create: null index");
  if (centroids->ctx != ctx || posting_lists->ctx != ctx) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_create: indices belong to another context");
  if (centroids->dim != posting_lists->dim) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_create: dimension mismatch");
  mgpu_spann *s = new mgpu_spann();
  s->ctx = ctx; s->centroids = centroids; s->lists = posting_lists;
  *out = s;
  return MGPU_OK;
}
void mgpu_spann_destroy(mgpu_spann *s) { delete s; }

// centroid-ratio pruning (spann/index.rs:233-246): keep centroids with score - nearest <= nearest * ratio
__global__ void k_spann_prune(const mgpu_u128 *__restrict__ cdocs, const float *__restrict__ cscores,
                              const uint32_t *__restrict__ ccounts, uint32_t B, uint32_t ne, float ratio, uint32_t nlist,
                              uint32_t *__restrict__ probes, uint32_t *__restrict__ pcounts) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint32_t n = ccounts[b];
  float nearest = 0.0f;
  for (uint32_t i = 0; i < n; i++) { float s = cscores[(size_t)b * ne + i]; if (i == 0 || s < nearest) nearest = s; }
  uint32_t kept = 0;
  for (uint32_t i = 0; i < n; i++) {
    float s = cscores[(size_t)b * ne + i];
    uint32_t cid = (uint32_t)cdocs[(size_t)b * ne + i].lo;
    if (__fsub_rn(s, nearest) <= __fmul_rn(nearest, ratio) && cid < nlist) probes[(size_t)b * ne + kept++] = cid;
  }
  pcounts[b] = kept;
}

__global__ void k_spann_mark_none(const uint32_t *__restrict__ ccounts, uint32_t B, uint32_t *__restrict__ out_counts) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && ccounts[b] == 0) out_counts[b] = 0xFFFFFFFFu;
}

static int spann_search_impl(mgpu_spann *sp, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                             uint32_t num_explored_centroids, float ratio, const uint32_t *filter_bits, uint64_t filter_stride,
                             mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem, bool caller_holds_lock,
                             const uint8_t *d_qcodes_ext) {
  if (!sp) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = sp->ctx;
  mgpu_ivf *ivf = sp->lists;
  mgpu_hnsw *hn = sp->centroids;
  if (B && (!Q || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_search: null buffer");
  if (top_k > MGPU_MAX_K) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", top_k, MGPU_MAX_K);
  std::unique_lock<std::mutex> g(ctx->mu, std::defer_lock);
  if (!caller_holds_lock) g.lock();
  cudaSetDevice(ctx->device);
  if (B == 0) return MGPU_OK;
  const uint32_t ne = num_explored_centroids;
  if (ne == 0 || hn->n == 0) {  // empty centroid result => None (spann/index.rs:229-231)
    if (mem == MGPU_HOST) memset(out_counts, 0xFF, (size_t)B * 4); else CUDA_TRY(ctx, cudaMemsetAsync(out_counts, 0xFF, (size_t)B * 4, ctx->stream));
    return MGPU_OK;
  }
  const uint32_t k = std::max<uint32_t>(top_k, 1);
  const uint32_t m = ivf->pq ? ivf->pq->m : 0;
  const uint64_t fwords = (ivf->n + 31) / 32;
  if (filter_bits && filter_stride != 0 && filter_stride < fwords)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_search: filter stride smaller than one bitmap");
  const size_t bF = filter_bits && mem == MGPU_HOST ? (filter_stride ? (size_t)B * filter_stride : (size_t)fwords) * 4 : 0;
  size_t need = 0;
  need = ws_need(need, bF);
  need = ws_need(need, (size_t)B * ivf->dim * 4);
  need = ws_need(need, (size_t)B * ne * 16); need = ws_need(need, (size_t)B * ne * 4); need = ws_need(need, (size_t)B * 4);
  need = ws_need(need, (size_t)B * ne * 4); need = ws_need(need, (size_t)B * 4);
  need = ws_need(need, (size_t)B * m); need = ws_need(need, (size_t)B * MGPU_NCAND * 8); need = ws_need(need, (size_t)B * MGPU_NCAND * 4);
  need = ws_need(need, (size_t)B * k * 16); need = ws_need(need, (size_t)B * k * 4); need = ws_need(need, (size_t)B * 4);
  MGPU_TRY(mgpu_ws_reserve(ctx, need));
  WsAlloc w(ctx->ws, ctx->ws_bytes);
  uint32_t *sF = w.get<uint32_t>(bF / 4);
  const uint32_t *dF = filter_bits;
  if (bF) { const void *t; MGPU_TRY(stage_in(ctx, filter_bits, bF, mem, sF, &t)); dF = (const uint32_t *)t; }
  float *sQ = w.get<float>((size_t)B * ivf->dim);
  mgpu_u128 *cD = w.get<mgpu_u128>((size_t)B * ne); float *cS = w.get<float>((size_t)B * ne); uint32_t *cC = w.get<uint32_t>(B);
  uint32_t *pr = w.get<uint32_t>((size_t)B * ne); uint32_t *pc = w.get<uint32_t>(B);
  uint8_t *dQC = w.get<uint8_t>((size_t)B * m); uint64_t *dCK = w.get<uint64_t>((size_t)B * MGPU_NCAND); uint32_t *dCS = w.get<uint32_t>((size_t)B * MGPU_NCAND);
  mgpu_u128 *sD = w.get<mgpu_u128>((size_t)B * k); float *sS = w.get<float>((size_t)B * k); uint32_t *sC = w.get<uint32_t>(B);
  const void *dQ;
  MGPU_TRY(stage_in(ctx, Q, (size_t)B * ivf->dim * 4, mem, sQ, &dQ));
  HnswSearchArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = (const float *)dQ; a.B = B; a.k = ne; a.ef = ef; a.out_docs = cD; a.out_scores = cS; a.out_counts = cC;
  MGPU_TRY(launch_hnsw_search(hn, a));
  {
    LaunchScope ls(ctx, MGPU_K_OTHER);
    k_spann_prune<<<(B + 127) / 128, 128, 0, ctx->stream>>>(cD, cS, cC, B, ne, ratio, ivf->nlist, pr, pc);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  mgpu_u128 *oD = mem == MGPU_DEVICE ? out_doc_ids : sD;
  float *oS = mem == MGPU_DEVICE ? out_scores : sS;
  uint32_t *oC = mem == MGPU_DEVICE ? out_counts : sC;
  if (top_k == 0) CUDA_TRY(ctx, cudaMemsetAsync(oC, 0, (size_t)B * 4, ctx->stream));
  else MGPU_TRY(ivf_scan_dev(ivf, (const float *)dQ, B, pr, ne, pc, top_k, d_qcodes_ext ? (uint8_t *)d_qcodes_ext : dQC, dCK, dCS, nullptr, nullptr,
                             oD, oS, oC, false, dF, filter_stride, false, d_qcodes_ext != nullptr));
  {
    LaunchScope ls(ctx, MGPU_K_OTHER);
    k_spann_mark_none<<<(B + 127) / 128, 128, 0, ctx->stream>>>(cC, B, oC);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  if (mem == MGPU_HOST) {
    CUDA_TRY(ctx, cudaMemcpyAsync(out_doc_ids, oD, (size_t)B * top_k * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_scores, oS, (size_t)B * top_k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_counts, oC, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MGPU_OK;
}

int mgpu_spann_search(mgpu_spann *sp, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                      uint32_t num_explored_centroids, float ratio, mgpu_u128 *out_doc_ids, float *out_scores,
                      uint32_t *out_counts, int mem) {
  return spann_search_impl(sp, Q, B, top_k, ef, num_explored_centroids, ratio, nullptr, 0, out_doc_ids, out_scores, out_counts, mem);
}

/* Pipelined form of mgpu_spann_search for page-locked host buffers (see mgpu_ivf_search_submit). */
int mgpu_spann_search_submit(mgpu_spann *sp, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids,
                             float ratio, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, uint64_t *ticket) {
  if (!sp || !ticket) return MGPU_ERR_INVALID_ARG;
  *ticket = 0;
  mgpu_ctx *ctx = sp->ctx;
  if (B && (!Q || !out_doc_ids || !out_scores || !out_counts)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "spann_search_submit: null buffer");
  if (top_k > MGPU_MAX_K) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", top_k, MGPU_MAX_K);
  if (B == 0 || top_k == 0)   // nothing to pipeline: the blocking call answers (None / empty per query), ticket 0
    return spann_search_impl(sp, Q, B, top_k, ef, num_explored_centroids, ratio, nullptr, 0, out_doc_ids, out_scores, out_counts, MGPU_HOST);
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  mgpu_ctx::Pipe *pp = nullptr;
  MGPU_TRY(pipe_begin(ctx, Q, (size_t)B * sp->lists->dim * 4, (size_t)B * top_k * 24 + (size_t)B * 4 + 768, &pp));
  WsAlloc wo(pp->out, pp->out_bytes);
  mgpu_u128 *dD = wo.get<mgpu_u128>((size_t)B * top_k); float *dS = wo.get<float>((size_t)B * top_k); uint32_t *dC = wo.get<uint32_t>(B);
  MGPU_TRY(spann_search_impl(sp, (const float *)pp->q, B, top_k, ef, num_explored_centroids, ratio, nullptr, 0, dD, dS, dC, MGPU_DEVICE, true));
  return pipe_end(ctx, pp, B, top_k, dD, dS, dC, out_doc_ids, out_scores, out_counts, ticket);
}

/* Spann::search with Some(planner): the filter reaches the posting-list scan (spann/index.rs:253-263). */
int mgpu_spann_search_filtered(mgpu_spann *sp, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                               uint32_t num_explored_centroids, float ratio, const uint32_t *filter_bits,
                               uint64_t filter_stride_words, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts,
                               int mem) {
  return spann_search_impl(sp, Q, B, top_k, ef, num_explored_centroids, ratio, filter_bits, filter_stride_words, out_doc_ids,
                           out_scores, out_counts, mem);
}

}  // extern "C"

// batcher.cu -- host-side micro-batcher (SURVEY.md 8b "Threading", 8f row 4).
//
// The reference answers one query per call: IndexServer::search (rs/index_server/src/index_server.rs:171-271) awaits
// Snapshot::search_for_users -> ... -> BlockBasedIvf::search / Spann::search for a single vector, from many tokio tasks at
// once (`&self` methods, Send + Sync).  The GPU path wants batches.  The batcher is the seam between the two: every
// caller blocks in mgpu_batcher_search with ONE query; a worker thread collects what has arrived, closes the batch when it
// holds max_batch queries or when its oldest query has waited max_wait_us, issues one batched search through the same
// C-ABI entry points a direct caller would use, and scatters the per-query results into the callers' own buffers.
// Three staging buffers rotate and unfiltered batches go through mgpu_ivf_search_submit / mgpu_spann_search_submit +
// mgpu_search_wait, so two batches are in flight (upload of one next to the kernels and download of the other) while a third
// fills; filtered batches use the blocking call.
#include <chrono>
#include <deque>
#include <condition_variable>
#include <thread>

#include "internal.cuh"

namespace {

struct Waiter {
  mgpu_u128 *out_docs; float *out_scores; uint32_t *out_count; int *status;
};

struct Staging {
  float *Q = nullptr;               // pinned: max_batch x dim
  mgpu_u128 *docs = nullptr;        // pinned: max_batch x k
  float *scores = nullptr;          // pinned: max_batch x k
  uint32_t *counts = nullptr;       // pinned: max_batch
  std::vector<uint32_t> filters;    // lazily max_batch x words (only when a query of the batch carries a planner filter)
  std::vector<Waiter> waiters;
  uint32_t n = 0;
  bool any_filter = false;
  uint64_t generation = 0;          // id of the batch currently filling this buffer
  std::chrono::steady_clock::time_point first_arrival;
};

}  // namespace

struct mgpu_batcher {
  mgpu_ivf *ivf = nullptr;
  mgpu_spann *spann = nullptr;
  uint32_t max_batch = 0, max_wait_us = 0, k = 0, nprobe = 0, ef = 0, nexp = 0, dim = 0;
  float ratio = 0.f;
  uint64_t words = 0;
  std::mutex mu;
  std::condition_variable cv_work, cv_done, cv_space;
  static constexpr int NBUF = 3;
  Staging st[NBUF];
  int fill = 0;                     // buffer callers are appending to
  bool busy[NBUF] = {false, false, false};   // buffer is closed / on the GPU (not appendable)
  uint64_t next_generation = 1, done_generation[NBUF] = {0, 0, 0};
  bool stop = false;
  std::thread worker;
  // stats
  uint64_t n_queries = 0, n_batches = 0, max_seen = 0, full_batches = 0;

  mgpu_ctx *ctx() const { return ivf ? ivf->ctx : spann->ctx; }
};

// scatter the results of the batch in buffer `cur` to its callers and release the buffer (lock NOT held on entry)
static void batcher_complete(mgpu_batcher *b, int cur, int status) {
  Staging &c = b->st[cur];
  const uint32_t n = c.n;
  if (status == MGPU_OK) {
    for (uint32_t i = 0; i < n; i++) {
      const Waiter &w = c.waiters[i];
      const uint32_t cnt = c.counts[i];
      const uint32_t ncopy = cnt == 0xFFFFFFFFu ? 0 : std::min(cnt, b->k);
      memcpy(w.out_docs, c.docs + (size_t)i * b->k, (size_t)ncopy * sizeof(mgpu_u128));
      memcpy(w.out_scores, c.scores + (size_t)i * b->k, (size_t)ncopy * sizeof(float));
      *w.out_count = cnt;
    }
  }
  for (uint32_t i = 0; i < n; i++) *c.waiters[i].status = status;
  std::lock_guard<std::mutex> g(b->mu);
  b->done_generation[cur] = c.generation;
  c.n = 0; c.any_filter = false; c.waiters.clear();
  b->busy[cur] = false;
  b->cv_done.notify_all();
  b->cv_space.notify_all();
}

static void batcher_run(mgpu_batcher *b) {
  std::deque<std::pair<int, uint64_t>> inflight;   // (buffer, ticket) of the submitted batches, oldest first
  std::unique_lock<std::mutex> lk(b->mu);
  auto finish_oldest = [&]() {                     // lock held on entry and on return
    const std::pair<int, uint64_t> f = inflight.front();
    inflight.pop_front();
    lk.unlock();
    const int status = mgpu_search_wait(b->ctx(), f.second);
    batcher_complete(b, f.first, status);
    lk.lock();
  };
  for (;;) {
    Staging &s = b->st[b->fill];
    // close the batch when it is full, when its oldest query has waited long enough, or on shutdown
    const auto deadline = s.first_arrival + std::chrono::microseconds(b->max_wait_us);
    const bool closable = s.n > 0 && (s.n >= b->max_batch || b->stop || std::chrono::steady_clock::now() >= deadline);
    if (!closable) {
      if (!inflight.empty()) { finish_oldest(); continue; }   // nothing to launch: deliver what is on the GPU
      if (s.n == 0) {
        if (b->stop) return;
        b->cv_work.wait(lk);
      } else {
        b->cv_work.wait_until(lk, deadline);
      }
      continue;
    }
    const int cur = b->fill;
    Staging &c = b->st[cur];
    const bool pipelined = !c.any_filter;   // the filtered searches have no submit form
    // two submissions per context; the blocking call runs behind whatever is in flight anyway, so deliver those first
    while (inflight.size() >= (pipelined ? 2u : 1u)) finish_oldest();
    b->busy[cur] = true;
    int next = -1;
    for (int i = 0; i < mgpu_batcher::NBUF; i++) if (!b->busy[i]) { next = i; break; }   // <= 1 in flight + cur: one is free
    b->fill = next;                          // new arrivals go to a free buffer
    b->st[b->fill].generation = b->next_generation++;
    b->cv_space.notify_all();
    const uint32_t n = c.n;
    b->n_queries += n; b->n_batches++; b->max_seen = std::max<uint64_t>(b->max_seen, n);
    if (n == b->max_batch) b->full_batches++;
    lk.unlock();
    if (pipelined) {
      uint64_t ticket = 0;
      const int status = b->ivf ? mgpu_ivf_search_submit(b->ivf, c.Q, n, b->k, b->nprobe, c.docs, c.scores, c.counts, &ticket)
                                : mgpu_spann_search_submit(b->spann, c.Q, n, b->k, b->ef, b->nexp, b->ratio, c.docs, c.scores, c.counts, &ticket);
      if (status != MGPU_OK) batcher_complete(b, cur, status);
      lk.lock();
      if (status == MGPU_OK) inflight.emplace_back(cur, ticket);
      continue;
    }
    const uint32_t *fb = c.any_filter ? c.filters.data() : nullptr;
    int status;
    if (b->ivf) status = mgpu_ivf_search_filtered(b->ivf, c.Q, n, b->k, b->nprobe, fb, fb ? b->words : 0, c.docs, c.scores, c.counts, MGPU_HOST);
    else status = mgpu_spann_search_filtered(b->spann, c.Q, n, b->k, b->ef, b->nexp, b->ratio, fb, fb ? b->words : 0, c.docs, c.scores, c.counts, MGPU_HOST);
    batcher_complete(b, cur, status);
    lk.lock();
  }
}

static int batcher_alloc(mgpu_batcher *b) {
  mgpu_ctx *ctx = b->ctx();
  cudaSetDevice(ctx->device);
  for (int i = 0; i < mgpu_batcher::NBUF; i++) {
    Staging &s = b->st[i];
    CUDA_TRY(ctx, cudaMallocHost((void **)&s.Q, (size_t)b->max_batch * b->dim * sizeof(float)));
    CUDA_TRY(ctx, cudaMallocHost((void **)&s.docs, (size_t)b->max_batch * b->k * sizeof(mgpu_u128)));
    CUDA_TRY(ctx, cudaMallocHost((void **)&s.scores, (size_t)b->max_batch * b->k * sizeof(float)));
    CUDA_TRY(ctx, cudaMallocHost((void **)&s.counts, (size_t)b->max_batch * sizeof(uint32_t)));
    s.waiters.reserve(b->max_batch);
  }
  b->st[0].generation = b->next_generation++;
  return MGPU_OK;
}

static int batcher_finish_create(mgpu_batcher *b, mgpu_batcher **out) {
  int s = batcher_alloc(b);
  if (s != MGPU_OK) { mgpu_batcher_destroy(b); return s; }
  b->worker = std::thread(batcher_run, b);
  *out = b;
  return MGPU_OK;
}

extern "C" {

int mgpu_batcher_create(mgpu_ivf *ivf, uint32_t max_batch, uint32_t max_wait_us, uint32_t k, uint32_t nprobe, mgpu_batcher **out) {
  if (!ivf || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  if (max_batch == 0 || k == 0) return mgpu_fail(ivf->ctx, MGPU_ERR_INVALID_ARG, "batcher: max_batch and k must be positive");
  if (k > MGPU_MAX_K) return mgpu_fail(ivf->ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", k, MGPU_MAX_K);
  if (nprobe == 0 || nprobe > ivf->nlist) return mgpu_fail(ivf->ctx, MGPU_ERR_OUT_OF_RANGE, "num_probes %u out of range 1..%u (the reference panics)", nprobe, ivf->nlist);
  mgpu_batcher *b = new mgpu_batcher();
  b->ivf = ivf; b->max_batch = max_batch; b->max_wait_us = max_wait_us; b->k = k; b->nprobe = nprobe; b->dim = ivf->dim;
  b->words = (ivf->n + 31) / 32;
  return batcher_finish_create(b, out);
}

int mgpu_batcher_create_spann(mgpu_spann *sp, uint32_t max_batch, uint32_t max_wait_us, uint32_t top_k, uint32_t ef,
                              uint32_t num_explored_centroids, float centroid_distance_ratio, mgpu_batcher **out) {
  if (!sp || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  if (max_batch == 0 || top_k == 0) return mgpu_fail(sp->ctx, MGPU_ERR_INVALID_ARG, "batcher: max_batch and top_k must be positive");
  if (top_k > MGPU_MAX_K) return mgpu_fail(sp->ctx, MGPU_ERR_UNSUPPORTED, "k = %u > %d is not supported", top_k, MGPU_MAX_K);
  mgpu_batcher *b = new mgpu_batcher();
  b->spann = sp; b->max_batch = max_batch; b->max_wait_us = max_wait_us; b->k = top_k; b->ef = ef; b->nexp = num_explored_centroids;
  b->ratio = centroid_distance_ratio; b->dim = sp->lists->dim; b->words = (sp->lists->n + 31) / 32;
  return batcher_finish_create(b, out);
}

void mgpu_batcher_destroy(mgpu_batcher *b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> g(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  b->cv_space.notify_all();
  if (b->worker.joinable()) b->worker.join();
  cudaSetDevice(b->ctx()->device);
  for (int i = 0; i < mgpu_batcher::NBUF; i++) {
    cudaFreeHost(b->st[i].Q); cudaFreeHost(b->st[i].docs); cudaFreeHost(b->st[i].scores); cudaFreeHost(b->st[i].counts);
  }
  delete b;
}

int mgpu_batcher_search_filtered(mgpu_batcher *b, const float *query, const uint32_t *filter_bits, mgpu_u128 *out_doc_ids,
                                 float *out_scores, uint32_t *out_count) {
  if (!b || !query || !out_doc_ids || !out_scores || !out_count) return MGPU_ERR_INVALID_ARG;
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return MGPU_ERR_INVALID_ARG;
  // wait for room in the buffer that is filling (only when it is full and the worker has not swapped yet)
  while (!b->stop && b->st[b->fill].n >= b->max_batch) b->cv_space.wait(lk);
  if (b->stop) return MGPU_ERR_INVALID_ARG;
  const int buf = b->fill;
  Staging &s = b->st[buf];
  const uint32_t slot = s.n;
  if (slot == 0) s.first_arrival = std::chrono::steady_clock::now();
  memcpy(s.Q + (size_t)slot * b->dim, query, (size_t)b->dim * sizeof(float));
  if (filter_bits || s.any_filter) {
    if (s.filters.size() < (size_t)b->max_batch * b->words) s.filters.resize((size_t)b->max_batch * b->words);
    if (!s.any_filter) {  // earlier queries of this batch carry no filter: all ids allowed
      memset(s.filters.data(), 0xFF, (size_t)slot * b->words * sizeof(uint32_t));
      s.any_filter = true;
    }
    if (filter_bits) memcpy(s.filters.data() + (size_t)slot * b->words, filter_bits, b->words * sizeof(uint32_t));
    else memset(s.filters.data() + (size_t)slot * b->words, 0xFF, b->words * sizeof(uint32_t));
  }
  int status = MGPU_OK;
  s.waiters.push_back(Waiter{out_doc_ids, out_scores, out_count, &status});
  s.n = slot + 1;
  const uint64_t gen = s.generation;
  if (slot == 0 || s.n == b->max_batch) b->cv_work.notify_one();
  while (b->done_generation[buf] < gen) b->cv_done.wait(lk);
  return status;
}

int mgpu_batcher_search(mgpu_batcher *b, const float *query, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_count) {
  return mgpu_batcher_search_filtered(b, query, nullptr, out_doc_ids, out_scores, out_count);
}

int mgpu_batcher_stats(mgpu_batcher *b, uint64_t stats[4]) {
  if (!b || !stats) return MGPU_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(b->mu);
  stats[0] = b->n_queries; stats[1] = b->n_batches; stats[2] = b->max_seen; stats[3] = b->full_batches;
  return MGPU_OK;
}

}  // extern "C"

/*
 * oracle.c -- CPU restatement of MuopDB's batched-ANN hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is the checker the
 * CUDA path is compared against (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing under
 * muopdb_b200/ may import, link or call it.
 *
 * Parity status: PINNED against the reference's own known-answer tests
 * (tests/test_oracle_golden.py restates them).  The
 * Rust reference itself cannot be compiled in this image (no rustc/cargo), so
 * this is a "port" oracle: every function cites the reference file:line it
 * follows (paths relative to the reference checkout root).
 *
 * Arithmetic contract: fp32 everywhere, multiply and
 * add never fused (build with -ffp-contract=off), 16/8/4-lane accumulators,
 * lanes reduced in index order starting from -0.0 (std::simd reduce_sum =
 * simd_reduce_add_ordered(self, -0.0)), partial sums added 16 -> 8 -> 4 -> tail.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Distance kernels: rs/utils/src/distance/{l2,dot_product,lane_conforming}.rs */
/* ------------------------------------------------------------------------- */

/* std::simd reduce_sum: ordered, lane 0 first, identity -0.0 */
static inline float reduce_sum(const float *v, int lanes) {
  float s = -0.0f;
  for (int i = 0; i < lanes; i++) s = s + v[i];
  return s;
}

/* L2DistanceCalculator::accumulate_lanes<LANES>  (l2.rs:77-89): diff, diff*diff, += */
static inline void l2_accumulate_lanes(const float *a, const float *b, size_t n, int lanes, float *acc) {
  size_t chunks = n / (size_t)lanes;
  for (size_t c = 0; c < chunks; c++) {
    const float *pa = a + c * lanes, *pb = b + c * lanes;
    for (int l = 0; l < lanes; l++) {
      float d = pa[l] - pb[l];
      float sq = d * d;
      acc[l] = acc[l] + sq;
    }
  }
}

/* DotProductDistanceCalculator::accumulate_lanes (dot_product.rs:74-87) */
static inline void dot_accumulate_lanes(const float *a, const float *b, size_t n, int lanes, float *acc) {
  size_t chunks = n / (size_t)lanes;
  for (size_t c = 0; c < chunks; c++) {
    const float *pa = a + c * lanes, *pb = b + c * lanes;
    for (int l = 0; l < lanes; l++) {
      float p = pa[l] * pb[l];
      acc[l] = acc[l] + p;
    }
  }
}

/* L2DistanceCalculator::calculate_squared (l2.rs:30-68) */
static inline __attribute__((always_inline)) float l2_squared_inl(const float *a, const float *b, uint64_t n) {
  if (n == 8) { /* same arithmetic as the general cascade (one 8-lane chunk: 0 + d*d per lane, ordered reduce, 0 + s) */
    float sq[8];
    for (int l = 0; l < 8; l++) { float d = a[l] - b[l]; sq[l] = 0.0f + d * d; }
    float s = -0.0f;
    for (int l = 0; l < 8; l++) s = s + sq[l];
    return 0.0f + s;
  }
  float ret = 0.0f;
  size_t rem = (size_t)n;
  static const int LANES[3] = {16, 8, 4};
  for (int li = 0; li < 3; li++) {
    int L = LANES[li];
    if (rem / (size_t)L > 0) {
      float acc[16] = {0};
      l2_accumulate_lanes(a, b, rem, L, acc);
      size_t used = (rem / (size_t)L) * (size_t)L;
      a += used; b += used; rem -= used;
      ret = ret + reduce_sum(acc, L);
    }
  }
  for (size_t i = 0; i < rem; i++) {
    float d = a[i] - b[i];
    ret = ret + d * d; /* powi(2) */
  }
  return ret;
}

ORC_API float orc_l2_squared(const float *a, const float *b, uint64_t n) { return l2_squared_inl(a, b, n); }

/* L2DistanceCalculator::calculate (l2.rs:72-74) */
ORC_API float orc_l2(const float *a, const float *b, uint64_t n) { return sqrtf(orc_l2_squared(a, b, n)); }

/* L2DistanceCalculator::calculate_scalar (l2.rs:21-27) */
ORC_API float orc_l2_scalar(const float *a, const float *b, uint64_t n) {
  float s = 0.0f;
  for (uint64_t i = 0; i < n; i++) { float d = a[i] - b[i]; s = s + d * d; }
  return sqrtf(s);
}

/* L2 accumulate_scalar (l2.rs:91-94): sum of squares, no sqrt */
static inline float l2_accumulate_scalar(const float *a, const float *b, size_t n) {
  float s = 0.0f;
  for (size_t i = 0; i < n; i++) { float d = a[i] - b[i]; s = s + d * d; }
  return s;
}

/* DotProductDistanceCalculator::calculate (dot_product.rs:38-71): strict '>' thresholds */
ORC_API float orc_dot(const float *a, const float *b, uint64_t n) {
  float res = 0.0f;
  size_t rem = (size_t)n;
  static const int LANES[3] = {16, 8, 4};
  for (int li = 0; li < 3; li++) {
    int L = LANES[li];
    if (rem > (size_t)L) {
      float acc[16] = {0};
      dot_accumulate_lanes(a, b, rem, L, acc);
      res = res + reduce_sum(acc, L);
      size_t used = (rem / (size_t)L) * (size_t)L;
      a += used; b += used; rem -= used;
    }
  }
  for (size_t i = 0; i < rem; i++) res = res + a[i] * b[i];
  return -res; /* neg_score (dot_product.rs:25-27) */
}

/* DotProductDistanceCalculator::calculate_scalar (dot_product.rs:10-16) */
ORC_API float orc_dot_scalar(const float *a, const float *b, uint64_t n) {
  float r = 0.0f;
  for (uint64_t i = 0; i < n; i++) r = r + a[i] * b[i];
  return -r;
}

static inline float dot_accumulate_scalar(const float *a, const float *b, size_t n) {
  float s = 0.0f;
  for (size_t i = 0; i < n; i++) s = s + a[i] * b[i];
  return s;
}

/* LaneConformingDistanceCalculator<LANES, D>::calculate_squared (lane_conforming.rs:16-27).
 * metric: 0 = L2 (outermost_op identity), 1 = dot (outermost_op neg). n % lanes == 0. */
ORC_API float orc_lane_conforming(const float *a, const float *b, uint64_t n, int lanes, int metric) {
  float acc[16] = {0};
  if (metric == 0) l2_accumulate_lanes(a, b, (size_t)n, lanes, acc);
  else dot_accumulate_lanes(a, b, (size_t)n, lanes, acc);
  float r = reduce_sum(acc, lanes);
  return metric == 0 ? r : -r;
}

/* Batched all-pairs form of DistanceCalculator::calculate / calculate_squared
 * (what mgpu_l2_batch / mgpu_dot_batch compute). out[i*nB + j]. */
ORC_API void orc_distance_batch(const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim,
                                int metric, int squared, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nA; i++)
    for (uint64_t j = 0; j < nB; j++) {
      const float *a = A + (size_t)i * dim, *b = B + j * dim;
      float v;
      if (metric == 0) v = squared ? orc_l2_squared(a, b, dim) : orc_l2(a, b, dim);
      else v = orc_dot(a, b, dim);
      out[(size_t)i * nB + j] = v;
    }
}

/* ------------------------------------------------------------------------- */
/* Product quantizer: rs/quantization/src/pq/mod.rs                           */
/* ------------------------------------------------------------------------- */

/* ProductQuantizer::quantize (pq/mod.rs:152-177): per subspace first-minimum argmin of
 * L2DistanceCalculator::calculate_squared, strict '<' against f32::MAX start. */
ORC_API void orc_pq_quantize(const float *cb, uint32_t dim, uint32_t dsub, uint32_t nbits, const float *v,
                             uint8_t *out) {
  uint32_t m = dim / dsub, K = 1u << nbits;
  for (uint32_t s = 0; s < m; s++) {
    const float *sub = v + (size_t)s * dsub;
    const float *base = cb + (size_t)s * dsub * K;
    uint32_t best = 0;
    float best_d = 3.40282347e+38f; /* f32::MAX */
    for (uint32_t i = 0; i < K; i++) {
      float d = l2_squared_inl(sub, base + (size_t)i * dsub, dsub);
      if (d < best_d) { best_d = d; best = i; }
    }
    out[s] = (uint8_t)best;
  }
}

ORC_API void orc_pq_quantize_batch(const float *cb, uint32_t dim, uint32_t dsub, uint32_t nbits,
                                   const float *X, uint64_t n, uint8_t *out) {
  uint32_t m = dim / dsub;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++)
    orc_pq_quantize(cb, dim, dsub, nbits, X + (size_t)i * dim, out + (size_t)i * m);
}

/* ProductQuantizer::original_vector (pq/mod.rs:184-200) */
ORC_API void orc_pq_original_vector(const float *cb, uint32_t dim, uint32_t dsub, uint32_t nbits,
                                    const uint8_t *codes, float *out) {
  uint32_t m = dim / dsub, K = 1u << nbits;
  for (uint32_t s = 0; s < m; s++)
    memcpy(out + (size_t)s * dsub, cb + ((size_t)s * K + codes[s]) * dsub, dsub * sizeof(float));
}

/* ProductQuantizer::distance (pq/mod.rs:202-278).
 * impl: 0 = Scalar (:220-229), 1 = SIMD (:267-276), 2 = StreamingSIMD (:231-266).
 * metric: 0 = L2 calculator, 1 = dot-product calculator as the generic D.
 * StreamingSIMD shares the lane accumulators across all subspaces, reduces once, and keeps the
 * reference's `sum_1 = ...` ASSIGNMENT for the scalar tail (:260). */
ORC_API float orc_pq_distance(const float *cb, uint32_t dim, uint32_t dsub, uint32_t nbits, const uint8_t *a,
                              const uint8_t *b, int impl, int metric) {
  uint32_t m = dim / dsub, K = 1u << nbits;
  if (impl == 0 || impl == 1) {
    float sum = 0.0f;
    for (uint32_t s = 0; s < m; s++) {
      const float *av = cb + ((size_t)s * K + a[s]) * dsub;
      const float *bv = cb + ((size_t)s * K + b[s]) * dsub;
      float d;
      if (impl == 0) d = orc_l2_scalar(av, bv, dsub);
      else d = (metric == 0) ? orc_l2(av, bv, dsub) : orc_dot(av, bv, dsub);
      sum = sum + d * d;
    }
    return sum;
  }
  float s16[16] = {0}, s8[8] = {0}, s4[4] = {0}, s1 = 0.0f;
  for (uint32_t s = 0; s < m; s++) {
    const float *av = cb + ((size_t)s * K + a[s]) * dsub;
    const float *bv = cb + ((size_t)s * K + b[s]) * dsub;
    size_t rem = dsub;
    if (rem / 16 > 0) {
      if (metric == 0) l2_accumulate_lanes(av, bv, rem, 16, s16); else dot_accumulate_lanes(av, bv, rem, 16, s16);
      size_t used = (rem / 16) * 16; av += used; bv += used; rem -= used;
    }
    if (rem / 8 > 0) {
      if (metric == 0) l2_accumulate_lanes(av, bv, rem, 8, s8); else dot_accumulate_lanes(av, bv, rem, 8, s8);
      size_t used = (rem / 8) * 8; av += used; bv += used; rem -= used;
    }
    if (rem / 4 > 0) {
      if (metric == 0) l2_accumulate_lanes(av, bv, rem, 4, s4); else dot_accumulate_lanes(av, bv, rem, 4, s4);
      size_t used = (rem / 4) * 4; av += used; bv += used; rem -= used;
    }
    if (rem > 0) s1 = (metric == 0) ? l2_accumulate_scalar(av, bv, rem) : dot_accumulate_scalar(av, bv, rem);
  }
  float r = reduce_sum(s16, 16) + reduce_sum(s8, 8);
  r = r + reduce_sum(s4, 4);
  r = r + s1;
  return metric == 0 ? r : -r; /* D::outermost_op */
}

/* ------------------------------------------------------------------------- */
/* Orderings: rs/index/src/utils.rs:71-76 (PointAndDistance) and :95-114      */
/* ------------------------------------------------------------------------- */

typedef struct { float distance; uint32_t point_id; } orc_pd;
typedef struct { uint64_t lo, hi; float score; } orc_ids;

/* f32::total_cmp */
static inline int total_cmp(float x, float y) {
  int32_t a, b;
  memcpy(&a, &x, 4); memcpy(&b, &y, 4);
  a ^= (int32_t)(((uint32_t)(a >> 31)) >> 1);
  b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
  return (a > b) - (a < b);
}

/* derive(Ord) on (NotNan<f32> distance, u32 point_id): NotNan compares by value (-0 == +0) */
static inline int pd_cmp(const orc_pd *a, const orc_pd *b) {
  if (a->distance < b->distance) return -1;
  if (a->distance > b->distance) return 1;
  return (a->point_id > b->point_id) - (a->point_id < b->point_id);
}
static int pd_cmp_q(const void *a, const void *b) { return pd_cmp((const orc_pd *)a, (const orc_pd *)b); }

static inline int u128_cmp(uint64_t alo, uint64_t ahi, uint64_t blo, uint64_t bhi) {
  if (ahi != bhi) return ahi < bhi ? -1 : 1;
  if (alo != blo) return alo < blo ? -1 : 1;
  return 0;
}

/* impl Ord for IdWithScore (utils.rs:95-114): NaN last, ties by doc_id */
static int ids_cmp(const orc_ids *a, const orc_ids *b) {
  int an = isnan(a->score), bn = isnan(b->score);
  if (an && bn) return u128_cmp(a->lo, a->hi, b->lo, b->hi);
  if (an) return 1;
  if (bn) return -1;
  if (a->score < b->score) return -1;
  if (a->score > b->score) return 1;
  return u128_cmp(a->lo, a->hi, b->lo, b->hi);
}
static int ids_cmp_q(const void *a, const void *b) { return ids_cmp((const orc_ids *)a, (const orc_ids *)b); }

/* Sort (doc_id, score) pairs by IdWithScore::cmp and keep the first k:
 * Snapshot::search_for_user / search_for_users merge (collection/snapshot.rs:60-61,105-106). */
ORC_API uint32_t orc_merge_topk(const uint64_t *doc_ids /* n x (lo,hi) */, const float *scores, uint32_t n,
                                uint32_t k, uint64_t *out_doc_ids, float *out_scores) {
  orc_ids *v = (orc_ids *)malloc(sizeof(orc_ids) * (n ? n : 1));
  for (uint32_t i = 0; i < n; i++) { v[i].lo = doc_ids[2 * i]; v[i].hi = doc_ids[2 * i + 1]; v[i].score = scores[i]; }
  qsort(v, n, sizeof(orc_ids), ids_cmp_q);
  uint32_t c = n < k ? n : k;
  for (uint32_t i = 0; i < c; i++) { out_doc_ids[2 * i] = v[i].lo; out_doc_ids[2 * i + 1] = v[i].hi; out_scores[i] = v[i].score; }
  free(v);
  return c;
}

/* ------------------------------------------------------------------------- */
/* Quantizer dispatch (rs/quantization/src/{noq,pq}/mod.rs, typing.rs)        */
/* ------------------------------------------------------------------------- */

/* quant kinds */
enum { ORC_FLAT = 0, ORC_PQ = 1 };
enum { ORC_L2 = 0, ORC_DOT = 1 };

typedef struct {
  int quant, metric;
  uint32_t dim;   /* original dimension */
  uint32_t qdim;  /* quantized_dimension(): dim (flat) or dim/dsub (pq) */
  const float *cb; uint32_t dsub, nbits;
} orc_quantizer;

/* VectorOps::process_vector -> Quantizer::quantize (typing.rs:17-19,31-33; noq:32-34; pq:152-177) */
static void *q_process_vector(const orc_quantizer *q, const float *v) {
  if (q->quant == ORC_FLAT) {
    float *o = (float *)malloc(sizeof(float) * q->dim);
    memcpy(o, v, sizeof(float) * q->dim); /* to_vec() */
    return o;
  }
  uint8_t *o = (uint8_t *)malloc(q->qdim);
  orc_pq_quantize(q->cb, q->dim, q->dsub, q->nbits, v, o);
  return o;
}

/* Quantizer::distance(query_q, point_q, StreamingSIMD) (noq:44-51 => D::calculate; pq:231-266) */
static inline float q_distance(const orc_quantizer *q, const void *qq, const void *row) {
  if (q->quant == ORC_FLAT)
    return q->metric == ORC_L2 ? orc_l2((const float *)qq, (const float *)row, q->dim)
                               : orc_dot((const float *)qq, (const float *)row, q->dim);
  return orc_pq_distance(q->cb, q->dim, q->dsub, q->nbits, (const uint8_t *)qq, (const uint8_t *)row, 2, q->metric);
}

static inline const void *q_row(const orc_quantizer *q, const void *rows, uint32_t pid) {
  return q->quant == ORC_FLAT ? (const void *)((const float *)rows + (size_t)pid * q->dim)
                              : (const void *)((const uint8_t *)rows + (size_t)pid * q->qdim);
}

/* ------------------------------------------------------------------------- */
/* IVF: rs/index/src/ivf/block_based/index.rs                                 */
/* ------------------------------------------------------------------------- */

typedef struct orc_ivf {
  orc_quantizer q;
  uint32_t nlist; uint64_t n;
  const float *centroids;          /* nlist x dim */
  const uint64_t *list_offsets;    /* nlist + 1 */
  const uint32_t *list_ids;        /* point ids, ascending within a list */
  const void *rows;                /* indexed by point id (vector_storage.get(point_id)) */
  const uint64_t *doc_ids;         /* n x (lo,hi); NULL => doc_id == point_id */
  uint8_t *invalid;                /* bitmap over point ids (DashSet<u32>, index.rs:30) */
  int hoist_quantize;              /* 0 = faithful: quantize query once per probed list (index.rs:193) */
} orc_ivf;

ORC_API orc_ivf *orc_ivf_new(uint32_t dim, uint32_t nlist, const float *centroids, const uint64_t *list_offsets,
                             const uint32_t *list_ids, int quant, int metric, const void *rows, uint64_t n,
                             const uint64_t *doc_ids, const float *cb, uint32_t dsub, uint32_t nbits) {
  orc_ivf *x = (orc_ivf *)calloc(1, sizeof(orc_ivf));
  x->q.quant = quant; x->q.metric = metric; x->q.dim = dim; x->q.cb = cb; x->q.dsub = dsub; x->q.nbits = nbits;
  x->q.qdim = quant == ORC_PQ ? dim / dsub : dim;
  x->nlist = nlist; x->n = n; x->centroids = centroids; x->list_offsets = list_offsets; x->list_ids = list_ids;
  x->rows = rows; x->doc_ids = doc_ids;
  x->invalid = (uint8_t *)calloc((n + 7) / 8 + 1, 1);
  return x;
}
ORC_API void orc_ivf_free(orc_ivf *x) { if (x) { free(x->invalid); free(x); } }
ORC_API void orc_ivf_set_hoist_quantize(orc_ivf *x, int on) { x->hoist_quantize = on; }

/* BlockBasedIvf::invalidate_batch: marks point ids (index.rs:338-394 family) */
ORC_API void orc_ivf_invalidate(orc_ivf *x, const uint32_t *pids, uint32_t cnt) {
  for (uint32_t i = 0; i < cnt; i++) if (pids[i] < x->n) x->invalid[pids[i] >> 3] |= (uint8_t)(1u << (pids[i] & 7));
}
static inline int ivf_is_invalid(const orc_ivf *x, uint32_t pid) { return (x->invalid[pid >> 3] >> (pid & 7)) & 1; }

typedef struct { uint32_t idx; float dist; } idx_dist;
static int idx_dist_cmp(const void *a, const void *b) {
  const idx_dist *x = (const idx_dist *)a, *y = (const idx_dist *)b;
  int c = total_cmp(x->dist, y->dist);
  if (c) return c;
  return (x->idx > y->idx) - (x->idx < y->idx); /* reference leaves tie order unspecified; we fix index order */
}

/* BlockBasedIvf::find_nearest_centroids (index.rs:147-163): sqrt-L2 to every centroid,
 * nprobe smallest by total_cmp, nearest first.  Returns -1 where the reference panics. */
ORC_API int orc_ivf_find_nearest_centroids(const orc_ivf *x, const float *q, uint32_t nprobe, uint32_t *out_ids,
                                           float *out_dists) {
  if (nprobe == 0 || nprobe > x->nlist) return -1;
  idx_dist *d = (idx_dist *)malloc(sizeof(idx_dist) * x->nlist);
  for (uint32_t i = 0; i < x->nlist; i++) {
    d[i].idx = i;
    d[i].dist = orc_l2(q, x->centroids + (size_t)i * x->q.dim, x->q.dim);
  }
  qsort(d, x->nlist, sizeof(idx_dist), idx_dist_cmp);
  for (uint32_t i = 0; i < nprobe; i++) { out_ids[i] = d[i].idx; if (out_dists) out_dists[i] = d[i].dist; }
  free(d);
  return (int)nprobe;
}

/* bounded max-heap on PointAndDistance, as std BinaryHeap is used in index.rs:257-275 */
static void heap_sift_up(orc_pd *h, uint32_t i) {
  while (i > 0) { uint32_t p = (i - 1) / 2; if (pd_cmp(&h[i], &h[p]) <= 0) break; orc_pd t = h[i]; h[i] = h[p]; h[p] = t; i = p; }
}
static void heap_sift_down(orc_pd *h, uint32_t n, uint32_t i) {
  for (;;) {
    uint32_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && pd_cmp(&h[l], &h[m]) > 0) m = l;
    if (r < n && pd_cmp(&h[r], &h[m]) > 0) m = r;
    if (m == i) break;
    orc_pd t = h[i]; h[i] = h[m]; h[m] = t; i = m;
  }
}

/* scan_posting_list (index.rs:175-237) + search_with_centroids (index.rs:250-285).
 * Returns number of results (<= k), sorted by (distance, point_id). */
/* `filter` (may be NULL) stands for Some(planner): planner.plan_with_ids(all_ids) yields the scanned ids that are in the
 * request's DocumentFilter (rs/index/src/query/planner.rs:43-60); here the filter is that id set as a bitmap over point
 * ids.  Distances are computed for every non-invalidated row BEFORE the filter is applied (index.rs:196-226). */
ORC_API int orc_ivf_search_with_centroids_f(const orc_ivf *x, const float *query, const uint32_t *cids, uint32_t ncids,
                                            uint32_t k, const uint32_t *filter, uint32_t *out_pids, float *out_dists) {
  orc_pd *heap = (orc_pd *)malloc(sizeof(orc_pd) * (k ? k : 1));
  uint32_t hn = 0;
  void *qq_hoisted = x->hoist_quantize ? q_process_vector(&x->q, query) : NULL;
  for (uint32_t ci = 0; ci < ncids; ci++) {
    uint32_t c = cids[ci];
    uint64_t b = x->list_offsets[c], e = x->list_offsets[c + 1];
    /* index.rs:193 -- the query is (re)quantized for every probed list */
    void *qq = qq_hoisted ? qq_hoisted : q_process_vector(&x->q, query);
    orc_pd *pds = (orc_pd *)malloc(sizeof(orc_pd) * (size_t)(e - b + 1));
    uint32_t np = 0;
    for (uint64_t i = b; i < e; i++) {
      uint32_t pid = x->list_ids[i];
      if (ivf_is_invalid(x, pid)) continue;                       /* index.rs:198-200 */
      float d = q_distance(&x->q, qq, q_row(&x->q, x->rows, pid)); /* index.rs:202-207 */
      pds[np].distance = d; pds[np].point_id = pid; np++;
    }
    if (filter) {                                                 /* index.rs:214-226: retain ids the planner yields */
      uint32_t w = 0;
      for (uint32_t i = 0; i < np; i++)
        if ((filter[pds[i].point_id >> 5] >> (pds[i].point_id & 31)) & 1u) pds[w++] = pds[i];
      np = w;
    }
    /* index.rs:212 sort by id, :228 stable sort by distance (total_cmp) == sort by (distance, id) */
    qsort(pds, np, sizeof(orc_pd), pd_cmp_q);
    for (uint32_t i = 0; i < np; i++) {                           /* index.rs:265-274 */
      if (hn < k) { heap[hn] = pds[i]; heap_sift_up(heap, hn); hn++; }
      else if (hn > 0 && pd_cmp(&pds[i], &heap[0]) < 0) { heap[0] = pds[i]; heap_sift_down(heap, hn, 0); }
    }
    free(pds);
    if (!qq_hoisted) free(qq);
  }
  free(qq_hoisted);
  qsort(heap, hn, sizeof(orc_pd), pd_cmp_q);                      /* results.sort() index.rs:278 */
  for (uint32_t i = 0; i < hn; i++) { out_pids[i] = heap[i].point_id; out_dists[i] = heap[i].distance; }
  free(heap);
  return (int)hn;
}

ORC_API int orc_ivf_search_with_centroids(const orc_ivf *x, const float *query, const uint32_t *cids, uint32_t ncids,
                                          uint32_t k, uint32_t *out_pids, float *out_dists) {
  return orc_ivf_search_with_centroids_f(x, query, cids, ncids, k, NULL, out_pids, out_dists);
}

static inline void ivf_doc_id(const orc_ivf *x, uint32_t pid, uint64_t *lo, uint64_t *hi) {
  if (x->doc_ids) { *lo = x->doc_ids[2 * (size_t)pid]; *hi = x->doc_ids[2 * (size_t)pid + 1]; }
  else { *lo = pid; *hi = 0; }
}

/* search_with_centroids_and_remap (index.rs:298-332): point ids -> doc ids, sort by (score, doc_id) */
ORC_API int orc_ivf_search_with_centroids_and_remap_f(const orc_ivf *x, const float *query, const uint32_t *cids,
                                                      uint32_t ncids, uint32_t k, const uint32_t *filter,
                                                      uint64_t *out_doc_ids, float *out_scores) {
  uint32_t *pids = (uint32_t *)malloc(sizeof(uint32_t) * (k ? k : 1));
  float *ds = (float *)malloc(sizeof(float) * (k ? k : 1));
  int n = orc_ivf_search_with_centroids_f(x, query, cids, ncids, k, filter, pids, ds);
  orc_ids *v = (orc_ids *)malloc(sizeof(orc_ids) * (n ? n : 1));
  for (int i = 0; i < n; i++) { ivf_doc_id(x, pids[i], &v[i].lo, &v[i].hi); v[i].score = ds[i]; }
  qsort(v, n, sizeof(orc_ids), ids_cmp_q);
  for (int i = 0; i < n; i++) { out_doc_ids[2 * i] = v[i].lo; out_doc_ids[2 * i + 1] = v[i].hi; out_scores[i] = v[i].score; }
  free(v); free(pids); free(ds);
  return n;
}

ORC_API int orc_ivf_search_with_centroids_and_remap(const orc_ivf *x, const float *query, const uint32_t *cids,
                                                    uint32_t ncids, uint32_t k, uint64_t *out_doc_ids,
                                                    float *out_scores) {
  return orc_ivf_search_with_centroids_and_remap_f(x, query, cids, ncids, k, NULL, out_doc_ids, out_scores);
}

/* BlockBasedIvf::search (index.rs:396-412); filter = Some(planner) or NULL */
ORC_API int orc_ivf_search_f(const orc_ivf *x, const float *query, uint32_t k, uint32_t nprobe, const uint32_t *filter,
                             uint64_t *out_doc_ids, float *out_scores) {
  uint32_t *cids = (uint32_t *)malloc(sizeof(uint32_t) * (nprobe ? nprobe : 1));
  int r = orc_ivf_find_nearest_centroids(x, query, nprobe, cids, NULL);
  if (r < 0) { free(cids); return -1; }
  r = orc_ivf_search_with_centroids_and_remap_f(x, query, cids, nprobe, k, filter, out_doc_ids, out_scores);
  free(cids);
  return r;
}
ORC_API int orc_ivf_search(const orc_ivf *x, const float *query, uint32_t k, uint32_t nprobe, uint64_t *out_doc_ids,
                           float *out_scores) {
  return orc_ivf_search_f(x, query, k, nprobe, NULL, out_doc_ids, out_scores);
}

/* filtered batch: query b uses filter + b * stride_words (stride 0: one bitmap for all) */
ORC_API void orc_ivf_search_batch_f(const orc_ivf *x, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe,
                                    const uint32_t *filter, uint64_t stride_words, uint64_t *out_doc_ids,
                                    float *out_scores, int32_t *out_counts, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < (int64_t)B; b++)
    out_counts[b] = orc_ivf_search_f(x, Q + (size_t)b * x->q.dim, k, nprobe, filter ? filter + (size_t)b * stride_words : NULL,
                                     out_doc_ids + (size_t)b * k * 2, out_scores + (size_t)b * k);
}

/* one query per thread over a batch (the reference has no batching; this is B independent calls) */
ORC_API void orc_ivf_search_batch(const orc_ivf *x, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe,
                                  uint64_t *out_doc_ids, float *out_scores, int32_t *out_counts, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < (int64_t)B; b++)
    out_counts[b] = orc_ivf_search(x, Q + (size_t)b * x->q.dim, k, nprobe, out_doc_ids + (size_t)b * k * 2,
                                   out_scores + (size_t)b * k);
}

/* Build-time assignment: IvfBuilder::find_nearest_centroids (ivf/builder.rs:268-282) +
 * the acceptance rule of build_posting_lists (:303-329).  SQUARED L2; top-`max_clusters`
 * by total_cmp (ties by centroid index here); accept if |d - dmin| <= dmin * threshold.
 * out_cids: n x max_clusters (padded with UINT32_MAX), out_counts: n. */
ORC_API void orc_ivf_assign(const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim,
                            uint32_t max_clusters, float threshold, uint32_t *out_cids, uint32_t *out_counts) {
#pragma omp parallel
  {
    idx_dist *d = (idx_dist *)malloc(sizeof(idx_dist) * nlist);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      for (uint32_t c = 0; c < nlist; c++) {
        d[c].idx = c;
        d[c].dist = orc_l2_squared(X + (size_t)i * dim, centroids + (size_t)c * dim, dim);
      }
      qsort(d, nlist, sizeof(idx_dist), idx_dist_cmp);
      uint32_t r = max_clusters < nlist ? max_clusters : nlist;
      float dmin = d[0].dist; /* min_by partial_cmp over the kept set */
      uint32_t cnt = 0;
      for (uint32_t j = 0; j < max_clusters; j++) out_cids[(size_t)i * max_clusters + j] = UINT32_MAX;
      for (uint32_t j = 0; j < r; j++)
        if (fabsf(d[j].dist - dmin) <= dmin * threshold) out_cids[(size_t)i * max_clusters + cnt++] = d[j].idx;
      out_counts[i] = cnt;
    }
    free(d);
  }
}

/* Assignment step of KMeansBuilder::run_lloyd (kmeans_builder.rs:199-221): per point
 * argmin_c (T::calculate_squared(x, c) + penalties[c]) folded from (0, f32::MAX) with a strict '<'
 * (first minimum wins; a NaN cost is never taken).  T is chosen by the dimension (kmeans_builder.rs:126-136):
 * LaneConformingDistanceCalculator<16|8|4, D> when dim % 16|8|4 == 0, else D itself.
 * metric: 0 = L2, 1 = dot.  out_costs (may be NULL) = the winning cost (distance + penalty). */
static inline float kmeans_cost(const float *x, const float *c, uint32_t dim, int metric) {
  int lanes = dim % 16 == 0 ? 16 : (dim % 8 == 0 ? 8 : (dim % 4 == 0 ? 4 : 0));
  if (lanes) return orc_lane_conforming(x, c, dim, lanes, metric);
  return metric == 0 ? orc_l2_squared(x, c, dim) : orc_dot(x, c, dim);
}

ORC_API void orc_kmeans_assign(const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim, int metric,
                               const float *penalties, uint32_t *out, float *out_costs) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    float best = 3.40282347e+38f; uint32_t bi = 0;
    for (uint32_t c = 0; c < nlist; c++) {
      float d = kmeans_cost(X + (size_t)i * dim, centroids + (size_t)c * dim, dim, metric);
      d = d + (penalties ? penalties[c] : 0.0f);   /* penalties default to 0.0 (kmeans_builder.rs:183-189) */
      if (d < best) { best = d; bi = c; }
    }
    out[i] = bi;
    if (out_costs) out_costs[i] = best;
  }
}

/* Batched all-pairs LaneConformingDistanceCalculator<LANES, D>::calculate_squared (what mgpu_distance_batch_lanes computes). */
ORC_API void orc_lane_conforming_batch(const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim, int lanes,
                                       int metric, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nA; i++)
    for (uint64_t j = 0; j < nB; j++)
      out[(size_t)i * nB + j] = orc_lane_conforming(A + (size_t)i * dim, B + j * dim, dim, lanes, metric);
}

/* ------------------------------------------------------------------------- */
/* HNSW: rs/index/src/hnsw/block_based/{index,graph_storage}.rs               */
/* ------------------------------------------------------------------------- */

typedef struct orc_hnsw {
  orc_quantizer q;
  uint32_t num_layers; uint64_t n;
  const uint32_t *edges; const uint32_t *points;
  const uint64_t *edge_offsets; uint64_t n_edge_offsets;
  const uint64_t *level_offsets; /* num_layers + 1, top layer first */
  const void *rows; const uint64_t *doc_ids;
} orc_hnsw;

ORC_API orc_hnsw *orc_hnsw_new(uint32_t dim, uint32_t num_layers, const uint32_t *edges, const uint32_t *points,
                               const uint64_t *edge_offsets, uint64_t n_edge_offsets, const uint64_t *level_offsets,
                               int quant, int metric, const void *rows, uint64_t n, const uint64_t *doc_ids,
                               const float *cb, uint32_t dsub, uint32_t nbits) {
  orc_hnsw *h = (orc_hnsw *)calloc(1, sizeof(orc_hnsw));
  h->q.quant = quant; h->q.metric = metric; h->q.dim = dim; h->q.cb = cb; h->q.dsub = dsub; h->q.nbits = nbits;
  h->q.qdim = quant == ORC_PQ ? dim / dsub : dim;
  h->num_layers = num_layers; h->n = n; h->edges = edges; h->points = points; h->edge_offsets = edge_offsets;
  h->n_edge_offsets = n_edge_offsets; h->level_offsets = level_offsets; h->rows = rows; h->doc_ids = doc_ids;
  return h;
}
ORC_API void orc_hnsw_free(orc_hnsw *h) { free(h); }

/* get_edges_for_point (graph_storage.rs:459-521).  Returns edge count, 0 => None. */
static uint32_t hnsw_edges(const orc_hnsw *h, uint32_t pid, uint32_t layer, const uint32_t **out) {
  uint32_t L = h->num_layers;
  if (layer >= L) return 0;
  uint64_t s = h->level_offsets[L - 1 - layer], e = h->level_offsets[L - layer];
  uint64_t idx;
  if (layer > 0) {
    uint64_t i = s;
    for (; i < e; i++) if (h->points[i] == pid) break; /* find_point_in_range: first match */
    if (i == e) return 0;
    idx = i - s;
  } else idx = pid;
  if (s + idx + 1 >= h->n_edge_offsets) return 0;
  uint64_t a = h->edge_offsets[s + idx], b = h->edge_offsets[s + idx + 1];
  if (a == b) return 0;
  *out = h->edges + a;
  return (uint32_t)(b - a);
}

/* get_entry_point_top_layer (graph_storage.rs:527-554) */
ORC_API uint32_t orc_hnsw_entry_point(const orc_hnsw *h) {
  if (h->num_layers == 1) {
    uint64_t np = h->n_edge_offsets - 1;
    for (uint64_t i = 0; i < np; i++) if (h->edge_offsets[i + 1] > h->edge_offsets[i]) return (uint32_t)i;
    return 0;
  }
  return h->points[h->level_offsets[0]];
}

/* generic binary heap over orc_pd with a sign: max-heap on pd_cmp */
typedef struct { orc_pd *v; uint32_t n, cap; } pdheap;
static void ph_push(pdheap *h, orc_pd x) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->v = (orc_pd *)realloc(h->v, sizeof(orc_pd) * h->cap); }
  h->v[h->n] = x; heap_sift_up(h->v, h->n); h->n++;
}
static orc_pd ph_pop(pdheap *h) {
  orc_pd top = h->v[0]; h->n--;
  if (h->n) { h->v[0] = h->v[h->n]; heap_sift_down(h->v, h->n, 0); }
  return top;
}

typedef struct { uint8_t *visited; uint64_t n_dist, n_expand; } hnsw_ctx;
static inline int ctx_visited(hnsw_ctx *c, uint32_t id) { return (c->visited[id >> 3] >> (id & 7)) & 1; }
static inline void ctx_set(hnsw_ctx *c, uint32_t id) { c->visited[id >> 3] |= (uint8_t)(1u << (id & 7)); }

/* BlockBasedHnsw::search_layer (hnsw/block_based/index.rs:212-287).  Returns the working list
 * sorted by (distance, point_id); caller frees. */
static orc_pd *hnsw_search_layer(const orc_hnsw *h, hnsw_ctx *ctx, const void *qq, uint32_t ep, uint32_t ef,
                                 uint32_t layer, uint32_t *out_n) {
  ctx_set(ctx, ep);
  pdheap cand = {0}, work = {0};
  float ed = q_distance(&h->q, qq, q_row(&h->q, h->rows, ep)); ctx->n_dist++;
  orc_pd c0 = {-ed, ep}, w0 = {ed, ep};
  ph_push(&cand, c0); ph_push(&work, w0);
  while (cand.n) {
    orc_pd c = ph_pop(&cand);
    float dist = -c.distance;
    if (work.n == 0) continue;
    if (dist > work.v[0].distance) break;                      /* :244 strict */
    const uint32_t *edges = NULL;
    uint32_t ne = hnsw_edges(h, c.point_id, layer, &edges);
    if (ne == 0) continue;
    ctx->n_expand++;
    for (uint32_t i = 0; i < ne; i++) {
      uint32_t e = edges[i];
      if (ctx_visited(ctx, e)) continue;
      ctx_set(ctx, e);
      if (work.n == 0) continue;                               /* :259-262 peek None => continue */
      float furthest = work.v[0].distance;
      float de = q_distance(&h->q, qq, q_row(&h->q, h->rows, e)); ctx->n_dist++;
      if (de < furthest || work.n < ef) {                      /* :266-268 */
        orc_pd a = {-de, e}, b = {de, e};
        ph_push(&cand, a); ph_push(&work, b);
        if (work.n > ef) ph_pop(&work);                        /* :277-279 */
      }
    }
  }
  qsort(work.v, work.n, sizeof(orc_pd), pd_cmp_q);             /* :284-286 */
  *out_n = work.n;
  free(cand.v);
  return work.v;
}

/* BlockBasedHnsw::ann_search (index.rs:159-210).  stats (may be NULL): [0]=#distance evals, [1]=#expansions.
 * out_pids may be NULL.  Returns result count (<= k). */
ORC_API int orc_hnsw_ann_search(const orc_hnsw *h, const float *query, uint32_t k, uint32_t ef, uint64_t *out_doc_ids,
                                float *out_scores, uint32_t *out_pids, uint64_t *stats) {
  void *qq = q_process_vector(&h->q, query);
  hnsw_ctx ctx; ctx.visited = (uint8_t *)calloc((h->n + 7) / 8 + 1, 1); ctx.n_dist = 0; ctx.n_expand = 0;
  int32_t layer = (int32_t)h->num_layers - 1;
  uint32_t ep = orc_hnsw_entry_point(h);
  orc_pd *ws; uint32_t wn;
  while (layer > 0) {
    ws = hnsw_search_layer(h, &ctx, qq, ep, ef, (uint32_t)layer, &wn);
    /* min_by distance: first minimum in iteration order of the sorted list */
    uint32_t bi = 0;
    for (uint32_t i = 1; i < wn; i++) if (ws[i].distance < ws[bi].distance) bi = i;
    if (wn) ep = ws[bi].point_id;
    free(ws);
    layer--;
  }
  ws = hnsw_search_layer(h, &ctx, qq, ep, ef, 0, &wn);
  /* stable sort by distance of an already (distance,id)-sorted list is a no-op; truncate k */
  uint32_t c = wn < k ? wn : k;
  for (uint32_t i = 0; i < c; i++) {
    uint32_t pid = ws[i].point_id;
    if (out_pids) out_pids[i] = pid;
    if (out_doc_ids) {
      if (h->doc_ids) { out_doc_ids[2 * i] = h->doc_ids[2 * (size_t)pid]; out_doc_ids[2 * i + 1] = h->doc_ids[2 * (size_t)pid + 1]; }
      else { out_doc_ids[2 * i] = pid; out_doc_ids[2 * i + 1] = 0; }
    }
    out_scores[i] = ws[i].distance;
  }
  if (stats) { stats[0] = ctx.n_dist; stats[1] = ctx.n_expand; }
  free(ws); free(ctx.visited); free(qq);
  return (int)c;
}

ORC_API void orc_hnsw_search_batch(const orc_hnsw *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef,
                                   uint64_t *out_doc_ids, float *out_scores, int32_t *out_counts, uint64_t *stats,
                                   int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < (int64_t)B; b++)
    out_counts[b] = orc_hnsw_ann_search(h, Q + (size_t)b * h->q.dim, k, ef, out_doc_ids + (size_t)b * k * 2,
                                        out_scores + (size_t)b * k, NULL, stats ? stats + 2 * b : NULL);
}

/* ------------------------------------------------------------------------- */
/* SPANN: rs/index/src/spann/index.rs:211-266                                 */
/* ------------------------------------------------------------------------- */

/* centroids: HNSW over the IVF centroids (NoQuantizer<L2>), whose doc ids are centroid indices.
 * Returns result count, or -1 for None. */
ORC_API int orc_spann_search_f(const orc_hnsw *centroids, const orc_ivf *lists, const float *query, uint32_t top_k,
                               uint32_t ef, uint32_t num_explored_centroids, float centroid_distance_ratio,
                               const uint32_t *filter, uint64_t *out_doc_ids, float *out_scores) {
  uint32_t ne = num_explored_centroids;
  uint64_t *cdoc = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (ne ? ne : 1));
  float *cs = (float *)malloc(sizeof(float) * (ne ? ne : 1));
  int nc = orc_hnsw_ann_search(centroids, query, ne, ef, cdoc, cs, NULL, NULL);
  if (nc <= 0) { free(cdoc); free(cs); return -1; }
  float nearest = cs[0];
  for (int i = 1; i < nc; i++) if (cs[i] < nearest) nearest = cs[i];     /* :233-237 */
  uint32_t *cids = (uint32_t *)malloc(sizeof(uint32_t) * nc);
  uint32_t kept = 0;
  for (int i = 0; i < nc; i++)
    if (cs[i] - nearest <= nearest * centroid_distance_ratio) cids[kept++] = (uint32_t)cdoc[2 * i]; /* :239-246 */
  int r = orc_ivf_search_with_centroids_and_remap_f(lists, query, cids, kept, top_k, filter, out_doc_ids, out_scores);
  free(cids); free(cdoc); free(cs);
  return r;
}
ORC_API int orc_spann_search(const orc_hnsw *centroids, const orc_ivf *lists, const float *query, uint32_t top_k,
                             uint32_t ef, uint32_t num_explored_centroids, float centroid_distance_ratio,
                             uint64_t *out_doc_ids, float *out_scores) {
  return orc_spann_search_f(centroids, lists, query, top_k, ef, num_explored_centroids, centroid_distance_ratio, NULL,
                            out_doc_ids, out_scores);
}
ORC_API void orc_spann_search_batch_f(const orc_hnsw *centroids, const orc_ivf *lists, const float *Q, uint32_t B,
                                      uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids, float ratio,
                                      const uint32_t *filter, uint64_t stride_words, uint64_t *out_doc_ids,
                                      float *out_scores, int32_t *out_counts, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < (int64_t)B; b++)
    out_counts[b] = orc_spann_search_f(centroids, lists, Q + (size_t)b * lists->q.dim, top_k, ef, num_explored_centroids, ratio,
                                       filter ? filter + (size_t)b * stride_words : NULL,
                                       out_doc_ids + (size_t)b * top_k * 2, out_scores + (size_t)b * top_k);
}

ORC_API void orc_spann_search_batch(const orc_hnsw *centroids, const orc_ivf *lists, const float *Q, uint32_t B,
                                    uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids, float ratio,
                                    uint64_t *out_doc_ids, float *out_scores, int32_t *out_counts, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < (int64_t)B; b++)
    out_counts[b] = orc_spann_search(centroids, lists, Q + (size_t)b * lists->q.dim, top_k, ef,
                                     num_explored_centroids, ratio, out_doc_ids + (size_t)b * top_k * 2,
                                     out_scores + (size_t)b * top_k);
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

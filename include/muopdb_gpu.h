/*
 * muopdb_gpu.h -- C ABI of the B200-native (sm_100a) batched ANN search path for MuopDB.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch
 * types, no exceptions.  Each entry point cites the reference interface it replaces (paths are
 * relative to the reference checkout, hicder/muopdb @ c520f017).  The reference has no FFI of its
 * own; INTEGRATION.md shows the Rust `extern "C"` binding a maintainer would add behind
 * `Quantizer`, `BlockBasedIvf`, `BlockBasedHnsw` and `Spann`.
 *
 * Conventions
 *   - Every function returns an int status: MGPU_OK or a negative MGPU_ERR_* code;
 *     mgpu_last_error(ctx) returns a human-readable message for the last failure on that ctx.
 *   - `mem` says where the caller's query/result buffers live: MGPU_HOST (pageable or pinned host
 *     memory; the call copies in/out and returns when results are in the host buffers) or
 *     MGPU_DEVICE (device pointers on ctx's device; the call only enqueues work on the ctx stream --
 *     call mgpu_sync before reading results).
 *   - Index-construction inputs (`*_create`) are copied to HBM in a scan-friendly layout; the
 *     caller keeps ownership of its buffers.  Handles are opaque and freed by `*_destroy`.
 *   - Scores follow the reference: lower is closer; flat L2 = sqrt(sum (a-b)^2)
 *     (rs/quantization/src/noq/mod.rs:44-51), PQ = squared symmetric distance between the two
 *     code words' centroids (rs/quantization/src/pq/mod.rs:231-266), dot = -sum a*b
 *     (rs/utils/src/distance/dot_product.rs:25-27).
 *   - Doc ids are u128 little-endian: mgpu_u128 {lo, hi}.
 *   - Result buffers have fixed stride k; out_counts[b] <= k entries are valid for query b.
 *   - There is no CPU fallback: without a CUDA device mgpu_init fails with MGPU_ERR_NO_DEVICE.
 *   - Calls on one ctx are serialised by an internal mutex (the reference's `&self` search
 *     methods are called concurrently from tokio tasks; use one ctx per worker for concurrency).
 */
#ifndef MUOPDB_GPU_H
#define MUOPDB_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGPU_OK 0
#define MGPU_ERR_INVALID_ARG (-1)
#define MGPU_ERR_OUT_OF_RANGE (-2) /* e.g. nprobe == 0 || nprobe > nlist: the reference panics (ivf/block_based/index.rs:158) */
#define MGPU_ERR_CUDA (-3)
#define MGPU_ERR_OOM (-4)
#define MGPU_ERR_UNSUPPORTED (-5)
#define MGPU_ERR_NO_DEVICE (-6)
#define MGPU_ERR_NCCL (-7)

enum { MGPU_L2 = 0, MGPU_DOT = 1 };            /* DistanceCalculator impls: rs/utils/src/distance/{l2,dot_product}.rs */
enum { MGPU_QUANT_NONE = 0, MGPU_QUANT_PQ = 1 }; /* NoQuantizer / ProductQuantizer: rs/quantization/src/{noq,pq}/mod.rs */
enum { MGPU_HOST = 0, MGPU_DEVICE = 1 };

/* kernel classes for mgpu_profile_* */
enum {
  MGPU_K_COARSE = 0,   /* query x centroid distances */
  MGPU_K_SELECT = 1,   /* per-query top-nprobe */
  MGPU_K_QUANTIZE = 2, /* PQ encode */
  MGPU_K_SCAN = 3,     /* posting-list scan (PQ LUT scan or flat) */
  MGPU_K_FINALIZE = 4, /* exact re-rank + ordering + doc-id remap */
  MGPU_K_HNSW = 5,     /* graph beam search */
  MGPU_K_MERGE = 6,    /* shard top-k merge */
  MGPU_K_OTHER = 7,
  MGPU_K_FALLBACK = 8, /* exact re-scan of the queries the 16-bit PQ scan could not certify (normally an empty launch) */
  MGPU_K_COUNT = 9
};

typedef struct { uint64_t lo, hi; } mgpu_u128;

typedef struct mgpu_ctx mgpu_ctx;
typedef struct mgpu_pq mgpu_pq;
typedef struct mgpu_ivf mgpu_ivf;
typedef struct mgpu_hnsw mgpu_hnsw;
typedef struct mgpu_spann mgpu_spann;
typedef struct mgpu_batcher mgpu_batcher;

/* ---- context ------------------------------------------------------------------------------ */
int mgpu_init(int device, mgpu_ctx **out);
void mgpu_destroy(mgpu_ctx *ctx);
const char *mgpu_last_error(mgpu_ctx *ctx);
const char *mgpu_version(void);
int mgpu_sync(mgpu_ctx *ctx);                    /* wait for everything enqueued on the ctx stream */
void *mgpu_stream(mgpu_ctx *ctx);                /* the cudaStream_t all kernels of this ctx run on */
/* The ctx stream is created non-blocking: it is NOT ordered against the legacy default stream or any other stream.  A caller
 * that produces MGPU_DEVICE inputs (or pre-fills outputs) on its own stream orders them with mgpu_stream_wait (ctx stream waits
 * for the work enqueued on `other` so far) and consumes results with mgpu_stream_signal (`other` waits for the work enqueued on
 * the ctx streams so far) -- or simply calls mgpu_sync.  `other` is a cudaStream_t (NULL = legacy default stream). */
int mgpu_stream_wait(mgpu_ctx *ctx, void *other);
int mgpu_stream_signal(mgpu_ctx *ctx, void *other);
int mgpu_device_sm_count(mgpu_ctx *ctx);

/* CUDA-event timing on the ctx stream (bench.py times device work with these). */
int mgpu_timer_start(mgpu_ctx *ctx);
int mgpu_timer_stop(mgpu_ctx *ctx, float *elapsed_ms); /* records, synchronises, returns elapsed */

/* Per-kernel-class profiling: launches of the enabled classes are bracketed by events.  on > 0: all classes; on < 0: -on is
 * a bit mask of classes (1 << MGPU_K_*); 0: off. */
int mgpu_profile_enable(mgpu_ctx *ctx, int on);
int mgpu_profile_reset(mgpu_ctx *ctx);
int mgpu_profile_get(mgpu_ctx *ctx, int kernel_class, float *total_ms, uint64_t *launches);
uint64_t mgpu_launch_count(mgpu_ctx *ctx);       /* kernels launched by this ctx since creation */
/* Name of the kernel this context launched last in a class (scan and HNSW classes; "" if none yet): which of the
 * alternative kernels of a class served the last call -- bench.py quotes it as roofline.kernel.  Static string. */
const char *mgpu_last_kernel(mgpu_ctx *ctx, int kernel_class);
/* Diagnostic of the tensor-core coarse selection (find_nearest_centroids, index.rs:147-163): out[0] = centroids that fell
 * into the uncertain band tau + 2 eps and were re-scored exactly, out[1] = queries, both summed over every selection on this
 * device since the last reset (reset != 0 clears them).  A band near C per query means un-centred data (|x|^2 >> distances):
 * results stay exact, the selection degenerates to a full exact scoring. */
int mgpu_coarse_band_stats(mgpu_ctx *ctx, uint64_t out[2], int reset);

/* ---- DistanceCalculator (rs/utils/src/lib.rs:17-40) ---------------------------------------- */
/* out[i*nB + j] = calculate(A[i], B[j]) (or calculate_squared when `squared` != 0; dot ignores it),
 * bit-identical to L2DistanceCalculator (distance/l2.rs:30-74) / DotProductDistanceCalculator
 * (distance/dot_product.rs:38-71) including their lane structure and summation order. */
int mgpu_distance_batch(mgpu_ctx *ctx, const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim,
                        int metric, int squared, float *out, int mem);

/* LaneConformingDistanceCalculator<LANES, D>::calculate_squared (rs/utils/src/distance/lane_conforming.rs:16-27), all pairs:
 * LANES accumulators over the whole vector (dim % lanes == 0, lanes in {1,2,4,8,16}), one ordered reduce_sum, then
 * D::outermost_op (L2: identity => squared distance; dot: negation).  Equals mgpu_distance_batch(squared) when dim % 16 == 0
 * and lanes == 16, differs in summation order otherwise.  k-means picks it by dimension (kmeans_builder.rs:126-136). */
int mgpu_distance_batch_lanes(mgpu_ctx *ctx, const float *A, uint64_t nA, const float *B, uint64_t nB, uint32_t dim, int metric,
                              int lanes, float *out, int mem);

/* ---- ProductQuantizer (rs/quantization/src/pq/mod.rs:23-286, Quantizer trait quantization.rs:6-38) */
/* codebook: m * 2^nbits * dsub floats laid out [subspace][centroid][dsub] (pq/mod.rs:155-167), host memory. */
int mgpu_pq_create(mgpu_ctx *ctx, uint32_t dim, uint32_t dsub, uint32_t nbits, const float *codebook, int metric,
                   mgpu_pq **out);
void mgpu_pq_destroy(mgpu_pq *pq);
/* Quantizer::quantize over a batch (pq/mod.rs:152-177): first-minimum argmin per subspace. codes: n x (dim/dsub). */
int mgpu_pq_quantize_batch(mgpu_pq *pq, const float *X, uint64_t n, uint8_t *codes, int mem);
/* Quantizer::distance(a[i], b[i], StreamingSIMD) for n code-word pairs (pq/mod.rs:231-266). */
int mgpu_pq_distance_batch(mgpu_pq *pq, const uint8_t *a, const uint8_t *b, uint64_t n, float *out, int mem);

/* ProductQuantizer::original_vector (pq/mod.rs:184-200) for n code words: out[i] = concatenation of the codebook centroids the
 * code word names (n x dim floats). */
int mgpu_pq_original_vector(mgpu_pq *pq, const uint8_t *codes, uint64_t n, float *out, int mem);

/* ---- BlockBasedIvf<Q> (rs/index/src/ivf/block_based/index.rs:22-471) ---------------------- */
/* Arrays are what BlockBasedIvf::new reads from `index` + `vectors` (ivf/writer.rs:300-353):
 *   centroids      nlist x dim f32
 *   list_offsets   nlist+1 prefix offsets into list_point_ids
 *   list_point_ids point ids (ascending inside a list; a point may appear in several lists)
 *   rows           N x dim f32 (MGPU_QUANT_NONE) or N x (dim/dsub) u8 codes (MGPU_QUANT_PQ), indexed by point id;
 *                  rows_mem says whether `rows` is a host or device pointer
 *   doc_ids        N u128 (NULL => doc id == point id)
 * All other array arguments are host pointers.  `pq` must outlive the index when quant == MGPU_QUANT_PQ. */
int mgpu_ivf_create(mgpu_ctx *ctx, uint32_t dim, uint32_t nlist, const float *centroids, const uint64_t *list_offsets,
                    const uint32_t *list_point_ids, int quant, int metric, mgpu_pq *pq, const void *rows, int rows_mem,
                    uint64_t n, const mgpu_u128 *doc_ids, mgpu_ivf **out);
void mgpu_ivf_destroy(mgpu_ivf *ivf);
uint64_t mgpu_ivf_num_vectors(mgpu_ivf *ivf);  /* index.rs:343-345 */
uint32_t mgpu_ivf_num_clusters(mgpu_ivf *ivf); /* index.rs:334-336 */
/* invalidate_batch by point id (index.rs:430-471 family): skipped before any distance work (index.rs:198-200). */
int mgpu_ivf_invalidate(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n);
int mgpu_ivf_is_invalidated(mgpu_ivf *ivf, uint32_t point_id, int *out);

/* The reference's doc-id keyed forms.  invalidate / invalidate_batch (index.rs:417-452): doc id -> point id through
 * doc_id_to_point_id (built in point-id order, a repeated doc id keeps its last point: index.rs:67-73); unknown doc ids are
 * skipped; out_ok[i] (may be NULL) = 1 when doc_ids[i] was newly invalidated, *out_num_ok (may be NULL) their number.
 * is_invalidated(doc_id) (index.rs:454-459): unknown doc ids report 0.  HOST pointers. */
int mgpu_ivf_invalidate_docs(mgpu_ivf *ivf, const mgpu_u128 *doc_ids, uint32_t n, uint8_t *out_ok, uint32_t *out_num_ok);
int mgpu_ivf_is_doc_invalidated(mgpu_ivf *ivf, const mgpu_u128 *doc_id, int *out);
/* get_point_id (index.rs:469-471): *found = 0 encodes None. */
int mgpu_ivf_get_point_id(mgpu_ivf *ivf, const mgpu_u128 *doc_id, int *found, uint32_t *point_id);
/* get_doc_id / get_doc_ids (index.rs:350-366): doc ids of n point ids in input order; an id >= num_vectors is an error. */
int mgpu_ivf_get_doc_ids(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n, mgpu_u128 *out_doc_ids);
/* get_vector (index.rs:372-384) for n point ids: the stored rows, n x quantized_dimension of u8 (PQ) or f32 (NoQuantizer),
 * read back out of the HBM scan layout. */
int mgpu_ivf_get_vectors(mgpu_ivf *ivf, const uint32_t *point_ids, uint32_t n, void *out_rows);

/* find_nearest_centroids (index.rs:147-163) for B queries: the nprobe smallest sqrt-L2 centroid distances,
 * nearest first, ties by centroid index.  out_ids: B x nprobe; out_dist (may be NULL): B x nprobe. */
int mgpu_ivf_coarse(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t nprobe, uint32_t *out_ids, float *out_dist,
                    int mem);
/* search_with_centroids (index.rs:250-285) for B queries with explicit probe lists.
 * probe_ids: B x max_probes; probe_counts: B (NULL => max_probes each).  Results ordered by (distance, point_id). */
int mgpu_ivf_scan(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                  const uint32_t *probe_counts, uint32_t k, uint32_t *out_point_ids, float *out_scores,
                  uint32_t *out_counts, int mem);
/* search_with_centroids_and_remap (index.rs:298-332): as above, then doc ids, ordered by (score, doc_id). */
int mgpu_ivf_scan_remap(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                        const uint32_t *probe_counts, uint32_t k, mgpu_u128 *out_doc_ids, float *out_scores,
                        uint32_t *out_counts, int mem);
/* BlockBasedIvf::search (index.rs:396-412): coarse + scan + remap. */
int mgpu_ivf_search(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, mgpu_u128 *out_doc_ids,
                    float *out_scores, uint32_t *out_counts, int mem);
/* Pipelined form of mgpu_ivf_search for HOST buffers (Q and the outputs should be page-locked): the call enqueues the H2D
 * copy of Q, the search and the D2H copies of the results and returns at once with a ticket; the outputs are valid after
 * mgpu_search_wait(ctx, ticket).  Two submissions may be in flight per context: batch i+1's upload and batch i-1's
 * download overlap batch i's kernels (the tokio tasks of the reference overlap their awaits the same way,
 * rs/index_server/src/index_server.rs:171-271).  Submitting a third one first completes the oldest.  Q and the output
 * buffers must stay untouched until the wait returns.  Results are identical to mgpu_ivf_search. */
int mgpu_ivf_search_submit(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, mgpu_u128 *out_doc_ids,
                           float *out_scores, uint32_t *out_counts, uint64_t *ticket);
int mgpu_search_wait(mgpu_ctx *ctx, uint64_t ticket);
/* Planner filter hook (index.rs:212-226; Planner::plan_with_ids rs/index/src/query/planner.rs:43-60): the caller evaluates
 * the request's DocumentFilter to a point-id set and passes it as a bitmap (bit p of word p/32 = point id p allowed).  A
 * scanned row is kept only if its bit is set; distances are still computed for every non-invalidated row, as in the
 * reference.  filter_bits: ceil(N/32) words per query, query b at filter_bits + b * filter_stride_words; stride 0 = one
 * bitmap shared by the batch; NULL = no filter.  Same memory space (`mem`) as Q. */
int mgpu_ivf_search_filtered(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, const uint32_t *filter_bits,
                             uint64_t filter_stride_words, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts,
                             int mem);
int mgpu_ivf_scan_remap_filtered(mgpu_ivf *ivf, const float *Q, uint32_t B, const uint32_t *probe_ids, uint32_t max_probes,
                                 const uint32_t *probe_counts, uint32_t k, const uint32_t *filter_bits,
                                 uint64_t filter_stride_words, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts,
                                 int mem);
/* Algorithmic bytes of the last scan on this index (sum over queries of L(q) x bytes/row + 4). */
uint64_t mgpu_ivf_last_scan_bytes(mgpu_ivf *ivf);
uint64_t mgpu_ivf_last_scan_rows(mgpu_ivf *ivf);
/* Queries of the last PQ search that went through the exact fallback scan (not certified by the 16-bit scan, or too many
 * chunks for its table); 0 for the other scan kernels. */
uint64_t mgpu_ivf_last_scan_fallbacks(mgpu_ivf *ivf);

/* ---- Build-time assignment (rs/index/src/ivf/builder.rs:268-366, kmeans_builder.rs:199-221) - */
/* Squared-L2 to every centroid; the max_clusters nearest; keep those with |d - dmin| <= dmin * threshold.
 * out_cids: n x max_clusters (UINT32_MAX padded), out_counts: n. */
int mgpu_ivf_assign(mgpu_ctx *ctx, const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim,
                    uint32_t max_clusters, float threshold, uint32_t *out_cids, uint32_t *out_counts, int mem);

/* Assignment step of KMeansBuilder::run_lloyd (rs/utils/src/kmeans_builder/kmeans_builder.rs:199-221): per row of X
 * argmin_c (T::calculate_squared(x, c) + penalties[c]), folded from (0, f32::MAX) with a strict '<' (first minimum wins).
 * T follows the dimension as in kmeans_builder.rs:126-136: LaneConforming<16|8|4, D> for dim % 16|8|4 == 0, else D.
 * penalties: nlist floats or NULL (= 0); out_labels: n; out_costs (may be NULL): n winning costs (distance + penalty).
 * Large L2 problems run the tensor-core estimate + exact re-score of the band (same labels and costs, bit for bit). */
int mgpu_kmeans_assign(mgpu_ctx *ctx, const float *X, uint64_t n, const float *centroids, uint32_t nlist, uint32_t dim, int metric,
                       const float *penalties, uint32_t *out_labels, float *out_costs, int mem);

/* ---- BlockBasedHnsw<Q> (rs/index/src/hnsw/block_based/index.rs:55-298) --------------------- */
/* Graph arrays exactly as in the `hnsw/index` file (hnsw/block_based/graph_storage.rs:122-193):
 * level_offsets has num_layers+1 entries, top layer first; layer 0 is addressed by point id. */
int mgpu_hnsw_create(mgpu_ctx *ctx, uint32_t dim, uint32_t num_layers, const uint32_t *edges, uint64_t n_edges,
                     const uint32_t *points, uint64_t n_points, const uint64_t *edge_offsets, uint64_t n_edge_offsets,
                     const uint64_t *level_offsets, int quant, int metric, mgpu_pq *pq, const void *rows, int rows_mem,
                     uint64_t n, const mgpu_u128 *doc_ids, mgpu_hnsw **out);
void mgpu_hnsw_destroy(mgpu_hnsw *h);
/* ann_search (index.rs:159-210) for B queries.  out_stats (may be NULL): B x 2 = {#distance evals, #expansions}. */
int mgpu_hnsw_search(mgpu_hnsw *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef, mgpu_u128 *out_doc_ids,
                     float *out_scores, uint32_t *out_counts, uint64_t *out_stats, int mem);
/* Pipelined form over page-locked HOST buffers, as mgpu_ivf_search_submit (ticket completed by mgpu_search_wait; a batch
 * with nothing to search returns ticket 0 and its counts at once).  No traversal statistics. */
int mgpu_hnsw_search_submit(mgpu_hnsw *h, const float *Q, uint32_t B, uint32_t k, uint32_t ef, mgpu_u128 *out_doc_ids,
                            float *out_scores, uint32_t *out_counts, uint64_t *ticket);

/* ---- Spann<Q> (rs/index/src/spann/index.rs:15-266) ---------------------------------------- */
/* centroids: HNSW over the IVF centroids with NoQuantizer<L2> whose doc ids are centroid indices. */
int mgpu_spann_create(mgpu_ctx *ctx, mgpu_hnsw *centroids, mgpu_ivf *posting_lists, mgpu_spann **out);
void mgpu_spann_destroy(mgpu_spann *s);
/* Spann::search (spann/index.rs:211-266) with SearchParams {top_k, ef_construction, num_explored_centroids,
 * centroid_distance_ratio} (rs/config/src/search_params.rs:2-34).  out_counts[b] == UINT32_MAX encodes `None`. */
int mgpu_spann_search(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                      uint32_t num_explored_centroids, float centroid_distance_ratio, mgpu_u128 *out_doc_ids,
                      float *out_scores, uint32_t *out_counts, int mem);
/* Pipelined form over page-locked HOST buffers, as mgpu_ivf_search_submit (ticket completed by mgpu_search_wait). */
int mgpu_spann_search_submit(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids,
                             float centroid_distance_ratio, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts,
                             uint64_t *ticket);

/* Spann::search with Some(planner) (spann/index.rs:253-263): filter semantics as mgpu_ivf_search_filtered. */
int mgpu_spann_search_filtered(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                               uint32_t num_explored_centroids, float centroid_distance_ratio, const uint32_t *filter_bits,
                               uint64_t filter_stride_words, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts,
                               int mem);

/* ---- Micro-batcher (SURVEY.md 8b "Threading", 8f row 4) ------------------------------------- */
/* The reference serves one query per call from many tokio tasks (IndexServer::search, rs/index_server/src/index_server.rs:
 * 171-271 -> Snapshot::search_for_users -> ... -> BlockBasedIvf::search index.rs:396-412 / Spann::search spann/index.rs:211).
 * A batcher keeps that call shape: mgpu_batcher_search blocks with ONE query (thread-safe, call it from spawn_blocking); a
 * worker thread closes a batch at max_batch queries or when its oldest query has waited max_wait_us, runs one batched
 * search and scatters the results into the callers' buffers (out_doc_ids/out_scores: k entries; *out_count as in the
 * batched call, UINT32_MAX = None for SPANN).  filter_bits: the request's planner filter (ceil(N/32) words) or NULL. */
int mgpu_batcher_create(mgpu_ivf *ivf, uint32_t max_batch, uint32_t max_wait_us, uint32_t k, uint32_t nprobe, mgpu_batcher **out);
int mgpu_batcher_create_spann(mgpu_spann *s, uint32_t max_batch, uint32_t max_wait_us, uint32_t top_k, uint32_t ef,
                              uint32_t num_explored_centroids, float centroid_distance_ratio, mgpu_batcher **out);
void mgpu_batcher_destroy(mgpu_batcher *b);  /* drains the queries already submitted, then stops the worker */
int mgpu_batcher_search(mgpu_batcher *b, const float *query, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_count);
int mgpu_batcher_search_filtered(mgpu_batcher *b, const float *query, const uint32_t *filter_bits, mgpu_u128 *out_doc_ids,
                                 float *out_scores, uint32_t *out_count);
/* stats = {queries served, batches launched, largest batch, batches closed because they were full} */
int mgpu_batcher_stats(mgpu_batcher *b, uint64_t stats[4]);

/* ---- Cross-segment / cross-shard merge (rs/index/src/collection/snapshot.rs:49-63,79-108) -- */
/* Concatenate S partial results per query, sort by (score, doc_id) (utils.rs:95-114), truncate to k.
 * doc_ids/scores: S x B x k, counts: S x B. */
int mgpu_merge_topk(mgpu_ctx *ctx, const mgpu_u128 *doc_ids, const float *scores, const uint32_t *counts, uint32_t S,
                    uint32_t B, uint32_t k, mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem);

/* Multi-GPU: one process per GPU, each holding one doc-shard (rs/aggregator/src/aggregator.rs:81-132 is the
 * semantic model).  mgpu_comm_* wraps an NCCL communicator (libnccl is dlopen'ed on first use). */
int mgpu_comm_unique_id(uint8_t out_id[128]);
int mgpu_comm_init(mgpu_ctx *ctx, int nranks, int rank, const uint8_t id[128]);
int mgpu_comm_destroy(mgpu_ctx *ctx);
/* All-gather every rank's B x k partial result over NVLink and merge locally (every rank gets the merged top-k).
 * Buffers are DEVICE pointers. */
int mgpu_shard_allgather_merge(mgpu_ctx *ctx, const mgpu_u128 *local_doc_ids, const float *local_scores,
                               const uint32_t *local_counts, uint32_t B, uint32_t k, mgpu_u128 *out_doc_ids,
                               float *out_scores, uint32_t *out_counts);

/* The sharded query path in one call: every rank passes the same (replicated) batch, searches its own shard with
 * BlockBasedIvf::search semantics, the per-shard top-k lists are all-gathered over NVLink and merged; every rank receives
 * the merged result.  shared_codebook != 0 states that all shards use one PQ codebook: the query encode (index.rs:193) is
 * then split across the ranks and its B x m code bytes all-gathered instead of being repeated on every rank.
 * With MGPU_HOST buffers each rank uploads only its 1/N slice of Q and the slices are all-gathered over NVLink, so the batch
 * must really be identical on every rank.
 * Collective: all ranks must call it with the same B, k.  Buffers in `mem` space. */
int mgpu_shard_ivf_search(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, int shared_codebook,
                          mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, int mem);
/* Pipelined form over page-locked HOST buffers (see mgpu_ivf_search_submit): returns a ticket, mgpu_search_wait completes it.
 * Still a collective: every rank submits the same sequence of batches. */
int mgpu_shard_ivf_search_submit(mgpu_ivf *ivf, const float *Q, uint32_t B, uint32_t k, uint32_t nprobe, int shared_codebook,
                                 mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, uint64_t *ticket);

/* The result exchange of the sharded calls (all-gather + merge) runs on its own stream (and communicator), so in the pipelined
 * *_submit forms batch i's exchange always overlaps batch i+1's kernels.  For DEVICE buffers the default keeps the ctx stream
 * ordered after the exchange (work enqueued on mgpu_stream() afterwards sees the merged result).  on != 0 drops that wait:
 * consecutive sharded calls then overlap the same way, and results are complete after mgpu_sync (which waits for both
 * streams) -- the caller must not reuse an output buffer before that. */
int mgpu_shard_overlap(mgpu_ctx *ctx, int on);

/* Config 5: every rank holds the SPANN index of one doc-shard (centroid HNSW + posting lists), all ranks pass
 * the same replicated batch; per rank Spann::search (spann/index.rs:211-266), then the per-shard results are all-gathered over
 * NVLink and merged by (score, doc_id) (collection/snapshot.rs:49-63; fan-out model rs/aggregator/src/aggregator.rs:81-132).
 * A shard that answers None (out_counts == UINT32_MAX) contributes nothing; the merged count is UINT32_MAX only if every shard
 * answered None.  shared_codebook as in mgpu_shard_ivf_search.  Collective; buffers in `mem` space. */
int mgpu_shard_spann_search(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef, uint32_t num_explored_centroids,
                            float centroid_distance_ratio, int shared_codebook, mgpu_u128 *out_doc_ids, float *out_scores,
                            uint32_t *out_counts, int mem);
/* Pipelined form over page-locked HOST buffers (ticket completed by mgpu_search_wait), as mgpu_shard_ivf_search_submit. */
int mgpu_shard_spann_search_submit(mgpu_spann *s, const float *Q, uint32_t B, uint32_t top_k, uint32_t ef,
                                   uint32_t num_explored_centroids, float centroid_distance_ratio, int shared_codebook,
                                   mgpu_u128 *out_doc_ids, float *out_scores, uint32_t *out_counts, uint64_t *ticket);

/* ---- Readers of the reference's on-disk formats -------------- */
/* One Elias-Fano posting-list payload (rs/compression/src/elias_fano/ef.rs:197-215; decode rule
 * block_based_decoder.rs:162-179,257-266) -> ascending values.  Host-only (no device needed).
 * Returns the element count or -1 on malformed input / insufficient capacity. */
int64_t mgpu_ef_decode(const uint8_t *payload, uint64_t len, uint64_t *out, uint64_t cap);
/* ProductQuantizerReader::read (rs/quantization/src/pq/mod.rs:52-136): `product_quantizer_config.yaml` + `codebook`. */
int mgpu_pq_load(mgpu_ctx *ctx, const char *quantizer_dir, int metric, mgpu_pq **out);
/* BlockBasedIvf::new_with_offset (rs/index/src/ivf/block_based/index.rs:95-138): `{base}/index` + `{base}/vectors`
 * (layouts: ivf/writer.rs:300-353, ivf/block_based/storage.rs:52-151); byte offsets address a user's section inside
 * the shared multi-user files (multi_spann/user_index_info.rs:4-14), 0 for single-index directories. */
int mgpu_ivf_load(mgpu_ctx *ctx, const char *base_dir, uint64_t index_offset, uint64_t vector_offset, int quant, int metric,
                  mgpu_pq *pq, mgpu_ivf **out);
/* BlockBasedHnsw::new_with_offsets (rs/index/src/hnsw/block_based/index.rs:94-140): `{base}/hnsw/index` +
 * `{base}/hnsw/vector_storage` (hnsw/writer.rs:206-265, graph_storage.rs:122-193).  dim = original dimension. */
int mgpu_hnsw_load(mgpu_ctx *ctx, const char *base_dir, uint64_t index_offset, uint64_t vector_offset, uint32_t dim, int quant,
                   int metric, mgpu_pq *pq, mgpu_hnsw **out);
/* Multi-user SPANN files (rs/index/src/multi_spann/{writer,reader,index}.rs): all users' sections are packed into
 * {base}/centroids/hnsw/{index,vector_storage} and {base}/ivf/{index,vectors,raw_vectors,quantizer/codebook}; the file
 * {base}/user_index_info maps user id -> byte offsets (UserIndexInfo, multi_spann/user_index_info.rs:4-14). */
typedef struct {
  mgpu_u128 user_id;
  uint64_t centroid_vector_offset, centroid_vector_len, centroid_index_offset, centroid_index_len;
  uint64_t ivf_vectors_offset, ivf_vectors_len, ivf_raw_vectors_offset, ivf_raw_vectors_len;
  uint64_t ivf_index_offset, ivf_index_len, ivf_pq_codebook_offset, ivf_pq_codebook_len;
} mgpu_user_index_info;
/* UserIndexInfo::from_le_bytes / to_le_bytes (user_index_info.rs:26-83): the 112-byte record. */
int mgpu_user_index_info_decode(const uint8_t bytes[112], mgpu_user_index_info *out);
int mgpu_user_index_info_encode(const mgpu_user_index_info *info, uint8_t out_bytes[112]);
/* Every record of a `user_index_info` table file (an odht 0.3.1 HashTableOwned<HashConfig> image: 32-byte header, slot_count
 * entries of 16-byte key + 112-byte value, then slot_count + 16 control bytes whose top bit marks an empty slot).  odht is a
 * third-party crate absent from the reference checkout: its layout is restated from the published format and checked by the
 * header's own item count and by key == value.user_id.  Returns the number of records (up to `cap` are stored), -1 on
 * malformed input. */
int64_t mgpu_user_index_info_read(const char *path, mgpu_user_index_info *out, uint64_t cap);
/* MultiSpannIndex::get_or_create_index (multi_spann/index.rs:100-128) -> SpannReader::new_with_offsets(...).read
 * (spann/reader.rs:43-82): opens one user's centroid HNSW (NoQuantizer<L2>) and posting lists at the recorded offsets.
 * quant == MGPU_QUANT_PQ reads {base}/ivf/quantizer/product_quantizer_config.yaml and the user's slice of `codebook`
 * (ivf_pq_codebook_offset/len); *out_pq then owns that quantizer (NULL otherwise).  Free with mgpu_spann_destroy,
 * mgpu_ivf_destroy, mgpu_hnsw_destroy, mgpu_pq_destroy. */
int mgpu_spann_load_user(mgpu_ctx *ctx, const char *base_dir, const mgpu_user_index_info *info, uint32_t dim, int quant, int metric,
                         mgpu_pq **out_pq, mgpu_hnsw **out_centroids, mgpu_ivf **out_lists, mgpu_spann **out_spann);
/* sizes = {num_layers, n_edges, n_points, n_edge_offsets, n, entry_point}; copy_graph reads the resident arrays back. */
int mgpu_hnsw_info(mgpu_hnsw *h, uint64_t sizes[6]);
int mgpu_hnsw_copy_graph(mgpu_hnsw *h, uint32_t *edges, uint32_t *points, uint64_t *edge_offsets, uint64_t *level_offsets);

#ifdef __cplusplus
}
#endif
#endif /* MUOPDB_GPU_H */

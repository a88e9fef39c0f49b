// internal.cuh -- handle layouts and kernel launcher prototypes (host side).
#pragma once
#include "common.cuh"

// ---- ProductQuantizer resident state ----------------------------------------------------------
struct mgpu_pq {
  mgpu_ctx *ctx;
  uint32_t dim, dsub, nbits, m, K;
  int metric;
  float *d_cb = nullptr;      // [m][K][dsub] as in pq/mod.rs:155-167
  float *d_table = nullptr;   // [m][K][K] score contribution of (query code a, row code b): L2 -> ||ca-cb||^2, dot -> -<ca,cb>
  float *d_rowmin = nullptr;  // [m][K] min_b table[m][a][b]
  float *d_rowmax = nullptr;  // [m][K]
  // 16-bit fixed-point image of the table with ONE scale for the whole quantizer (K == 256 only):
  //   table16[s][a][b] = rn((table[s][a][b] - rowmin[s][a]) * gscale),  gscale = 65535 / max_{s,a} (rowmax - rowmin)
  // A query's LUT is then a pure gather of m rows of 512 bytes (no per-query scale, no arithmetic): scan_pq16.cu.
  uint16_t *d_table16 = nullptr;
  float gscale = 0.0f;
};

// ---- IVF resident state -------------------------------------------------------------------------
// HBM layout: every posting list is padded to a multiple of 32 rows ("chunks"); slot = chunk*32 + lane.
//   slot_pid[slot]          point id of the row in that slot, MGPU_EMPTY_SLOT for padding
//   PQ fast layout (m % 32 == 0, nbits == 8): per chunk, per 32-subspace group g, two 512-byte units;
//       unit u holds for lane l the 16 bytes t = 16u .. 16u+15 where byte t = code of subspace 32g + (l ^ t).
//       A warp reads a unit with one fully coalesced 16-byte-per-lane load, and at step t the 32 lanes touch 32
//       different LUT columns (l ^ t) -> shared-memory lookups are bank-conflict free for ANY code values.
//       The chunk's 32 point ids follow its code units (one 128-byte line), so a chunk is ONE contiguous record of
//       pq_fast_chunk_bytes(ng) = ng * 1024 + 128 bytes: the scan fetches it with a single bulk copy (cp.async.bulk).
//   PQ generic layout: codes_rm[slot][m] row-major.
//   flat layout: per chunk [dim4][32 lanes] float4 (dim padded to a multiple of 4): lane = row, coalesced 512 B reads.
struct mgpu_ivf {
  mgpu_ctx *ctx;
  uint32_t dim, nlist;
  uint64_t n;
  int quant, metric;
  mgpu_pq *pq = nullptr;
  float *d_centroids = nullptr;
  void *d_csplit = nullptr;           // centroids as [hi | lo | hi] bf16 rows (tensor-core coarse scoring), may be null
  float *d_cn = nullptr;              // |c|^2 per centroid
  float cn_max = 0.f;
  uint32_t *d_chunk_start = nullptr;  // nlist+1, in chunks
  uint32_t *d_list_len = nullptr;     // nlist
  uint64_t total_chunks = 0;
  uint32_t *d_slot_pid = nullptr;
  bool pq_fast = false;
  uint32_t ng = 0;                    // m / 32 when pq_fast
  uint8_t *d_codes = nullptr;         // fast or generic layout
  float *d_rows = nullptr;            // flat interleaved
  uint32_t dim4 = 0;
  mgpu_u128 *d_doc_ids = nullptr;     // null => identity
  uint32_t *d_invalid = nullptr;      // bitmap over point ids
  uint64_t n_invalid = 0;
  unsigned long long *d_scan_rows = nullptr;  // [0] rows scanned by the last scan, [1] query scheduler counter
  uint64_t bytes_per_row = 0;
  std::vector<uint32_t> h_list_len;
  uint32_t *d_scan_overflow = nullptr;   // [0] count, [1..] query ids deferred by the table-driven PQ scan
  uint32_t scan_overflow_cap = 0;
  uint32_t scan_bound_probes = 0;        // cache: upper bound of chunks any `scan_bound_probes` lists can hold
  uint64_t scan_bound_chunks = 0;
  // accessors (index.rs:350-384,469-471), built on first use
  uint32_t *d_qstate = nullptr;          // per query: 1 = deferred by the 16-bit scan (scan_pq16.cu)
  uint32_t qstate_cap = 0;
  bool last_scan_was16 = false;
  uint32_t *d_pid_slot = nullptr;        // point id -> one slot holding its row (MGPU_EMPTY_SLOT: in no list)
  void *doc_map = nullptr;               // host std::unordered_map doc id -> point id (doc_id_to_point_id, index.rs:67-73)
};

// PQ fast layout: bytes / 16-byte words of one chunk record (code units of the ng groups, then the 32 point ids)
__host__ __device__ inline size_t pq_fast_chunk_bytes(uint32_t ng) { return (size_t)ng * 1024 + 128; }
__host__ __device__ inline size_t pq_fast_chunk_u4(uint32_t ng) { return (size_t)ng * 64 + 8; }
// byte offset of the code of subspace s of the row in lane l of chunk `chunk`
__host__ __device__ inline size_t pq_fast_code_offset(size_t chunk, uint32_t ng, uint32_t l, uint32_t s) {
  const uint32_t g = s >> 5, t = (s & 31) ^ l;
  return chunk * pq_fast_chunk_bytes(ng) + ((size_t)g * 2 + (t >> 4)) * 512 + l * 16 + (t & 15);
}

struct mgpu_hnsw {
  mgpu_ctx *ctx;
  uint32_t dim, num_layers;
  uint64_t n, n_edges, n_points, n_edge_offsets;
  int quant, metric;
  mgpu_pq *pq = nullptr;
  uint32_t *d_edges = nullptr, *d_points = nullptr;
  uint64_t *d_edge_offsets = nullptr, *d_level_offsets = nullptr;
  void *d_rows = nullptr;  // row-major, indexed by point id
  uint32_t qdim = 0;
  mgpu_u128 *d_doc_ids = nullptr;
  uint32_t entry_point = 0;
  uint32_t max_degree = 0;
  std::vector<uint64_t> h_level_offsets;
  // upper-layer point -> position lookup: sorted (point, pos) pairs per layer
  uint32_t *d_upper_sorted_pid = nullptr, *d_upper_sorted_pos = nullptr;
  int32_t *d_upper_dense = nullptr;  // dense (layer, point) -> position map for the upper layers (optional)
  // layer-0 adjacency in fixed-stride rows (optional): edges0[p * deg0 .. ) = the edges of point p in stored order, padded
  // with 0xFFFFFFFF.  One load instead of edge_offsets -> edges (two dependent loads) per expansion.
  uint32_t *d_edges0 = nullptr;
  uint32_t deg0 = 0;
};

struct mgpu_spann {
  mgpu_ctx *ctx;
  mgpu_hnsw *centroids;
  mgpu_ivf *lists;
};

// ---- candidates produced by a scan: per query 32 (composite key, slot) pairs -------------------
#define MGPU_NCAND 32
#define MGPU_MAX_K 2048   /* k > 32 (PQ: k > 16) runs scan rounds until k + 16 candidates are reported (api.cu: ivf_scan_dev) */
#define MGPU_ROUND_SPARE 16

struct ScanArgs {
  // index
  const uint32_t *chunk_start, *list_len, *slot_pid, *invalid;
  const uint8_t *codes; const float *rows;
  uint32_t dim, dim4, m, K, ng;
  // pq
  const float *table, *rowmin, *rowmax;
  const uint8_t *qcodes;  // B x m
  // queries
  const float *Q; uint32_t B;
  const uint32_t *probes; uint32_t max_probes; const uint32_t *probe_counts;
  // out
  uint64_t *cand_key; uint32_t *cand_slot;  // B x 32
  unsigned long long *rows_scanned;
  unsigned int *next_query;  // zeroed before every launch: dynamic query scheduler
  const uint32_t *order;     // optional: queries in descending-work order (k_plan_queries)
  // planner filter hook (index.rs:212-226): when non-null a scanned row survives only if bit `point id` of its query's
  // bitmap is set; query q uses filter + q * filter_stride (stride 0 = one bitmap shared by the whole batch)
  const uint32_t *filter; uint64_t filter_stride;
  // k > 32 (multi-round top-k, api.cu): when non-null, a row whose composite (key << 32 | point id) is below
  // lower_bound[q] is invisible to this scan (it was reported by an earlier round)
  const uint64_t *lower_bound;
  // PQ table-driven scan: queries with more chunks than the shared-memory chunk table are deferred to a second launch
  unsigned int *overflow_count; uint32_t *overflow_list;
  int metric;
  // exact fallback (scan.cu k_scan_pq_exact): rows scored with ProductQuantizer::distance itself; the queries come from
  // overflow_list[0 .. *overflow_count)
  int from_list;
  const float *cb; uint32_t dsub;
};

int launch_pq_build_table(mgpu_pq *pq);
int launch_pq_build_table16(mgpu_pq *pq);
int launch_pq_quantize(mgpu_pq *pq, const float *dX, uint64_t n, uint8_t *dcodes, cudaStream_t st = nullptr);
int launch_pq_distance_pairs(mgpu_pq *pq, const uint8_t *da, const uint8_t *db, uint64_t n, float *dout);
int launch_build_layout(mgpu_ivf *ivf, const void *d_rows_by_pid);
int launch_scan(mgpu_ivf *ivf, const ScanArgs &a);
int launch_plan_queries(mgpu_ivf *ivf, const uint32_t *d_probes, uint32_t max_probes, const uint32_t *d_counts, uint32_t B,
                        uint32_t *d_order, uint32_t *d_work, bool have_work = false);
int launch_scan_pq_db(mgpu_ivf *ivf, const ScanArgs &a);  // MGPU_ERR_UNSUPPORTED => use launch_scan's generic kernels
bool scan_pq16_applicable(mgpu_ivf *ivf, const ScanArgs &a);
int launch_scan_pq16(mgpu_ivf *ivf, const ScanArgs &a, uint32_t *d_qstate);
size_t scan_max_probes_supported(mgpu_ivf *ivf);

struct FinalizeArgs {
  const uint64_t *cand_key; const uint32_t *cand_slot; uint32_t B, k;
  const uint32_t *slot_pid;
  // exact PQ re-rank inputs (null for flat: keys are already exact)
  const float *cb; const uint8_t *codes; const uint8_t *qcodes; uint32_t m, K, dsub, ng; bool pq_fast; int metric;
  const mgpu_u128 *doc_ids;
  // L2 product quantizer only: the fixed-point ranking key of a row is its exact score times the query's scale to within
  // ~1e-5 relative (every LUT row's minimum is table[s][a][a] = 0, so no offset is involved), hence a candidate whose key
  // exceeds the k-th smallest key by more than 2^-12 relative + 8192 units cannot be among the exact top k and is not
  // re-scored
  bool prune;
  // Candidates of the 16-bit-LUT scan (scan_pq16.cu): key = sum of u16 table entries with the quantizer-wide scale, L2 only.
  //   |key - gscale * exact score| <= E(key) = 0.51 m + 2 + 1.2e-5 key   (entry rounding + fp32 summation-order slack)
  // so a candidate with key > (k-th smallest key) + 2 E cannot reach the exact top k (prune), and the answer is CERTIFIED
  // exact when gscale * (exact score of the k-th best candidate) + E < (32nd key): no row outside the 32 candidates can beat
  // it.  A query that cannot be certified is appended to uncert_list (count in uncert_count) for the exact fallback scan.
  // qstate[q] != 0: the scan deferred this query itself (its candidate list was never written).
  int key16;
  float gscale;                // table16 scale of the quantizer
  float cert_slack;            // multiplies E (test hook: a huge value sends every full query through the fallback)
  const uint32_t *qstate;
  uint32_t *uncert_count, *uncert_list;
  // outputs (either may be null)
  uint32_t *out_pids; mgpu_u128 *out_docs; float *out_scores; uint32_t *out_counts;
};
__host__ __device__ inline uint64_t key16_err(uint32_t m, uint32_t key, float slack) {
  const float e = (0.51f * (float)m + 2.0f + 1.2e-5f * (float)key) * slack;
  return e >= 4.0e9f ? 0xFFFFFFFFull : (uint64_t)e + 1ull;
}
int launch_finalize(mgpu_ctx *ctx, const FinalizeArgs &a);
int launch_scan_pq_exact_list(mgpu_ivf *ivf, const ScanArgs &a, const FinalizeArgs &f);

int launch_distance_matrix(mgpu_ctx *ctx, const float *dA, uint64_t nA, const float *dB, uint64_t nB, uint32_t dim,
                           int metric, int mode /*0 squared,1 sqrt*/, float *dout, int kernel_class);
int launch_select_smallest(mgpu_ctx *ctx, const float *dD, uint32_t B, uint32_t C, uint32_t nsel, uint32_t *out_ids,
                           float *out_vals);
int launch_assign_filter(mgpu_ctx *ctx, const uint32_t *sel_ids, const float *sel_vals, uint64_t n, uint32_t r,
                         float threshold, uint32_t *out_cids, uint32_t *out_counts);
int launch_merge_topk(mgpu_ctx *ctx, const mgpu_u128 *docs, const float *scores, const uint32_t *counts, uint32_t S,
                      uint32_t B, uint32_t k, mgpu_u128 *out_docs, float *out_scores, uint32_t *out_counts, cudaStream_t st = nullptr);
// multi-round top-k helpers (finalize.cu)
int launch_round_prepare(mgpu_ctx *ctx, uint64_t *cand_key, uint32_t *cand_slot, uint32_t B, uint64_t *lower_bound,
                         uint32_t *reported, uint32_t want, uint32_t *unfinished);
int launch_merge_rounds(mgpu_ctx *ctx, const uint32_t *pids, const float *scores, const uint32_t *counts, uint32_t R, uint32_t B,
                        uint32_t k, const mgpu_u128 *doc_ids, uint32_t *out_pids, mgpu_u128 *out_docs, float *out_scores,
                        uint32_t *out_counts);

// tensor-core coarse scoring (coarse_tc.cu)
uint32_t coarse_tc_kp(uint32_t dim);
bool coarse_tc_applicable(mgpu_ctx *ctx, uint32_t dim, uint32_t C, uint32_t nprobe);
int launch_split_bf16(mgpu_ctx *ctx, const float *dX, uint64_t n, uint32_t dim, int is_centroid, void *d_out, float *d_norms);
int coarse_band_stats(mgpu_ctx *ctx, uint64_t out[2], int reset);
int launch_coarse_tc(mgpu_ctx *ctx, const float *dQ, uint32_t B, const float *d_centroids, const void *d_csplit, const float *d_cn,
                     float cn_max, uint32_t C, uint32_t dim, uint32_t nprobe, void *d_qsplit, float *d_qn, float *d_Dt,
                     uint32_t *d_overflow, uint32_t *d_flags, int need_order, uint32_t *out_ids, float *out_dist,
                     const uint32_t *chunk_start = nullptr, uint32_t *d_work = nullptr, bool *work_done = nullptr,
                     cudaEvent_t after_gemm = nullptr);

int launch_tc_distances(mgpu_ctx *ctx, const float *dX, uint32_t B, const void *d_csplit, const float *d_cn, uint32_t C, uint32_t dim,
                        void *d_xsplit, float *d_xn, float *d_Dt);

// build-time kernels (assign.cu)
int launch_distance_lanes(mgpu_ctx *ctx, const float *dA, uint64_t nA, const float *dB, uint64_t nB, uint32_t dim, int metric,
                          int lanes, float *dout, int kernel_class);
int launch_kmeans_argmin(mgpu_ctx *ctx, const float *dD, uint64_t n, uint32_t C, const float *d_pen, uint32_t *d_labels, float *d_costs);
int launch_kmeans_pick_tc(mgpu_ctx *ctx, const float *d_Dt, const float *dX, const float *d_centroids, const float *d_xn,
                          const float *d_cn_max, uint64_t n, uint32_t C, uint32_t dim, const float *d_pen, uint32_t *d_labels,
                          float *d_costs);
int launch_max_f32(mgpu_ctx *ctx, const float *d_v, uint32_t n, float *d_out);
int launch_pq_original(mgpu_pq *pq, const uint8_t *d_codes, uint64_t n, float *d_out);
int launch_gather_docs(mgpu_ctx *ctx, const mgpu_u128 *d_doc_ids, const uint32_t *d_pids, uint32_t n, uint64_t nvec, mgpu_u128 *d_out,
                       uint32_t *d_bad);
int launch_gather_rows(mgpu_ivf *ivf, const uint32_t *d_pids, uint32_t n, void *d_out, uint32_t *d_bad);
int ivf_ensure_pid_slot(mgpu_ivf *ivf);
void ivf_free_doc_map(mgpu_ivf *ivf);

struct HnswSearchArgs {
  const float *Q; uint32_t B, k, ef;
  mgpu_u128 *out_docs; float *out_scores; uint32_t *out_counts; uint64_t *out_stats;
  uint32_t *out_pids;  // optional
};
int launch_hnsw_search(mgpu_hnsw *h, const HnswSearchArgs &a);

//! Rust binding of include/muopdb_gpu.h -- the shim a MuopDB maintainer adds (e.g. as `rs/gpu/src/lib.rs`) to put the
//! B200 search path behind the existing traits and query structs.
//!
//!   * the `extern "C"` block binds EVERY entry point of the header; it is generated from the header by
//!     tools/gen_rust_ffi.py and tests/test_abi.py fails when the two disagree;
//!   * `GpuL2` / `GpuDot` implement `utils::DistanceCalculator` + `CalculateSquared` (rs/utils/src/lib.rs:17-40),
//!     `GpuLaneConforming<LANES, D>` implements `CalculateSquared` (rs/utils/src/distance/lane_conforming.rs:16-27);
//!   * `GpuProductQuantizer` implements `quantization::Quantizer` (rs/quantization/src/quantization.rs:6-38);
//!   * `GpuIvf` / `GpuHnsw` / `GpuSpann` carry the inherent methods of `BlockBasedIvf<Q>` / `BlockBasedHnsw<Q>` / `Spann<Q>`
//!     with the reference's signatures (batch forms next to them), `GpuBatcher` is the per-request front door.
//!
//! SOURCE ONLY: this build image has no rustc/cargo, so the file is kept in step with the header mechanically but was not
//! compiled here.  The scalar trait methods (`calculate(a, b)`, `distance(a, b)`) go through the
//! batch entry points with n = 1: they exist for API completeness (tests, tooling); the hot path calls the batched forms.
#![allow(non_camel_case_types, non_snake_case, clippy::too_many_arguments)]
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_float, c_int, c_void};
use std::sync::OnceLock;

use anyhow::{anyhow, Result};

#[repr(C)] pub struct mgpu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_pq { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_ivf { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_hnsw { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_spann { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_batcher { _p: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy, Default, PartialEq, Eq, Debug)] pub struct mgpu_u128 { pub lo: u64, pub hi: u64 }
/// multi_spann/user_index_info.rs:4-18
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct mgpu_user_index_info {
    pub user_id: mgpu_u128,
    pub centroid_vector_offset: u64, pub centroid_vector_len: u64, pub centroid_index_offset: u64, pub centroid_index_len: u64,
    pub ivf_vectors_offset: u64, pub ivf_vectors_len: u64, pub ivf_raw_vectors_offset: u64, pub ivf_raw_vectors_len: u64,
    pub ivf_index_offset: u64, pub ivf_index_len: u64, pub ivf_pq_codebook_offset: u64, pub ivf_pq_codebook_len: u64,
}

pub const MGPU_OK: c_int = 0;
pub const MGPU_ERR_INVALID_ARG: c_int = -1;
pub const MGPU_ERR_OUT_OF_RANGE: c_int = -2;
pub const MGPU_ERR_CUDA: c_int = -3;
pub const MGPU_ERR_OOM: c_int = -4;
pub const MGPU_ERR_UNSUPPORTED: c_int = -5;
pub const MGPU_ERR_NO_DEVICE: c_int = -6;
pub const MGPU_ERR_NCCL: c_int = -7;
pub const MGPU_L2: c_int = 0;
pub const MGPU_DOT: c_int = 1;
pub const MGPU_QUANT_NONE: c_int = 0;
pub const MGPU_QUANT_PQ: c_int = 1;
pub const MGPU_HOST: c_int = 0;
pub const MGPU_DEVICE: c_int = 1;

#[link(name = "muopdb_gpu")]
extern "C" {
    // ---- BEGIN GENERATED (tools/gen_rust_ffi.py) ----
    pub fn mgpu_init(device: c_int, out: *mut *mut mgpu_ctx) -> c_int;
    pub fn mgpu_destroy(ctx: *mut mgpu_ctx);
    pub fn mgpu_last_error(ctx: *mut mgpu_ctx) -> *const c_char;
    pub fn mgpu_version() -> *const c_char;
    pub fn mgpu_sync(ctx: *mut mgpu_ctx) -> c_int;
    pub fn mgpu_stream(ctx: *mut mgpu_ctx) -> *mut c_void;
    pub fn mgpu_stream_wait(ctx: *mut mgpu_ctx, other: *mut c_void) -> c_int;
    pub fn mgpu_stream_signal(ctx: *mut mgpu_ctx, other: *mut c_void) -> c_int;
    pub fn mgpu_device_sm_count(ctx: *mut mgpu_ctx) -> c_int;
    pub fn mgpu_timer_start(ctx: *mut mgpu_ctx) -> c_int;
    pub fn mgpu_timer_stop(ctx: *mut mgpu_ctx, elapsed_ms: *mut c_float) -> c_int;
    pub fn mgpu_profile_enable(ctx: *mut mgpu_ctx, on: c_int) -> c_int;
    pub fn mgpu_profile_reset(ctx: *mut mgpu_ctx) -> c_int;
    pub fn mgpu_profile_get(ctx: *mut mgpu_ctx, kernel_class: c_int, total_ms: *mut c_float, launches: *mut u64) -> c_int;
    pub fn mgpu_launch_count(ctx: *mut mgpu_ctx) -> u64;
    pub fn mgpu_last_kernel(ctx: *mut mgpu_ctx, kernel_class: c_int) -> *const c_char;
    pub fn mgpu_coarse_band_stats(ctx: *mut mgpu_ctx, out: *mut u64, reset: c_int) -> c_int;
    pub fn mgpu_distance_batch(ctx: *mut mgpu_ctx, A: *const c_float, nA: u64, B: *const c_float, nB: u64, dim: u32,
        metric: c_int, squared: c_int, out: *mut c_float, mem: c_int) -> c_int;
    pub fn mgpu_distance_batch_lanes(ctx: *mut mgpu_ctx, A: *const c_float, nA: u64, B: *const c_float, nB: u64, dim: u32,
        metric: c_int, lanes: c_int, out: *mut c_float, mem: c_int) -> c_int;
    pub fn mgpu_pq_create(ctx: *mut mgpu_ctx, dim: u32, dsub: u32, nbits: u32, codebook: *const c_float, metric: c_int,
        out: *mut *mut mgpu_pq) -> c_int;
    pub fn mgpu_pq_destroy(pq: *mut mgpu_pq);
    pub fn mgpu_pq_quantize_batch(pq: *mut mgpu_pq, X: *const c_float, n: u64, codes: *mut u8, mem: c_int) -> c_int;
    pub fn mgpu_pq_distance_batch(pq: *mut mgpu_pq, a: *const u8, b: *const u8, n: u64, out: *mut c_float,
        mem: c_int) -> c_int;
    pub fn mgpu_pq_original_vector(pq: *mut mgpu_pq, codes: *const u8, n: u64, out: *mut c_float, mem: c_int) -> c_int;
    pub fn mgpu_ivf_create(ctx: *mut mgpu_ctx, dim: u32, nlist: u32, centroids: *const c_float, list_offsets: *const u64,
        list_point_ids: *const u32, quant: c_int, metric: c_int, pq: *mut mgpu_pq, rows: *const c_void, rows_mem: c_int,
        n: u64, doc_ids: *const mgpu_u128, out: *mut *mut mgpu_ivf) -> c_int;
    pub fn mgpu_ivf_destroy(ivf: *mut mgpu_ivf);
    pub fn mgpu_ivf_num_vectors(ivf: *mut mgpu_ivf) -> u64;
    pub fn mgpu_ivf_num_clusters(ivf: *mut mgpu_ivf) -> u32;
    pub fn mgpu_ivf_invalidate(ivf: *mut mgpu_ivf, point_ids: *const u32, n: u32) -> c_int;
    pub fn mgpu_ivf_is_invalidated(ivf: *mut mgpu_ivf, point_id: u32, out: *mut c_int) -> c_int;
    pub fn mgpu_ivf_invalidate_docs(ivf: *mut mgpu_ivf, doc_ids: *const mgpu_u128, n: u32, out_ok: *mut u8,
        out_num_ok: *mut u32) -> c_int;
    pub fn mgpu_ivf_is_doc_invalidated(ivf: *mut mgpu_ivf, doc_id: *const mgpu_u128, out: *mut c_int) -> c_int;
    pub fn mgpu_ivf_get_point_id(ivf: *mut mgpu_ivf, doc_id: *const mgpu_u128, found: *mut c_int,
        point_id: *mut u32) -> c_int;
    pub fn mgpu_ivf_get_doc_ids(ivf: *mut mgpu_ivf, point_ids: *const u32, n: u32, out_doc_ids: *mut mgpu_u128) -> c_int;
    pub fn mgpu_ivf_get_vectors(ivf: *mut mgpu_ivf, point_ids: *const u32, n: u32, out_rows: *mut c_void) -> c_int;
    pub fn mgpu_ivf_coarse(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, nprobe: u32, out_ids: *mut u32,
        out_dist: *mut c_float, mem: c_int) -> c_int;
    pub fn mgpu_ivf_scan(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, probe_ids: *const u32, max_probes: u32,
        probe_counts: *const u32, k: u32, out_point_ids: *mut u32, out_scores: *mut c_float, out_counts: *mut u32,
        mem: c_int) -> c_int;
    pub fn mgpu_ivf_scan_remap(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, probe_ids: *const u32, max_probes: u32,
        probe_counts: *const u32, k: u32, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32,
        mem: c_int) -> c_int;
    pub fn mgpu_ivf_search(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, k: u32, nprobe: u32, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_ivf_search_submit(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, k: u32, nprobe: u32,
        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32, ticket: *mut u64) -> c_int;
    pub fn mgpu_search_wait(ctx: *mut mgpu_ctx, ticket: u64) -> c_int;
    pub fn mgpu_ivf_search_filtered(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, k: u32, nprobe: u32,
        filter_bits: *const u32, filter_stride_words: u64, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
        out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_ivf_scan_remap_filtered(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, probe_ids: *const u32,
        max_probes: u32, probe_counts: *const u32, k: u32, filter_bits: *const u32, filter_stride_words: u64,
        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_ivf_last_scan_bytes(ivf: *mut mgpu_ivf) -> u64;
    pub fn mgpu_ivf_last_scan_rows(ivf: *mut mgpu_ivf) -> u64;
    pub fn mgpu_ivf_last_scan_fallbacks(ivf: *mut mgpu_ivf) -> u64;
    pub fn mgpu_ivf_assign(ctx: *mut mgpu_ctx, X: *const c_float, n: u64, centroids: *const c_float, nlist: u32, dim: u32,
        max_clusters: u32, threshold: c_float, out_cids: *mut u32, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_kmeans_assign(ctx: *mut mgpu_ctx, X: *const c_float, n: u64, centroids: *const c_float, nlist: u32,
        dim: u32, metric: c_int, penalties: *const c_float, out_labels: *mut u32, out_costs: *mut c_float,
        mem: c_int) -> c_int;
    pub fn mgpu_hnsw_create(ctx: *mut mgpu_ctx, dim: u32, num_layers: u32, edges: *const u32, n_edges: u64,
        points: *const u32, n_points: u64, edge_offsets: *const u64, n_edge_offsets: u64, level_offsets: *const u64,
        quant: c_int, metric: c_int, pq: *mut mgpu_pq, rows: *const c_void, rows_mem: c_int, n: u64,
        doc_ids: *const mgpu_u128, out: *mut *mut mgpu_hnsw) -> c_int;
    pub fn mgpu_hnsw_destroy(h: *mut mgpu_hnsw);
    pub fn mgpu_hnsw_search(h: *mut mgpu_hnsw, Q: *const c_float, B: u32, k: u32, ef: u32, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, out_stats: *mut u64, mem: c_int) -> c_int;
    pub fn mgpu_hnsw_search_submit(h: *mut mgpu_hnsw, Q: *const c_float, B: u32, k: u32, ef: u32,
        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32, ticket: *mut u64) -> c_int;
    pub fn mgpu_spann_create(ctx: *mut mgpu_ctx, centroids: *mut mgpu_hnsw, posting_lists: *mut mgpu_ivf,
        out: *mut *mut mgpu_spann) -> c_int;
    pub fn mgpu_spann_destroy(s: *mut mgpu_spann);
    pub fn mgpu_spann_search(s: *mut mgpu_spann, Q: *const c_float, B: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_spann_search_submit(s: *mut mgpu_spann, Q: *const c_float, B: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, ticket: *mut u64) -> c_int;
    pub fn mgpu_spann_search_filtered(s: *mut mgpu_spann, Q: *const c_float, B: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, filter_bits: *const u32, filter_stride_words: u64,
        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_batcher_create(ivf: *mut mgpu_ivf, max_batch: u32, max_wait_us: u32, k: u32, nprobe: u32,
        out: *mut *mut mgpu_batcher) -> c_int;
    pub fn mgpu_batcher_create_spann(s: *mut mgpu_spann, max_batch: u32, max_wait_us: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, out: *mut *mut mgpu_batcher) -> c_int;
    pub fn mgpu_batcher_destroy(b: *mut mgpu_batcher);
    pub fn mgpu_batcher_search(b: *mut mgpu_batcher, query: *const c_float, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_count: *mut u32) -> c_int;
    pub fn mgpu_batcher_search_filtered(b: *mut mgpu_batcher, query: *const c_float, filter_bits: *const u32,
        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_count: *mut u32) -> c_int;
    pub fn mgpu_batcher_stats(b: *mut mgpu_batcher, stats: *mut u64) -> c_int;
    pub fn mgpu_merge_topk(ctx: *mut mgpu_ctx, doc_ids: *const mgpu_u128, scores: *const c_float, counts: *const u32,
        S: u32, B: u32, k: u32, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32,
        mem: c_int) -> c_int;
    pub fn mgpu_comm_unique_id(out_id: *mut u8) -> c_int;
    pub fn mgpu_comm_init(ctx: *mut mgpu_ctx, nranks: c_int, rank: c_int, id: *const u8) -> c_int;
    pub fn mgpu_comm_destroy(ctx: *mut mgpu_ctx) -> c_int;
    pub fn mgpu_shard_allgather_merge(ctx: *mut mgpu_ctx, local_doc_ids: *const mgpu_u128, local_scores: *const c_float,
        local_counts: *const u32, B: u32, k: u32, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
        out_counts: *mut u32) -> c_int;
    pub fn mgpu_shard_ivf_search(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, k: u32, nprobe: u32,
        shared_codebook: c_int, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32,
        mem: c_int) -> c_int;
    pub fn mgpu_shard_ivf_search_submit(ivf: *mut mgpu_ivf, Q: *const c_float, B: u32, k: u32, nprobe: u32,
        shared_codebook: c_int, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32,
        ticket: *mut u64) -> c_int;
    pub fn mgpu_shard_overlap(ctx: *mut mgpu_ctx, on: c_int) -> c_int;
    pub fn mgpu_shard_spann_search(s: *mut mgpu_spann, Q: *const c_float, B: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, shared_codebook: c_int, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_shard_spann_search_submit(s: *mut mgpu_spann, Q: *const c_float, B: u32, top_k: u32, ef: u32,
        num_explored_centroids: u32, centroid_distance_ratio: c_float, shared_codebook: c_int, out_doc_ids: *mut mgpu_u128,
        out_scores: *mut c_float, out_counts: *mut u32, ticket: *mut u64) -> c_int;
    pub fn mgpu_ef_decode(payload: *const u8, len: u64, out: *mut u64, cap: u64) -> i64;
    pub fn mgpu_pq_load(ctx: *mut mgpu_ctx, quantizer_dir: *const c_char, metric: c_int, out: *mut *mut mgpu_pq) -> c_int;
    pub fn mgpu_ivf_load(ctx: *mut mgpu_ctx, base_dir: *const c_char, index_offset: u64, vector_offset: u64, quant: c_int,
        metric: c_int, pq: *mut mgpu_pq, out: *mut *mut mgpu_ivf) -> c_int;
    pub fn mgpu_hnsw_load(ctx: *mut mgpu_ctx, base_dir: *const c_char, index_offset: u64, vector_offset: u64, dim: u32,
        quant: c_int, metric: c_int, pq: *mut mgpu_pq, out: *mut *mut mgpu_hnsw) -> c_int;
    pub fn mgpu_user_index_info_decode(bytes: *const u8, out: *mut mgpu_user_index_info) -> c_int;
    pub fn mgpu_user_index_info_encode(info: *const mgpu_user_index_info, out_bytes: *mut u8) -> c_int;
    pub fn mgpu_user_index_info_read(path: *const c_char, out: *mut mgpu_user_index_info, cap: u64) -> i64;
    pub fn mgpu_spann_load_user(ctx: *mut mgpu_ctx, base_dir: *const c_char, info: *const mgpu_user_index_info, dim: u32,
        quant: c_int, metric: c_int, out_pq: *mut *mut mgpu_pq, out_centroids: *mut *mut mgpu_hnsw,
        out_lists: *mut *mut mgpu_ivf, out_spann: *mut *mut mgpu_spann) -> c_int;
    pub fn mgpu_hnsw_info(h: *mut mgpu_hnsw, sizes: *mut u64) -> c_int;
    pub fn mgpu_hnsw_copy_graph(h: *mut mgpu_hnsw, edges: *mut u32, points: *mut u32, edge_offsets: *mut u64,
        level_offsets: *mut u64) -> c_int;
    // ---- END GENERATED ----
}

impl From<u128> for mgpu_u128 { fn from(v: u128) -> Self { mgpu_u128 { lo: v as u64, hi: (v >> 64) as u64 } } }
impl From<mgpu_u128> for u128 { fn from(v: mgpu_u128) -> Self { ((v.hi as u128) << 64) | v.lo as u128 } }

fn check(ctx: *mut mgpu_ctx, status: c_int) -> Result<()> {
    if status == MGPU_OK { return Ok(()); }
    let msg = if ctx.is_null() { String::new() } else { unsafe { CStr::from_ptr(mgpu_last_error(ctx)) }.to_string_lossy().into_owned() };
    Err(anyhow!("mgpu error {status}: {msg}"))
}

/// One context per GPU / shard (calls on a context are serialised inside the library).
pub struct GpuContext { pub raw: *mut mgpu_ctx }
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(std::ptr::null_mut(), unsafe { mgpu_init(device, &mut raw) })?;
        Ok(Self { raw })
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { mgpu_destroy(self.raw) } } }

/// The trait impls below are associated functions without `self` (as in the reference), so they use a process-wide context.
static DEFAULT_CTX: OnceLock<GpuContext> = OnceLock::new();
pub fn default_ctx() -> *mut mgpu_ctx { DEFAULT_CTX.get_or_init(|| GpuContext::new(0).expect("no CUDA device: there is no CPU fallback")).raw }

// ---- utils::DistanceCalculator / CalculateSquared (rs/utils/src/lib.rs:17-40) -------------------------------------------------
fn distance_1(a: &[f32], b: &[f32], metric: c_int, squared: bool) -> f32 {
    assert_eq!(a.len(), b.len());
    let mut out = 0f32;
    let st = unsafe { mgpu_distance_batch(default_ctx(), a.as_ptr(), 1, b.as_ptr(), 1, a.len() as u32, metric, squared as c_int, &mut out, MGPU_HOST) };
    check(default_ctx(), st).expect("mgpu_distance_batch");
    out
}
/// All pairs: out[i * nb + j] = calculate(a_i, b_j); the form the index builders and the coarse quantizer want.
pub fn distance_batch(a: &[f32], b: &[f32], dim: usize, metric: c_int, squared: bool) -> Result<Vec<f32>> {
    let (na, nb) = (a.len() / dim, b.len() / dim);
    let mut out = vec![0f32; na * nb];
    let st = unsafe { mgpu_distance_batch(default_ctx(), a.as_ptr(), na as u64, b.as_ptr(), nb as u64, dim as u32, metric, squared as c_int, out.as_mut_ptr(), MGPU_HOST) };
    check(default_ctx(), st)?;
    Ok(out)
}

pub struct GpuL2;
pub struct GpuDot;
/// Mirrors `utils::DistanceCalculator`.  `accumulate_lanes` / `accumulate_scalar` are the CPU SIMD building blocks of the
/// reference's generic code (k-means, PQ distance); on the GPU path those callers are replaced as a whole
/// (`mgpu_kmeans_assign`, `mgpu_pq_distance_batch`), so the two methods forward to the reference's own CPU implementations.
macro_rules! impl_distance {
    ($t:ty, $cpu:ty, $metric:expr) => {
        impl utils::DistanceCalculator for $t {
            #[inline] fn calculate(a: &[f32], b: &[f32]) -> f32 { distance_1(a, b, $metric, false) }
            #[inline] fn accumulate_lanes<const LANES: usize>(a: &[f32], b: &[f32], acc: &mut std::simd::Simd<f32, LANES>)
            where std::simd::LaneCount<LANES>: std::simd::SupportedLaneCount { <$cpu as utils::DistanceCalculator>::accumulate_lanes::<LANES>(a, b, acc) }
            #[inline] fn accumulate_scalar(a: &[f32], b: &[f32]) -> f32 { <$cpu as utils::DistanceCalculator>::accumulate_scalar(a, b) }
            #[inline] fn outermost_op(x: f32) -> f32 { <$cpu as utils::DistanceCalculator>::outermost_op(x) }
        }
        impl utils::CalculateSquared for $t {
            #[inline] fn calculate_squared(a: &[f32], b: &[f32]) -> f32 { distance_1(a, b, $metric, true) }
        }
    };
}
impl_distance!(GpuL2, utils::distance::l2::L2DistanceCalculator, MGPU_L2);
impl_distance!(GpuDot, utils::distance::dot_product::DotProductDistanceCalculator, MGPU_DOT);

/// `LaneConformingDistanceCalculator<LANES, D>` (lane_conforming.rs:9-28) -> mgpu_distance_batch_lanes.
pub struct GpuLaneConforming<const LANES: usize, const METRIC: i32>;
impl<const LANES: usize, const METRIC: i32> utils::CalculateSquared for GpuLaneConforming<LANES, METRIC> {
    fn calculate_squared(a: &[f32], b: &[f32]) -> f32 {
        let mut out = 0f32;
        let st = unsafe { mgpu_distance_batch_lanes(default_ctx(), a.as_ptr(), 1, b.as_ptr(), 1, a.len() as u32, METRIC, LANES as c_int, &mut out, MGPU_HOST) };
        check(default_ctx(), st).expect("mgpu_distance_batch_lanes");
        out
    }
}

/// Assignment step of `KMeansBuilder::run_lloyd` (kmeans_builder.rs:199-221): (label, cost) per row; the calculator follows the
/// dimension exactly like kmeans_builder.rs:126-136.
pub fn kmeans_assign(points: &[f32], centroids: &[f32], dim: usize, penalties: Option<&[f32]>, metric: c_int) -> Result<Vec<(usize, f32)>> {
    let (n, c) = (points.len() / dim, centroids.len() / dim);
    let (mut labels, mut costs) = (vec![0u32; n], vec![0f32; n]);
    let st = unsafe {
        mgpu_kmeans_assign(default_ctx(), points.as_ptr(), n as u64, centroids.as_ptr(), c as u32, dim as u32, metric,
                           penalties.map_or(std::ptr::null(), |p| p.as_ptr()), labels.as_mut_ptr(), costs.as_mut_ptr(), MGPU_HOST)
    };
    check(default_ctx(), st)?;
    Ok(labels.into_iter().map(|l| l as usize).zip(costs).collect())
}

// ---- quantization::Quantizer (rs/quantization/src/quantization.rs:6-38) -----------------------------------------------------------
pub struct GpuProductQuantizer { pub ctx: *mut mgpu_ctx, pub raw: *mut mgpu_pq, pub dimension: usize, pub subvector_dimension: usize, pub num_bits: u8 }
unsafe impl Send for GpuProductQuantizer {}
unsafe impl Sync for GpuProductQuantizer {}
impl GpuProductQuantizer {
    /// `ProductQuantizer::new` (pq/mod.rs:139-149): codebook laid out [subspace][centroid][dsub].
    pub fn new(ctx: *mut mgpu_ctx, dimension: usize, subvector_dimension: usize, num_bits: u8, codebook: &[f32], metric: c_int) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(ctx, unsafe { mgpu_pq_create(ctx, dimension as u32, subvector_dimension as u32, num_bits as u32, codebook.as_ptr(), metric, &mut raw) })?;
        Ok(Self { ctx, raw, dimension, subvector_dimension, num_bits })
    }
    /// Batched `quantize`: n x dim floats -> n x m codes (what `IvfWriter::quantize_and_write_vectors` needs, ivf/writer.rs:205).
    pub fn quantize_batch(&self, x: &[f32]) -> Result<Vec<u8>> {
        let n = x.len() / self.dimension;
        let mut codes = vec![0u8; n * self.dimension / self.subvector_dimension];
        check(self.ctx, unsafe { mgpu_pq_quantize_batch(self.raw, x.as_ptr(), n as u64, codes.as_mut_ptr(), MGPU_HOST) })?;
        Ok(codes)
    }
    /// Batched `distance(a_i, b_i, StreamingSIMD)` (pq/mod.rs:231-266).
    pub fn distance_batch(&self, a: &[u8], b: &[u8]) -> Result<Vec<f32>> {
        let m = self.dimension / self.subvector_dimension;
        let n = a.len() / m;
        let mut out = vec![0f32; n];
        check(self.ctx, unsafe { mgpu_pq_distance_batch(self.raw, a.as_ptr(), b.as_ptr(), n as u64, out.as_mut_ptr(), MGPU_HOST) })?;
        Ok(out)
    }
}
impl quantization::quantization::Quantizer for GpuProductQuantizer {
    type QuantizedT = u8;
    fn quantize(&self, value: &[f32]) -> Vec<u8> { self.quantize_batch(value).expect("mgpu_pq_quantize_batch") }
    fn quantized_dimension(&self) -> usize { self.dimension / self.subvector_dimension }
    fn original_vector(&self, quantized_vector: &[u8]) -> Vec<f32> {
        let mut out = vec![0f32; self.dimension];
        check(self.ctx, unsafe { mgpu_pq_original_vector(self.raw, quantized_vector.as_ptr(), 1, out.as_mut_ptr(), MGPU_HOST) }).expect("mgpu_pq_original_vector");
        out
    }
    /// Every `implem` returns the StreamingSIMD value: the index code only ever asks for that one (typing.rs:25,38).
    fn distance(&self, query: &[u8], point: &[u8], _implem: utils::distance::l2::L2DistanceCalculatorImpl) -> f32 {
        self.distance_batch(query, point).expect("mgpu_pq_distance_batch")[0]
    }
    /// `ProductQuantizerReader::read` (pq/mod.rs:52-136): yaml config + raw f32 codebook, on the default context.
    fn read(dir: String) -> Result<Self> {
        let ctx = default_ctx();
        let mut raw = std::ptr::null_mut();
        let c = CString::new(dir.clone())?;
        check(ctx, unsafe { mgpu_pq_load(ctx, c.as_ptr(), MGPU_L2, &mut raw) })?;
        let cfg: quantization::pq::ProductQuantizerConfig = serde_yaml::from_reader(std::fs::File::open(format!("{dir}/product_quantizer_config.yaml"))?)?;
        Ok(Self { ctx, raw, dimension: cfg.dimension, subvector_dimension: cfg.subvector_dimension, num_bits: cfg.num_bits })
    }
}
impl Drop for GpuProductQuantizer { fn drop(&mut self) { unsafe { mgpu_pq_destroy(self.raw) } } }

// ---- results (rs/index/src/utils.rs:89-93,152-155) ---------------------------------------------------------------------------------
#[derive(Clone, Debug, PartialEq)]
pub struct IdWithScore { pub doc_id: u128, pub score: f32 }
#[derive(Clone, Debug, Default)]
pub struct SearchResult { pub id_with_scores: Vec<IdWithScore> }

fn unpack(b: usize, k: usize, ids: &[mgpu_u128], scores: &[f32], counts: &[u32]) -> Vec<Option<SearchResult>> {
    (0..b).map(|q| {
        if counts[q] == u32::MAX { return None; }   // Spann::search answered None
        Some(SearchResult { id_with_scores: (0..counts[q] as usize).map(|i| IdWithScore { doc_id: ids[q * k + i].into(), score: scores[q * k + i] }).collect() })
    }).collect()
}

// ---- BlockBasedIvf<Q> (rs/index/src/ivf/block_based/index.rs) ------------------------------------------------------------------------
pub struct GpuIvf { pub ctx: *mut mgpu_ctx, pub raw: *mut mgpu_ivf, pub dim: usize, pub quantized_dimension: usize, pub is_pq: bool }
unsafe impl Send for GpuIvf {}
unsafe impl Sync for GpuIvf {}
impl GpuIvf {
    /// `BlockBasedIvf::new_with_offset` (index.rs:95-138): `{base}/index` + `{base}/vectors` of an unmodified reference build.
    pub fn new_with_offset(ctx: *mut mgpu_ctx, base_directory: &str, index_offset: usize, vector_offset: usize, dim: usize,
                           pq: Option<&GpuProductQuantizer>) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        let c = CString::new(base_directory)?;
        let (quant, pqp, qd) = match pq { Some(p) => (MGPU_QUANT_PQ, p.raw, dim / p.subvector_dimension), None => (MGPU_QUANT_NONE, std::ptr::null_mut(), dim) };
        check(ctx, unsafe { mgpu_ivf_load(ctx, c.as_ptr(), index_offset as u64, vector_offset as u64, quant, MGPU_L2, pqp, &mut raw) })?;
        Ok(Self { ctx, raw, dim, quantized_dimension: qd, is_pq: pq.is_some() })
    }
    pub fn num_clusters(&self) -> usize { unsafe { mgpu_ivf_num_clusters(self.raw) as usize } }            // index.rs:338-340
    pub fn num_vectors(&self) -> usize { unsafe { mgpu_ivf_num_vectors(self.raw) as usize } }              // index.rs:346-348
    /// index.rs:350-366
    pub fn get_doc_ids(&self, point_ids: &[u32]) -> Result<Vec<u128>> {
        let mut out = vec![mgpu_u128::default(); point_ids.len()];
        check(self.ctx, unsafe { mgpu_ivf_get_doc_ids(self.raw, point_ids.as_ptr(), point_ids.len() as u32, out.as_mut_ptr()) })?;
        Ok(out.into_iter().map(Into::into).collect())
    }
    pub fn get_doc_id(&self, point_id: u32) -> Result<u128> { Ok(self.get_doc_ids(&[point_id])?[0]) }
    /// index.rs:469-471
    pub fn get_point_id(&self, doc_id: u128) -> Result<Option<u32>> {
        let (d, mut found, mut pid) = (mgpu_u128::from(doc_id), 0 as c_int, 0u32);
        check(self.ctx, unsafe { mgpu_ivf_get_point_id(self.raw, &d, &mut found, &mut pid) })?;
        Ok(if found != 0 { Some(pid) } else { None })
    }
    /// index.rs:372-384 for a PQ index (`Vec<u8>`); `get_vector_f32` is the NoQuantizer form.
    pub fn get_vector(&self, point_id: u32) -> Result<Vec<u8>> {
        let mut out = vec![0u8; self.quantized_dimension * if self.is_pq { 1 } else { 4 }];
        check(self.ctx, unsafe { mgpu_ivf_get_vectors(self.raw, &point_id, 1, out.as_mut_ptr() as *mut c_void) })?;
        Ok(out)
    }
    pub fn get_vector_f32(&self, point_id: u32) -> Result<Vec<f32>> {
        let mut out = vec![0f32; self.quantized_dimension];
        check(self.ctx, unsafe { mgpu_ivf_get_vectors(self.raw, &point_id, 1, out.as_mut_ptr() as *mut c_void) })?;
        Ok(out)
    }
    /// index.rs:417-429
    pub fn invalidate(&self, doc_id: u128) -> Result<bool> { Ok(self.invalidate_batch(&[doc_id])?.len() == 1) }
    /// index.rs:439-452: the doc ids that were successfully invalidated
    pub fn invalidate_batch(&self, doc_ids: &[u128]) -> Result<Vec<u128>> {
        let d: Vec<mgpu_u128> = doc_ids.iter().map(|&x| x.into()).collect();
        let mut ok = vec![0u8; d.len()];
        let mut n = 0u32;
        check(self.ctx, unsafe { mgpu_ivf_invalidate_docs(self.raw, d.as_ptr(), d.len() as u32, ok.as_mut_ptr(), &mut n) })?;
        Ok(doc_ids.iter().zip(ok).filter(|(_, o)| *o != 0).map(|(d, _)| *d).collect())
    }
    /// index.rs:454-459
    pub fn is_invalidated(&self, doc_id: u128) -> Result<bool> {
        let (d, mut out) = (mgpu_u128::from(doc_id), 0 as c_int);
        check(self.ctx, unsafe { mgpu_ivf_is_doc_invalidated(self.raw, &d, &mut out) })?;
        Ok(out != 0)
    }
    /// index.rs:147-163
    pub fn find_nearest_centroids(&self, vector: &[f32], num_probes: usize) -> Result<Vec<usize>> {
        let mut ids = vec![0u32; num_probes.max(1)];
        check(self.ctx, unsafe { mgpu_ivf_coarse(self.raw, vector.as_ptr(), 1, num_probes as u32, ids.as_mut_ptr(), std::ptr::null_mut(), MGPU_HOST) })?;
        Ok(ids.into_iter().take(num_probes).map(|x| x as usize).collect())
    }
    /// index.rs:298-332 (`planner`: the allowed point ids of `Planner::plan_with_ids` as a bitmap, query/planner.rs:43-60)
    pub fn search_with_centroids_and_remap(&self, query: &[f32], nearest_centroid_ids: Vec<usize>, k: usize, planner: Option<&[u32]>) -> Result<SearchResult> {
        let probes: Vec<u32> = nearest_centroid_ids.iter().map(|&c| c as u32).collect();
        let (mut ids, mut scores, mut counts) = (vec![mgpu_u128::default(); k.max(1)], vec![0f32; k.max(1)], vec![0u32; 1]);
        let st = unsafe {
            mgpu_ivf_scan_remap_filtered(self.raw, query.as_ptr(), 1, probes.as_ptr(), probes.len() as u32, std::ptr::null(), k as u32,
                                         planner.map_or(std::ptr::null(), |f| f.as_ptr()), 0, ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr(), MGPU_HOST)
        };
        check(self.ctx, st)?;
        Ok(unpack(1, k.max(1), &ids, &scores, &counts).pop().flatten().unwrap_or_default())
    }
    /// index.rs:396-412
    pub fn search(&self, query: &[f32], k: usize, num_probes: u32, planner: Option<&[u32]>) -> Result<Option<SearchResult>> {
        Ok(self.search_batch(query, k, num_probes, planner)?.pop().flatten())
    }
    /// B x dim row-major queries -> one result per query (the batched form the GPU path is built for).
    pub fn search_batch(&self, queries: &[f32], k: usize, num_probes: u32, planner: Option<&[u32]>) -> Result<Vec<Option<SearchResult>>> {
        let b = queries.len() / self.dim;
        let (mut ids, mut scores, mut counts) = (vec![mgpu_u128::default(); b * k.max(1)], vec![0f32; b * k.max(1)], vec![0u32; b]);
        let st = unsafe {
            mgpu_ivf_search_filtered(self.raw, queries.as_ptr(), b as u32, k as u32, num_probes, planner.map_or(std::ptr::null(), |f| f.as_ptr()), 0,
                                     ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr(), MGPU_HOST)
        };
        check(self.ctx, st)?;
        Ok(unpack(b, k.max(1), &ids, &scores, &counts))
    }
}
impl Drop for GpuIvf { fn drop(&mut self) { unsafe { mgpu_ivf_destroy(self.raw) } } }

// ---- BlockBasedHnsw<Q> (rs/index/src/hnsw/block_based/index.rs) -------------------------------------------------------------------------
pub struct GpuHnsw { pub ctx: *mut mgpu_ctx, pub raw: *mut mgpu_hnsw, pub dim: usize }
unsafe impl Send for GpuHnsw {}
unsafe impl Sync for GpuHnsw {}
impl GpuHnsw {
    /// `BlockBasedHnsw::new_with_offsets` (index.rs:94-140)
    pub fn new_with_offsets(ctx: *mut mgpu_ctx, base_directory: &str, index_offset: usize, vector_offset: usize, dim: usize,
                            pq: Option<&GpuProductQuantizer>) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        let c = CString::new(base_directory)?;
        let (quant, pqp) = pq.map_or((MGPU_QUANT_NONE, std::ptr::null_mut()), |p| (MGPU_QUANT_PQ, p.raw));
        check(ctx, unsafe { mgpu_hnsw_load(ctx, c.as_ptr(), index_offset as u64, vector_offset as u64, dim as u32, quant, MGPU_L2, pqp, &mut raw) })?;
        Ok(Self { ctx, raw, dim })
    }
    /// index.rs:159-210
    pub fn ann_search(&self, query: &[f32], k: usize, ef: u32) -> Result<SearchResult> {
        let (mut ids, mut scores, mut counts) = (vec![mgpu_u128::default(); k.max(1)], vec![0f32; k.max(1)], vec![0u32; 1]);
        check(self.ctx, unsafe { mgpu_hnsw_search(self.raw, query.as_ptr(), 1, k as u32, ef, ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr(), std::ptr::null_mut(), MGPU_HOST) })?;
        Ok(unpack(1, k.max(1), &ids, &scores, &counts).pop().flatten().unwrap_or_default())
    }
}
impl Drop for GpuHnsw { fn drop(&mut self) { unsafe { mgpu_hnsw_destroy(self.raw) } } }

// ---- Spann<Q> (rs/index/src/spann/index.rs) ------------------------------------------------------------------------------------------------
/// rs/config/src/search_params.rs:2-34
pub struct SearchParams { pub top_k: usize, pub ef_construction: u32, pub record_pages: bool, pub num_explored_centroids: Option<usize>, pub centroid_distance_ratio: f32 }
pub struct GpuSpann { pub ctx: *mut mgpu_ctx, pub raw: *mut mgpu_spann, pub centroids: GpuHnsw, pub posting_lists: GpuIvf }
unsafe impl Send for GpuSpann {}
unsafe impl Sync for GpuSpann {}
impl GpuSpann {
    /// `Spann::new` (spann/index.rs:21-30)
    pub fn new(centroids: GpuHnsw, posting_lists: GpuIvf) -> Result<Self> {
        let (ctx, mut raw) = (posting_lists.ctx, std::ptr::null_mut());
        check(ctx, unsafe { mgpu_spann_create(ctx, centroids.raw, posting_lists.raw, &mut raw) })?;
        Ok(Self { ctx, raw, centroids, posting_lists })
    }
    /// `MultiSpannIndex::get_or_create_index` -> `SpannReader::new_with_offsets(..).read` (multi_spann/index.rs:100-128,
    /// spann/reader.rs:43-82): one user's index inside the shared multi-user files.
    pub fn open_user(ctx: *mut mgpu_ctx, base_directory: &str, info: &mgpu_user_index_info, dim: usize, with_pq: bool)
                     -> Result<(Self, Option<GpuProductQuantizer>)> {
        let c = CString::new(base_directory)?;
        let (mut pq, mut hn, mut ivf, mut sp) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, unsafe { mgpu_spann_load_user(ctx, c.as_ptr(), info, dim as u32, if with_pq { MGPU_QUANT_PQ } else { MGPU_QUANT_NONE }, MGPU_L2, &mut pq, &mut hn, &mut ivf, &mut sp) })?;
        let q = if with_pq { Some(GpuProductQuantizer { ctx, raw: pq, dimension: dim, subvector_dimension: 0, num_bits: 0 }) } else { None };
        Ok((Self { ctx, raw: sp, centroids: GpuHnsw { ctx, raw: hn, dim }, posting_lists: GpuIvf { ctx, raw: ivf, dim, quantized_dimension: 0, is_pq: with_pq } }, q))
    }
    /// spann/index.rs:211-266
    pub fn search(&self, query: Vec<f32>, params: &SearchParams, planner: Option<&[u32]>) -> Option<SearchResult> {
        let k = params.top_k.max(1);
        let (mut ids, mut scores, mut counts) = (vec![mgpu_u128::default(); k], vec![0f32; k], vec![0u32; 1]);
        let st = unsafe {
            mgpu_spann_search_filtered(self.raw, query.as_ptr(), 1, params.top_k as u32, params.ef_construction,
                                       params.num_explored_centroids.unwrap_or(params.top_k) as u32, params.centroid_distance_ratio,
                                       planner.map_or(std::ptr::null(), |f| f.as_ptr()), 0, ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr(), MGPU_HOST)
        };
        check(self.ctx, st).ok()?;    // the reference maps errors to None as well (spann/index.rs:263)
        unpack(1, k, &ids, &scores, &counts).pop().flatten()
    }
}
impl Drop for GpuSpann { fn drop(&mut self) { unsafe { mgpu_spann_destroy(self.raw) } } }

/// Per-request front door: what `Spann::search` / `BlockBasedIvf::search` call for ONE query (from `spawn_blocking`); the
/// native worker thread behind `mgpu_batcher_*` forms the batches.
pub struct GpuBatcher { ctx: *mut mgpu_ctx, b: *mut mgpu_batcher, k: usize }
unsafe impl Send for GpuBatcher {}
unsafe impl Sync for GpuBatcher {}
impl GpuBatcher {
    pub fn for_ivf(ivf: &GpuIvf, max_batch: u32, max_wait_us: u32, k: usize, num_probes: u32) -> Result<Self> {
        let mut b = std::ptr::null_mut();
        check(ivf.ctx, unsafe { mgpu_batcher_create(ivf.raw, max_batch, max_wait_us, k as u32, num_probes, &mut b) })?;
        Ok(Self { ctx: ivf.ctx, b, k })
    }
    pub fn for_spann(s: &GpuSpann, max_batch: u32, max_wait_us: u32, p: &SearchParams) -> Result<Self> {
        let mut b = std::ptr::null_mut();
        check(s.ctx, unsafe { mgpu_batcher_create_spann(s.raw, max_batch, max_wait_us, p.top_k as u32, p.ef_construction, p.num_explored_centroids.unwrap_or(p.top_k) as u32, p.centroid_distance_ratio, &mut b) })?;
        Ok(Self { ctx: s.ctx, b, k: p.top_k })
    }
    /// `filter`: the planner's allowed point ids as a bitmap (ceil(N/32) words), or None (index.rs:212-226).
    pub fn search(&self, query: &[f32], filter: Option<&[u32]>) -> Result<Option<Vec<(u128, f32)>>> {
        let (mut ids, mut scores, mut count) = (vec![mgpu_u128::default(); self.k], vec![0f32; self.k], 0u32);
        let fp = filter.map_or(std::ptr::null(), |f| f.as_ptr());
        check(self.ctx, unsafe { mgpu_batcher_search_filtered(self.b, query.as_ptr(), fp, ids.as_mut_ptr(), scores.as_mut_ptr(), &mut count) })?;
        if count == u32::MAX { return Ok(None); }                      // Spann::search returned None
        Ok(Some((0..count as usize).map(|i| (ids[i].into(), scores[i])).collect()))
    }
}
impl Drop for GpuBatcher { fn drop(&mut self) { unsafe { mgpu_batcher_destroy(self.b) } } }

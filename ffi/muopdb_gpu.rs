//! Rust binding of include/muopdb_gpu.h -- the shim a MuopDB maintainer adds (e.g. as `rs/gpu/src/lib.rs`) to put the
//! B200 search path behind the existing query structs.  SOURCE ONLY: this build image has no rustc/cargo, so the file is
//! reviewed against the header but was not compiled here.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)] pub struct mgpu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_pq { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_ivf { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_hnsw { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_spann { _p: [u8; 0] }
#[repr(C)] pub struct mgpu_batcher { _p: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct mgpu_u128 { pub lo: u64, pub hi: u64 }

pub const MGPU_OK: c_int = 0;
pub const MGPU_HOST: c_int = 0;
pub const MGPU_L2: c_int = 0;
pub const MGPU_QUANT_NONE: c_int = 0;
pub const MGPU_QUANT_PQ: c_int = 1;

#[link(name = "muopdb_gpu")]
extern "C" {
    pub fn mgpu_init(device: c_int, out: *mut *mut mgpu_ctx) -> c_int;
    pub fn mgpu_destroy(ctx: *mut mgpu_ctx);
    pub fn mgpu_last_error(ctx: *mut mgpu_ctx) -> *const c_char;
    pub fn mgpu_pq_create(ctx: *mut mgpu_ctx, dim: u32, dsub: u32, nbits: u32, codebook: *const c_float, metric: c_int,
                          out: *mut *mut mgpu_pq) -> c_int;
    pub fn mgpu_pq_destroy(pq: *mut mgpu_pq);
    pub fn mgpu_pq_quantize_batch(pq: *mut mgpu_pq, x: *const c_float, n: u64, codes: *mut u8, mem: c_int) -> c_int;
    pub fn mgpu_ivf_create(ctx: *mut mgpu_ctx, dim: u32, nlist: u32, centroids: *const c_float, list_offsets: *const u64,
                           list_point_ids: *const u32, quant: c_int, metric: c_int, pq: *mut mgpu_pq, rows: *const c_void,
                           rows_mem: c_int, n: u64, doc_ids: *const mgpu_u128, out: *mut *mut mgpu_ivf) -> c_int;
    pub fn mgpu_ivf_destroy(ivf: *mut mgpu_ivf);
    pub fn mgpu_ivf_invalidate(ivf: *mut mgpu_ivf, point_ids: *const u32, n: u32) -> c_int;
    pub fn mgpu_ivf_coarse(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, nprobe: u32, out_ids: *mut u32,
                           out_dist: *mut c_float, mem: c_int) -> c_int;
    pub fn mgpu_ivf_scan_remap(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, probe_ids: *const u32, max_probes: u32,
                               probe_counts: *const u32, k: u32, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
                               out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_ivf_search(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, k: u32, nprobe: u32, out_doc_ids: *mut mgpu_u128,
                           out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    /// pipelined host-buffer search: returns a ticket, `mgpu_search_wait` completes it (two batches in flight per context)
    pub fn mgpu_ivf_search_submit(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, k: u32, nprobe: u32, out_doc_ids: *mut mgpu_u128,
                                  out_scores: *mut c_float, out_counts: *mut u32, ticket: *mut u64) -> c_int;
    pub fn mgpu_search_wait(ctx: *mut mgpu_ctx, ticket: u64) -> c_int;
    /// sharded search in one collective call (query encode split across ranks when the codebook is shared)
    pub fn mgpu_shard_ivf_search(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, k: u32, nprobe: u32, shared_codebook: c_int,
                                 out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_shard_ivf_search_submit(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, k: u32, nprobe: u32, shared_codebook: c_int,
                                        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_counts: *mut u32,
                                        ticket: *mut u64) -> c_int;
    pub fn mgpu_hnsw_search(h: *mut mgpu_hnsw, q: *const c_float, b: u32, k: u32, ef: u32, out_doc_ids: *mut mgpu_u128,
                            out_scores: *mut c_float, out_counts: *mut u32, out_stats: *mut u64, mem: c_int) -> c_int;
    pub fn mgpu_spann_search(s: *mut mgpu_spann, q: *const c_float, b: u32, top_k: u32, ef: u32, num_explored_centroids: u32,
                             centroid_distance_ratio: c_float, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
                             out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_ivf_search_filtered(ivf: *mut mgpu_ivf, q: *const c_float, b: u32, k: u32, nprobe: u32, filter_bits: *const u32,
                                    filter_stride_words: u64, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
                                    out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_spann_search_filtered(s: *mut mgpu_spann, q: *const c_float, b: u32, top_k: u32, ef: u32,
                                      num_explored_centroids: u32, centroid_distance_ratio: c_float, filter_bits: *const u32,
                                      filter_stride_words: u64, out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float,
                                      out_counts: *mut u32, mem: c_int) -> c_int;
    pub fn mgpu_batcher_create(ivf: *mut mgpu_ivf, max_batch: u32, max_wait_us: u32, k: u32, nprobe: u32,
                               out: *mut *mut mgpu_batcher) -> c_int;
    pub fn mgpu_batcher_create_spann(s: *mut mgpu_spann, max_batch: u32, max_wait_us: u32, top_k: u32, ef: u32,
                                     num_explored_centroids: u32, centroid_distance_ratio: c_float,
                                     out: *mut *mut mgpu_batcher) -> c_int;
    pub fn mgpu_batcher_destroy(b: *mut mgpu_batcher);
    pub fn mgpu_batcher_search_filtered(b: *mut mgpu_batcher, query: *const c_float, filter_bits: *const u32,
                                        out_doc_ids: *mut mgpu_u128, out_scores: *mut c_float, out_count: *mut u32) -> c_int;
    // ... the remaining entry points of include/muopdb_gpu.h bind the same way
}

fn check(ctx: *mut mgpu_ctx, status: c_int) -> anyhow::Result<()> {
    if status == MGPU_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(mgpu_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!("mgpu error {status}: {msg}"))
}

/// What `BlockBasedIvf::<Q>::search` (rs/index/src/ivf/block_based/index.rs:396-412) becomes for a micro-batch of queries.
pub struct GpuIvf { ctx: *mut mgpu_ctx, ivf: *mut mgpu_ivf, dim: usize }
unsafe impl Send for GpuIvf {}
unsafe impl Sync for GpuIvf {}

pub struct IdWithScore { pub doc_id: u128, pub score: f32 }

impl GpuIvf {
    /// `queries`: B x dim row-major.  Returns one `Vec<IdWithScore>` per query, ordered by (score, doc_id) like
    /// `search_with_centroids_and_remap` (index.rs:298-332).
    pub fn search_batch(&self, queries: &[f32], k: usize, num_probes: u32) -> anyhow::Result<Vec<Vec<IdWithScore>>> {
        let b = queries.len() / self.dim;
        let mut ids = vec![mgpu_u128::default(); b * k];
        let mut scores = vec![0f32; b * k];
        let mut counts = vec![0u32; b];
        let st = unsafe {
            mgpu_ivf_search(self.ivf, queries.as_ptr(), b as u32, k as u32, num_probes, ids.as_mut_ptr(), scores.as_mut_ptr(),
                            counts.as_mut_ptr(), MGPU_HOST)
        };
        check(self.ctx, st)?;
        Ok((0..b).map(|q| (0..counts[q] as usize).map(|i| {
            let d = ids[q * k + i];
            IdWithScore { doc_id: (d.hi as u128) << 64 | d.lo as u128, score: scores[q * k + i] }
        }).collect()).collect())
    }
}

/// Per-request front door: what `Spann::search` / `BlockBasedIvf::search` call for ONE query (from `spawn_blocking`); the
/// native worker thread behind `mgpu_batcher_*` forms the batches.
pub struct GpuBatcher { ctx: *mut mgpu_ctx, b: *mut mgpu_batcher, k: usize }
unsafe impl Send for GpuBatcher {}
unsafe impl Sync for GpuBatcher {}

impl GpuBatcher {
    /// `filter`: the planner's allowed point ids as a bitmap (ceil(N/32) words), or None (index.rs:212-226).
    pub fn search(&self, query: &[f32], filter: Option<&[u32]>) -> anyhow::Result<Option<Vec<(u128, f32)>>> {
        let mut ids = vec![mgpu_u128::default(); self.k];
        let mut scores = vec![0f32; self.k];
        let mut count = 0u32;
        let fp = filter.map_or(std::ptr::null(), |f| f.as_ptr());
        let st = unsafe { mgpu_batcher_search_filtered(self.b, query.as_ptr(), fp, ids.as_mut_ptr(), scores.as_mut_ptr(), &mut count) };
        check(self.ctx, st)?;
        if count == u32::MAX { return Ok(None); }                      // Spann::search returned None
        Ok(Some((0..count as usize).map(|i| (((ids[i].hi as u128) << 64) | ids[i].lo as u128, scores[i])).collect()))
    }
}
impl Drop for GpuBatcher { fn drop(&mut self) { unsafe { mgpu_batcher_destroy(self.b) } } }

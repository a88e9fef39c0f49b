// scan.cu -- posting-list scan kernels (the hot loop of BlockBasedIvf::scan_posting_list /
// search_with_centroids, rs/index/src/ivf/block_based/index.rs:175-285) for a batch of queries.
//
// One persistent CTA per SM; a CTA owns one query at a time (LUT / query vector resident in shared memory), its 16
// warps stride over the 32-row chunks of the query's probed lists, one row per lane.  Every warp keeps its 32 best
// (score key, point id) pairs sorted across lanes (warp-shuffle insertion); a CTA-wide threshold in shared memory
// prunes rows that cannot enter any warp's list; the 16 lists are merged with shuffle bitonic merges.
//
// Modes
//   SCAN_PQ_FAST    m % 32 == 0, 8-bit codes: conflict-free fixed-point LUT (see internal.cuh for the code layout)
//   SCAN_PQ_GENERIC any m / nbits <= 8: natural [m][K] fixed-point LUT, row-major codes
//   SCAN_FLAT_L2 / SCAN_FLAT_DOT  bit-exact fp32 distances in the reference's 16/8/4-lane order
//
// PQ modes rank rows by an order-independent fixed-point sum of LUT entries (exact integer adds, so equal code
// words always get equal keys, like the reference's exact ties); the 32 survivors per query are re-scored bit-exactly
// in finalize.cu, which also applies the reference's (distance, point_id) ordering.
#include "internal.cuh"
#include "scan_common.cuh"
#include "pq_device.cuh"
#include "finalize_device.cuh"

#ifndef SCAN_FAST_NT
#define SCAN_FAST_NT 768
#endif

struct ScanSmemLayout {
  uint32_t pol_bytes, tmp_bytes, total;
  uint32_t off_tmp, off_pref, off_pcs, off_mkey, off_mpay, off_misc;
};

__host__ __device__ inline ScanSmemLayout scan_layout(int mode, uint32_t dim, uint32_t m, uint32_t K, uint32_t ng,
                                                      uint32_t max_probes) {
  ScanSmemLayout L;
  if (mode == SCAN_PQ_FAST) { L.pol_bytes = ((ng + 1) / 2) * 65536u; L.tmp_bytes = 32u * 257u * 4u; }
  else if (mode == SCAN_PQ_GENERIC) { L.pol_bytes = m * K * 4u; L.tmp_bytes = 0; }
  else { L.pol_bytes = ((dim + 3) / 4) * 16u; L.tmp_bytes = 0; }
  L.pol_bytes = (L.pol_bytes + 15u) & ~15u;
  L.off_tmp = L.pol_bytes;
  L.off_pref = L.off_tmp + L.tmp_bytes;
  L.off_pcs = L.off_pref + (max_probes + 1) * 4u;
  L.off_mkey = (L.off_pcs + max_probes * 4u + 15u) & ~15u;
  L.off_mpay = L.off_mkey + SCAN_MAX_WARPS * 32u * 8u;
  L.off_misc = L.off_mpay + SCAN_MAX_WARPS * 32u * 4u;
  L.total = L.off_misc + 384u + 1024u * 4u + 1024u * 4u;  // flags, reduction scratch, b2[32] | soff[1024] | scode[1024]
  return L;
}

template <int MODE, int NG, int NT>
__global__ void __launch_bounds__(NT, 1) k_scan(ScanArgs a, ScanSmemLayout L) {
  constexpr int SCAN_THREADS = NT, SCAN_WARPS = NT / 32;
  // the "max over warps of their 2nd best" threshold bounds the global 32nd best only with >= 16 warps
  static_assert(2 * SCAN_WARPS >= MGPU_NCAND && SCAN_WARPS <= SCAN_MAX_WARPS, "threshold rule needs 16..32 warps");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *pol = smem;
  float *tmp = (float *)(smem + L.off_tmp);
  uint32_t *pref = (uint32_t *)(smem + L.off_pref);
  uint32_t *pcs = (uint32_t *)(smem + L.off_pcs);
  uint64_t *mkey = (uint64_t *)(smem + L.off_mkey);
  uint32_t *mpay = (uint32_t *)(smem + L.off_mpay);
  uint32_t *sh = (uint32_t *)(smem + L.off_misc);  // [0] threshold key, [1] total chunks, [2] next query
  float *shf = (float *)(sh + 32);                 // 32 floats of reduction scratch
  uint32_t *b2 = sh + 64;                          // per-warp 2nd-best key (threshold tightening)
  float *soff = (float *)(smem + L.off_misc + 384);
  uint32_t *scode = (uint32_t *)(smem + L.off_misc + 384 + 4096);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t m = a.m, K = a.K;

  // lane-dependent LUT column offsets for the conflict-free scan: x_t = ((lane ^ t) * 4) <= 124, four per register.
  // PRMT builds the LUT byte offset (code << 8) | x_t in ONE instruction: byte0 = x_t, byte1 = code, bytes 2/3 = the
  // sign replication of x_t (msb is 0) = 0.
  uint32_t xr[8];
  if (MODE == SCAN_PQ_FAST) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      xr[j] = (uint32_t)((lane ^ (4 * j)) << 2) | ((uint32_t)((lane ^ (4 * j + 1)) << 2) << 8) |
              ((uint32_t)((lane ^ (4 * j + 2)) << 2) << 16) | ((uint32_t)((lane ^ (4 * j + 3)) << 2) << 24);
  }

  for (;;) {
    // dynamic query scheduling: one atomic per query keeps the 148 persistent CTAs balanced on ragged probe lists
    if (tid == 0) { uint32_t t = atomicAdd(a.next_query, 1u); sh[2] = t < a.B ? (a.order ? a.order[t] : t) : 0xFFFFFFFFu; }
    __syncthreads();
    const uint32_t q = sh[2];
    if (q >= a.B) break;
    uint32_t np = a.probe_counts ? min(a.probe_counts[q], a.max_probes) : a.max_probes;
    // ---- 1. prefix of chunk counts over this query's probe list -------------------------------------------------
    if (warp == 0) {
      uint32_t run = 0;
      unsigned long long rows = 0;
      for (uint32_t base = 0; base < np; base += 32) {
        uint32_t i = base + lane, cnt = 0, cs = 0, len = 0;
        if (i < np) {
          uint32_t c = a.probes[(size_t)q * a.max_probes + i];
          cs = a.chunk_start[c];
          cnt = a.chunk_start[c + 1] - cs;
          len = a.list_len[c];
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        if (i < np) { pref[i] = run + incl - cnt; pcs[i] = cs; }
        run += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
        rows += len;
      }
      if (lane == 0) {
        pref[np] = run;
        sh[0] = 0xFFFFFFFFu;
        sh[1] = run;
      }
      if (lane < SCAN_WARPS) b2[lane] = 0xFFFFFFFFu;
      if (lane == 0) {
        if (a.rows_scanned) atomicAdd(a.rows_scanned, rows);
      }
    }
    // ---- 2. per-query state into shared memory -------------------------------------------------------------------
    if (MODE == SCAN_PQ_FAST || MODE == SCAN_PQ_GENERIC) {
      // fixed-point scale: sum over subspaces of the LUT row range must fit 32 bits
      float part = 0.0f;
      for (uint32_t s = tid; s < m; s += SCAN_THREADS) {
        uint32_t code = a.qcodes[(size_t)q * m + s];
        float mn = a.rowmin[(size_t)s * K + code], mx = a.rowmax[(size_t)s * K + code];
        soff[s] = mn; scode[s] = code;
        part += mx - mn;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) shf[warp] = part;
      __syncthreads();
      float R = 0.0f;
#pragma unroll
      for (int w = 0; w < SCAN_WARPS; w++) R += shf[w];
      float scale = R > 0.0f ? 4.0e9f / R : 0.0f;
      if (MODE == SCAN_PQ_FAST) {
        for (int g = 0; g < NG; g++) {
          // stage 32 table rows (1 KB each, coalesced) into a padded tile
          for (int r = warp; r < 32; r += SCAN_WARPS) {
            uint32_t s = g * 32 + r;
            const float *src = a.table + ((size_t)s * K + scode[s]) * K;
#pragma unroll
            for (int j = lane; j < 256; j += 32) tmp[r * 257 + j] = src[j];
          }
          __syncthreads();
          // transpose + convert: lane <-> LUT column (subspace), warps stride over codes; both sides conflict free
          float off = soff[g * 32 + lane];
          uint8_t *dst = pol + (g >> 1) * 65536 + (g & 1) * 128 + lane * 4;
          for (int j = warp; j < 256; j += SCAN_WARPS) {
            float v = tmp[lane * 257 + j];
            *(uint32_t *)(dst + j * 256) = __float2uint_rn(__fmul_rn(__fsub_rn(v, off), scale));
          }
          __syncthreads();
        }
      } else {
        uint32_t *lut = (uint32_t *)pol;
        for (uint32_t s = warp; s < m; s += SCAN_WARPS) {
          const float *src = a.table + ((size_t)s * K + scode[s]) * K;
          float off = soff[s];
          for (uint32_t j = lane; j < K; j += 32) lut[s * K + j] = __float2uint_rn(__fmul_rn(__fsub_rn(src[j], off), scale));
        }
      }
    } else {
      float *sq = (float *)pol;
      for (uint32_t d = tid; d < a.dim4 * 4; d += SCAN_THREADS) sq[d] = d < a.dim ? a.Q[(size_t)q * a.dim + d] : 0.0f;
    }
    __syncthreads();

    // ---- 3. scan ------------------------------------------------------------------------------------------------
    const uint32_t total = sh[1];
    WarpTop32 top;
    top.init();
    bool first = true;
    uint32_t p = 0;
    for (uint32_t it = warp; it < total; it += SCAN_WARPS) {
      while (it >= pref[p + 1]) p++;
      const uint32_t chunk = pcs[p] + (it - pref[p]);
      const uint32_t slot = chunk * 32 + lane;
      const uint32_t pid = a.slot_pid[slot];
      bool valid = pid != MGPU_EMPTY_SLOT;
      if (a.invalid && valid) valid = !((a.invalid[pid >> 5] >> (pid & 31)) & 1u);  // index.rs:198-200
      if (a.filter && valid) valid = (a.filter[(size_t)q * a.filter_stride + (pid >> 5)] >> (pid & 31)) & 1u;  // index.rs:212-226

      uint32_t key;
      if (MODE == SCAN_PQ_FAST) {
        const uint4 *base = (const uint4 *)a.codes + (size_t)chunk * (NG * 64 + 8) + lane;
        uint4 u[NG * 2];
#pragma unroll
        for (int i = 0; i < NG * 2; i++) u[i] = ldg_stream16(base + i * 32);
        uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          const uint8_t *lut_g = pol + (g >> 1) * 65536 + (g & 1) * 128;
#pragma unroll
          for (int wi = 0; wi < 8; wi++) {
            const uint4 &uu = u[g * 2 + (wi >> 2)];
            uint32_t w = (wi & 3) == 0 ? uu.x : ((wi & 3) == 1 ? uu.y : ((wi & 3) == 2 ? uu.z : uu.w));
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int t = wi * 4 + k;
              uint32_t idx = prmt(w, xr[t >> 2], ((12 + (t & 3)) << 12) | ((12 + (t & 3)) << 8) | (k << 4) | (4 + (t & 3)));
              uint32_t v = *(const uint32_t *)(lut_g + idx);
              if (t & 1) acc1 += v; else acc0 += v;
            }
          }
        }
        key = acc0 + acc1;
      } else if (MODE == SCAN_PQ_GENERIC) {
        const uint32_t *lut = (const uint32_t *)pol;
        const uint8_t *cr = a.codes + (size_t)slot * m;
        uint32_t acc = 0;
        for (uint32_t s = 0; s < m; s++) acc += lut[s * K + cr[s]];
        key = acc;
      } else {
        constexpr int METRIC = MODE == SCAN_FLAT_L2 ? MGPU_L2 : MGPU_DOT;
        const float *sq = (const float *)pol;
        const float4 *base = (const float4 *)a.rows + (size_t)chunk * a.dim4 * 32 + lane;
        const int n = (int)a.dim;
        float ret = 0.0f;
        int pdim = 0;
        const bool go16 = METRIC == MGPU_L2 ? (n / 16 > 0) : (n > 16);
        if (go16) {
          float acc[16];
#pragma unroll
          for (int l = 0; l < 16; l++) acc[l] = 0.0f;
          const int chunks = n / 16;
#pragma unroll 2
          for (int c = 0; c < chunks; c++) {
            float4 r[4], qq[4];
#pragma unroll
            for (int j = 0; j < 4; j++) r[j] = ldg_stream16f(base + (size_t)(c * 4 + j) * 32);
#pragma unroll
            for (int j = 0; j < 4; j++) qq[j] = *(const float4 *)(sq + c * 16 + j * 4);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float rv[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
              const float qv[4] = {qq[j].x, qq[j].y, qq[j].z, qq[j].w};
#pragma unroll
              for (int e = 0; e < 4; e++) {
                if (METRIC == MGPU_L2) { float d = __fsub_rn(qv[e], rv[e]); acc[j * 4 + e] = __fadd_rn(acc[j * 4 + e], __fmul_rn(d, d)); }
                else acc[j * 4 + e] = __fadd_rn(acc[j * 4 + e], __fmul_rn(qv[e], rv[e]));
              }
            }
          }
          ret = __fadd_rn(ret, ordered_reduce(acc, 16));
          pdim = chunks * 16;
        }
        // 8-lane / 4-lane / scalar remainder (l2.rs:45-66, dot_product.rs:51-69)
        auto rowacc = [&](int i) { return ((const float *)(base + (size_t)(i >> 2) * 32))[i & 3]; };
        auto qacc = [&](int i) { return sq[i]; };
#pragma unroll
        for (int li = 1; li < 3; li++) {
          const int LN = li == 1 ? 8 : 4;
          int rem = n - pdim;
          bool go = METRIC == MGPU_L2 ? (rem / LN > 0) : (rem > LN);
          if (go) {
            float acc[8];
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] = 0.0f;
            int chunks = rem / LN;
            for (int c = 0; c < chunks; c++) {
#pragma unroll
              for (int l = 0; l < 8; l++) {
                if (l < LN) {
                  float x = qacc(pdim + c * LN + l), y = rowacc(pdim + c * LN + l);
                  if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); acc[l] = __fadd_rn(acc[l], __fmul_rn(d, d)); }
                  else acc[l] = __fadd_rn(acc[l], __fmul_rn(x, y));
                }
              }
            }
            float s = -0.0f;
#pragma unroll
            for (int l = 0; l < 8; l++) if (l < LN) s = __fadd_rn(s, acc[l]);
            ret = __fadd_rn(ret, s);
            pdim += chunks * LN;
          }
        }
        for (; pdim < n; pdim++) {
          float x = qacc(pdim), y = rowacc(pdim);
          if (METRIC == MGPU_L2) { float d = __fsub_rn(x, y); ret = __fadd_rn(ret, __fmul_rn(d, d)); }
          else ret = __fadd_rn(ret, __fmul_rn(x, y));
        }
        key = METRIC == MGPU_L2 ? f2key(sqrtf(ret)) : f2key(-ret);  // noq/mod.rs:44-51 -> D::calculate
      }

      const uint32_t thr = *(volatile uint32_t *)&sh[0];
      bool pass = valid && key <= thr;
      if (a.lower_bound && pass && (((uint64_t)key << 32) | pid) < a.lower_bound[q]) pass = false;  // reported by an earlier round
      if (__any_sync(0xffffffffu, pass)) {
        uint32_t worst;
        if (first) {
          // empty list: one bitonic sort instead of up to 32 sequential insertions
          top.key = pass ? (((uint64_t)key << 32) | pid) : MGPU_EMPTY_KEY;
          top.pay = pass ? slot : MGPU_EMPTY_SLOT;
          top.sort();
          worst = (uint32_t)(shfl64(top.key, 31) >> 32);
          first = false;
        } else {
          worst = top.offer(pass, ((uint64_t)key << 32) | pid, slot);
        }
        // Threshold tightening.  (a) this warp's 32nd best bounds the global 32nd best; (b) so does the maximum over
        // warps of their 2nd best (the union of every warp's two best already holds >= 32 rows).  (b) converges ~5x
        // faster because each warp only sees 1/NW of the rows.
        uint32_t second = (uint32_t)(shfl64(top.key, 1) >> 32);
        if (lane == 0) b2[warp] = second;
        __syncwarp();
        uint32_t v = lane < SCAN_WARPS ? *(volatile uint32_t *)&b2[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        v = min(v, worst);
        if (lane == 0 && v < thr) atomicMin(&sh[0], v);
      }
    }

    // ---- 4. merge the warp lists -----------------------------------------------------------------------------
    mkey[warp * 32 + lane] = top.key;
    mpay[warp * 32 + lane] = top.pay;
    __syncthreads();
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
      if (warp < half && warp + half < SCAN_WARPS) {
        top.merge(mkey[(warp + half) * 32 + lane], mpay[(warp + half) * 32 + lane]);
        mkey[warp * 32 + lane] = top.key;
        mpay[warp * 32 + lane] = top.pay;
      }
      __syncthreads();
    }
    if (warp == 0) {
      a.cand_key[(size_t)q * MGPU_NCAND + lane] = top.key;
      a.cand_slot[(size_t)q * MGPU_NCAND + lane] = top.pay;
    }
    __syncthreads();
  }
}

// ---- exact fallback -------------------------------------------------------------------------------------------------------------
// Queries the 16-bit scan could not certify (finalize.cu) or did not scan (more chunks than its chunk table) are answered here
// with NO approximation: every row of the probed lists is scored with ProductQuantizer::distance itself (pq/mod.rs:231-266),
// the per-warp top-32 lists hold exact (score key, point id) composites, and warp 0 runs the ordinary epilogue
// (finalize_tail: (distance, point_id) order -> top k -> doc ids -> (score, doc_id) order).  Persistent CTAs take queries from
// a.overflow_list[0 .. *a.overflow_count); the launch is a no-op when the list is empty.  Slow by design (96 codebook gathers
// per row) -- it only has to be right.
#define EXACT_WARPS 16
template <int METRIC>
__global__ void __launch_bounds__(EXACT_WARPS * 32, 1) k_scan_pq_exact(ScanArgs a, FinalizeArgs f) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t *pref = (uint32_t *)smem;                    // maxp + 1
  uint32_t *pcs = pref + a.max_probes + 1;              // maxp
  uint64_t *mkey = (uint64_t *)(((uintptr_t)(pcs + a.max_probes) + 15) & ~(uintptr_t)15);
  uint32_t *mpay = (uint32_t *)(mkey + EXACT_WARPS * 32);
  uint32_t *sh = mpay + EXACT_WARPS * 32;               // [0] threshold, [1] total chunks, [2] query
  uint32_t *b2 = sh + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) { uint32_t t = atomicAdd(a.next_query, 1u); sh[2] = t < *a.overflow_count ? a.overflow_list[t] : 0xFFFFFFFFu; }
    __syncthreads();
    const uint32_t q = sh[2];
    if (q >= a.B) break;
    const uint32_t np = a.probe_counts ? min(a.probe_counts[q], a.max_probes) : a.max_probes;
    if (warp == 0) {
      uint32_t run = 0;
      for (uint32_t base = 0; base < np; base += 32) {
        uint32_t i = base + lane, cnt = 0, cs = 0;
        if (i < np) {
          uint32_t c = a.probes[(size_t)q * a.max_probes + i];
          cs = a.chunk_start[c];
          cnt = a.chunk_start[c + 1] - cs;
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        if (i < np) { pref[i] = run + incl - cnt; pcs[i] = cs; }
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) { pref[np] = run; sh[0] = 0xFFFFFFFFu; sh[1] = run; }
      if (lane < EXACT_WARPS) b2[lane] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t total = sh[1];
    WarpTop32 top;
    top.init();
    uint32_t p = 0;
    const RowMajorCode qc{a.qcodes + (size_t)q * a.m};
    for (uint32_t it = warp; it < total; it += EXACT_WARPS) {
      while (it >= pref[p + 1]) p++;
      const uint32_t chunk = pcs[p] + (it - pref[p]);
      const uint32_t slot = chunk * 32 + lane;
      const uint32_t pid = a.slot_pid[slot];
      bool valid = pid != MGPU_EMPTY_SLOT;
      if (a.invalid && valid) valid = !((a.invalid[pid >> 5] >> (pid & 31)) & 1u);  // index.rs:198-200
      if (a.filter && valid) valid = (a.filter[(size_t)q * a.filter_stride + (pid >> 5)] >> (pid & 31)) & 1u;  // index.rs:212-226
      uint32_t key = 0xFFFFFFFFu;
      if (valid) key = f2key(pq_distance_streaming<METRIC>(a.cb, a.m, a.K, a.dsub, qc, FastLayoutCode{a.codes, slot, a.ng}));
      const uint32_t thr = *(volatile uint32_t *)&sh[0];
      const bool pass = valid && key <= thr;
      if (__any_sync(0xffffffffu, pass)) {
        const uint32_t worst = top.offer(pass, ((uint64_t)key << 32) | pid, slot);
        uint32_t second = (uint32_t)(shfl64(top.key, 1) >> 32);
        if (lane == 0) b2[warp] = second;
        __syncwarp();
        uint32_t v = lane < EXACT_WARPS ? *(volatile uint32_t *)&b2[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        v = min(v, worst);
        if (lane == 0 && v < thr) atomicMin(&sh[0], v);
      }
    }
    mkey[warp * 32 + lane] = top.key;
    mpay[warp * 32 + lane] = top.pay;
    __syncthreads();
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
      if (warp < half && warp + half < EXACT_WARPS) {
        top.merge(mkey[(warp + half) * 32 + lane], mpay[(warp + half) * 32 + lane]);
        mkey[warp * 32 + lane] = top.key;
        mpay[warp * 32 + lane] = top.pay;
      }
      __syncthreads();
    }
    if (warp == 0) {
      const bool valid = top.pay != MGPU_EMPTY_SLOT;
      finalize_tail(f, q, lane, valid, (uint32_t)(top.key >> 32), (uint32_t)top.key, top.pay);
    }
    __syncthreads();
  }
}

int launch_scan_pq_exact_list(mgpu_ivf *ivf, const ScanArgs &a0, const FinalizeArgs &f0) {
  mgpu_ctx *ctx = ivf->ctx;
  ScanArgs a = a0;
  FinalizeArgs f = f0;
  f.key16 = 0; f.prune = false; f.qstate = nullptr; f.cb = nullptr;   // keys are exact: no certification, no re-score
  a.from_list = 1; a.cb = ivf->pq->d_cb; a.dsub = ivf->pq->dsub;
  const size_t smem = ((size_t)2 * a.max_probes + 1) * 4 + 16 + (size_t)EXACT_WARPS * 32 * 12 + 16 + 32 * 4 + 64;
  if (smem > ctx->smem_optin) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "exact fallback scan: max_probes too large");
  CUDA_TRY(ctx, cudaMemsetAsync(a.next_query, 0, 4, ctx->stream));
  unsigned grid = a.B < (uint32_t)ctx->sm_count ? a.B : (unsigned)ctx->sm_count;
  LaunchScope ls(ctx, MGPU_K_FALLBACK);
  if (ivf->metric == MGPU_L2) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan_pq_exact<MGPU_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_scan_pq_exact<MGPU_L2><<<grid, EXACT_WARPS * 32, smem, ctx->stream>>>(a, f);
  } else {
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan_pq_exact<MGPU_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_scan_pq_exact<MGPU_DOT><<<grid, EXACT_WARPS * 32, smem, ctx->stream>>>(a, f);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

// Longest-processing-time-first order for the persistent CTAs' dynamic scheduler (removes the ragged tail of the last
// wave: ~3 % of the scan).  work(q) = chunks of q's probed lists (one warp per query), then a 64-bucket counting sort
// by descending work in one CTA -- an exact sort is not needed, only "big queries first".
__global__ void __launch_bounds__(256) k_query_work(const uint32_t *__restrict__ probes, uint32_t max_probes,
                                                    const uint32_t *__restrict__ probe_counts,
                                                    const uint32_t *__restrict__ chunk_start, uint32_t B,
                                                    uint32_t *__restrict__ work) {
  const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= B) return;
  const uint32_t np = probe_counts ? min(probe_counts[q], max_probes) : max_probes;
  uint32_t w = 0;
  for (uint32_t i = threadIdx.x & 31; i < np; i += 32) {
    uint32_t c = probes[(size_t)q * max_probes + i];
    w += chunk_start[c + 1] - chunk_start[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0) work[q] = w;
}

__global__ void __launch_bounds__(1024) k_plan_bucket(const uint32_t *__restrict__ work, uint32_t B, uint32_t *__restrict__ order) {
  __shared__ uint32_t hist[64], base[64], wmax_s;
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) wmax_s = 0;
  __syncthreads();
  uint32_t mx = 0;
  for (uint32_t q = threadIdx.x; q < B; q += blockDim.x) mx = max(mx, work[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&wmax_s, mx);
  __syncthreads();
  const uint64_t wmax = (uint64_t)wmax_s + 1;
  // pass 1: histogram; pass 2 (after the prefix): scatter -- positions inside a bucket come from a second atomic counter
  for (uint32_t q = threadIdx.x; q < B; q += blockDim.x) atomicAdd(&hist[63 - (uint32_t)(((uint64_t)work[q] * 64) / wmax)], 1u);
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t a0 = hist[2 * threadIdx.x], a1 = hist[2 * threadIdx.x + 1], sum = a0 + a1, incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += v; }
    base[2 * threadIdx.x] = incl - sum;
    base[2 * threadIdx.x + 1] = incl - sum + a0;
  }
  __syncthreads();
  for (uint32_t q = threadIdx.x; q < B; q += blockDim.x) {
    uint32_t b = 63 - (uint32_t)(((uint64_t)work[q] * 64) / wmax);
    order[atomicAdd(&base[b], 1u)] = q;
  }
}

int launch_plan_queries(mgpu_ivf *ivf, const uint32_t *d_probes, uint32_t max_probes, const uint32_t *d_counts, uint32_t B,
                        uint32_t *d_order, uint32_t *d_work, bool have_work) {
  mgpu_ctx *ctx = ivf->ctx;
  LaunchScope ls(ctx, MGPU_K_OTHER);
  // the coarse selection already emitted work[q] when it chose the probes itself
  if (!have_work) k_query_work<<<(B + 7) / 8, 256, 0, ctx->stream>>>(d_probes, max_probes, d_counts, ivf->d_chunk_start, B, d_work);
  k_plan_bucket<<<1, 1024, 0, ctx->stream>>>(d_work, B, d_order);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return MGPU_OK;
}

static int scan_mode(const mgpu_ivf *ivf) {
  if (ivf->quant == MGPU_QUANT_PQ) return ivf->pq_fast ? SCAN_PQ_FAST : SCAN_PQ_GENERIC;
  return ivf->metric == MGPU_L2 ? SCAN_FLAT_L2 : SCAN_FLAT_DOT;
}

size_t scan_max_probes_supported(mgpu_ivf *ivf) {
  int mode = scan_mode(ivf);
  uint32_t m = ivf->pq ? ivf->pq->m : 0, K = ivf->pq ? ivf->pq->K : 0;
  ScanSmemLayout L0 = scan_layout(mode, ivf->dim, m, K, ivf->ng, 0);
  if (L0.total > ivf->ctx->smem_optin) return 0;
  return (ivf->ctx->smem_optin - L0.total) / 8;
}

template <int MODE, int NG, int NT = 512>
static int launch_scan_t(mgpu_ivf *ivf, const ScanArgs &a, const ScanSmemLayout &L) {
  mgpu_ctx *ctx = ivf->ctx;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan<MODE, NG, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  unsigned grid = a.B < (uint32_t)ctx->sm_count ? a.B : (unsigned)ctx->sm_count;
  CUDA_TRY(ctx, cudaMemsetAsync(a.next_query, 0, 4, ctx->stream));
  LaunchScope ls(ctx, MGPU_K_SCAN, nullptr, (MODE == SCAN_FLAT_L2 || MODE == SCAN_FLAT_DOT) ? "k_scan<flat> (scan.cu)" : "k_scan<pq> (scan.cu)");
  k_scan<MODE, NG, NT><<<grid, NT, L.total, ctx->stream>>>(a, L);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

int launch_scan(mgpu_ivf *ivf, const ScanArgs &a) {
  mgpu_ctx *ctx = ivf->ctx;
  if (a.B == 0) return MGPU_OK;
  int mode = scan_mode(ivf);
  ScanSmemLayout L = scan_layout(mode, a.dim, a.m, a.K, a.ng, a.max_probes);
  if (L.total > ctx->smem_optin)
    return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "scan needs %u bytes of shared memory (max %zu): m*K or max_probes too large", L.total, ctx->smem_optin);
  if (a.m > 1024) return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "scan supports at most 1024 PQ subspaces");
  switch (mode) {
    case SCAN_PQ_FAST: {
      // headline path: warp-specialised, double-buffered LUT (scan_pq.cu); single-buffer kernel for m = 128
      static const bool use_db = !(getenv("MGPU_SCAN_DB") && getenv("MGPU_SCAN_DB")[0] == '0');
      if (use_db) {
        int s = launch_scan_pq_db(ivf, a);
        if (s != MGPU_ERR_UNSUPPORTED) return s;
      }
    }
      switch (a.ng) {
        case 1: return launch_scan_t<SCAN_PQ_FAST, 1>(ivf, a, L);
        case 2: return launch_scan_t<SCAN_PQ_FAST, 2>(ivf, a, L);
        case 3: return launch_scan_t<SCAN_PQ_FAST, 3, SCAN_FAST_NT>(ivf, a, L);
        case 4: return launch_scan_t<SCAN_PQ_FAST, 4>(ivf, a, L);
        default: return mgpu_fail(ctx, MGPU_ERR_UNSUPPORTED, "fast PQ scan supports m in {32,64,96,128}");
      }
    case SCAN_PQ_GENERIC: return launch_scan_t<SCAN_PQ_GENERIC, 1>(ivf, a, L);
    case SCAN_FLAT_L2: return launch_scan_t<SCAN_FLAT_L2, 1>(ivf, a, L);
    default: return launch_scan_t<SCAN_FLAT_DOT, 1>(ivf, a, L);
  }
}

// ---- layout builders --------------------------------------------------------------------------------------------------
// slot -> point id, then scatter the rows (indexed by point id) into the chunked layouts described in internal.cuh.
__global__ void k_layout_pq_fast(const uint8_t *__restrict__ codes_by_pid, const uint32_t *__restrict__ slot_pid,
                                 uint64_t nslots, uint32_t m, uint32_t ng, uint8_t *__restrict__ out) {
  // one thread per (slot, group, unit): writes 16 bytes
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t total = nslots * ng * 2;
  if (i >= total) return;
  uint32_t u = (uint32_t)(i % 2);
  uint32_t g = (uint32_t)((i / 2) % ng);
  uint64_t slot = i / (2 * ng);
  uint32_t lane = (uint32_t)(slot & 31);
  uint64_t chunk = slot >> 5;
  uint32_t pid = slot_pid[slot];
  uint32_t w[4] = {0, 0, 0, 0};
  if (pid != MGPU_EMPTY_SLOT) {
    const uint8_t *row = codes_by_pid + (size_t)pid * m + g * 32;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      uint32_t t = u * 16 + b;
      uint32_t code = row[lane ^ t];
      w[b >> 2] |= code << ((b & 3) * 8);
    }
  }
  uint4 *dst = (uint4 *)out + chunk * pq_fast_chunk_u4(ng) + ((size_t)g * 2 + u) * 32 + lane;
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
  // the chunk's point ids ride behind its code units (one thread per row writes its own)
  if (g == 0 && u == 0) ((uint32_t *)((uint4 *)out + chunk * pq_fast_chunk_u4(ng) + (size_t)ng * 64))[lane] = pid;
}

__global__ void k_layout_pq_generic(const uint8_t *__restrict__ codes_by_pid, const uint32_t *__restrict__ slot_pid,
                                    uint64_t nslots, uint32_t m, uint8_t *__restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots * m) return;
  uint64_t slot = i / m;
  uint32_t s = (uint32_t)(i % m);
  uint32_t pid = slot_pid[slot];
  out[i] = pid != MGPU_EMPTY_SLOT ? codes_by_pid[(size_t)pid * m + s] : 0;
}

__global__ void k_layout_flat(const float *__restrict__ rows_by_pid, const uint32_t *__restrict__ slot_pid,
                              uint64_t nslots, uint32_t dim, uint32_t dim4, float *__restrict__ out) {
  // one thread per (chunk, d4, lane): writes one float4; consecutive threads -> consecutive lanes (coalesced writes)
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots * dim4) return;
  uint32_t lane = (uint32_t)(i & 31);
  uint64_t rest = i >> 5;
  uint32_t d4 = (uint32_t)(rest % dim4);
  uint64_t chunk = rest / dim4;
  uint32_t pid = slot_pid[chunk * 32 + lane];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (pid != MGPU_EMPTY_SLOT) {
    const float *row = rows_by_pid + (size_t)pid * dim;
#pragma unroll
    for (int e = 0; e < 4; e++) if (d4 * 4 + e < dim) v[e] = row[d4 * 4 + e];
  }
  ((float4 *)out)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

int launch_build_layout(mgpu_ivf *ivf, const void *d_rows_by_pid) {
  mgpu_ctx *ctx = ivf->ctx;
  uint64_t nslots = ivf->total_chunks * 32;
  if (nslots == 0) return MGPU_OK;
  LaunchScope ls(ctx, MGPU_K_OTHER);
  if (ivf->quant == MGPU_QUANT_PQ) {
    uint32_t m = ivf->pq->m;
    if (ivf->pq_fast) {
      uint64_t total = nslots * ivf->ng * 2;
      k_layout_pq_fast<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t *)d_rows_by_pid, ivf->d_slot_pid, nslots, m, ivf->ng, ivf->d_codes);
    } else {
      uint64_t total = nslots * m;
      k_layout_pq_generic<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t *)d_rows_by_pid, ivf->d_slot_pid, nslots, m, ivf->d_codes);
    }
  } else {
    uint64_t total = nslots * ivf->dim4;
    k_layout_flat<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const float *)d_rows_by_pid, ivf->d_slot_pid, nslots, ivf->dim, ivf->dim4, ivf->d_rows);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

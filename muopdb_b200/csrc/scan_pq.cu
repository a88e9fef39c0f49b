// scan_pq.cu -- warp-specialised, double-buffered PQ posting-list scan (the headline kernel).
//
// Reference work replaced: BlockBasedIvf::scan_posting_list + search_with_centroids with a ProductQuantizer
// (rs/index/src/ivf/block_based/index.rs:175-285; distance = pq/mod.rs:231-266), for a batch of queries.
//
// One persistent CTA per SM, two roles:
//   * NPW producer warps build the NEXT query's fixed-point LUT (gather of m table rows of 1 KB, scale, transpose into the
//     conflict-free [code][column] image) and its probe-list prefix into the idle half of a double buffer;
//   * NCW consumer warps scan the CURRENT query's probed lists, one 32-row chunk per warp iteration, one row per lane:
//     PRMT (code byte -> LUT byte offset) + LDS + IADD3 per lookup, zero bank conflicts by construction (lane l reads column
//     l ^ t at step t; the HBM code layout is pre-permuted to match, see internal.cuh).
// full[2]/empty[2] mbarriers hand the buffers over, so the latency-bound LUT build never stalls the shared-memory
// pipe the scan is bound by.
//
// Shared memory (NG = m/32 groups, 32 KB of LUT per group): the two LUTs are interleaved at 128 B granularity inside
// 64 KB blocks with a 256 B pitch per code, because PRMT can only place the code byte at a byte boundary (code << 8):
//   slot s = parity * NG + g  ->  block s/2, half s%2;  address = block*64K + code*256 + half*128 + column*4.
// NG = 3 (m = 96): 3 blocks = 192 KB for both LUTs.
#include "internal.cuh"
#include "scan_common.cuh"

#include <algorithm>
#include <functional>

#ifdef MGPU_SCAN_DBG
// experiment build only (make DBG=1): bit0 = no row ever passes, bit1 = skip the merge rounds; g_dbg = pass statistics
__device__ unsigned long long g_dbg[8];
__constant__ unsigned int c_dbg;
#endif

struct DbLayout {
  uint32_t lut_bytes, off_ctab, off_pref, off_pcs, off_plen, off_mkey, off_mpay, off_misc, total, maxp;
};

// Chunk table: one word per 32-row chunk of the query's probed lists, in scan order:
//   (global chunk index << 5) | (rows in the chunk - 1).
// The consumers get their chunk with one shared-memory load (no prefix search over the probe list) and know which lanes
// hold a row without touching slot_pid.  Queries with more chunks than the table holds use the prefix-search path.
#define DB_CTAB_CAP 2048

__host__ __device__ inline DbLayout db_layout(uint32_t ng, uint32_t ncw, uint32_t max_probes) {
  DbLayout L;
  L.maxp = max_probes;
  L.lut_bytes = ((2 * ng + 1) / 2) * 65536u;
  L.off_ctab = L.lut_bytes;
  L.off_pref = L.off_ctab + 2u * DB_CTAB_CAP * 4u;
  L.off_pcs = L.off_pref + 2u * (max_probes + 1) * 4u;
  L.off_plen = L.off_pcs + 2u * max_probes * 4u;
  L.off_mkey = (L.off_plen + 2u * max_probes * 4u + 15u) & ~15u;
  L.off_mpay = L.off_mkey + ncw * 32u * 8u;
  L.off_misc = L.off_mpay + ncw * 32u * 4u;
  // misc: bars[4] (32 B) | qinfo[2][4] (32 B) | thr (4) pad | b2[32] | shf[32] | pq[4] | soff[128] | scode[128]
  L.total = L.off_misc + 64u + 64u + 128u + 128u + 16u + 512u + 512u;
  return L;
}

// Consumer side of k_scan_pq_db for one query; P = buffer parity (compile time => LUT offsets are LDS immediates).
template <int NG, int NCW>
struct DbConsumerCtx {
  const ScanArgs &a; const DbLayout &L;
  uint8_t *lut; uint32_t *pref, *pcs, *ctab; uint64_t *mkey; uint32_t *mpay; uint64_t *bars; uint32_t *qinfo, *thr_p, *b2;
  const uint32_t *xr; int lane, warp;

  // scores one 32-row chunk held in u[] against LUT buffer P
  template <int P>
  __device__ __forceinline__ uint32_t score_group(const uint4 (&u)[NG * 2], int g) const {
    uint32_t acc0 = 0, acc1 = 0;
    const uint32_t slotg = P * NG + g;
    const uint8_t *lut_g = lut + (slotg >> 1) * 65536u + (slotg & 1) * 128u;
#pragma unroll
    for (int wi = 0; wi < 8; wi++) {
      const uint4 &uu = u[g * 2 + (wi >> 2)];
      const uint32_t w = (wi & 3) == 0 ? uu.x : ((wi & 3) == 1 ? uu.y : ((wi & 3) == 2 ? uu.z : uu.w));
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int t = wi * 4 + k;
        const uint32_t idx = prmt(w, xr[t >> 2], ((12 + (t & 3)) << 12) | ((12 + (t & 3)) << 8) | (k << 4) | (4 + (t & 3)));
        const uint32_t v = *(const uint32_t *)(lut_g + idx);
        if (t & 1) acc1 += v; else acc0 += v;
      }
    }
    return acc0 + acc1;
  }

  // threshold maintenance shared by both scan paths: offer the passing rows to the warp's top-32, publish the new bound
  __device__ __forceinline__ void offer_rows(WarpTop32 &top, bool &first, bool pass, uint32_t key, uint32_t pid, uint32_t slot,
                                             uint32_t thr) {
    constexpr int NCW_ = NCW;
    // "max over warps of their 2nd best" bounds the global 32nd best only if the warps' two best already are >= 32 rows
    static_assert(2 * NCW >= MGPU_NCAND && NCW <= 32, "threshold rule needs 16..32 consumer warps");
    uint32_t worst;
#ifdef MGPU_SCAN_DBG
    { unsigned mm = __ballot_sync(0xffffffffu, pass); if (lane == 0) { atomicAdd(&g_dbg[0], 1ull); atomicAdd(&g_dbg[1], (unsigned long long)__popc(mm)); if (first) atomicAdd(&g_dbg[2], 1ull);} }
#endif
    if (first) {
      top.key = pass ? (((uint64_t)key << 32) | pid) : MGPU_EMPTY_KEY;
      top.pay = pass ? slot : MGPU_EMPTY_SLOT;
      top.sort();
      worst = (uint32_t)(shfl64(top.key, 31) >> 32);
      first = false;
    } else {
      worst = top.offer(pass, ((uint64_t)key << 32) | pid, slot);
    }
    // threshold: min(this warp's 32nd best, max over warps of their 2nd best) -- both bound the global 32nd best
    uint32_t second = (uint32_t)(shfl64(top.key, 1) >> 32);
    if (lane == 0) b2[warp] = second;
    __syncwarp();
    uint32_t v = lane < NCW_ ? *(volatile uint32_t *)&b2[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    v = min(v, worst);
    if (lane == 0 && v < thr) atomicMin(thr_p, v);
  }

  // table-driven scan of one query (total <= DB_CTAB_CAP chunks): one LDS per chunk for its address and row count; point
  // ids, the invalidation bitmap and the planner filter are only touched for rows that pass the running threshold
  template <int P>
  __device__ __forceinline__ void scan_table(uint32_t q, uint32_t total, WarpTop32 &top) {
    const uint32_t *ct = ctab + P * DB_CTAB_CAP;
    bool first = true;
    uint4 u[NG * 2];
    uint32_t it = warp, e_n = 0, pid_n = MGPU_EMPTY_SLOT;
    if (it < total) {
      e_n = ct[it];
      const uint4 *base = (const uint4 *)a.codes + (size_t)(e_n >> 5) * (NG * 64 + 8) + lane;
#pragma unroll
      for (int i = 0; i < NG * 2; i++) u[i] = ldg_stream16(base + i * 32);
      pid_n = a.slot_pid[(e_n >> 5) * 32 + lane];
    }
#pragma unroll 1
    while (it < total) {
      const uint32_t e = e_n;
      // the point ids travel with the codes (one coalesced 128-byte load per chunk): a row that passes the threshold does
      // not wait a memory round trip for its id (with a few chunks per warp and query those waits add up)
      uint32_t pid = pid_n;
      it += NCW;
      const bool more = it < total;
      const uint4 *nbase = (const uint4 *)a.codes;
      if (more) {
        e_n = ct[it];
        nbase = (const uint4 *)a.codes + (size_t)(e_n >> 5) * (NG * 64 + 8) + lane;
        pid_n = a.slot_pid[(e_n >> 5) * 32 + lane];
      }
      uint32_t key = 0;
#pragma unroll
      for (int g = 0; g < NG; g++) {
        key += score_group<P>(u, g);
        if (more) {
          u[g * 2] = ldg_stream16(nbase + (g * 2) * 32);
          u[g * 2 + 1] = ldg_stream16(nbase + (g * 2 + 1) * 32);
        }
      }
      const uint32_t thr = *(volatile uint32_t *)thr_p;
      bool pass = (uint32_t)lane <= (e & 31u) && key <= thr;
      if (__any_sync(0xffffffffu, pass)) {
        const uint32_t slot = (e >> 5) * 32 + lane;
        if (pass) {
          if (a.invalid && ((a.invalid[pid >> 5] >> (pid & 31)) & 1u)) pass = false;            // index.rs:198-200
          if (a.filter && pass && !((a.filter[(size_t)q * a.filter_stride + (pid >> 5)] >> (pid & 31)) & 1u)) pass = false;  // :212-226
          if (a.lower_bound && pass && (((uint64_t)key << 32) | pid) < a.lower_bound[q]) pass = false;  // reported by an earlier round
        }
        if (__any_sync(0xffffffffu, pass)) offer_rows(top, first, pass, key, pid, slot, thr);
      }
    }
  }

  // prefix-search scan of one query (more chunks than the table holds; rare): lives in its own kernel instantiation
  // (TAB = false) so that its registers do not weigh on the table-driven loop
  template <int P>
  __device__ __forceinline__ void scan_search(uint32_t q, uint32_t total, const uint32_t *prefp, const uint32_t *pcsp, WarpTop32 &top) {

    bool first = true;
    uint32_t pi = 0;
    uint32_t hi_it = 0, base_off = 0;  // current probed list covers chunk iterations [.., hi_it); chunk = base_off + it
    // software pipeline with NO extra registers: the 2 x 16 B of group g are re-loaded for the NEXT chunk right after group
    // g of the current chunk has been scored, so every load has ~a full chunk of lookups (2 groups + top-k) to land.
    uint4 u[NG * 2];
    uint32_t pid_n = MGPU_EMPTY_SLOT, slot_n = 0;
    uint32_t it = warp;
    if (it < total) {
      while (it >= prefp[pi + 1]) pi++;
      hi_it = prefp[pi + 1]; base_off = pcsp[pi] - prefp[pi];
      const uint32_t chunk = base_off + it;
      slot_n = chunk * 32 + lane;
      const uint4 *base = (const uint4 *)a.codes + (size_t)chunk * (NG * 64 + 8) + lane;
#pragma unroll
      for (int i = 0; i < NG * 2; i++) u[i] = ldg_stream16(base + i * 32);
      pid_n = a.slot_pid[slot_n];
    }
#pragma unroll 1
    while (it < total) {
      const uint32_t pid = pid_n, slot = slot_n;
      it += NCW;
      const bool more = it < total;
      const uint4 *nbase = (const uint4 *)a.codes;
      if (more) {
        if (it >= hi_it) {  // crossed into another probed list: refresh the cached bounds (rare)
          while (it >= prefp[pi + 1]) pi++;
          hi_it = prefp[pi + 1]; base_off = pcsp[pi] - prefp[pi];
        }
        const uint32_t chunk = base_off + it;
        slot_n = chunk * 32 + lane;
        nbase = (const uint4 *)a.codes + (size_t)chunk * (NG * 64 + 8) + lane;
        pid_n = a.slot_pid[slot_n];
      }
      bool valid = pid != MGPU_EMPTY_SLOT;
      if (a.invalid && valid) valid = !((a.invalid[pid >> 5] >> (pid & 31)) & 1u);  // index.rs:198-200
      if (a.filter && valid) valid = (a.filter[(size_t)q * a.filter_stride + (pid >> 5)] >> (pid & 31)) & 1u;  // index.rs:212-226

      uint32_t key = 0;
#pragma unroll
      for (int g = 0; g < NG; g++) {
        key += score_group<P>(u, g);
        if (more) {
          u[g * 2] = ldg_stream16(nbase + (g * 2) * 32);
          u[g * 2 + 1] = ldg_stream16(nbase + (g * 2 + 1) * 32);
        }
      }
      const uint32_t thr = *(volatile uint32_t *)thr_p;
      bool pass = valid && key <= thr;
      if (a.lower_bound && pass && (((uint64_t)key << 32) | pid) < a.lower_bound[q]) pass = false;  // reported by an earlier round
      if (__any_sync(0xffffffffu, pass)) offer_rows(top, first, pass, key, pid, slot, thr);
    }
  }

  template <int P, bool TAB>
  __device__ __forceinline__ bool query(uint32_t phase) {
    constexpr int NCT = NCW * 32;
    warp_mbar_wait(&bars[P], phase);  // buffer P is full
    const uint32_t q = qinfo[P * 4 + 0];
    if (q == 0xFFFFFFFFu) return false;
    const uint32_t total = qinfo[P * 4 + 1];
    const uint32_t *prefp = pref + P * (L.maxp + 1), *pcsp = pcs + P * L.maxp;
    if (warp == 0) {
#ifdef MGPU_SCAN_DBG
      if (lane == 0) *thr_p = (c_dbg & 1) ? 0u : 0xFFFFFFFFu;
#else
      if (lane == 0) *thr_p = 0xFFFFFFFFu;
#endif
      b2[lane] = 0xFFFFFFFFu;
    }
    named_bar_sync(1, NCT);

    WarpTop32 top;
    top.init();
    if (TAB) scan_table<P>(q, total, top);
    else scan_search<P>(q, total, prefp, pcsp, top);
    // this warp no longer needs LUT[P] / pref[P]: hand the buffer back to the producers before the merge
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[2 + P]);

    // ---- merge the warp lists (shuffle bitonic merges, log2 rounds) -------------------------------------------------
    mkey[warp * 32 + lane] = top.key;
    mpay[warp * 32 + lane] = top.pay;
    named_bar_sync(1, NCT);
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
#ifdef MGPU_SCAN_DBG
      if (c_dbg & 2) break;
#endif
      if (warp < half && warp + half < NCW) {
        top.merge(mkey[(warp + half) * 32 + lane], mpay[(warp + half) * 32 + lane]);
        mkey[warp * 32 + lane] = top.key;
        mpay[warp * 32 + lane] = top.pay;
      }
      named_bar_sync(1, NCT);
    }
    if (warp == 0) {
      a.cand_key[(size_t)q * MGPU_NCAND + lane] = top.key;
      a.cand_slot[(size_t)q * MGPU_NCAND + lane] = top.pay;
    }
    return true;
  }
};

// TAB = true : table-driven consumers; a query with more than DB_CTAB_CAP chunks is appended to a.overflow_list instead of
//              being scanned.  TAB = false: prefix-search consumers over the queries of a.overflow_list (second launch, only
//              issued when the index could produce such a query at all).
template <int NG, int NCW, int NPW, bool TAB>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1) k_scan_pq_db(ScanArgs a, DbLayout L) {
  constexpr int NCT = NCW * 32, NPT = NPW * 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *lut = smem;
  uint32_t *ctab = (uint32_t *)(smem + L.off_ctab);  // [2][DB_CTAB_CAP]
  uint32_t *pref = (uint32_t *)(smem + L.off_pref);  // [2][maxp + 1]
  uint32_t *pcs = (uint32_t *)(smem + L.off_pcs);    // [2][maxp]
  uint32_t *plen = (uint32_t *)(smem + L.off_plen);  // [2][maxp]
  uint64_t *mkey = (uint64_t *)(smem + L.off_mkey);
  uint32_t *mpay = (uint32_t *)(smem + L.off_mpay);
  uint64_t *bars = (uint64_t *)(smem + L.off_misc);        // full[0], full[1], empty[0], empty[1]
  uint32_t *qinfo = (uint32_t *)(smem + L.off_misc + 32);  // [2][4]: query, total chunks, nprobe
  uint32_t *thr_p = (uint32_t *)(smem + L.off_misc + 64);
  uint32_t *b2 = (uint32_t *)(smem + L.off_misc + 128);
  float *shf = (float *)(smem + L.off_misc + 256);
  uint32_t *pq = (uint32_t *)(smem + L.off_misc + 384);
  float *soff = (float *)(smem + L.off_misc + 400);
  uint32_t *scode = (uint32_t *)(smem + L.off_misc + 400 + 512);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t K = 256;
  const uint32_t m = NG * 32;

  if (tid == 0) {
    mbar_init(&bars[0], NPT); mbar_init(&bars[1], NPT);
    mbar_init(&bars[2], NCW); mbar_init(&bars[3], NCW);
  }
  __syncthreads();

  if (warp >= NCW) {
    // =========================================== PRODUCER ===========================================================
    const int ptid = tid - NCT, pwarp = warp - NCW;
#pragma unroll 1
    for (uint32_t n = 0;; n++) {
      const uint32_t p = n & 1;
      warp_mbar_wait(&bars[2 + p], ((n >> 1) & 1) ^ 1, 2000);  // buffer p is free
    next_query:
      if (ptid == 0) {  // dynamic query scheduling over the persistent CTAs, longest queries first (a.order)
        uint32_t t = atomicAdd(a.next_query, 1u);
        if (TAB) pq[0] = t < a.B ? (a.order ? a.order[t] : t) : 0xFFFFFFFFu;
        else pq[0] = t < *a.overflow_count ? a.overflow_list[t] : 0xFFFFFFFFu;
      }
      named_bar_sync(2, NPT);
      const uint32_t q = pq[0];
      if (q >= a.B) {
        if (ptid == 0) qinfo[p * 4] = 0xFFFFFFFFu;
        mbar_arrive(&bars[p]);
        break;
      }
      uint32_t *prefp = pref + p * (L.maxp + 1), *pcsp = pcs + p * L.maxp, *plenp = plen + p * L.maxp;
      const uint32_t np = a.probe_counts ? min(a.probe_counts[q], a.max_probes) : a.max_probes;
      // ---- probe-list prefix (chunk counts) --------------------------------------------------------------------------
      if (pwarp == 0) {
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
          if (i < np) { prefp[i] = run + incl - cnt; pcsp[i] = cs; plenp[i] = len; }
          run += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
          rows += len;
        }
        if (lane == 0) {
          prefp[np] = run;
          qinfo[p * 4 + 0] = q; qinfo[p * 4 + 1] = run; qinfo[p * 4 + 2] = np;
          const bool mine = TAB ? run <= DB_CTAB_CAP : true;
          if (a.rows_scanned && mine) atomicAdd(a.rows_scanned, rows);
          if (TAB && !mine) a.overflow_list[atomicAdd(a.overflow_count, 1u)] = q;  // left to the prefix-search launch
        }
      }
      // ---- fixed-point scale: the sum over subspaces of the LUT row ranges must fit 32 bits ---------------------------
      float part = 0.0f;
      for (uint32_t s = ptid; s < m; s += NPT) {
        uint32_t code = a.qcodes[(size_t)q * m + s];
        float mn = a.rowmin[(size_t)s * K + code], mx = a.rowmax[(size_t)s * K + code];
        soff[s] = mn; scode[s] = code;
        part += mx - mn;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) shf[pwarp] = part;
      named_bar_sync(2, NPT);
      float R = 0.0f;
#pragma unroll
      for (int w = 0; w < NPW; w++) R += shf[w];
      const float scale = R > 0.0f ? 4.0e9f / R : 0.0f;
      // ---- chunk table (the prefix arrays were published by the barrier above) -----------------------------------------
      if (TAB) {
        if (prefp[np] > DB_CTAB_CAP) {   // uniform over the producer warps: skip this query, fetch another one
          named_bar_sync(2, NPT);        // everybody has read prefp[np] before pwarp 0 overwrites it
          goto next_query;
        }
        uint32_t *ct = ctab + p * DB_CTAB_CAP;
        for (uint32_t i = pwarp; i < np; i += NPW) {
          const uint32_t first_it = prefp[i], cnt = prefp[i + 1] - first_it, cs = pcsp[i], len = plenp[i];
          for (uint32_t j = lane; j < cnt; j += 32) ct[first_it + j] = ((cs + j) << 5) | (min(32u, len - 32u * j) - 1u);
        }
      }
      // ---- LUT image, no staging: lane <-> LUT column (subspace), each lane streams its own 1 KB table row with 16-byte
      // loads and stores four codes' entries; every store is one conflict-free 128-byte wavefront (32 lanes = 32
      // consecutive columns of one code row).  The producer warps split the 256 codes.
#ifdef MGPU_SCAN_DBG
      if (!(c_dbg & 4))
#endif
#ifdef MGPU_SCAN_DBG
      if (!(c_dbg & 4))
#endif
#ifdef MGPU_SCAN_DBG
      if (!(c_dbg & 4))
#endif
      {
        constexpr int CODES_PER_WARP = 256 / NPW;
#pragma unroll 1
        for (int g = 0; g < NG; g++) {
          const uint32_t slot = p * NG + g;
          uint8_t *dstg = lut + (slot >> 1) * 65536u + (slot & 1) * 128u + lane * 4;
          const uint32_t s = g * 32 + lane;
          const float off = soff[s];
          const float4 *src = (const float4 *)(a.table + ((size_t)s * K + scode[s]) * K) + pwarp * (CODES_PER_WARP / 4);
#pragma unroll 1
          for (int j4 = 0; j4 < CODES_PER_WARP / 4; j4 += 8) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __ldg(src + j4 + i);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int j = pwarp * CODES_PER_WARP + (j4 + i) * 4;
              *(uint32_t *)(dstg + (j + 0) * 256) = __float2uint_rn(__fmul_rn(__fsub_rn(v[i].x, off), scale));
              *(uint32_t *)(dstg + (j + 1) * 256) = __float2uint_rn(__fmul_rn(__fsub_rn(v[i].y, off), scale));
              *(uint32_t *)(dstg + (j + 2) * 256) = __float2uint_rn(__fmul_rn(__fsub_rn(v[i].z, off), scale));
              *(uint32_t *)(dstg + (j + 3) * 256) = __float2uint_rn(__fmul_rn(__fsub_rn(v[i].w, off), scale));
            }
          }
        }
      }
      mbar_arrive(&bars[p]);  // buffer p is full (release: every producer thread's writes precede its arrive)
    }
    return;
  }

  // ============================================= CONSUMER ===========================================================
  // lane-dependent LUT column offsets: x_t = ((lane ^ t) * 4) <= 124, four per register.  PRMT builds (code << 8) | x_t in
  // ONE instruction: byte0 = x_t, byte1 = code, bytes 2/3 = sign replication of x_t (msb 0) = 0.
  uint32_t xr[8];
#pragma unroll
  for (int j = 0; j < 8; j++)
    xr[j] = (uint32_t)((lane ^ (4 * j)) << 2) | ((uint32_t)((lane ^ (4 * j + 1)) << 2) << 8) |
            ((uint32_t)((lane ^ (4 * j + 2)) << 2) << 16) | ((uint32_t)((lane ^ (4 * j + 3)) << 2) << 24);

  DbConsumerCtx<NG, NCW> cc{a, L, lut, pref, pcs, ctab, mkey, mpay, bars, qinfo, thr_p, b2, xr, lane, warp};
#pragma unroll 1
  for (uint32_t n2 = 0;; n2++) {
    // the two buffer parities are unrolled so that every LUT offset is an immediate of the LDS
    if (!cc.template query<0, TAB>(n2 & 1)) break;
    if (!cc.template query<1, TAB>(n2 & 1)) break;
  }
}

template <int NG, int NCW, int NPW>
static int launch_db_t(mgpu_ivf *ivf, const ScanArgs &a0) {
  mgpu_ctx *ctx = ivf->ctx;
  ScanArgs a = a0;
  DbLayout L = db_layout(NG, NCW, a.max_probes);
  if (L.total > ctx->smem_optin) return MGPU_ERR_UNSUPPORTED;
  // overflow list for queries whose probed lists hold more chunks than the shared-memory chunk table
  if (!ivf->d_scan_overflow || ivf->scan_overflow_cap < a.B) {
    if (ivf->d_scan_overflow) { cudaStreamSynchronize(ctx->stream); cudaFree(ivf->d_scan_overflow); ivf->d_scan_overflow = nullptr; }
    CUDA_TRY(ctx, cudaMalloc((void **)&ivf->d_scan_overflow, ((size_t)a.B + 1) * 4));
    ivf->scan_overflow_cap = a.B;
  }
  a.overflow_count = ivf->d_scan_overflow;
  a.overflow_list = ivf->d_scan_overflow + 1;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan_pq_db<NG, NCW, NPW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  unsigned grid = a.B < (uint32_t)ctx->sm_count ? a.B : (unsigned)ctx->sm_count;
  CUDA_TRY(ctx, cudaMemsetAsync(a.next_query, 0, 4, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(a.overflow_count, 0, 4, ctx->stream));
#ifdef MGPU_SCAN_DBG
  {
    static int n = 0;
    unsigned int d = getenv("MGPU_SCAN_DBG") ? atoi(getenv("MGPU_SCAN_DBG")) : 0;
    cudaMemcpyToSymbolAsync(c_dbg, &d, 4, 0, cudaMemcpyHostToDevice, ctx->stream);
    if (++n == 8) { unsigned long long h[8]; cudaStreamSynchronize(ctx->stream); cudaMemcpyFromSymbol(h, g_dbg, 64);
      fprintf(stderr, "[scan dbg] after 7 launches: offer calls %llu rows offered %llu first %llu\n", h[0], h[1], h[2]); }
  }
#endif
  {
    LaunchScope ls(ctx, MGPU_K_SCAN, nullptr, "k_scan_pq_db (32-bit LUT, scan_pq.cu)");
    k_scan_pq_db<NG, NCW, NPW, true><<<grid, (NCW + NPW) * 32, L.total, ctx->stream>>>(a, L);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  // can any max_probes lists of this index exceed the table?  (sum of the largest chunk counts; cached per probe count)
  if (ivf->scan_bound_probes != a.max_probes) {
    std::vector<uint32_t> c(ivf->h_list_len.size());
    for (size_t i = 0; i < c.size(); i++) c[i] = (ivf->h_list_len[i] + 31) / 32;
    const size_t k = std::min<size_t>(a.max_probes, c.size());
    std::partial_sort(c.begin(), c.begin() + k, c.end(), std::greater<uint32_t>());
    uint64_t sum = 0;
    for (size_t i = 0; i < k; i++) sum += c[i];
    ivf->scan_bound_probes = a.max_probes; ivf->scan_bound_chunks = sum;
  }
  if (ivf->scan_bound_chunks > DB_CTAB_CAP) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan_pq_db<NG, NCW, NPW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    CUDA_TRY(ctx, cudaMemsetAsync(a.next_query, 0, 4, ctx->stream));
    LaunchScope ls(ctx, MGPU_K_SCAN);
    k_scan_pq_db<NG, NCW, NPW, false><<<grid, (NCW + NPW) * 32, L.total, ctx->stream>>>(a, L);
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return MGPU_OK;
}

// Returns MGPU_ERR_UNSUPPORTED when the double-buffered kernel does not apply (caller falls back to k_scan).
int launch_scan_pq_db(mgpu_ivf *ivf, const ScanArgs &a) {
  if (!ivf->pq_fast || a.m > 128) return MGPU_ERR_UNSUPPORTED;
  static const int cfg = getenv("MGPU_DB_CFG") ? atoi(getenv("MGPU_DB_CFG")) : 0;
  switch (a.ng) {
    case 1: return launch_db_t<1, 16, 4>(ivf, a);
    case 2: return launch_db_t<2, 16, 4>(ivf, a);
    case 3:
      // measured on B200 (profiles/): 16 consumer + 4 producer warps (96 registers, no spills) is the best split
      switch (cfg) {
        case 1: return launch_db_t<3, 20, 4>(ivf, a);
        case 4: return launch_db_t<3, 16, 8>(ivf, a);
        default: return launch_db_t<3, 16, 4>(ivf, a);
      }
    default: return MGPU_ERR_UNSUPPORTED;
  }
}

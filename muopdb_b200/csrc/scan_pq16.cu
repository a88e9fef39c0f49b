// scan_pq16.cu -- PQ posting-list scan, third generation: 16-bit LUT entries with ONE scale per quantizer, the LUT image
// gathered by bulk copies (TMA, cp.async.bulk) and transposed out of a staging tile.  Same roles as scan_pq.cu (persistent CTA per SM, producer warps prepare the NEXT query
// while consumer warps scan the CURRENT one, full/empty mbarriers), different data movement:
//
//   * table16 (pq.cu) holds rn((table[s][a][b] - rowmin[s][a]) * gscale) as u16, so a query's LUT is a pure GATHER of m rows
//     of 512 bytes: the producers issue one cp.async.bulk per row into a padded staging tile (no LSU cycles for the global
//     side), then transpose it into the [code][column] image with conflict-free LDS.128 / STS.32 (two groups' entries packed
//     in one word).  The old build streamed 96 KB per query through LDG with 32 different lines per instruction (~6.9 k
//     L1 data-pipe cycles per query); this one costs ~0.9 k.
//   * the producer's chain of dependent global loads (query id -> codes -> probes -> list bounds) is overlapped with the
//     gather: next query id fetched one iteration ahead, rows requested as soon as the codes are known, probe-list prefix and
//     chunk table computed while the copies fly.  At the sharded shape (few chunks per query) the producer set the pace.
//   * lookups are LDS.U16 (one wavefront each, bank = column, conflict free as before); keys are u32 sums of u16 entries.
//
// Exactness: keys rank candidates, finalize.cu re-scores the 32 survivors with the reference's arithmetic and CERTIFIES the
// answer (the k-th exact candidate's key + 2 E < the 32nd key, E = rigorous bound of |key - gscale * exact score|); a query
// that cannot be certified is re-scanned by the exact kernel (scan.cu: SCAN_PQ_EXACT).  L2 metric only (rowmin == 0, all
// terms non-negative: the fp32 error bound is relative to the score).
//
// Shared memory (NG = m / 32 groups; NP = NG / 2 pair planes per parity, NS = NG % 2 single plane):
//   pair planes   [code][128 B]: word (code, col) = entry(group 2i) | entry(group 2i+1) << 16; two planes interleaved per
//                 64 KB block at a 256 B pitch so that PRMT builds the byte offset (code << 8) | (col << 2)
//   single plane  (odd NG) [code][128 B]: one u16 entry per word, the two parities interleaved in one 64 KB block
//   NG = 3: 128 KB of LUT (both parities), 50 KB staging tile (all 96 rows of a query at once), 16 KB chunk tables.
#include "internal.cuh"
#include "scan_common.cuh"

#include <algorithm>
#include <functional>

#ifdef MGPU_SCAN_DBG
// experiment build only (make DBG=1): cycle counters per role/phase, summed over CTAs (lane 0 of producer warp 0 / consumer warp 0)
__device__ unsigned long long g_dbg16[16];
#define DBG_T(var) const long long var = clock64()
#define DBG_ADD(i, t0, t1) do { if (lane == 0) atomicAdd(&g_dbg16[i], (unsigned long long)((t1) - (t0))); } while (0)
#else
#define DBG_T(var)
#define DBG_ADD(i, t0, t1)
#endif

#define P16_CTAB_CAP 2048
#define P16_STAGE_PITCH 528   /* 512 B row + 16: quarter-warp LDS.128 over 8 rows hits 32 different banks */

struct P16Layout {
  uint32_t lut_bytes, off_single, off_stage, off_ctab, off_pref, off_pcs, off_plen, off_mkey, off_mpay, off_misc, total, maxp;
};

__host__ __device__ inline P16Layout p16_layout(uint32_t ng, uint32_t ncw, uint32_t max_probes) {
  P16Layout L;
  L.maxp = max_probes;
  const uint32_t npair_planes = 2 * (ng / 2);          // both parities
  const uint32_t pair_bytes = ((npair_planes + 1) / 2) * 65536u;
  L.off_single = pair_bytes;
  L.lut_bytes = pair_bytes + (ng & 1u) * 65536u;
  L.off_stage = L.lut_bytes;
  L.off_ctab = L.off_stage + ng * 32u * P16_STAGE_PITCH;
  L.off_pref = L.off_ctab + 2u * P16_CTAB_CAP * 4u;
  L.off_pcs = L.off_pref + 2u * (max_probes + 1) * 4u;
  L.off_plen = L.off_pcs + 2u * max_probes * 4u;
  L.off_mkey = (L.off_plen + 2u * max_probes * 4u + 15u) & ~15u;
  L.off_mpay = L.off_mkey + ncw * 32u * 8u;
  L.off_misc = L.off_mpay + ncw * 32u * 4u;
  // misc: bars full[2] empty[2] stage (8 B each) | qinfo[2][4] | thr | b2[32] | pq[4] | scode[128]
  L.total = L.off_misc + 320u + 32u + 32u + 128u + 16u + 512u;
  return L;
}

__device__ __forceinline__ void p16_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p16_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void p16_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NG, int NCW>
struct P16Consumer {
  static constexpr int NP = NG / 2, NS = NG % 2;
  const ScanArgs &a; const P16Layout &L;
  uint8_t *lut; uint32_t *ctab; uint64_t *mkey; uint32_t *mpay; uint64_t *bars; uint32_t *qinfo, *thr_p, *b2;
  const uint32_t *xr; int lane, warp;

  // one 32-subspace group of the chunk in u[] against LUT parity P
  template <int P>
  __device__ __forceinline__ uint32_t score_group(const uint4 (&u)[NG * 2], int g) const {
    uint32_t acc0 = 0, acc1 = 0;
    // pair planes: two groups' entries share a word; the single plane (odd NG) holds one u16 per word, parity at +128 B
    const bool paired = g < 2 * NP;
    const uint32_t plane = P * NP + (g >> 1);
    const uint8_t *base = paired ? lut + (plane >> 1) * 65536u + (plane & 1) * 128u + (g & 1) * 2u : lut + L.off_single + P * 128u;
#pragma unroll
    for (int wi = 0; wi < 8; wi++) {
      const uint4 &uu = u[g * 2 + (wi >> 2)];
      const uint32_t w = (wi & 3) == 0 ? uu.x : ((wi & 3) == 1 ? uu.y : ((wi & 3) == 2 ? uu.z : uu.w));
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int t = wi * 4 + k;
        const uint32_t idx = prmt(w, xr[t >> 2], ((12 + (t & 3)) << 12) | ((12 + (t & 3)) << 8) | (k << 4) | (4 + (t & 3)));  // (code << 8) | (col << 2)
        const uint32_t v = *(const uint16_t *)(base + idx);
        if (t & 1) acc1 += v; else acc0 += v;
      }
    }
    return acc0 + acc1;
  }

  __device__ __forceinline__ void offer_rows(WarpTop32 &top, bool &first, bool pass, uint32_t key, uint32_t pid, uint32_t slot,
                                             uint32_t thr) {
    static_assert(2 * NCW >= MGPU_NCAND && NCW <= 32, "threshold rule needs 16..32 consumer warps");
    uint32_t worst;
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
    uint32_t v = lane < NCW ? *(volatile uint32_t *)&b2[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    v = min(v, worst);
    if (lane == 0 && v < thr) atomicMin(thr_p, v);
  }

  // Table-driven scan of one query: one LDS per chunk for its record index and row count.  The code units are fetched with
  // plain 16-byte loads, software-pipelined with no extra registers (the 2 x 16 B of group g are re-loaded for the NEXT chunk
  // right after group g of the current chunk has been scored); the point ids ride in the chunk record's trailing line.
  // (A per-warp cp.async.bulk ring for the chunk records was built and measured: 0.66 ms vs 0.39 ms per launch at the headline
  // shape -- one 3.2 KB bulk copy per chunk occupies the SM's copy engine for ~300 cycles, more than the 16 consumer warps
  // leave between requests; profiles/r2_scan16_ring_ncu.json.  The bulk copies stayed where requests are few: the LUT rows.)
  template <int P>
  __device__ __forceinline__ void scan(uint32_t q, uint32_t total, WarpTop32 &top) {
    const uint32_t *ct = ctab + P * P16_CTAB_CAP;
    constexpr uint32_t CU4 = NG * 64 + 8;   // 16-byte words per chunk record
    bool first = true;
    uint4 u[NG * 2];
    uint32_t it = warp, e_n = 0, pid_n = MGPU_EMPTY_SLOT;
    if (it < total) {
      e_n = ct[it];
      const uint4 *base = (const uint4 *)a.codes + (size_t)(e_n >> 5) * CU4 + lane;
#pragma unroll
      for (int i = 0; i < NG * 2; i++) u[i] = ldg_stream16(base + i * 32);
      pid_n = ((const uint32_t *)((const uint4 *)a.codes + (size_t)(e_n >> 5) * CU4 + NG * 64))[lane];
    }
#pragma unroll 1
    while (it < total) {
      const uint32_t e = e_n;
      uint32_t pid = pid_n;
      it += NCW;
      const bool more = it < total;
      const uint4 *nbase = (const uint4 *)a.codes;
      if (more) {
        e_n = ct[it];
        nbase = (const uint4 *)a.codes + (size_t)(e_n >> 5) * CU4 + lane;
        pid_n = ((const uint32_t *)((const uint4 *)a.codes + (size_t)(e_n >> 5) * CU4 + NG * 64))[lane];
      }
      // (Early exit -- skipping the remaining groups once all 32 partial sums exceed the threshold, every LUT entry being
      // >= 0 -- was built and measured: -1 % at the headline shape, +3..5 % at the sharded shapes, where the extra vote and
      // threshold read per group outweigh the rare skip.  Not kept)
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
        }
        if (__any_sync(0xffffffffu, pass)) offer_rows(top, first, pass, key, pid, slot, thr);
      }
    }
  }

  template <int P>
  __device__ __forceinline__ bool query(uint32_t phase) {
    constexpr int NCT = NCW * 32;
    DBG_T(t0);
    warp_mbar_wait(&bars[P], phase);  // buffer P is full
    DBG_T(t1);
    const uint32_t q = qinfo[P * 4 + 0];
    if (q == 0xFFFFFFFFu) return false;
    const uint32_t total = qinfo[P * 4 + 1];
    if (warp == 0) {
      if (lane == 0) *thr_p = 0xFFFFFFFFu;
      b2[lane] = 0xFFFFFFFFu;
    }
    named_bar_sync(1, NCT);
    WarpTop32 top;
    top.init();
    DBG_T(t2);
    scan<P>(q, total, top);
    DBG_T(t3);
    // this warp no longer needs LUT[P] / ctab[P]: hand the buffer back to the producers before the merge
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[2 + P]);
    mkey[warp * 32 + lane] = top.key;
    mpay[warp * 32 + lane] = top.pay;
    named_bar_sync(1, NCT);
    DBG_T(t4);
    if (warp == 0) { DBG_ADD(0, t0, t1); DBG_ADD(1, t2, t3); DBG_ADD(2, t3, t4); DBG_ADD(3, t0, t0 + 1); }
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
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

// a.overflow_list / a.overflow_count: queries this kernel does not scan (more chunks than the chunk table) are appended for the
// exact fallback; a.qstate[q] = 1 marks them so that the finalize kernel skips their (unwritten) candidate lists.
template <int NG, int NCW, int NPW>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1) k_scan_pq16(ScanArgs a, P16Layout L, const uint16_t *__restrict__ table16,
                                                                   uint32_t *__restrict__ qstate) {
  constexpr int NCT = NCW * 32, NPT = NPW * 32, NP = NG / 2, NS = NG % 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *lut = smem;
  uint8_t *stage = smem + L.off_stage;
  uint32_t *ctab = (uint32_t *)(smem + L.off_ctab);  // [2][P16_CTAB_CAP]
  uint32_t *pref = (uint32_t *)(smem + L.off_pref);  // [2][maxp + 1]
  uint32_t *pcs = (uint32_t *)(smem + L.off_pcs);    // [2][maxp]
  uint32_t *plen = (uint32_t *)(smem + L.off_plen);  // [2][maxp]
  uint64_t *mkey = (uint64_t *)(smem + L.off_mkey);
  uint32_t *mpay = (uint32_t *)(smem + L.off_mpay);
  uint64_t *bars = (uint64_t *)(smem + L.off_misc);         // full[0..1], empty[0..1], stage
  uint32_t *qinfo = (uint32_t *)(smem + L.off_misc + 320);  // [2][4]: query, total chunks, nprobe
  uint32_t *thr_p = (uint32_t *)(smem + L.off_misc + 352);
  uint32_t *b2 = (uint32_t *)(smem + L.off_misc + 384);
  uint32_t *pq = (uint32_t *)(smem + L.off_misc + 512);
  uint32_t *scode = (uint32_t *)(smem + L.off_misc + 528);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t m = NG * 32;

  if (tid == 0) {
    mbar_init(&bars[0], NPT); mbar_init(&bars[1], NPT);
    mbar_init(&bars[2], NCW); mbar_init(&bars[3], NCW);
    mbar_init(&bars[4], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // =========================================== PRODUCER ===========================================================
    const int ptid = tid - NCT, pwarp = warp - NCW;
    uint32_t stage_phase = 0, nb = 0;
    bool have_buffer = false;
    auto fetch_query = [&]() -> uint32_t {   // dynamic query scheduling over the persistent CTAs, longest queries first (a.order)
      const uint32_t t = atomicAdd(a.next_query, 1u);
      return t < a.B ? (a.order ? a.order[t] : t) : 0xFFFFFFFFu;
    };
    if (ptid == 0) pq[0] = fetch_query();
    // The producer's per-query work is a chain of dependent global loads (query id -> codes -> probes -> list bounds) next to
    // the LUT gather; at the sharded shape (few chunks per query) it, not the scan, sets the pace.  So the chain is overlapped:
    // the NEXT query id is fetched one iteration ahead, the table rows are requested (bulk copies) as soon as the query's
    // codes are known, and the probe-list prefix + chunk table are computed while those copies are in flight.
#pragma unroll 1
    for (uint32_t ns = 0;; ns++) {   // ns: position in this CTA's query stream; nb: LUT buffers produced so far
      const uint32_t p = nb & 1;
      DBG_T(pt0);
      if (!have_buffer) { warp_mbar_wait(&bars[2 + p], ((nb >> 1) & 1) ^ 1, 200); have_buffer = true; }  // buffer p is free
      DBG_T(pt1);
      if (pwarp == 0) DBG_ADD(8, pt0, pt1);
      named_bar_sync(2, NPT);          // pq[ns & 1] is published; everybody is done with the previous staging tile
      const uint32_t q = pq[ns & 1];
      if (q >= a.B) {
        if (ptid == 0) qinfo[p * 4] = 0xFFFFFFFFu;
        mbar_arrive(&bars[p]);
        break;
      }
      for (uint32_t s = ptid; s < m; s += NPT) scode[s] = a.qcodes[(size_t)q * m + s];
      if (ptid == 0) p16_expect_tx(&bars[4], m * 512u);
      named_bar_sync(2, NPT);          // scode visible; expect_tx precedes every complete_tx
      // ---- LUT rows: one bulk copy per subspace (512 B of table16) into the padded staging tile
      if ((uint32_t)ptid < m) {
        p16_fence_proxy_async();
        p16_bulk_g2s(stage + ptid * P16_STAGE_PITCH, table16 + ((size_t)ptid * 256u + scode[ptid]) * 256u, 512u, &bars[4]);
      }
      if (ptid == 0) pq[(ns + 1) & 1] = fetch_query();
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
          const bool mine = run <= P16_CTAB_CAP;
          if (a.rows_scanned && mine) atomicAdd(a.rows_scanned, rows);
          if (!mine) { a.overflow_list[atomicAdd(a.overflow_count, 1u)] = q; qstate[q] = 1u; }  // left to the exact fallback
          else qstate[q] = 0u;
        }
      }
      named_bar_sync(2, NPT);
      const bool skip = prefp[np] > P16_CTAB_CAP;   // uniform over the producer warps
      // ---- chunk table ----------------------------------------------------------------------------------------------
      if (!skip) {
        uint32_t *ct = ctab + p * P16_CTAB_CAP;
        for (uint32_t i = pwarp; i < np; i += NPW) {
          const uint32_t first_it = prefp[i], cnt = prefp[i + 1] - first_it, cs = pcsp[i], len = plenp[i];
          for (uint32_t j = lane; j < cnt; j += 32) {
            ct[first_it + j] = ((cs + j) << 5) | (min(32u, len - 32u * j) - 1u);
            // every consumer warp starts a query with a cold load of its first chunk record: bring those records to L2 now,
            // one query ahead (the producers idle half of the time at the sharded shape)
            if (first_it + j < (uint32_t)NCW) {
              const char *rec = (const char *)a.codes + (size_t)(cs + j) * (NG * 1024u + 128u);
#pragma unroll 1
              for (uint32_t o = 0; o < NG * 1024u + 128u; o += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + o));
            }
          }
        }
      }
      DBG_T(pt2);
      if (pwarp == 0) DBG_ADD(9, pt1, pt2);
      warp_mbar_wait(&bars[4], stage_phase, 20);   // the rows have landed (the phase is consumed even for a skipped query)
      stage_phase ^= 1u;
      DBG_T(pt4);
      if (pwarp == 0) DBG_ADD(10, pt2, pt4);
      if (skip) continue;                          // buffer p stays ours: the next query of the stream takes it
      // ---- transpose the staging tile into the [code][column] image ------------------------------------------------------
#pragma unroll
      for (int pl = 0; pl < NP; pl++) {
        const uint32_t plane = p * NP + pl;
        uint8_t *dst = lut + (plane >> 1) * 65536u + (plane & 1) * 128u + lane * 4;
        const uint8_t *r0 = stage + (64 * pl + lane) * P16_STAGE_PITCH, *r1 = r0 + 32 * P16_STAGE_PITCH;
        // work item = (column = lane, piece j of 8 codes); the producer warps split the 32 pieces
        for (uint32_t j = pwarp; j < 32; j += NPW) {
          const uint4 x = *(const uint4 *)(r0 + j * 16);
          const uint4 y = *(const uint4 *)(r1 + j * 16);
          const uint32_t xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            *(uint32_t *)(dst + (j * 8 + 2 * c) * 256) = (xv[c] & 0xFFFFu) | (yv[c] << 16);
            *(uint32_t *)(dst + (j * 8 + 2 * c + 1) * 256) = (xv[c] >> 16) | (yv[c] & 0xFFFF0000u);
          }
        }
      }
      if (NS) {
        uint8_t *dst = lut + L.off_single + p * 128u + lane * 4;
        const uint8_t *r0 = stage + (64 * NP + lane) * P16_STAGE_PITCH;
        for (uint32_t j = pwarp; j < 32; j += NPW) {
          const uint4 x = *(const uint4 *)(r0 + j * 16);
          const uint32_t xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            *(uint16_t *)(dst + (j * 8 + 2 * c) * 256) = (uint16_t)(xv[c] & 0xFFFFu);
            *(uint16_t *)(dst + (j * 8 + 2 * c + 1) * 256) = (uint16_t)(xv[c] >> 16);
          }
        }
      }
      DBG_T(pt5);
      if (pwarp == 0) { DBG_ADD(11, pt4, pt5); DBG_ADD(12, pt0, pt0 + 1); }
      mbar_arrive(&bars[p]);  // buffer p is full (release: every producer thread's writes precede its arrive)
      nb++;
      have_buffer = false;
    }
    return;
  }

  // ============================================= CONSUMER ===========================================================
  // lane-dependent LUT column offsets, four per register: x_t = (lane ^ t) << 2 <= 124.  PRMT builds (code << 8) | x_t in ONE
  // instruction: byte 0 = x_t, byte 1 = code, bytes 2/3 = the sign replication of x_t (msb 0) = 0.
  uint32_t xr[8];
#pragma unroll
  for (int j = 0; j < 8; j++)
    xr[j] = (uint32_t)((lane ^ (4 * j)) << 2) | ((uint32_t)((lane ^ (4 * j + 1)) << 2) << 8) |
            ((uint32_t)((lane ^ (4 * j + 2)) << 2) << 16) | ((uint32_t)((lane ^ (4 * j + 3)) << 2) << 24);
  P16Consumer<NG, NCW> cc{a, L, lut, ctab, mkey, mpay, bars, qinfo, thr_p, b2, xr, lane, warp};
#pragma unroll 1
  for (uint32_t n2 = 0;; n2++) {
    // the two buffer parities are unrolled so that every LUT offset is an immediate of the LDS
    if (!cc.template query<0>(n2 & 1)) break;
    if (!cc.template query<1>(n2 & 1)) break;
  }
}

template <int NG, int NCW, int NPW>
static int launch_p16_t(mgpu_ivf *ivf, const ScanArgs &a0, uint32_t *d_qstate) {
  mgpu_ctx *ctx = ivf->ctx;
  ScanArgs a = a0;
  P16Layout L = p16_layout(NG, NCW, a.max_probes);
  if (L.total > ctx->smem_optin) return MGPU_ERR_UNSUPPORTED;
  CUDA_TRY(ctx, cudaFuncSetAttribute(k_scan_pq16<NG, NCW, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  unsigned grid = a.B < (uint32_t)ctx->sm_count ? a.B : (unsigned)ctx->sm_count;
  CUDA_TRY(ctx, cudaMemsetAsync(a.next_query, 0, 4, ctx->stream));
#ifdef MGPU_SCAN_DBG
  {
    static int nl = 0;
    if (++nl == 12) {
      unsigned long long h[16];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpyFromSymbol(h, g_dbg16, sizeof(h));
      fprintf(stderr, "[scan16 dbg] 11 launches, cycles summed over CTAs: consumer wait_full %llu scan %llu merge %llu (queries %llu) | producer "
                      "wait_empty %llu codes+prefix+ctab %llu stage_wait %llu transpose %llu (queries %llu)\n",
              h[0], h[1], h[2], h[3], h[8], h[9], h[10], h[11], h[12]);
    }
  }
#endif
  LaunchScope ls(ctx, MGPU_K_SCAN, nullptr, NG == 3 ? "k_scan_pq16<3,16,4> (scan_pq16.cu)" : (NG == 2 ? "k_scan_pq16<2,16,4> (scan_pq16.cu)" : "k_scan_pq16<1,16,4> (scan_pq16.cu)"));
  k_scan_pq16<NG, NCW, NPW><<<grid, (NCW + NPW) * 32, L.total, ctx->stream>>>(a, L, ivf->pq->d_table16, d_qstate);
  CUDA_TRY(ctx, cudaGetLastError());
  return MGPU_OK;
}

bool scan_pq16_applicable(mgpu_ivf *ivf, const ScanArgs &a) {
  static const bool off = getenv("MGPU_SCAN16") && getenv("MGPU_SCAN16")[0] == '0';
  if (off || !ivf->pq_fast || !ivf->pq || !ivf->pq->d_table16 || ivf->pq->gscale <= 0.0f) return false;
  if (ivf->metric != MGPU_L2 || a.lower_bound != nullptr || a.ng < 1 || a.ng > 3) return false;
  return p16_layout(a.ng, 16, a.max_probes).total <= ivf->ctx->smem_optin;
}

// a.overflow_count / a.overflow_list must be set (the deferred-query list shared with the finalize kernel)
int launch_scan_pq16(mgpu_ivf *ivf, const ScanArgs &a, uint32_t *d_qstate) {
  switch (a.ng) {
    case 1: return launch_p16_t<1, 16, 4>(ivf, a, d_qstate);
    case 2: return launch_p16_t<2, 16, 4>(ivf, a, d_qstate);
    case 3: return launch_p16_t<3, 16, 4>(ivf, a, d_qstate);
    default: return MGPU_ERR_UNSUPPORTED;
  }
}

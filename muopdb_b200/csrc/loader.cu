// loader.cu -- host-side readers of the reference's on-disk index formats, so that an
// index built by the unmodified reference can be made HBM-resident:
//   Elias-Fano posting lists   rs/compression/src/elias_fano/ef.rs:197-215 (layout), block_based_decoder.rs:162-266 (decode)
//   IVF `index` + `vectors`    rs/index/src/ivf/writer.rs:300-353, rs/index/src/ivf/block_based/storage.rs:52-151
//   HNSW `hnsw/index` + `hnsw/vector_storage`  rs/index/src/hnsw/writer.rs:206-265, hnsw/block_based/graph_storage.rs:122-193
//   PQ quantizer directory     rs/quantization/src/pq/mod.rs:18-19,101-136 (yaml config + raw f32 codebook)
//   vector files               u64 count, then row-major rows (rs/index/src/vector/file.rs:213-225)
// Pure byte/integer work on the host; the parsed arrays go through the same mgpu_*_create as caller-provided arrays.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "internal.cuh"

static bool read_file(const std::string &path, std::vector<uint8_t> &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? (size_t)sz : 0);
  size_t got = sz > 0 ? fread(out.data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  return got == out.size();
}

template <class T> static T rd(const uint8_t *p) { T v; memcpy(&v, p, sizeof(T)); return v; }  // little-endian host
static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) & ~(a - 1); }
// overflow-safe "section [off, off + len) lies inside a file of `size` bytes"
static inline bool fits(uint64_t off, uint64_t len, uint64_t size) { return off <= size && len <= size - off; }
// overflow-safe a * b <= limit
static inline bool mul_le(uint64_t a, uint64_t b, uint64_t limit) { return a == 0 || b <= limit / a; }

// Header fields of a (possibly truncated or corrupt) file size the allocations below: every loader body runs inside this
// guard so that a std::bad_alloc / length_error never crosses the extern "C" frame.
#define LOADER_GUARD_BEGIN try {
#define LOADER_GUARD_END(ctx, tag)                                                                       \
  } catch (const std::bad_alloc &) { return mgpu_fail(ctx, MGPU_ERR_OOM, tag ": out of host memory"); } \
  catch (const std::exception &e) { return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, tag ": %s", e.what()); }

extern "C" {

// Decode one Elias-Fano payload: u64 num_elem | u64 lower_bit_length | u64 lower_vec_len | u64 upper_vec_len |
// lower words | upper words; bits are Lsb0 inside little-endian u64 words (BitVec<u64, Lsb0>, ef.rs:13-14).
// value_i = (zeros before the i-th one in the upper bits) << L | low_i.  Returns the element count, or -1 on malformed input.
int64_t mgpu_ef_decode(const uint8_t *payload, uint64_t len, uint64_t *out, uint64_t cap) {
  if (!payload || len < 32) return -1;
  const uint64_t n = rd<uint64_t>(payload), L = rd<uint64_t>(payload + 8), lw = rd<uint64_t>(payload + 16), uw = rd<uint64_t>(payload + 24);
  if (L > 64 || lw > (len - 32) / 8 || uw > (len - 32) / 8 - lw) return -1;
  if (n > cap || (n && !out)) return -1;
  if (!mul_le(uw, 64, ~0ull) || n > uw * 64) return -1;            // every element owns a one bit of the upper vector
  if (n && L && (!mul_le(lw, 64, ~0ull) || n > (lw * 64) / L)) return -1;  // n * L low bits must exist (no u64 wrap)
  const uint8_t *lower = payload + 32, *upper = payload + 32 + lw * 8;
  uint64_t high = 0, i = 0;
  for (uint64_t w = 0; w < uw && i < n; w++) {
    uint64_t word = rd<uint64_t>(upper + w * 8);
    for (int b = 0; b < 64 && i < n; b++) {
      if ((word >> b) & 1ull) {
        uint64_t low = 0;
        if (L) {
          uint64_t bit = i * L, wi = bit >> 6, sh = bit & 63;
          low = rd<uint64_t>(lower + wi * 8) >> sh;
          if (sh + L > 64) low |= rd<uint64_t>(lower + (wi + 1) * 8) << (64 - sh);
          if (L < 64) low &= (1ull << L) - 1;
        }
        out[i++] = (L < 64 ? (high << L) : 0) | low;
      } else {
        high++;
      }
    }
  }
  return i == n ? (int64_t)n : -1;
}

// product_quantizer_config.yaml {dimension, subvector_dimension, num_bits} + raw little-endian f32 `codebook`
static int pq_load_impl(mgpu_ctx *ctx, const char *quantizer_dir, bool slice, uint64_t cb_off, uint64_t cb_len, int metric, mgpu_pq **out) {
  if (!ctx || !quantizer_dir || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  LOADER_GUARD_BEGIN
  std::string dir(quantizer_dir);
  std::vector<uint8_t> cfg, cb;
  if (!read_file(dir + "/product_quantizer_config.yaml", cfg)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Config file does not exist");  // pq/mod.rs:58-60
  if (!read_file(dir + "/codebook", cb)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Codebook file does not exist");
  long dim = -1, dsub = -1, nbits = -1;
  std::string text(cfg.begin(), cfg.end());
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    std::string line = text.substr(pos, eol - pos);
    pos = eol + 1;
    size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string key = line.substr(0, c);
    while (!key.empty() && (key.back() == ' ' || key.back() == '\t')) key.pop_back();
    while (!key.empty() && (key[0] == ' ' || key[0] == '\t')) key.erase(0, 1);
    long v = strtol(line.c_str() + c + 1, nullptr, 10);
    if (key == "dimension") dim = v;
    else if (key == "subvector_dimension") dsub = v;
    else if (key == "num_bits") nbits = v;
  }
  if (dim <= 0 || dsub <= 0 || nbits <= 0 || nbits > 8 || dim > (1l << 24)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_load: malformed product_quantizer_config.yaml");
  if (dim % dsub != 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Dimensions are not valid");
  size_t want = (size_t)dim * ((size_t)1 << nbits) * 4;
  if (slice) {  // one user's codebook inside the multi-user file (multi_spann/writer.rs: ivf_pq_codebook_offset/len)
    if (!fits(cb_off, cb_len, cb.size()) || cb_len != want)
      return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_load: codebook slice [%llu, +%llu) does not hold the %zu bytes of one codebook",
                       (unsigned long long)cb_off, (unsigned long long)cb_len, want);
    std::vector<float> part(want / 4);
    memcpy(part.data(), cb.data() + cb_off, want);
    return mgpu_pq_create(ctx, (uint32_t)dim, (uint32_t)dsub, (uint32_t)nbits, part.data(), metric, out);
  }
  if (cb.size() != want) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "pq_load: codebook has %zu bytes, expected %zu", cb.size(), want);
  return mgpu_pq_create(ctx, (uint32_t)dim, (uint32_t)dsub, (uint32_t)nbits, (const float *)cb.data(), metric, out);
  LOADER_GUARD_END(ctx, "pq_load")
}

int mgpu_pq_load(mgpu_ctx *ctx, const char *quantizer_dir, int metric, mgpu_pq **out) {
  return pq_load_impl(ctx, quantizer_dir, false, 0, 0, metric, out);
}

// ---- multi-user SPANN: the user -> byte-offset table and one user's index ----------------------------------------------------------
// UserIndexInfo::from_le_bytes (multi_spann/user_index_info.rs:59-83)
int mgpu_user_index_info_decode(const uint8_t bytes[112], mgpu_user_index_info *out) {
  if (!bytes || !out) return MGPU_ERR_INVALID_ARG;
  out->user_id.lo = rd<uint64_t>(bytes); out->user_id.hi = rd<uint64_t>(bytes + 8);
  uint64_t *f = &out->centroid_vector_offset;   // the 12 u64 fields follow in declaration order, as in the byte layout
  for (int i = 0; i < 12; i++) f[i] = rd<uint64_t>(bytes + 16 + 8 * i);
  return MGPU_OK;
}
// UserIndexInfo::to_le_bytes (user_index_info.rs:26-42)
int mgpu_user_index_info_encode(const mgpu_user_index_info *info, uint8_t out_bytes[112]) {
  if (!info || !out_bytes) return MGPU_ERR_INVALID_ARG;
  memcpy(out_bytes, &info->user_id.lo, 8); memcpy(out_bytes + 8, &info->user_id.hi, 8);
  const uint64_t *f = &info->centroid_vector_offset;
  for (int i = 0; i < 12; i++) memcpy(out_bytes + 16 + 8 * i, &f[i], 8);
  return MGPU_OK;
}

// odht 0.3.1 table image (crate absent from the reference checkout; Cargo.lock pins 0.3.1): 32-byte header
//   "ODHT" | size_of_metadata u8 | size_of_key u8 | size_of_value u8 | size_of_header u8 | item_count u64 | slot_count u64 |
//   file_format_version u32 | max_load_factor u16 | padding u16
// then slot_count entries (key bytes, value bytes), then slot_count + 16 control bytes; a control byte with the top bit set
// marks an empty slot (full slots store 7 hash bits).  The scan below needs no hash function.
int64_t mgpu_user_index_info_read(const char *path, mgpu_user_index_info *out, uint64_t cap) {
  if (!path) return -1;
  try {
    std::vector<uint8_t> f;
    if (!read_file(path, f) || f.size() < 32) return -1;
    if (memcmp(f.data(), "ODHT", 4) != 0) return -1;
    const uint32_t meta_sz = f[4], key_sz = f[5], val_sz = f[6], hdr_sz = f[7];
    const uint64_t items = rd<uint64_t>(f.data() + 8), slots = rd<uint64_t>(f.data() + 16);
    if (meta_sz != 1 || key_sz != 16 || val_sz != 112 || hdr_sz != 32) return -1;
    if (!mul_le(slots, 128, f.size()) || !fits(32, slots * 128, f.size()) || !fits(32 + slots * 128, slots, f.size()) || items > slots) return -1;
    const uint8_t *entries = f.data() + 32, *ctrl = f.data() + 32 + slots * 128;
    uint64_t n = 0;
    for (uint64_t s = 0; s < slots; s++) {
      if (ctrl[s] & 0x80u) continue;
      const uint8_t *e = entries + s * 128;
      if (memcmp(e, e + 16, 16) != 0) return -1;   // the key is the record's own user id
      if (n < cap && out) mgpu_user_index_info_decode(e + 16, &out[n]);
      n++;
    }
    return n == items ? (int64_t)n : -1;
  } catch (...) { return -1; }
}

int mgpu_spann_load_user(mgpu_ctx *ctx, const char *base_dir, const mgpu_user_index_info *info, uint32_t dim, int quant, int metric,
                         mgpu_pq **out_pq, mgpu_hnsw **out_centroids, mgpu_ivf **out_lists, mgpu_spann **out_spann) {
  if (!ctx || !base_dir || !info || !out_pq || !out_centroids || !out_lists || !out_spann) return MGPU_ERR_INVALID_ARG;
  *out_pq = nullptr; *out_centroids = nullptr; *out_lists = nullptr; *out_spann = nullptr;
  const std::string base(base_dir);
  mgpu_pq *pq = nullptr;
  mgpu_hnsw *hn = nullptr;
  mgpu_ivf *ivf = nullptr;
  mgpu_spann *sp = nullptr;
  int s = MGPU_OK;
  if (quant == MGPU_QUANT_PQ)
    s = pq_load_impl(ctx, (base + "/ivf/quantizer").c_str(), true, info->ivf_pq_codebook_offset, info->ivf_pq_codebook_len, metric, &pq);
  // the centroid index is an HNSW over the centroids with NoQuantizer<L2> (spann/reader.rs:56-64, multi_spann/writer.rs:124-129)
  if (s == MGPU_OK) s = mgpu_hnsw_load(ctx, (base + "/centroids").c_str(), info->centroid_index_offset, info->centroid_vector_offset, dim,
                                       MGPU_QUANT_NONE, MGPU_L2, nullptr, &hn);
  if (s == MGPU_OK) s = mgpu_ivf_load(ctx, (base + "/ivf").c_str(), info->ivf_index_offset, info->ivf_vectors_offset, quant, metric, pq, &ivf);
  if (s == MGPU_OK) s = mgpu_spann_create(ctx, hn, ivf, &sp);
  if (s != MGPU_OK) {
    if (ivf) mgpu_ivf_destroy(ivf);
    if (hn) mgpu_hnsw_destroy(hn);
    if (pq) mgpu_pq_destroy(pq);
    return s;
  }
  *out_pq = pq; *out_centroids = hn; *out_lists = ivf; *out_spann = sp;
  return MGPU_OK;
}

// IVF: {base}/index (from byte `index_offset`) + {base}/vectors (from byte `vector_offset`), as BlockBasedIvf::new_with_offset
// (ivf/block_based/index.rs:95-138).  quant/metric/pq describe the quantizer the index was written with.
int mgpu_ivf_load(mgpu_ctx *ctx, const char *base_dir, uint64_t index_offset, uint64_t vector_offset, int quant, int metric,
                  mgpu_pq *pq, mgpu_ivf **out) {
  if (!ctx || !base_dir || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  LOADER_GUARD_BEGIN
  std::string dir(base_dir);
  std::vector<uint8_t> f, vf;
  if (!read_file(dir + "/index", f)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Failed to open index file: %s/index", base_dir);
  if (!read_file(dir + "/vectors", vf)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Failed to open vectors file: %s/vectors", base_dir);
  if (!fits(index_offset, 48, f.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: index file too short");
  const uint8_t *h = f.data() + index_offset;
  if (h[0] != 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Unknown version: %d", (int)h[0]);  // storage.rs:101-104
  const uint32_t num_features = rd<uint32_t>(h + 1), qdim = rd<uint32_t>(h + 5), num_clusters = rd<uint32_t>(h + 9);
  const uint64_t num_vectors = rd<uint64_t>(h + 13), doc_len = rd<uint64_t>(h + 21), cent_len = rd<uint64_t>(h + 29);
  const uint64_t doc_off = index_offset + align_up(45, 16);                       // storage.rs:64-65,134
  const uint64_t cent_off = align_up(doc_off + doc_len, 8);                       // :67-70
  const uint64_t meta_off = align_up(cent_off + cent_len, 8);                     // :72-73
  if (!fits(doc_off, doc_len, f.size()) || cent_off < doc_off || !fits(cent_off, cent_len, f.size()) || meta_off < cent_off ||
      !fits(meta_off, 8, f.size()))
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: truncated index file");
  if (doc_len < 16 || !mul_le(num_vectors, 16, doc_len - 16) || cent_len < 8 || num_features == 0 ||
      !mul_le((uint64_t)num_clusters * num_features, 4, cent_len - 8))
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: section lengths inconsistent with the header");
  const mgpu_u128 *docs = (const mgpu_u128 *)(f.data() + doc_off + 16);           // skip the u128 count (storage.rs:165,203)
  std::vector<mgpu_u128> docs_aligned(num_vectors);
  memcpy(docs_aligned.data(), docs, num_vectors * 16);
  std::vector<float> cents((size_t)num_clusters * num_features);
  memcpy(cents.data(), f.data() + cent_off + 8, cents.size() * 4);                // skip the u64 count (storage.rs:260-262)
  const uint64_t npl = rd<uint64_t>(f.data() + meta_off);
  if (npl != num_clusters) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: %llu posting lists for %u clusters", (unsigned long long)npl, num_clusters);
  const uint64_t plm_off = meta_off + 8;                                            // storage.rs:78-80
  if (!mul_le(npl, 16, f.size()) || !fits(plm_off, npl * 16, f.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: truncated posting-list table");
  const uint64_t pl_start = plm_off + npl * 16;
  std::vector<uint64_t> offsets(npl + 1, 0);
  std::vector<uint32_t> ids;
  std::vector<uint64_t> tmp;
  for (uint64_t i = 0; i < npl; i++) {
    const uint64_t plen = rd<uint64_t>(f.data() + plm_off + i * 16), poff = rd<uint64_t>(f.data() + plm_off + i * 16 + 8);
    if (!fits(pl_start, poff, f.size()) || !fits(pl_start + poff, plen, f.size()) || plen < 32)
      return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: posting list %llu out of bounds", (unsigned long long)i);
    const uint8_t *pl = f.data() + pl_start + poff;
    const uint64_t n = rd<uint64_t>(pl);
    if (n > plen * 8) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: posting list %llu claims %llu ids in %llu bytes", (unsigned long long)i, (unsigned long long)n, (unsigned long long)plen);
    tmp.resize(n ? n : 1);
    if (mgpu_ef_decode(pl, plen, tmp.data(), n) != (int64_t)n) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: malformed Elias-Fano list %llu", (unsigned long long)i);
    for (uint64_t j = 0; j < n; j++) {
      if (tmp[j] >= num_vectors) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: list %llu holds point id %llu >= num_vectors", (unsigned long long)i, (unsigned long long)tmp[j]);
      ids.push_back((uint32_t)tmp[j]);                                             // point_id_u64 as u32 (index.rs:197)
    }
    offsets[i + 1] = offsets[i] + n;
  }
  // vectors: u64 count + rows of qdim elements
  if (!fits(vector_offset, 8, vf.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: vectors file too short");
  const uint64_t vcount = rd<uint64_t>(vf.data() + vector_offset);
  const size_t esz = quant == MGPU_QUANT_PQ ? 1 : 4;
  if (quant == MGPU_QUANT_PQ && (!pq || pq->m != qdim)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: quantized_dimension %u does not match the quantizer", qdim);
  if (quant == MGPU_QUANT_NONE && qdim != num_features) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: quantized_dimension %u != num_features %u for NoQuantizer", qdim, num_features);
  if (vcount < num_vectors || qdim == 0 || !mul_le(vcount, (uint64_t)qdim * esz, vf.size() - vector_offset - 8))
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "ivf_load: vectors file holds %llu rows, index expects %llu", (unsigned long long)vcount, (unsigned long long)num_vectors);
  std::vector<uint8_t> rows(num_vectors * qdim * esz + 16);
  memcpy(rows.data(), vf.data() + vector_offset + 8, num_vectors * qdim * esz);
  return mgpu_ivf_create(ctx, num_features, num_clusters, cents.data(), offsets.data(), ids.data(), quant, metric, pq, rows.data(),
                         MGPU_HOST, num_vectors, docs_aligned.data(), out);
  LOADER_GUARD_END(ctx, "ivf_load")
}

// HNSW: {base}/hnsw/index (from `index_offset`) + {base}/hnsw/vector_storage (from `vector_offset`), as
// BlockBasedHnsw::new_with_offsets (hnsw/block_based/index.rs:94-140).  `dim` is the original dimension (the header only
// stores the quantized one).  Files written before doc ids became u128 (8 bytes per id, e.g. the sample committed under
// rs/index_writer/test_output/hnsw) are accepted: their ids are widened.
int mgpu_hnsw_load(mgpu_ctx *ctx, const char *base_dir, uint64_t index_offset, uint64_t vector_offset, uint32_t dim, int quant,
                   int metric, mgpu_pq *pq, mgpu_hnsw **out) {
  if (!ctx || !base_dir || !out) return MGPU_ERR_INVALID_ARG;
  *out = nullptr;
  LOADER_GUARD_BEGIN
  std::string dir(base_dir);
  std::vector<uint8_t> f, vf;
  if (!read_file(dir + "/hnsw/index", f)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Failed to open %s/hnsw/index", base_dir);
  if (!read_file(dir + "/hnsw/vector_storage", vf)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Failed to open %s/hnsw/vector_storage", base_dir);
  if (!fits(index_offset, 49, f.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: index file too short");
  const uint8_t *h = f.data() + index_offset;
  if (h[0] != 0) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "Unknown version: %d", (int)h[0]);
  const uint32_t qdim = rd<uint32_t>(h + 1), num_layers = rd<uint32_t>(h + 5);
  const uint64_t edges_len = rd<uint64_t>(h + 9), points_len = rd<uint64_t>(h + 17), eo_len = rd<uint64_t>(h + 25),
                 lo_len = rd<uint64_t>(h + 33), doc_len = rd<uint64_t>(h + 41);
  // calculate_offsets (graph_storage.rs:168-193)
  uint64_t off = index_offset + 49;
  const uint64_t edges_off = off + (4 - off % 4) % 4;
  const uint64_t points_off = edges_off + edges_len;
  uint64_t t = points_off + points_len;
  const uint64_t eo_off = t + (8 - t % 8) % 8;
  const uint64_t lo_off = eo_off + eo_len;
  t = lo_off + lo_len;
  uint64_t doc_off = t + (16 - t % 16) % 16;
  // every section inside the file, in order, without u64 wrap-around
  if (!fits(edges_off, edges_len, f.size()) || !fits(points_off, points_len, f.size()) || eo_off < points_off ||
      !fits(eo_off, eo_len, f.size()) || !fits(lo_off, lo_len, f.size()) || (edges_len | points_len) % 4 != 0 || eo_len % 8 != 0)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: truncated or corrupt index file (a section exceeds the file)");
  if (num_layers == 0 || lo_len != ((uint64_t)num_layers + 1) * 8)
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: header inconsistent (layers %u, level_offsets_len %llu)", num_layers, (unsigned long long)lo_len);
  std::vector<uint32_t> edges(edges_len / 4 + 1), points(points_len / 4 + 1);
  std::vector<uint64_t> eo(eo_len / 8), lo(lo_len / 8);
  memcpy(edges.data(), f.data() + edges_off, edges_len);
  memcpy(points.data(), f.data() + points_off, points_len);
  memcpy(eo.data(), f.data() + eo_off, eo_len);
  memcpy(lo.data(), f.data() + lo_off, lo_len);
  // layer 0 is addressed by point id and edge_offsets ends with one terminal entry (hnsw/writer.rs:122-140); the writer's
  // last level offset counts that terminal entry too, so the point count comes from edge_offsets, not from level_offsets
  if (eo.empty() || lo[num_layers - 1] > eo.size() - 1) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: edge_offsets shorter than the upper layers");
  const uint64_t n = eo.size() - 1 - lo[num_layers - 1];
  std::vector<mgpu_u128> docs(n ? n : 1);
  if (doc_len == n * 16 && fits(doc_off, doc_len, f.size())) memcpy(docs.data(), f.data() + doc_off, n * 16);
  else if (doc_len == n * 8) {  // legacy 8-byte ids, written right after the level offsets (8-byte aligned)
    doc_off = t + (8 - t % 8) % 8;
    if (!fits(doc_off, doc_len, f.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: truncated legacy doc-id section");
    for (uint64_t i = 0; i < n; i++) { docs[i].lo = rd<uint64_t>(f.data() + doc_off + i * 8); docs[i].hi = 0; }
  } else return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: doc_id_mapping_len %llu does not match %llu points", (unsigned long long)doc_len, (unsigned long long)n);
  if (!fits(vector_offset, 8, vf.size())) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: vector_storage too short");
  const uint64_t vcount = rd<uint64_t>(vf.data() + vector_offset);
  const size_t esz = quant == MGPU_QUANT_PQ ? 1 : 4;
  if (quant == MGPU_QUANT_PQ && (!pq || pq->m != qdim || pq->dim != dim)) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: quantized_dimension %u does not match the quantizer", qdim);
  if (quant == MGPU_QUANT_NONE && qdim != dim) return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: quantized_dimension %u != dimension %u", qdim, dim);
  if (vcount < n || qdim == 0 || !mul_le(vcount, (uint64_t)qdim * esz, vf.size() - vector_offset - 8))
    return mgpu_fail(ctx, MGPU_ERR_INVALID_ARG, "hnsw_load: vector_storage holds %llu rows, graph has %llu points", (unsigned long long)vcount, (unsigned long long)n);
  std::vector<uint8_t> rows(n * qdim * esz + 16);
  memcpy(rows.data(), vf.data() + vector_offset + 8, n * qdim * esz);
  return mgpu_hnsw_create(ctx, dim, num_layers, edges.data(), edges_len / 4, points.data(), points_len / 4, eo.data(), eo.size(),
                          lo.data(), quant, metric, pq, rows.data(), MGPU_HOST, n, docs.data(), out);
  LOADER_GUARD_END(ctx, "hnsw_load")
}

// Graph sections of a loaded index, for inspection/tests: sizes = {num_layers, n_edges, n_points, n_edge_offsets, n, entry_point}
int mgpu_hnsw_info(mgpu_hnsw *h, uint64_t sizes[6]) {
  if (!h || !sizes) return MGPU_ERR_INVALID_ARG;
  sizes[0] = h->num_layers; sizes[1] = h->n_edges; sizes[2] = h->n_points; sizes[3] = h->n_edge_offsets; sizes[4] = h->n; sizes[5] = h->entry_point;
  return MGPU_OK;
}
int mgpu_hnsw_copy_graph(mgpu_hnsw *h, uint32_t *edges, uint32_t *points, uint64_t *edge_offsets, uint64_t *level_offsets) {
  if (!h) return MGPU_ERR_INVALID_ARG;
  mgpu_ctx *ctx = h->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if (edges) CUDA_TRY(ctx, cudaMemcpyAsync(edges, h->d_edges, h->n_edges * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (points) CUDA_TRY(ctx, cudaMemcpyAsync(points, h->d_points, h->n_points * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (edge_offsets) CUDA_TRY(ctx, cudaMemcpyAsync(edge_offsets, h->d_edge_offsets, h->n_edge_offsets * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (level_offsets) CUDA_TRY(ctx, cudaMemcpyAsync(level_offsets, h->d_level_offsets, (h->num_layers + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return MGPU_OK;
}

}  // extern "C"

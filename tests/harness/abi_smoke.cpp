// abi_smoke.cpp -- calls libmuopdb_gpu.so through its C ABI from plain C++ (no Python, no torch, no CUDA headers): what a
// non-Python host (the Rust shim of ffi/muopdb_gpu.rs) does.  Build: g++ -std=c++17 -I include tests/harness/abi_smoke.cpp -ldl
// Usage: abi_smoke <path to libmuopdb_gpu.so>
//   exit 0 + "NO_DEVICE"  : the library loads, every symbol resolves, mgpu_init reports MGPU_ERR_NO_DEVICE (no CPU fallback)
//   exit 0 + "OK ..."     : on a GPU box, the reference's own golden case (spann/index.rs:335-366: 1000 vectors [i,i,i,i],
//                           query [2.4,3.4,4.4,5.4], k = 2 -> doc ids [4, 3]) through mgpu_ivf_search, the flat scores compared
//                           bit for bit with L2DistanceCalculator::calculate restated inline (l2.rs:30-74, 4-lane phase), and
//                           the doc-id accessors / invalidation (after invalidating doc 4 -> [3, 5], spann/index.rs:412-444)
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "muopdb_gpu.h"

#define SYM(name) auto p_##name = (decltype(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static float ref_l2_dim4(const float *a, const float *b) {   // l2.rs:30-74 for n = 4: one 4-lane chunk, ordered reduce, sqrt
  float acc[4];
  for (int l = 0; l < 4; l++) { float d = a[l] - b[l]; acc[l] = 0.0f + d * d; }
  float s = -0.0f;
  for (int l = 0; l < 4; l++) s = s + acc[l];
  return sqrtf(0.0f + s);
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: abi_smoke <libmuopdb_gpu.so>\n"); return 2; }
  void *lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 2; }
  SYM(mgpu_version) SYM(mgpu_init) SYM(mgpu_destroy) SYM(mgpu_last_error) SYM(mgpu_ivf_create) SYM(mgpu_ivf_destroy)
  SYM(mgpu_ivf_search) SYM(mgpu_ivf_get_doc_ids) SYM(mgpu_ivf_get_point_id) SYM(mgpu_ivf_invalidate_docs) SYM(mgpu_ivf_get_vectors)
  SYM(mgpu_distance_batch) SYM(mgpu_user_index_info_decode) SYM(mgpu_user_index_info_encode) SYM(mgpu_ef_decode)
  if (!strstr(p_mgpu_version(), "sm_100a")) { fprintf(stderr, "unexpected version %s\n", p_mgpu_version()); return 2; }
  // host-only entry points work without a device: UserIndexInfo round trip (user_index_info.rs:143-203)
  mgpu_user_index_info info{}, back{};
  info.user_id.lo = 1234567890; info.centroid_vector_offset = 100; info.ivf_pq_codebook_len = 1200;
  uint8_t bytes[112];
  if (p_mgpu_user_index_info_encode(&info, bytes) != MGPU_OK || p_mgpu_user_index_info_decode(bytes, &back) != MGPU_OK ||
      memcmp(&info, &back, sizeof(info)) != 0) { fprintf(stderr, "user_index_info round trip failed\n"); return 1; }
  mgpu_ctx *ctx = nullptr;
  int st = p_mgpu_init(0, &ctx);
  if (st == MGPU_ERR_NO_DEVICE) { printf("NO_DEVICE\n"); return 0; }
  if (st != MGPU_OK) { fprintf(stderr, "mgpu_init failed: %d\n", st); return 1; }
  const uint32_t n = 1000, dim = 4, nlist = 4;
  std::vector<float> X(n * dim), cents(nlist * dim);
  for (uint32_t i = 0; i < n; i++) for (uint32_t d = 0; d < dim; d++) X[i * dim + d] = (float)i;
  for (uint32_t c = 0; c < nlist; c++) for (uint32_t d = 0; d < dim; d++) cents[c * dim + d] = 125.0f + 250.0f * c;
  std::vector<uint64_t> offs(nlist + 1);
  std::vector<uint32_t> ids;
  for (uint32_t c = 0; c < nlist; c++) { offs[c] = ids.size(); for (uint32_t i = c * 250; i < (c + 1) * 250; i++) ids.push_back(i); }
  offs[nlist] = ids.size();
  std::vector<mgpu_u128> docs(n);
  for (uint32_t i = 0; i < n; i++) docs[i] = mgpu_u128{i, 0};
  mgpu_ivf *ivf = nullptr;
  st = p_mgpu_ivf_create(ctx, dim, nlist, cents.data(), offs.data(), ids.data(), MGPU_QUANT_NONE, MGPU_L2, nullptr, X.data(), MGPU_HOST, n,
                         docs.data(), &ivf);
  if (st != MGPU_OK) { fprintf(stderr, "ivf_create: %s\n", p_mgpu_last_error(ctx)); return 1; }
  const float q[4] = {2.4f, 3.4f, 4.4f, 5.4f};
  mgpu_u128 out_docs[2]; float out_scores[2]; uint32_t cnt = 0;
  st = p_mgpu_ivf_search(ivf, q, 1, 2, nlist, out_docs, out_scores, &cnt, MGPU_HOST);
  if (st != MGPU_OK) { fprintf(stderr, "ivf_search: %s\n", p_mgpu_last_error(ctx)); return 1; }
  const float want4 = ref_l2_dim4(q, &X[4 * dim]), want3 = ref_l2_dim4(q, &X[3 * dim]);
  if (cnt != 2 || out_docs[0].lo != 4 || out_docs[1].lo != 3 || memcmp(&out_scores[0], &want4, 4) || memcmp(&out_scores[1], &want3, 4)) {
    fprintf(stderr, "golden mismatch: cnt %u ids %llu %llu scores %.9g %.9g (want 4 3 %.9g %.9g)\n", cnt, (unsigned long long)out_docs[0].lo,
            (unsigned long long)out_docs[1].lo, out_scores[0], out_scores[1], want4, want3);
    return 1;
  }
  // accessors + invalidation by doc id
  uint32_t pids[2] = {4, 999}; mgpu_u128 got[2]; float row[4]; int found = 0; uint32_t pid = 0, nok = 0; uint8_t ok[2];
  mgpu_u128 d4{4, 0}, dmiss{123456, 0}, inv[2] = {d4, dmiss};
  if (p_mgpu_ivf_get_doc_ids(ivf, pids, 2, got) != MGPU_OK || got[0].lo != 4 || got[1].lo != 999) { fprintf(stderr, "get_doc_ids\n"); return 1; }
  if (p_mgpu_ivf_get_point_id(ivf, &d4, &found, &pid) != MGPU_OK || !found || pid != 4) { fprintf(stderr, "get_point_id\n"); return 1; }
  if (p_mgpu_ivf_get_point_id(ivf, &dmiss, &found, &pid) != MGPU_OK || found) { fprintf(stderr, "get_point_id(miss)\n"); return 1; }
  if (p_mgpu_ivf_get_vectors(ivf, pids, 1, row) != MGPU_OK || row[0] != 4.0f || row[3] != 4.0f) { fprintf(stderr, "get_vectors\n"); return 1; }
  if (p_mgpu_ivf_invalidate_docs(ivf, inv, 2, ok, &nok) != MGPU_OK || nok != 1 || ok[0] != 1 || ok[1] != 0) { fprintf(stderr, "invalidate_docs\n"); return 1; }
  st = p_mgpu_ivf_search(ivf, q, 1, 2, nlist, out_docs, out_scores, &cnt, MGPU_HOST);
  if (st != MGPU_OK || cnt != 2 || out_docs[0].lo != 3 || out_docs[1].lo != 5) { fprintf(stderr, "after invalidation: %llu %llu\n", (unsigned long long)out_docs[0].lo, (unsigned long long)out_docs[1].lo); return 1; }
  p_mgpu_ivf_destroy(ivf);
  p_mgpu_destroy(ctx);
  printf("OK golden [4,3] -> [3,5], scores bit-exact\n");
  return 0;
}

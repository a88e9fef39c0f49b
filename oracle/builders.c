/*
 * builders.c -- seeded generators of synthetic INDEX INPUTS for the oracle and the CUDA path.
 *
 * TEST INFRASTRUCTURE (part of liboracle.so; see oracle.c header for who may use it).
 *
 * The reference builds its indices with unseeded RNGs (rand::thread_rng in
 * rs/index/src/hnsw/builder.rs:332-337, rs/index/src/ivf/builder.rs:409,476,
 * rs/utils/src/kmeans_builder/kmeans_builder.rs:151), so built artefacts are not reproducible
 * and are INPUTS to parity, not outputs.  These generators follow the same
 * construction rules with a seeded RNG so tests and benchmarks get deterministic, structurally
 * valid indices in the reference's array layout:
 *   - orc_hnsw_build: HnswBuilder::insert / select_neighbors_heuristic / get_random_layer
 *     (hnsw/builder.rs:221-375) over NoQuantizer<L2>, emitted as the graph arrays that
 *     hnsw/writer.rs:43-265 serialises (edges, points, edge_offsets, level_offsets).
 *   - orc_kmeans: plain Lloyd iterations with first-minimum assignment
 *     (kmeans_builder.rs:199-221 assignment rule, no balance penalty).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

float orc_l2(const float *a, const float *b, uint64_t n);
float orc_l2_squared(const float *a, const float *b, uint64_t n);

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { /* splitmix64 */
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline float rng_f32(rng_t *r) { return (float)((rng_next(r) >> 40) + 1) * (1.0f / 16777217.0f); } /* (0,1) */

/* ---------------------------------------------------------------- k-means */
ORC_API void orc_kmeans(const float *X, uint64_t n, uint32_t dim, uint32_t k, uint32_t iters, uint64_t seed,
                        float *centroids /* k x dim out */) {
  rng_t r = {seed};
  /* init: k distinct random rows (partial Fisher-Yates over indices) */
  uint64_t *perm = (uint64_t *)malloc(sizeof(uint64_t) * n);
  for (uint64_t i = 0; i < n; i++) perm[i] = i;
  for (uint32_t c = 0; c < k; c++) {
    uint64_t j = c + rng_next(&r) % (n - c > 0 ? n - c : 1);
    if (j >= n) j = n - 1;
    uint64_t t = perm[c % n]; perm[c % n] = perm[j]; perm[j] = t;
    memcpy(centroids + (size_t)c * dim, X + (size_t)perm[c % n] * dim, sizeof(float) * dim);
  }
  free(perm);
  uint32_t *assign = (uint32_t *)malloc(sizeof(uint32_t) * n);
  double *sums = (double *)malloc(sizeof(double) * (size_t)k * dim);
  uint64_t *cnt = (uint64_t *)malloc(sizeof(uint64_t) * k);
  for (uint32_t it = 0; it < iters; it++) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      float best = 3.40282347e+38f; uint32_t bi = 0;
      for (uint32_t c = 0; c < k; c++) {
        float d = orc_l2_squared(X + (size_t)i * dim, centroids + (size_t)c * dim, dim);
        if (d < best) { best = d; bi = c; }
      }
      assign[i] = bi;
    }
    memset(sums, 0, sizeof(double) * (size_t)k * dim);
    memset(cnt, 0, sizeof(uint64_t) * k);
    for (uint64_t i = 0; i < n; i++) {
      cnt[assign[i]]++;
      double *s = sums + (size_t)assign[i] * dim;
      const float *x = X + (size_t)i * dim;
      for (uint32_t d = 0; d < dim; d++) s[d] += x[d];
    }
    for (uint32_t c = 0; c < k; c++)
      if (cnt[c])
        for (uint32_t d = 0; d < dim; d++) centroids[(size_t)c * dim + d] = (float)(sums[(size_t)c * dim + d] / (double)cnt[c]);
  }
  free(assign); free(sums); free(cnt);
}

/* ------------------------------------------------------------------- HNSW */
typedef struct { uint32_t id; float d; } nb_t;

typedef struct {
  uint32_t n, dim, M, max_layer, efc;
  const float *X;
  uint8_t *level;        /* per node */
  int32_t **slot;        /* slot[l][node] = index into adj[l] (l>=1), -1 if absent; layer 0 uses node id */
  nb_t **adj;            /* adj[l][slot*(M+1) + j] */
  uint32_t **deg;        /* deg[l][slot] */
  uint32_t *nslots;      /* nodes present in layer l */
  uint32_t *caps;
  uint32_t top; uint32_t entry;
  /* scratch */
  uint32_t *visit_tag; uint32_t tag;
} hb_t;

static inline uint32_t hb_slot(hb_t *h, uint32_t l, uint32_t node) { return l == 0 ? node : (uint32_t)h->slot[l][node]; }

static void hb_add_node(hb_t *h, uint32_t l, uint32_t node) {
  if (l == 0) return;
  if (h->nslots[l] == h->caps[l]) {
    h->caps[l] = h->caps[l] ? h->caps[l] * 2 : 1024;
    h->adj[l] = (nb_t *)realloc(h->adj[l], sizeof(nb_t) * (size_t)h->caps[l] * (h->M + 1));
    h->deg[l] = (uint32_t *)realloc(h->deg[l], sizeof(uint32_t) * h->caps[l]);
  }
  h->slot[l][node] = (int32_t)h->nslots[l];
  h->deg[l][h->nslots[l]] = 0;
  h->nslots[l]++;
}

/* simple binary heaps on nb_t keyed by (d, id); max-heap with sign trick handled by caller */
typedef struct { nb_t *v; uint32_t n, cap; } nheap;
static inline int nb_less(nb_t a, nb_t b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }
static void nh_push(nheap *h, nb_t x) { /* max-heap by nb_less */
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 256; h->v = (nb_t *)realloc(h->v, sizeof(nb_t) * h->cap); }
  uint32_t i = h->n++; h->v[i] = x;
  while (i) { uint32_t p = (i - 1) / 2; if (!nb_less(h->v[p], h->v[i])) break; nb_t t = h->v[p]; h->v[p] = h->v[i]; h->v[i] = t; i = p; }
}
static nb_t nh_pop(nheap *h) {
  nb_t top = h->v[0]; h->n--;
  if (h->n) {
    h->v[0] = h->v[h->n];
    uint32_t i = 0;
    for (;;) {
      uint32_t l = 2 * i + 1, r = l + 1, m = i;
      if (l < h->n && nb_less(h->v[m], h->v[l])) m = l;
      if (r < h->n && nb_less(h->v[m], h->v[r])) m = r;
      if (m == i) break;
      nb_t t = h->v[m]; h->v[m] = h->v[i]; h->v[i] = t; i = m;
    }
  }
  return top;
}

static int nb_cmp_q(const void *a, const void *b) {
  nb_t x = *(const nb_t *)a, y = *(const nb_t *)b;
  return nb_less(x, y) ? -1 : (nb_less(y, x) ? 1 : 0);
}

/* GraphTraversal::search_layer (hnsw/utils.rs:58-129): same beam rule as the query path */
static uint32_t hb_search_layer(hb_t *h, const float *q, uint32_t ep, uint32_t ef, uint32_t l, nb_t **out) {
  nheap cand = {0}, work = {0};
  h->visit_tag[ep] = h->tag;
  float ed = orc_l2(q, h->X + (size_t)ep * h->dim, h->dim);
  nb_t c0 = {ep, -ed}, w0 = {ep, ed};
  nh_push(&cand, c0); nh_push(&work, w0);
  while (cand.n) {
    nb_t c = nh_pop(&cand);
    float dist = -c.d;
    if (work.n == 0) continue;
    if (dist > work.v[0].d) break;
    uint32_t s = hb_slot(h, l, c.id);
    uint32_t deg = h->deg[l][s];
    const nb_t *e = h->adj[l] + (size_t)s * (h->M + 1);
    for (uint32_t i = 0; i < deg; i++) {
      uint32_t id = e[i].id;
      if (h->visit_tag[id] == h->tag) continue;
      h->visit_tag[id] = h->tag;
      float furthest = work.v[0].d;
      float de = orc_l2(q, h->X + (size_t)id * h->dim, h->dim);
      if (de < furthest || work.n < ef) {
        nb_t a = {id, -de}, b = {id, de};
        nh_push(&cand, a); nh_push(&work, b);
        if (work.n > ef) nh_pop(&work);
      }
    }
  }
  qsort(work.v, work.n, sizeof(nb_t), nb_cmp_q);
  *out = work.v;
  free(cand.v);
  return work.n;
}

/* select_neighbors_heuristic (hnsw/builder.rs:339-375) */
static uint32_t hb_select(hb_t *h, const nb_t *cands, uint32_t nc, uint32_t want, nb_t *out) {
  nb_t *s = (nb_t *)malloc(sizeof(nb_t) * (nc ? nc : 1));
  memcpy(s, cands, sizeof(nb_t) * nc);
  /* min-heap pop order on (-d, id) max-heap == ascending d, ties by larger id first */
  for (uint32_t i = 0; i < nc; i++) s[i].d = -s[i].d;
  nheap hp = {0};
  for (uint32_t i = 0; i < nc; i++) nh_push(&hp, s[i]);
  uint32_t no = 0;
  while (hp.n && no < want) {
    nb_t e = nh_pop(&hp);
    float deq = -e.d;
    int good = 1;
    for (uint32_t j = 0; j < no; j++) {
      float dxe = orc_l2(h->X + (size_t)e.id * h->dim, h->X + (size_t)out[j].id * h->dim, h->dim);
      if (dxe < deq) { good = 0; break; }
    }
    if (good) { out[no].id = e.id; out[no].d = deq; no++; }
  }
  free(hp.v); free(s);
  return no;
}

static void hb_push_edge(hb_t *h, uint32_t l, uint32_t from, nb_t e) {
  uint32_t s = hb_slot(h, l, from);
  nb_t *a = h->adj[l] + (size_t)s * (h->M + 1);
  uint32_t d = h->deg[l][s];
  if (d <= h->M) { a[d] = e; h->deg[l][s] = d + 1; }
}

typedef struct {
  uint32_t num_layers; uint64_t n_edges, n_points, n_edge_offsets;
  uint32_t *edges, *points; uint64_t *edge_offsets, *level_offsets;
} orc_graph;

ORC_API orc_graph *orc_hnsw_build(const float *X, uint32_t n, uint32_t dim, uint32_t M, uint32_t max_layer,
                                  uint32_t ef_construction, uint64_t seed) {
  hb_t h; memset(&h, 0, sizeof(h));
  h.n = n; h.dim = dim; h.M = M; h.max_layer = max_layer; h.efc = ef_construction; h.X = X;
  uint32_t NL = max_layer + 1;
  h.level = (uint8_t *)calloc(n, 1);
  h.slot = (int32_t **)calloc(NL, sizeof(int32_t *));
  h.adj = (nb_t **)calloc(NL, sizeof(nb_t *));
  h.deg = (uint32_t **)calloc(NL, sizeof(uint32_t *));
  h.nslots = (uint32_t *)calloc(NL, sizeof(uint32_t));
  h.caps = (uint32_t *)calloc(NL, sizeof(uint32_t));
  h.adj[0] = (nb_t *)malloc(sizeof(nb_t) * (size_t)n * (M + 1));
  h.deg[0] = (uint32_t *)calloc(n, sizeof(uint32_t));
  for (uint32_t l = 1; l < NL; l++) { h.slot[l] = (int32_t *)malloc(sizeof(int32_t) * n); memset(h.slot[l], 0xFF, sizeof(int32_t) * n); }
  h.visit_tag = (uint32_t *)calloc(n, sizeof(uint32_t));
  rng_t r = {seed};
  nb_t *sel = (nb_t *)malloc(sizeof(nb_t) * (M + 2));
  nb_t *sel2 = (nb_t *)malloc(sizeof(nb_t) * (M + 2));
  for (uint32_t p = 0; p < n; p++) {
    const float *q = X + (size_t)p * dim;
    /* get_random_layer (builder.rs:332-337) */
    float rnd = rng_f32(&r);
    uint32_t layer = (uint32_t)floorf(-logf(rnd) / logf((float)M));
    if (layer > max_layer) layer = max_layer;
    h.level[p] = (uint8_t)layer;
    h.tag++;
    if (p == 0) {
      for (uint32_t l = 1; l <= layer; l++) hb_add_node(&h, l, p);
      h.top = layer; h.entry = p;
      continue;
    }
    uint32_t ep = h.entry;
    if (layer < h.top) {
      for (uint32_t l = h.top; l > layer; l--) {
        nb_t *ne; uint32_t nn = hb_search_layer(&h, q, ep, 1, l, &ne);
        if (nn) ep = ne[0].id;
        free(ne);
      }
    }
    for (uint32_t l = 1; l <= layer; l++) hb_add_node(&h, l, p);
    uint32_t lstart = layer < h.top ? layer : h.top;
    for (int32_t l = (int32_t)lstart; l >= 0; l--) {
      nb_t *ne; uint32_t nn = hb_search_layer(&h, q, ep, h.efc, (uint32_t)l, &ne);
      uint32_t ns = hb_select(&h, ne, nn, M, sel);
      for (uint32_t i = 0; i < ns; i++) {
        nb_t back = {p, sel[i].d};
        hb_push_edge(&h, (uint32_t)l, sel[i].id, back);
        hb_push_edge(&h, (uint32_t)l, p, sel[i]);
      }
      for (uint32_t i = 0; i < ns; i++) {
        uint32_t s = hb_slot(&h, (uint32_t)l, sel[i].id);
        if (h.deg[l][s] > M) {
          nb_t *a = h.adj[l] + (size_t)s * (M + 1);
          uint32_t k2 = hb_select(&h, a, h.deg[l][s], M, sel2);
          memcpy(a, sel2, sizeof(nb_t) * k2);
          h.deg[l][s] = k2;
        }
      }
      if (nn) ep = ne[0].id;
      free(ne);
    }
    if (layer > h.top) { h.top = layer; h.entry = p; }
  }
  /* emit arrays: layers top first; upper-layer point lists ascending by id with the entry point FIRST
   * in the top layer (get_entry_point_top_layer reads points[level_offsets[0]], graph_storage.rs:549-553) */
  orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
  uint32_t L = h.top + 1;
  g->num_layers = L;
  g->level_offsets = (uint64_t *)calloc(L + 1, sizeof(uint64_t));
  uint64_t tot_upper = 0, tot_edges = 0;
  for (uint32_t l = L - 1; l >= 1; l--) tot_upper += h.nslots[l];
  for (uint32_t l = 0; l < L; l++) {
    uint32_t cnt = l == 0 ? n : h.nslots[l];
    for (uint32_t s = 0; s < cnt; s++) tot_edges += h.deg[l][s];
  }
  g->n_points = tot_upper;
  g->points = (uint32_t *)malloc(sizeof(uint32_t) * (tot_upper ? tot_upper : 1));
  g->n_edge_offsets = tot_upper + n + 1;
  g->edge_offsets = (uint64_t *)malloc(sizeof(uint64_t) * g->n_edge_offsets);
  g->n_edges = tot_edges;
  g->edges = (uint32_t *)malloc(sizeof(uint32_t) * (tot_edges ? tot_edges : 1));
  uint64_t pi = 0, ei = 0;
  for (uint32_t li = 0; li < L; li++) {
    uint32_t l = L - 1 - li;
    g->level_offsets[li] = pi;
    if (l == 0) {
      for (uint32_t node = 0; node < n; node++) {
        g->edge_offsets[pi + node] = ei;
        const nb_t *a = h.adj[0] + (size_t)node * (M + 1);
        for (uint32_t j = 0; j < h.deg[0][node]; j++) g->edges[ei++] = a[j].id;
      }
      pi += n;
    } else {
      /* nodes of this layer ascending by id, entry point first on the top layer */
      uint64_t start = pi;
      if (l == L - 1) g->points[pi++] = h.entry;
      for (uint32_t node = 0; node < n; node++)
        if (h.level[node] >= l && !(l == L - 1 && node == h.entry)) g->points[pi++] = node;
      for (uint64_t x = start; x < pi; x++) {
        uint32_t node = g->points[x];
        g->edge_offsets[x] = ei;
        uint32_t s = hb_slot(&h, l, node);
        const nb_t *a = h.adj[l] + (size_t)s * (M + 1);
        for (uint32_t j = 0; j < h.deg[l][s]; j++) g->edges[ei++] = a[j].id;
      }
    }
  }
  g->level_offsets[L] = pi;
  g->edge_offsets[pi] = ei;
  for (uint32_t l = 0; l < NL; l++) { free(h.slot[l]); free(h.adj[l]); free(h.deg[l]); }
  free(h.slot); free(h.adj); free(h.deg); free(h.nslots); free(h.caps); free(h.level); free(h.visit_tag);
  free(sel); free(sel2);
  return g;
}

ORC_API void orc_graph_sizes(const orc_graph *g, uint64_t *out /* num_layers, n_edges, n_points, n_edge_offsets */) {
  out[0] = g->num_layers; out[1] = g->n_edges; out[2] = g->n_points; out[3] = g->n_edge_offsets;
}
ORC_API void orc_graph_copy(const orc_graph *g, uint32_t *edges, uint32_t *points, uint64_t *edge_offsets,
                            uint64_t *level_offsets) {
  memcpy(edges, g->edges, sizeof(uint32_t) * g->n_edges);
  memcpy(points, g->points, sizeof(uint32_t) * g->n_points);
  memcpy(edge_offsets, g->edge_offsets, sizeof(uint64_t) * g->n_edge_offsets);
  memcpy(level_offsets, g->level_offsets, sizeof(uint64_t) * (g->num_layers + 1));
}
ORC_API void orc_graph_free(orc_graph *g) {
  if (!g) return;
  free(g->edges); free(g->points); free(g->edge_offsets); free(g->level_offsets); free(g);
}

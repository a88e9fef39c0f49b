"""Pins the CPU oracle against the reference's own known-answer tests.

Each test restates a `#[test]` of the reference (file:line cited) on the oracle.  Where the reference test
builds an index with unseeded RNGs (k-means, HNSW levels) we build the same structure with the seeded
builders in oracle/builders.c -- the asserted answers do not depend on the RNG.
"""
import numpy as np
import pytest

import oracle as O
from tests import synth


# ---- distance kernels ------------------------------------------------------------------------
def test_l2_impls():
    """rs/utils/src/distance/l2.rs:107-117 -- SIMD vs scalar within 1e-5 on random 128-d."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        a, b = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
        assert abs(O.l2(a, b) - O.l2_scalar(a, b)) < 1e-5


def test_dot_impls():
    """rs/utils/src/distance/dot_product.rs:106-113 -- within 2e-5."""
    rng = np.random.default_rng(2)
    for _ in range(20):
        a, b = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
        assert abs(O.dot(a, b) - O.dot_scalar(a, b)) < 2e-5 * 2  # values ~ -32; the reference eps is absolute


def test_lane_conforming():
    """rs/utils/src/distance/lane_conforming.rs:36-57 -- lane-conforming == generic within 1e-5 (16-d, 4 lanes)."""
    rng = np.random.default_rng(3)
    a, b = rng.random(16, dtype=np.float32), rng.random(16, dtype=np.float32)
    assert abs(O.lane_conforming(a, b, 4, O.L2) - O.l2_squared(a, b)) < 1e-5
    assert abs(O.lane_conforming(a, b, 4, O.DOT) - O.dot(a, b)) < 1e-5


def test_l2_lane_structure_exact():
    """l2.rs:30-68 -- the 16/8/4/tail cascade, checked bit-exactly against an independent numpy restatement."""
    rng = np.random.default_rng(4)

    def ref(a, b):
        ret = np.float32(0)
        p = 0
        n = len(a)
        for L in (16, 8, 4):
            if (n - p) // L > 0:
                acc = np.zeros(L, dtype=np.float32)
                chunks = (n - p) // L
                for c in range(chunks):
                    d = a[p + c * L:p + (c + 1) * L] - b[p + c * L:p + (c + 1) * L]
                    acc = acc + d * d
                s = np.float32(-0.0)
                for x in acc:
                    s = np.float32(s + x)
                ret = np.float32(ret + s)
                p += chunks * L
        for i in range(p, n):
            d = np.float32(a[i] - b[i])
            ret = np.float32(ret + d * d)
        return float(ret)

    for n in (1, 3, 4, 7, 8, 15, 16, 17, 30, 31, 128, 768, 771):
        a, b = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        assert O.l2_squared(a, b) == ref(a, b), n


def test_dot_strict_thresholds():
    """dot_product.rs:43,51,59 -- thresholds are strict '>' (len 16 uses two 8-lane chunks, not one 16-lane)."""
    rng = np.random.default_rng(5)
    a, b = rng.random(16, dtype=np.float32), rng.random(16, dtype=np.float32)
    acc = np.zeros(8, dtype=np.float32)
    for c in range(2):
        acc = acc + a[c * 8:(c + 1) * 8] * b[c * 8:(c + 1) * 8]
    s = np.float32(-0.0)
    for x in acc:
        s = np.float32(s + x)
    assert O.dot(a, b) == float(-np.float32(np.float32(0) + s))


# ---- product quantizer -----------------------------------------------------------------------
def test_product_quantizer_known_codes():
    """rs/quantization/src/pq/mod.rs:321-351."""
    cb = []
    for s in range(5):
        for i in range(2):
            cb += [float(s * 2 + i)] * 2
    pq = O.ProductQuantizer(10, 2, 1, cb)
    assert pq.quantize([1, 1, 3, 3, 5, 5, 7, 7, 9, 9]).tolist() == [1, 1, 1, 1, 1]


def test_ivf_writer_known_codes():
    """rs/index/src/ivf/writer.rs:526-614 -- D=3, dsub=1, 1 bit."""
    pq = O.ProductQuantizer(3, 1, 1, [1.5, 4.5, 2.3, 5.3, 3.1, 6.1])
    assert pq.quantize([1.0, 2.0, 3.0]).tolist() == [0, 0, 0]
    assert pq.quantize([4.0, 5.0, 6.0]).tolist() == [1, 1, 1]
    assert pq.quantize(np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32)).tolist() == [[0, 0, 0], [1, 1, 1]]


@pytest.mark.parametrize("dim,dsub,nbits", [(128, 8, 8), (128, 4, 4), (64, 16, 8), (120, 24, 6), (30, 5, 3)])
def test_pq_distance_impls_agree(dim, dsub, nbits):
    """rs/quantization/src/pq/pq_builder.rs:149-188 -- Scalar / SIMD / StreamingSIMD within 1e-5 (relative here,
    dsub%4 != 0 exercises the reference's `sum_1 =` assignment quirk so that case only checks Scalar==SIMD)."""
    rng = np.random.default_rng(6)
    cb = rng.random(dim * (1 << nbits), dtype=np.float32)
    pq = O.ProductQuantizer(dim, dsub, nbits, cb)
    for _ in range(10):
        a = pq.quantize(rng.random(dim, dtype=np.float32))
        b = pq.quantize(rng.random(dim, dtype=np.float32))
        s, v, st = (pq.distance(a, b, i) for i in (O.IMPL_SCALAR, O.IMPL_SIMD, O.IMPL_STREAMING))
        assert abs(s - v) < 1e-5 * max(1.0, s)
        if dsub % 4 == 0:
            assert abs(s - st) < 1e-5 * max(1.0, s)


def test_pq_distance_streaming_tail_quirk():
    """pq/mod.rs:259-261 -- `sum_1 = accumulate_scalar(..)` overwrites (does not accumulate) the tail."""
    rng = np.random.default_rng(7)
    dim, dsub, nbits = 15, 5, 2  # dsub 5 -> 4-lane chunk + 1 scalar tail per subspace
    cb = rng.random(dim * 4, dtype=np.float32)
    pq = O.ProductQuantizer(dim, dsub, nbits, cb)
    a = np.array([0, 1, 2], dtype=np.uint8)
    b = np.array([3, 2, 1], dtype=np.uint8)
    c = pq.codebook.reshape(3, 4, 5)
    s4 = np.zeros(4, dtype=np.float32)
    s1 = np.float32(0)
    for s in range(3):
        d = c[s, a[s]] - c[s, b[s]]
        s4 = s4 + d[:4] * d[:4]
        s1 = np.float32(d[4] * d[4])  # assignment: only the LAST subspace's tail survives
    r = np.float32(-0.0)
    for x in s4:
        r = np.float32(r + x)
    expect = np.float32(np.float32(np.float32(np.float32(0.0) + np.float32(0.0)) + r) + s1)
    assert pq.distance(a, b, O.IMPL_STREAMING) == float(expect)


# ---- IVF assignment --------------------------------------------------------------------------
def test_build_posting_lists_golden():
    """rs/index/src/ivf/builder.rs:810-872 -- 1-d points 1..6, centroids {2.5, 5.5}, 2 clusters/vector, thr 0.1."""
    X = np.arange(1, 7, dtype=np.float32).reshape(6, 1)
    cents = np.array([[2.5], [5.5]], dtype=np.float32)
    offsets, ids = O.build_posting_lists(X, cents, max_clusters=2, threshold=0.1)
    assert ids[offsets[0]:offsets[1]].tolist() == [0, 1, 2, 3]
    assert ids[offsets[1]:offsets[2]].tolist() == [3, 4, 5]


# ---- orderings -------------------------------------------------------------------------------
def test_id_with_score_sorting():
    """rs/index/src/utils.rs:229-296 -- NaN last, ties by doc id."""
    nan = float("nan")
    docs = [5, 2, 1, 0, 4, 1]
    scores = [nan, 1.0, 1.0, 3.0, nan, 2.0]
    d, s = O.merge_topk(docs, scores, 6)
    assert d == [1, 2, 1, 0, 4, 5]
    assert s[:4].tolist() == [1.0, 1.0, 2.0, 3.0] and np.isnan(s[4:]).all()


def test_id_with_score_u128_tiebreak():
    """utils.rs:95-114 with doc ids that differ only above bit 64."""
    big = [(1 << 100) + 7, (1 << 64) + 7, 7]
    d, s = O.merge_topk(big, [1.0, 1.0, 1.0], 2)
    assert d == [7, (1 << 64) + 7]


# ---- SPANN / IVF / HNSW end to end -----------------------------------------------------------
def _spann_flat(extra=None, nlist=10):
    n = 1000
    X = np.repeat(np.arange(n, dtype=np.float32)[:, None], 4, axis=1)
    docs = list(range(n))
    if extra is not None:
        X = np.vstack([X, np.asarray(extra, dtype=np.float32)[None]])
        docs.append(n)
    cents = O.kmeans(X, nlist, iters=25, seed=7)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    ivf = O.Ivf(cents, offsets, ids, X, doc_ids=docs)
    g = O.hnsw_build(cents, max_neighbors=10, max_layer=2, ef_construction=100, seed=3)
    hn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents)
    return O.Spann(hn, ivf), ivf, X


def test_spann_search_golden():
    """rs/index/src/spann/index.rs:293-366 -- q=[2.4,3.4,4.4,5.4], k=2, ef=2 -> doc ids [4, 3]."""
    spann, _, _ = _spann_flat()
    ids, scores = spann.search([2.4, 3.4, 4.4, 5.4], top_k=2, ef_construction=2)
    assert ids == [4, 3]
    assert scores[0] <= scores[1]


def test_spann_search_with_invalidation_golden():
    """rs/index/src/spann/index.rs:369-445 -- after invalidating doc 4 -> [3, 5]."""
    spann, ivf, _ = _spann_flat()
    ivf.invalidate_batch([4])
    ids, _ = spann.search([2.4, 3.4, 4.4, 5.4], top_k=2, ef_construction=2)
    assert ids == [3, 5]


def test_spann_search_with_pq_golden():
    """rs/index/src/spann/index.rs:448-525 -- dsub 2, 2 bits: top-5 scores are all exactly 0.0
    (pins: squared, symmetric, query quantized with the same codebook)."""
    n = 1000
    X = np.repeat(np.arange(n, dtype=np.float32)[:, None], 4, axis=1)
    cents = O.kmeans(X, 10, iters=25, seed=7)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    cb = O.train_pq_codebook(X[np.random.default_rng(0).choice(n, 200, replace=False)], 2, 2, iters=20, seed=5)
    pq = O.ProductQuantizer(4, 2, 2, cb)
    codes = pq.quantize(X)
    ivf = O.Ivf(cents, offsets, ids, codes, doc_ids=list(range(n)), pq=pq)
    g = O.hnsw_build(cents, 10, 2, 100, seed=3)
    hn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents)
    res = O.Spann(hn, ivf).search([2.4, 3.4, 4.4, 5.4], top_k=5, ef_construction=2)
    assert res is not None
    ids_, scores = res
    assert len(ids_) == 5 and (scores == 0.0).all()
    # ties are ordered by doc id (utils.rs:95-114)
    assert ids_ == sorted(ids_)


def _bits(ids, n):
    b = np.zeros((n + 31) // 32, dtype=np.uint32)
    for i in ids:
        b[i >> 5] |= np.uint32(1 << (i & 31))
    return b


def test_spann_search_with_where_document_golden():
    """rs/index/src/multi_spann/index.rs:787-882 -- 10 docs [i,i,i,i], filter `field contains even` (= point ids 0,2,4,6,8),
    q=[4.4]*4, k=10, ef=5: every result is an even doc; and :884-980 -- a filter that matches nothing -> 0 results.
    The planner filter reaches the scan as the id set `plan_with_ids` yields (ivf/block_based/index.rs:212-226)."""
    n = 10
    X = np.repeat(np.arange(n, dtype=np.float32)[:, None], 4, axis=1)
    cents = O.kmeans(X, 3, iters=10, seed=7)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    ivf = O.Ivf(cents, offsets, ids, X, doc_ids=list(range(n)))
    g = O.hnsw_build(cents, 10, 2, 100, seed=3)
    hn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents)
    sp = O.Spann(hn, ivf)
    q = np.full((1, 4), 4.4, dtype=np.float32)
    od, os_, oc = sp.search_batch(q, 10, 5, 3, 10.0, filter_bits=_bits([0, 2, 4, 6, 8], n))
    got = [int(x) for x in od[0, :oc[0], 0]]
    assert oc[0] > 0 and all(d % 2 == 0 for d in got)
    assert got[0] == 4 and (np.diff(os_[0, :oc[0]]) >= 0).all()
    od, os_, oc = sp.search_batch(np.full((1, 4), 2.4, dtype=np.float32), 10, 2, 3, 10.0, filter_bits=_bits([], n))
    assert oc[0] == 0


def test_ivf_filter_equals_bruteforce_over_allowed_ids():
    """index.rs:212-226: with every list probed, the filtered search is the exact top-k over the allowed ids."""
    X = synth.clustered(1500, 32, n_blobs=6, seed=3)
    cents = O.kmeans(X, 12, iters=8, seed=1)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    ivf = O.Ivf(cents, offsets, ids, X)
    rng = np.random.default_rng(5)
    allowed = np.sort(rng.choice(len(X), 300, replace=False))
    Q = X[:20] + 0.01
    od, os_, oc = ivf.search_batch(Q, 7, 12, filter_bits=_bits(allowed.tolist(), len(X)))
    for b in range(len(Q)):
        d = np.array([O.l2(Q[b], X[i]) for i in allowed], dtype=np.float32)
        order = np.lexsort((allowed, d))[:7]
        assert oc[b] == 7
        assert [int(x) for x in od[b, :7, 0]] == [int(allowed[i]) for i in order]


def test_multi_spann_search_golden():
    """rs/index/src/multi_spann/index.rs:358-412 -- extra doc 1000=[1.2,2.2,3.2,4.2]; q=[1.4,2.4,3.4,4.4], k=3
    -> [1000, 3, 2]."""
    spann, _, _ = _spann_flat(extra=[1.2, 2.2, 3.2, 4.2])
    ids, _ = spann.search([1.4, 2.4, 3.4, 4.4], top_k=3, ef_construction=2)
    assert ids == [1000, 3, 2]


def test_ivf_search_structural():
    """rs/index/src/ivf/block_based/index.rs:504-650 -- k results, ascending scores; invalidation removes ids."""
    rng = np.random.default_rng(11)
    X = rng.random((1000, 4), dtype=np.float32)
    cents = O.kmeans(X, 10, 10, seed=1)
    offsets, ids = O.build_posting_lists(X, cents)
    ivf = O.Ivf(cents, offsets, ids, X, doc_ids=[100 + i for i in range(1000)])
    q = rng.random(4, dtype=np.float32)
    d, s = ivf.search(q, 5, 2)
    assert len(d) == 5 and all(s[i] <= s[i + 1] for i in range(4))
    assert all(100 <= x < 1100 for x in d)
    ivf.invalidate_batch([d[0] - 100])
    d2, _ = ivf.search(q, 5, 2)
    assert d[0] not in d2 and d2[:4] == d[1:]
    with pytest.raises(ValueError):
        ivf.search(q, 5, 0)  # reference panics (select_nth_unstable_by(num_probes - 1), index.rs:158)
    with pytest.raises(ValueError):
        ivf.search(q, 5, 11)


def test_ivf_full_probe_equals_bruteforce():
    """With nprobe == nlist the IVF result is the exact top-k by (sqrt L2, point id) (index.rs:147-332)."""
    rng = np.random.default_rng(12)
    X = rng.random((500, 8), dtype=np.float32)
    cents = O.kmeans(X, 8, 5, seed=2)
    offsets, ids = O.build_posting_lists(X, cents)
    ivf = O.Ivf(cents, offsets, ids, X)
    for _ in range(5):
        q = rng.random(8, dtype=np.float32)
        d, s = ivf.search(q, 10, 8)
        ref = sorted((O.l2(q, X[i]), i) for i in range(500))[:10]
        assert d == [i for _, i in ref]
        assert s.tolist() == [np.float32(x) for x, _ in ref]


def test_hnsw_search_structural():
    """rs/index/src/hnsw/block_based/index.rs:368-712 -- <= k results, sorted, ids in range; plus high recall."""
    rng = np.random.default_rng(13)
    X = rng.random((3000, 16), dtype=np.float32)
    g = O.hnsw_build(X, 16, 5, 100, seed=9)
    hn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X)
    hits = 0
    for _ in range(20):
        q = rng.random(16, dtype=np.float32)
        ids, sc, st = hn.ann_search(q, 10, 100, with_stats=True)
        assert len(ids) == 10 and all(sc[i] <= sc[i + 1] for i in range(9)) and all(0 <= i < 3000 for i in ids)
        assert st[0] >= st[1] > 0
        bf = set(np.argsort(((X - q) ** 2).sum(1))[:10].tolist())
        hits += len(bf & set(ids))
    assert hits / 200 > 0.9


def test_hnsw_hand_built_graph():
    """A hand-built 2-layer graph in the array layout of hnsw/block_based/graph_storage.rs:122-193,459-554:
    entry point = points[level_offsets[0]]; layer-0 rows are addressed by point id."""
    X = np.array([[0.0], [1.0], [2.0], [3.0], [4.0]], dtype=np.float32)
    # top layer (layer 1): points [4, 0], edges 4->[0], 0->[4];  layer 0: chain 0-1-2-3-4
    points = [4, 0]
    edges = [0, 4, 1, 0, 2, 1, 3, 2, 4, 3]
    edge_offsets = [0, 1, 2, 3, 5, 7, 9, 10]
    level_offsets = [0, 2, 7]
    hn = O.Hnsw(2, edges, points, edge_offsets, level_offsets, X)
    assert hn.entry_point() == 4
    ids, sc = hn.ann_search([0.9], 2, 4)
    assert ids == [1, 0] and np.allclose(sc, [0.1, 0.9], atol=1e-6)


# ---- k-means assignment step / lane-conforming calculators (round 2 oracle functions) -----------------------------------
def _np_lane_conforming_l2(a, b, lanes):
    """Independent numpy restatement of LaneConformingDistanceCalculator<LANES, L2>::calculate_squared
    (lane_conforming.rs:16-27 -> l2.rs:77-89 accumulate_lanes, ordered reduce_sum starting from -0.0)."""
    acc = np.zeros(lanes, dtype=np.float32)
    for c in range(len(a) // lanes):
        d = a[c * lanes:(c + 1) * lanes] - b[c * lanes:(c + 1) * lanes]
        acc = acc + d * d
    s = np.float32(-0.0)
    for x in acc:
        s = np.float32(s + x)
    return s


def test_lane_conforming_batch_exact():
    """All LANES of kmeans_builder.rs:126-136, bit for bit against the numpy restatement (dimension a multiple of LANES)."""
    rng = np.random.default_rng(21)
    for lanes, dim in ((16, 64), (8, 24), (4, 12), (4, 4), (8, 8), (16, 16)):
        A, B = rng.random((5, dim), dtype=np.float32), rng.random((7, dim), dtype=np.float32)
        got = O.lane_conforming_batch(A, B, lanes, O.L2)
        for i in range(5):
            for j in range(7):
                assert got[i, j] == _np_lane_conforming_l2(A[i], B[j], lanes), (lanes, dim, i, j)


def test_kmeans_assign_rule_exact():
    """kmeans_builder.rs:199-221: cost = calculate_squared(point, centroid) + penalty with the calculator chosen by the
    dimension (:126-136), strict '<' fold from (0, f32::MAX) -> the FIRST minimum wins.  Labels and costs bit for bit against
    numpy, with penalties, exact ties (duplicated centroids) included."""
    rng = np.random.default_rng(22)
    for dim in (16, 24, 12, 6):        # LaneConforming<16>, <8>, <4>, plain D
        lanes = 16 if dim % 16 == 0 else 8 if dim % 8 == 0 else 4 if dim % 4 == 0 else 0
        X = rng.random((40, dim), dtype=np.float32)
        Cn = rng.random((9, dim), dtype=np.float32)
        Cn[5] = Cn[2]                  # exact tie: label 2 must win over 5
        pen = (rng.random(9, dtype=np.float32) * np.float32(0.05)).astype(np.float32)
        pen[5] = pen[2]
        lab, cost = O.kmeans_assign(X, Cn, penalties=pen, with_costs=True)
        for i in range(len(X)):
            best, bl = np.float32(np.finfo(np.float32).max), 0
            for c in range(len(Cn)):
                d = _np_lane_conforming_l2(X[i], Cn[c], lanes) if lanes else np.float32(O.l2_squared(X[i], Cn[c]))
                v = np.float32(d + pen[c])
                if v < best:
                    best, bl = v, c
            assert lab[i] == bl and cost[i] == best, (dim, i)
        assert 5 not in set(lab.tolist())


def _lloyd(data, k, tol, init, iters=100):
    """KMeansBuilder::run_lloyd (kmeans_builder.rs:162-358) with the oracle's assignment step; centroid update and penalties
    as in the reference (no empty-cluster repair: the restated cases never produce one)."""
    X = np.asarray(data, dtype=np.float32)
    cents = X[init].copy()
    sizes = np.zeros(k, dtype=np.int64)
    pen = np.zeros(k, dtype=np.float32)
    labels = np.zeros(len(X), dtype=np.int64)
    for it in range(iters + 1):
        last = labels.copy()
        labels = O.kmeans_assign(X, cents, penalties=pen if tol > 0 else None).astype(np.int64)
        sizes = np.bincount(labels, minlength=k)
        assert (sizes > 0).all()
        cents = np.stack([X[labels == c].sum(0) / np.float32(sizes[c]) for c in range(k)]).astype(np.float32)
        if tol > 0:
            pen = (np.float32(tol) * sizes.astype(np.float32)).astype(np.float32)
        if np.array_equal(labels, last):
            break
    return labels


def test_kmeans_lloyd_reference_case():
    """kmeans_builder.rs:374-415 (test_kmeans_lloyd): three well separated groups, unbalanced penalty 1e-4."""
    data = [[0, 0], [40, 40], [90, 90], [1, 1], [41, 41], [91, 91], [2, 2], [42, 42], [92, 92]]
    a = _lloyd(data, 3, 1e-4, [0, 1, 2])
    assert a[0] == a[3] == a[6] and a[1] == a[4] == a[7] and a[2] == a[5] == a[8]


def test_kmeans_no_distance_penalty_reference_case():
    """kmeans_builder.rs:417-454 (test_kmeans_no_distance_penalty): with tolerance 0 point 7 = (5, 5) joins the cluster of
    points 0, 3, 6."""
    data = [[0, 0], [40, 40], [90, 90], [1, 1], [41, 41], [91, 91], [2, 2], [5, 5], [92, 92]]
    a = _lloyd(data, 3, 0.0, [0, 1, 2])
    assert a[0] == a[3] == a[6] == a[7] and a[1] == a[4] and a[2] == a[5] == a[8]

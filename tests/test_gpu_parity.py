"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar: bit-exact for codes, ids and scores (the kernels restate the reference's lane order with non-fused fp32 ops);
doc-id sets compared exactly, including ties.  Each test cites the reference behaviour it pins.
"""
import numpy as np
import pytest

import oracle as O
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import muopdb_b200 as M
    return M


def _same_f32(a, b):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- DistanceCalculator -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 3, 4, 7, 8, 15, 16, 17, 30, 31, 32, 128, 768, 771])
def test_distance_batch_bit_exact(M, dim):
    """l2.rs:30-74, dot_product.rs:38-71: every lane-cascade shape (16/8/4/tail), bit-exact."""
    rng = np.random.default_rng(dim)
    A = rng.random((37, dim), dtype=np.float32)
    B = rng.random((53, dim), dtype=np.float32)
    assert _same_f32(M.L2DistanceCalculator.calculate_batch(A, B), O.distance_batch(A, B, O.L2, False))
    assert _same_f32(M.L2DistanceCalculator.calculate_batch(A, B, squared=True), O.distance_batch(A, B, O.L2, True))
    assert _same_f32(M.DotProductDistanceCalculator.calculate_batch(A, B), O.distance_batch(A, B, O.DOT))


def test_distance_single_pair_and_device_buffers(M):
    import torch
    rng = np.random.default_rng(5)
    a, b = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
    assert M.L2DistanceCalculator.calculate(a, b) == O.l2(a, b)
    assert M.L2DistanceCalculator.calculate_squared(a, b) == O.l2_squared(a, b)
    assert M.DotProductDistanceCalculator.calculate(a, b) == O.dot(a, b)
    A = rng.random((64, 128), dtype=np.float32)
    B = rng.random((96, 128), dtype=np.float32)
    out = M.L2DistanceCalculator.calculate_batch(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
    M.default_context().sync()
    assert _same_f32(out.cpu().numpy(), O.distance_batch(A, B))


# ---- ProductQuantizer -------------------------------------------------------------------------------------------------
def test_pq_known_codes(M):
    """pq/mod.rs:321-351 and ivf/writer.rs:526-614 golden codes, on the GPU."""
    cb = []
    for s in range(5):
        for i in range(2):
            cb += [float(s * 2 + i)] * 2
    pq = M.ProductQuantizer(10, 2, 1, cb)
    assert pq.quantize([1, 1, 3, 3, 5, 5, 7, 7, 9, 9]).tolist() == [1, 1, 1, 1, 1]
    pq2 = M.ProductQuantizer(3, 1, 1, [1.5, 4.5, 2.3, 5.3, 3.1, 6.1])
    assert pq2.quantize(np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32)).tolist() == [[0, 0, 0], [1, 1, 1]]
    with pytest.raises(M.InvalidArgument):
        M.ProductQuantizer(10, 3, 1, [0.0] * 20)  # "Dimensions are not valid" (pq/mod.rs:41-46)


@pytest.mark.parametrize("dim,dsub,nbits", [(128, 8, 8), (768, 8, 8), (64, 16, 8), (120, 24, 6), (30, 5, 3), (32, 2, 4), (16, 1, 8)])
def test_pq_quantize_and_distance_bit_exact(M, dim, dsub, nbits):
    """pq/mod.rs:152-177 (first-minimum argmin) and :231-266 (StreamingSIMD, incl. the sum_1 assignment quirk)."""
    rng = np.random.default_rng(dim * 31 + dsub)
    cb = rng.random(dim * (1 << nbits), dtype=np.float32)
    # duplicate a few centroids so that exact argmin ties exist (first minimum must win)
    c3 = cb.reshape(dim // dsub, 1 << nbits, dsub)
    c3[:, -1] = c3[:, 0]
    X = rng.random((300, dim), dtype=np.float32)
    X[:10] = c3[:, 0].reshape(-1)  # rows exactly on a duplicated centroid
    opq = O.ProductQuantizer(dim, dsub, nbits, cb)
    gpq = M.ProductQuantizer(dim, dsub, nbits, cb)
    oc, gc = opq.quantize(X), gpq.quantize(X)
    assert np.array_equal(oc, gc)
    a, b = oc[:150], oc[150:]
    od = np.array([opq.distance(a[i], b[i]) for i in range(150)], dtype=np.float32)
    assert _same_f32(gpq.distance(a, b), od)
    # dot-product calculator as the generic D (only reachable from the offline index_writer)
    opd = O.ProductQuantizer(dim, dsub, nbits, cb, metric=O.DOT)
    gpd = M.ProductQuantizer(dim, dsub, nbits, cb, distance=M.DotProductDistanceCalculator)
    od = np.array([opd.distance(a[i], b[i]) for i in range(150)], dtype=np.float32)
    assert _same_f32(gpd.distance(a, b), od)


# ---- IVF --------------------------------------------------------------------------------------------------------------
def _check_ivf(M, X, nlist, nprobe, k, pq_params=None, max_clusters=1, metric="l2", nq=64, invalidate=None, seed=0,
               queries=None):
    rng = np.random.default_rng(seed)
    cents, offsets, ids = synth.build_ivf_arrays(X, nlist, seed=seed + 1, max_clusters=max_clusters)
    docs = synth.doc_ids_for(len(X), seed=seed + 2)
    dim = X.shape[1]
    if pq_params:
        dsub, nbits = pq_params
        cb = O.train_pq_codebook(X[rng.choice(len(X), min(len(X), 2000), replace=False)], dsub, nbits, iters=4, seed=seed)
        opq = O.ProductQuantizer(dim, dsub, nbits, cb, metric=O.DOT if metric == "dot" else O.L2)
        gq = M.ProductQuantizer(dim, dsub, nbits, cb, distance=M.DotProductDistanceCalculator if metric == "dot" else M.L2DistanceCalculator)
        rows = opq.quantize(X)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
    else:
        rows = X
        gq = M.NoQuantizer(dim, M.DotProductDistanceCalculator if metric == "dot" else M.L2DistanceCalculator)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, metric=O.DOT if metric == "dot" else O.L2)
    givf = M.BlockBasedIvf(cents, offsets, ids, rows, gq, doc_ids=docs)
    assert givf.num_clusters() == nlist and givf.num_vectors() == len(X)
    if invalidate is not None:
        oivf.invalidate_batch(invalidate)
        givf.invalidate_points(invalidate)
        assert givf.is_point_invalidated(int(invalidate[0]))
    Q = queries if queries is not None else np.vstack([X[:nq // 2] + 0.01 * rng.standard_normal((nq // 2, dim)).astype(np.float32),
                                                         rng.random((nq - nq // 2, dim), dtype=np.float32)])
    # coarse: identical probe lists (distances are bit-exact, ties by centroid index on both sides)
    gp, gd = givf.find_nearest_centroids_batch(Q, nprobe, with_distances=True)
    for b in range(len(Q)):
        op, od = oivf.find_nearest_centroids(Q[b], nprobe, with_dist=True)
        assert np.array_equal(op, gp[b]), b
        assert _same_f32(od, gd[b])
    # full search
    od, os_, oc = oivf.search_batch(Q, k, nprobe)
    res = givf.search_batch(Q, k, nprobe)
    assert np.array_equal(np.asarray(res.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(len(Q)):
        n = oc[b]
        assert np.array_equal(res.doc_ids[b, :n], od[b, :n]), (b, res.doc_ids[b, :n], od[b, :n], res.scores[b, :n], os_[b, :n])
        assert _same_f32(res.scores[b, :n], os_[b, :n]), b
    # search_with_centroids (point ids, ordered by (distance, point_id)) for a few queries
    r2 = givf.search_with_centroids_batch(Q[:8], gp[:8], k, remap=False)
    for b in range(8):
        pids, ds = oivf.search_with_centroids(Q[b], gp[b], k)
        assert np.array_equal(r2.doc_ids[b, :len(pids)], pids)
        assert _same_f32(r2.scores[b, :len(pids)], ds)
    return givf, oivf, Q


def test_ivf_flat_config1_fixture(M):
    """BASELINE config 1 on the reference's own committed rows (rs/index/resources/10000_rows_128_dim, all 10 000 rows
    committed under tests/golden): IVF flat-L2, 10k x 128, nlist 64, nprobe 8."""
    import os
    X = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "rows_10000x128.f32"), dtype="<f4").reshape(10000, 128)
    _check_ivf(M, X, nlist=64, nprobe=8, k=10, nq=100)


@pytest.mark.parametrize("dim,nlist,nprobe,k", [(4, 10, 2, 5), (30, 16, 16, 10), (128, 32, 5, 1), (768, 24, 6, 10), (20, 7, 7, 32)])
def test_ivf_flat_l2_parity(M, dim, nlist, nprobe, k):
    """ivf/block_based/index.rs:147-332 with NoQuantizer<L2>: sqrt scores, (score, doc_id) order."""
    X = synth.clustered(3000, dim, n_blobs=12, seed=dim)
    _check_ivf(M, X, nlist, nprobe, k, seed=dim)


def test_ivf_flat_dot_parity(M):
    """NoQuantizer<DotProduct> (index_writer only): negated dot scores; coarse scoring stays sqrt-L2 (index.rs:155)."""
    X = synth.uniform(2000, 48, seed=9)
    _check_ivf(M, X, 16, 4, 10, metric="dot", seed=9)


@pytest.mark.parametrize("dim,dsub,nbits,nlist,nprobe", [(256, 8, 8, 16, 4), (768, 8, 8, 24, 8), (512, 8, 8, 8, 8), (1024, 8, 8, 6, 3),
                                                         (128, 8, 8, 32, 8), (64, 16, 4, 8, 3), (30, 5, 3, 8, 8), (128, 4, 8, 10, 4)])
def test_ivf_pq_parity(M, dim, dsub, nbits, nlist, nprobe):
    """ProductQuantizer<L2>: symmetric, squared scores (pq/mod.rs:231-266); query quantized with the same codebook
    (index.rs:193).  m in {32,64,96,128} with 8 bits runs the conflict-free fast scan, the rest the generic one."""
    X = synth.clustered(4000, dim, n_blobs=16, seed=dim + dsub)
    _check_ivf(M, X, nlist, nprobe, 10, pq_params=(dsub, nbits), seed=dim + dsub)


def test_ivf_pq_exact_ties_and_duplicates(M):
    """Many rows share a code word (exact score ties) and points live in two lists (max_clusters_per_vector = 2):
    ties resolve by point id then doc id, duplicates are kept (index.rs:265-274, utils.rs:71-76,95-114)."""
    rng = np.random.default_rng(77)
    base = synth.clustered(400, 256, n_blobs=6, seed=77)
    X = np.repeat(base, 5, axis=0)  # 5 exact copies of every vector
    X = X[rng.permutation(len(X))]
    _check_ivf(M, X, nlist=8, nprobe=8, k=16, pq_params=(8, 8), max_clusters=2, seed=78)
    _check_ivf(M, X, nlist=8, nprobe=3, k=16, max_clusters=2, seed=79)


def test_ivf_pq_heavy_queries_take_the_prefix_search_launch(M):
    """The table-driven PQ scan keeps one shared-memory word per 32-row chunk of a query (2048 of them); a query whose
    probed lists are longer is deferred to the second, prefix-search launch.  Mixed batch: light and heavy queries, plus
    invalidations (applied only to rows that pass the running threshold on the table path)."""
    rng = np.random.default_rng(5)
    dim = 256
    big = (rng.standard_normal((70000, dim)) * 0.05 + 3.0).astype(np.float32)      # one huge, tight cluster
    small = synth.clustered(3000, dim, n_blobs=6, seed=9)
    X = np.vstack([big, small]).astype(np.float32)
    inv = np.arange(5, 73000, 97, dtype=np.uint32)
    Q = np.vstack([big[:12] + 0.001, small[:12] + 0.01]).astype(np.float32)
    givf, _, _ = _check_ivf(M, X, nlist=8, nprobe=3, k=10, pq_params=(8, 8), invalidate=inv, seed=31, queries=Q)
    assert max(givf.ctx.lib.mgpu_ivf_last_scan_rows(givf.handle), 0) > 0


def test_ivf_invalidation_and_short_results(M):
    """index.rs:198-200: invalidated ids are skipped; fewer than k results are possible."""
    X = synth.clustered(600, 64, n_blobs=4, seed=5)
    inv = np.arange(0, 600, 3, dtype=np.uint32)
    _check_ivf(M, X, nlist=40, nprobe=2, k=32, invalidate=inv, seed=6)
    _check_ivf(M, X, nlist=40, nprobe=1, k=32, pq_params=(2, 8), invalidate=inv, seed=7)


def test_ivf_empty_lists_and_probe_range(M):
    """Empty posting lists are legal; num_probes == 0 or > num_clusters panics in the reference (index.rs:158)."""
    X = synth.uniform(50, 16, seed=3)
    cents = np.vstack([X[:5], 10.0 + synth.uniform(3, 16, seed=4)]).astype(np.float32)  # last 3 lists stay empty
    offsets, ids = O.build_posting_lists(X, cents)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(16))
    oivf = O.Ivf(cents, offsets, ids, X)
    Q = synth.uniform(9, 16, seed=5)
    Q[0] += 10.0  # probes only empty lists with nprobe small
    for nprobe in (1, 3, 8):
        od, os_, oc = oivf.search_batch(Q, 10, nprobe)
        r = givf.search_batch(Q, 10, nprobe)
        assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
        for b in range(len(Q)):
            assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]) and _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]])
    with pytest.raises(M.OutOfRange):
        givf.search_batch(Q, 10, 0)
    with pytest.raises(M.OutOfRange):
        givf.search_batch(Q, 10, 9)
    r = givf.search_batch(Q, 0, 2)  # k == 0: a heap of capacity 0 keeps nothing
    assert not np.asarray(r.counts).any()
    assert givf.search(Q[1], 3, 2).id_with_scores[0].score == O.Ivf(cents, offsets, ids, X).search(Q[1], 3, 2)[1][0]


def test_ivf_device_buffers_and_scan_accounting(M):
    """MGPU_DEVICE path: torch CUDA tensors in and out; algorithmic-bytes accounting = sum of probed list lengths."""
    import torch
    X = synth.clustered(5000, 256, n_blobs=20, seed=21)
    cents, offsets, ids = synth.build_ivf_arrays(X, 32, seed=22)
    cb = O.train_pq_codebook(X[:2000], 8, 8, iters=3, seed=1)
    opq = O.ProductQuantizer(256, 8, 8, cb)
    gpq = M.ProductQuantizer(256, 8, 8, cb)
    codes = opq.quantize(X)
    oivf = O.Ivf(cents, offsets, ids, codes, pq=opq)
    givf = M.BlockBasedIvf(cents, offsets, ids, torch.from_numpy(codes).cuda(), gpq)  # rows already resident in HBM
    Q = X[100:164] + 0.01
    od, os_, oc = oivf.search_batch(Q, 10, 6)
    r = givf.search_batch(torch.from_numpy(Q).cuda(), 10, 6)
    givf.ctx.sync()
    d = r.doc_ids.cpu().numpy().view(np.uint64)
    assert np.array_equal(d, od) and _same_f32(r.scores.cpu().numpy(), os_)
    lens = np.diff(offsets).astype(np.int64)
    expect_rows = sum(int(lens[oivf.find_nearest_centroids(Q[b], 6)].sum()) for b in range(len(Q)))
    assert givf.last_scan_rows() == expect_rows
    assert givf.last_scan_bytes() == expect_rows * (32 + 4)


# ---- build-time assignment / merge --------------------------------------------------------------------------------------
def test_assign_matches_builder_rule(M):
    """ivf/builder.rs:268-329 incl. the golden case of :810-872."""
    X = np.arange(1, 7, dtype=np.float32).reshape(6, 1)
    cents = np.array([[2.5], [5.5]], dtype=np.float32)
    cids, cnt = M.assign_to_centroids(X, cents, 2, 0.1)
    assert cnt.tolist() == [1, 1, 1, 2, 1, 1] and cids[3].tolist() == [0, 1] and cids[0, 0] == 0 and cids[5, 0] == 1
    Xb = synth.clustered(3000, 128, seed=8)
    cb = O.kmeans(Xb, 50, 4, seed=2)
    for r, thr in ((1, 0.1), (3, 0.2)):
        oc, on = O.ivf_assign(Xb, cb, r, thr)
        gc, gn = M.assign_to_centroids(Xb, cb, r, thr)
        assert np.array_equal(on, gn) and np.array_equal(oc, gc)


def test_merge_topk(M):
    """collection/snapshot.rs:60-61,105-106 + utils.rs:95-114 (NaN last, doc-id tie-break on the full 128 bits)."""
    rng = np.random.default_rng(4)
    S, B, k = 5, 33, 7
    docs = np.zeros((S, B, k, 2), dtype=np.uint64)
    docs[..., 0] = rng.integers(0, 50, (S, B, k))
    docs[..., 1] = rng.integers(0, 2, (S, B, k))
    scores = rng.integers(0, 6, (S, B, k)).astype(np.float32)  # many ties
    scores[0, 0, 0] = np.nan
    counts = rng.integers(0, k + 1, (S, B)).astype(np.uint32)
    r = M.merge_topk(docs, scores, counts, k)
    for b in range(B):
        dl, sl = [], []
        for s in range(S):
            n = counts[s, b]
            dl += [int(lo) | (int(hi) << 64) for lo, hi in docs[s, b, :n]]
            sl += scores[s, b, :n].tolist()
        ed, es = O.merge_topk(dl, sl, k) if dl else ([], np.zeros(0, np.float32))
        assert int(r.counts[b]) == len(ed)
        got = [int(lo) | (int(hi) << 64) for lo, hi in r.doc_ids[b, :len(ed)]]
        assert got == ed
        assert np.array_equal(np.isnan(es), np.isnan(r.scores[b, :len(ed)]))
        assert np.array_equal(es[~np.isnan(es)], r.scores[b, :len(ed)][~np.isnan(es)])


# ---- HNSW / SPANN ---------------------------------------------------------------------------------------------------------
def _graph(X, M_=16, layers=5, efc=60, seed=9):
    g = O.hnsw_build(X, M_, layers, efc, seed=seed)
    return g


@pytest.mark.parametrize("dim,n,ef,k", [(16, 3000, 64, 10), (128, 2000, 128, 10), (768, 800, 32, 5), (20, 1500, 100, 32), (4, 500, 1, 1)])
def test_hnsw_flat_parity(M, dim, n, ef, k):
    """hnsw/block_based/index.rs:159-298: identical ids, scores and traversal counts (#distance evals, #expansions)."""
    X = synth.clustered(n, dim, n_blobs=10, seed=dim)
    g = _graph(X, seed=dim)
    docs = synth.doc_ids_for(n, seed=1)
    oh = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X, doc_ids=docs)
    gh = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X,
                          M.NoQuantizer(dim), doc_ids=docs)
    rng = np.random.default_rng(3)
    Q = np.vstack([X[:20] + 0.02 * rng.standard_normal((20, dim)).astype(np.float32), rng.random((20, dim), dtype=np.float32)])
    od, os_, oc, ost = oh.search_batch(Q, k, ef)
    r, st = gh.ann_search_batch(Q, k, ef, with_stats=True)
    assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(len(Q)):
        assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]), b
        assert _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]]), b
    assert np.array_equal(st, ost)


def test_hnsw_pq_parity(M):
    """BlockBasedHnsw<ProductQuantizer<L2>>: quantized query (index.rs:168), squared SDC scores, many exact ties."""
    dim, n = 64, 2500
    X = synth.clustered(n, dim, n_blobs=8, seed=31)
    g = _graph(X, seed=31)
    cb = O.train_pq_codebook(X[:1500], 8, 4, iters=4, seed=2)  # 4 bits -> coarse codes -> lots of ties
    opq, gpq = O.ProductQuantizer(dim, 8, 4, cb), M.ProductQuantizer(dim, 8, 4, cb)
    codes = opq.quantize(X)
    oh = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], codes, pq=opq)
    gh = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], codes, gpq)
    Q = X[50:90] + 0.01
    od, os_, oc, ost = oh.search_batch(Q, 10, 48)
    r, st = gh.ann_search_batch(Q, 10, 48, with_stats=True)
    assert np.array_equal(r.doc_ids, od) and _same_f32(r.scores, os_) and np.array_equal(st, ost)


def test_hnsw_hand_built_graph(M):
    """graph_storage.rs:459-554 on a hand-built 2-layer file image (same case as tests/test_oracle_golden.py)."""
    X = np.array([[0.0], [1.0], [2.0], [3.0], [4.0]], dtype=np.float32)
    gh = M.BlockBasedHnsw(2, [0, 4, 1, 0, 2, 1, 3, 2, 4, 3], [4, 0], [0, 1, 2, 3, 5, 7, 9, 10], [0, 2, 7], X, M.NoQuantizer(1))
    res = gh.ann_search([0.9], 2, 4)
    assert [x.doc_id for x in res.id_with_scores] == [1, 0]


def _spann_pair(M, X, docs, nlist, pq_params=None, seed=7):
    cents = O.kmeans(X, nlist, iters=25, seed=seed)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    g = O.hnsw_build(cents, 10, 2, 100, seed=3)
    dim = X.shape[1]
    if pq_params:
        dsub, nbits, ntrain = pq_params
        cb = O.train_pq_codebook(X[np.random.default_rng(0).choice(len(X), ntrain, replace=False)], dsub, nbits, iters=20, seed=5)
        opq, gpq = O.ProductQuantizer(dim, dsub, nbits, cb), M.ProductQuantizer(dim, dsub, nbits, cb)
        rows = opq.quantize(X)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
        givf = M.BlockBasedIvf(cents, offsets, ids, rows, gpq, doc_ids=docs)
    else:
        oivf = O.Ivf(cents, offsets, ids, X, doc_ids=docs)
        givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(dim), doc_ids=docs)
    ohn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents)
    ghn = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents, M.NoQuantizer(dim))
    return O.Spann(ohn, oivf), M.Spann(ghn, givf), oivf, givf


def test_spann_reference_golden_cases(M):
    """spann/index.rs:293-366 ([4,3]), :369-445 (after invalidating 4 -> [3,5]), :448-525 (PQ: five exact zeros),
    multi_spann/index.rs:358-412 ([1000,3,2]) -- the reference's own known answers, on the GPU."""
    n = 1000
    X = np.repeat(np.arange(n, dtype=np.float32)[:, None], 4, axis=1)
    _, gs, _, givf = _spann_pair(M, X, list(range(n)), 10)
    res = gs.search([2.4, 3.4, 4.4, 5.4], M.SearchParams(2, 2))
    assert [x.doc_id for x in res.id_with_scores] == [4, 3]
    givf.invalidate(4)
    res = gs.search([2.4, 3.4, 4.4, 5.4], M.SearchParams(2, 2))
    assert [x.doc_id for x in res.id_with_scores] == [3, 5]
    _, gs, _, _ = _spann_pair(M, X, list(range(n)), 10, pq_params=(2, 2, 200))
    res = gs.search([2.4, 3.4, 4.4, 5.4], M.SearchParams(5, 2))
    assert len(res.id_with_scores) == 5 and all(x.score == 0.0 for x in res.id_with_scores)
    X2 = np.vstack([X, np.array([[1.2, 2.2, 3.2, 4.2]], dtype=np.float32)])
    _, gs, _, _ = _spann_pair(M, X2, list(range(n + 1)), 10)
    res = gs.search([1.4, 2.4, 3.4, 4.4], M.SearchParams(3, 2))
    assert [x.doc_id for x in res.id_with_scores] == [1000, 3, 2]


@pytest.mark.parametrize("pq", [None, (8, 8, 1500)])
def test_spann_parity(M, pq):
    """spann/index.rs:211-266: centroid HNSW -> ratio prune -> list scan -> remap, vs the oracle, default and loose ratio."""
    X = synth.clustered(4000, 256, n_blobs=24, seed=41)
    docs = synth.doc_ids_for(len(X), seed=4)
    osp, gsp, _, _ = _spann_pair(M, X, docs, 48, pq_params=pq)
    Q = X[200:260] + 0.01
    for ne, ratio in ((10, 0.1), (16, 10.0), (1, 0.0)):
        od, os_, oc = osp.search_batch(Q, 10, 50, ne, ratio)
        r = gsp.search_batch(Q, M.SearchParams(10, 50, False, ne, ratio))
        assert np.array_equal(np.asarray(r.counts).astype(np.int32), oc)  # UINT32_MAX == -1 encodes None
        for b in range(len(Q)):
            n = max(int(oc[b]), 0)
            assert np.array_equal(r.doc_ids[b, :n], od[b, :n]), (ne, ratio, b)
            assert _same_f32(r.scores[b, :n], os_[b, :n])


# ---- planner filter hook ---------------------------------------------------------------------------------
def _bits(ids, n):
    b = np.zeros((n + 31) // 32, dtype=np.uint32)
    ids = np.asarray(ids, dtype=np.int64)
    np.bitwise_or.at(b, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    return b


def test_spann_where_document_reference_golden(M):
    """multi_spann/index.rs:787-882 (only even docs come back) and :884-980 (a filter matching nothing -> no results)."""
    n = 10
    X = np.repeat(np.arange(n, dtype=np.float32)[:, None], 4, axis=1)
    _, gs, _, _ = _spann_pair(M, X, list(range(n)), 3)
    res = gs.search([4.4] * 4, M.SearchParams(10, 5, False, 3, 10.0), planner=M.Planner([0, 2, 4, 6, 8]))
    got = [x.doc_id for x in res.id_with_scores]
    assert got and all(d % 2 == 0 for d in got) and got[0] == 4
    res = gs.search([2.4] * 4, M.SearchParams(10, 2, False, 3, 10.0), planner=M.Planner([]))
    assert res is not None and res.id_with_scores == []


@pytest.mark.parametrize("pq", [None, (8, 8, 1500)])
@pytest.mark.parametrize("on_device", [False, True])
def test_ivf_planner_filter_parity(M, pq, on_device):
    """BlockBasedIvf::search with Some(planner) (index.rs:212-226, 396-412): one shared filter, one filter per query (with a
    None entry = unfiltered), a filter that leaves fewer than k rows, together with invalidations -- vs the oracle."""
    X = synth.clustered(6000, 256, n_blobs=20, seed=11)
    docs = synth.doc_ids_for(len(X), seed=2)
    _, _, oivf, givf = _spann_pair(M, X, docs, 32, pq_params=pq)
    rng = np.random.default_rng(9)
    Q = (X[100:164] + 0.01).astype(np.float32)
    B, n = len(Q), len(X)
    oivf.invalidate_batch([101, 105, 3000]); givf.invalidate_points([101, 105, 3000])
    Qg = Q
    if on_device:
        import torch
        Qg = torch.from_numpy(Q).cuda()
    shared = np.sort(rng.choice(n, 900, replace=False))
    od, os_, oc = oivf.search_batch(Q, 10, 8, filter_bits=_bits(shared, n))
    r = givf.search_batch(Qg, 10, 8, planner=M.Planner(shared))
    _cmp_batch(r, od, os_, oc)
    sets = [np.sort(rng.choice(n, int(sz), replace=False)) for sz in rng.integers(3, 2000, B)]
    sets[5] = np.array([7], dtype=np.int64)   # fewer than k survivors
    per_q = np.stack([_bits(s_, n) for s_ in sets])
    planners = [M.Planner(s_) for s_ in sets]
    per_q[9] = 0xFFFFFFFF
    planners[9] = None                        # no filter for this query
    od, os_, oc = oivf.search_batch(Q, 10, 8, filter_bits=per_q)
    r = givf.search_batch(Qg, 10, 8, planner=planners)
    _cmp_batch(r, od, os_, oc)


def _cmp_batch(r, od, os_, oc):
    d = r.doc_ids.cpu().numpy().view(np.uint64) if hasattr(r.doc_ids, "cpu") else r.doc_ids
    s = r.scores.cpu().numpy() if hasattr(r.scores, "cpu") else r.scores
    c = r.counts.cpu().numpy() if hasattr(r.counts, "cpu") else r.counts
    assert np.array_equal(np.asarray(c).astype(np.int64), oc.astype(np.int64))
    for b in range(len(oc)):
        n = int(oc[b])
        assert np.array_equal(d[b, :n], od[b, :n]), b
        assert _same_f32(s[b, :n], os_[b, :n])


def test_spann_planner_filter_parity(M):
    X = synth.clustered(4000, 128, n_blobs=24, seed=41)
    docs = synth.doc_ids_for(len(X), seed=4)
    osp, gsp, _, _ = _spann_pair(M, X, docs, 48, pq_params=(8, 8, 1500))
    Q = X[200:240] + 0.01
    allowed = np.sort(np.random.default_rng(1).choice(len(X), 700, replace=False))
    od, os_, oc = osp.search_batch(Q, 10, 50, 8, 1.0, filter_bits=_bits(allowed, len(X)))
    r = gsp.search_batch(Q, M.SearchParams(10, 50, False, 8, 1.0), planner=M.Planner(allowed))
    assert np.array_equal(np.asarray(r.counts).astype(np.int32), oc)
    for b in range(len(Q)):
        n = max(int(oc[b]), 0)
        assert np.array_equal(r.doc_ids[b, :n], od[b, :n]), b
        assert _same_f32(r.scores[b, :n], os_[b, :n])


# ---- micro-batcher ---------------------------------------------------------------------------------------
def test_micro_batcher_concurrent_callers_match_oracle(M):
    """Many threads call the single-query search (the reference's call shape, index.rs:396-412); the native batcher groups
    them into batched GPU searches.  Every caller must get exactly the oracle's answer for ITS query (also with a per-request
    planner filter), and the calls must really have been batched."""
    import threading
    X = synth.clustered(5000, 128, n_blobs=16, seed=21)
    docs = synth.doc_ids_for(len(X), seed=6)
    _, _, oivf, givf = _spann_pair(M, X, docs, 32, pq_params=(8, 8, 1500))
    nthreads, per_thread = 24, 12
    Q = (X[:nthreads * per_thread] + 0.01).astype(np.float32)
    rng = np.random.default_rng(3)
    allowed = [np.sort(rng.choice(len(X), 800, replace=False)) if i % 3 == 0 else None for i in range(len(Q))]
    mb = M.MicroBatcher(givf, k=10, num_probes=8, max_batch=16, max_wait_us=20000)
    got, errs = [None] * len(Q), []

    def worker(t):
        try:
            for j in range(per_thread):
                i = t * per_thread + j
                got[i] = mb.search(Q[i], planner=M.Planner(allowed[i]) if allowed[i] is not None else None)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    st = mb.stats()
    mb.close()
    assert st["queries"] == len(Q) and st["batches"] < len(Q) and st["largest_batch"] > 1
    for i in range(len(Q)):
        fb = None if allowed[i] is None else _bits(allowed[i], len(X))
        od, os_, oc = oivf.search_batch(Q[i:i + 1], 10, 8, filter_bits=fb)
        n = int(oc[0])
        exp_ids = [int(lo) | (int(hi) << 64) for lo, hi in od[0, :n]]
        assert [x.doc_id for x in got[i].id_with_scores] == exp_ids, i
        assert _same_f32(np.array([x.score for x in got[i].id_with_scores], dtype=np.float32), os_[0, :n])


def test_micro_batcher_spann_and_errors(M):
    X = np.repeat(np.arange(1000, dtype=np.float32)[:, None], 4, axis=1)
    _, gs, _, givf = _spann_pair(M, X, list(range(1000)), 10)
    mb = M.MicroBatcher(gs, k=2, params=M.SearchParams(2, 2), max_batch=8, max_wait_us=100)
    res = mb.search([2.4, 3.4, 4.4, 5.4])     # spann/index.rs:293-366
    assert [x.doc_id for x in res.id_with_scores] == [4, 3]
    mb.close()
    with pytest.raises(M.OutOfRange):          # num_probes == 0 panics in the reference (index.rs:158)
        M.MicroBatcher(givf, k=2, num_probes=0)


# ---- tensor-core coarse scoring --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,nlist,nprobe", [(64, 1024, 16), (128, 2048, 64), (768, 1280, 32), (80, 1028, 7), (768, 512, 8), (96, 256, 256)])
def test_coarse_tensor_core_path_matches_exact(M, dim, nlist, nprobe):
    """find_nearest_centroids (index.rs:147-163) through the tcgen05 GEMM + margin + exact re-score: the probe lists and
    distances must be IDENTICAL to the oracle's (the tensor-core pass only proposes candidates).  Includes duplicated
    centroids (exact ties -> index order) and near-ties far below bf16/tf32 resolution."""
    rng = np.random.default_rng(dim + nlist)
    cents = synth.clustered(nlist, dim, n_blobs=9, sigma=0.3, seed=dim).astype(np.float32)
    cents[1::97] = cents[0::97][: len(cents[1::97])]                    # exact duplicates
    cents[2::101] = cents[3::101][: len(cents[2::101])] * np.float32(1.0 + 1e-6)  # near-ties ~1e-6 apart
    X = cents[rng.integers(0, nlist, 3000)] + 0.05 * rng.standard_normal((3000, dim)).astype(np.float32)
    offsets, ids = O.build_posting_lists(X, cents)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(dim))
    oivf = O.Ivf(cents, offsets, ids, X)
    Q = np.vstack([X[:100] + 0.01, cents[:28], rng.random((100, dim), dtype=np.float32) * 3 - 1]).astype(np.float32)
    gp, gd = givf.find_nearest_centroids_batch(Q, nprobe, with_distances=True)
    for b in range(len(Q)):
        op, od = oivf.find_nearest_centroids(Q[b], nprobe, with_dist=True)
        assert np.array_equal(op, gp[b]), (b, op[:8], gp[b][:8])
        assert _same_f32(od, gd[b])
    assert givf.ctx.profile_get(0)[1] > 0
    # the search path only asks the selection for the probe SET (certain members are not re-scored, only the uncertain
    # band is): results must still be the oracle's
    od, os_, oc = oivf.search_batch(Q, 10, nprobe)
    r = givf.search_batch(Q, 10, nprobe)
    assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(len(Q)):
        n = int(oc[b])
        assert np.array_equal(r.doc_ids[b, :n], od[b, :n]), b
        assert _same_f32(r.scores[b, :n], os_[b, :n])


def test_pipelined_submit_wait_matches_blocking_search(M):
    """mgpu_ivf_search_submit / mgpu_search_wait: batches in flight two at a time through their own staging buffers must
    return exactly what the blocking call (and the oracle) returns, in any wait order, including reuse of a slot whose
    ticket was never waited for."""
    import torch
    X = synth.clustered(6000, 256, n_blobs=12, seed=3)
    givf, oivf, _ = _check_ivf(M, X, nlist=24, nprobe=6, k=10, pq_params=(8, 8), seed=4)
    rng = np.random.default_rng(11)
    k, nprobe, nb, B = 10, 6, 7, 96
    Qs = [torch.from_numpy((X[rng.integers(0, len(X), B)] + 0.02 * rng.standard_normal((B, 256))).astype(np.float32)).pin_memory()
          for _ in range(nb)]
    outs = [(torch.zeros((B, k, 2), dtype=torch.int64).pin_memory(), torch.zeros((B, k), dtype=torch.float32).pin_memory(),
             torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(nb)]
    tickets = []
    for i in range(nb):
        tickets.append(givf.search_batch_submit(Qs[i], k, nprobe, outs[i]))
        if i >= 1 and i != 4:          # batch 3's ticket is only waited for at the end: its slot is reused by batch 5
            givf.search_wait(tickets[i - 1])
    for t in reversed(tickets):
        givf.search_wait(t)
    givf.search_wait(tickets[0])       # waiting twice is harmless
    for i in range(nb):
        od, os_, oc = oivf.search_batch(Qs[i].numpy(), k, nprobe)
        ids, sc, cn = (x.numpy() for x in outs[i])
        assert np.array_equal(cn.astype(np.int64), oc.astype(np.int64)), i
        for b in range(B):
            n = int(oc[b])
            assert np.array_equal(ids[b, :n].view(np.uint64), np.asarray(od[b, :n], dtype=np.uint64)), (i, b)
            assert _same_f32(sc[b, :n], os_[b, :n]), (i, b)
    with pytest.raises(M.InvalidArgument):
        givf.search_wait(10 ** 9)


def test_shard_search_single_rank_world(M):
    """mgpu_shard_ivf_search on a 1-rank NCCL world (all-gather + merge of one shard) must equal the plain search and the
    oracle, for device and host buffers.  Multi-rank equivalence is checked by tools/check_shard_search.py under torchrun."""
    import torch
    X = synth.clustered(5000, 256, n_blobs=10, seed=21)
    ctx = M.Context(0)
    ctx.comm_init(1, 0, M.Context.comm_unique_id())
    rng = np.random.default_rng(2)
    cents = O.kmeans(X, 20, iters=4, seed=1)
    offsets, ids = O.build_posting_lists(X, cents)
    cb = O.train_pq_codebook(X[:2000], 8, 8, iters=3, seed=2)
    opq, gpq = O.ProductQuantizer(256, 8, 8, cb), M.ProductQuantizer(256, 8, 8, cb, ctx=ctx)
    codes = opq.quantize(X)
    oivf = O.Ivf(cents, offsets, ids, codes, pq=opq)
    givf = M.BlockBasedIvf(cents, offsets, ids, codes, gpq, ctx=ctx)
    Q = (X[rng.integers(0, len(X), 130)] + 0.02 * rng.standard_normal((130, 256))).astype(np.float32)
    od, os_, oc = oivf.search_batch(Q, 10, 5)
    for dev in (False, True):
        r = givf.shard_search_batch(torch.from_numpy(Q).cuda() if dev else Q, 10, 5)
        ctx.sync()   # device-buffer calls are asynchronous on the library stream
        ids_, sc, cn = ((x.cpu().numpy() if dev else np.asarray(x)) for x in (r.doc_ids, r.scores, r.counts))
        assert np.array_equal(cn.astype(np.int64), oc.astype(np.int64))
        for b in range(len(Q)):
            n = int(oc[b])
            assert np.array_equal(ids_[b, :n].view(np.uint64), np.asarray(od[b, :n], dtype=np.uint64)), (dev, b)
            assert _same_f32(sc[b, :n], os_[b, :n]), (dev, b)


@pytest.mark.parametrize("k", [33, 64, 100, 257])
@pytest.mark.parametrize("pq", [None, (8, 8)])
def test_ivf_large_k_multi_round(M, k, pq):
    """k > 32 runs several scan rounds (each hides the rows reported so far) and merges their exact results: same ids,
    order and scores as the reference's bounded heap of size k (index.rs:265-274), through search (doc ids) and
    search_with_centroids (point ids), with fewer than k rows available for some queries."""
    X = synth.clustered(3000, 128, n_blobs=8, seed=k)
    _check_ivf(M, X, nlist=12, nprobe=5, k=k, pq_params=pq, seed=k + 1)


def test_ivf_large_k_with_duplicates_and_invalidation(M):
    """Exact code-word ties, points living in two probed lists (same composite twice) and invalidated ids, k = 48 and 120:
    the group sharing a round's last composite is carried over to the next round as a whole."""
    rng = np.random.default_rng(7)
    base = synth.clustered(300, 256, n_blobs=5, seed=17)
    X = np.repeat(base, 4, axis=0)[rng.permutation(1200)]
    inv = np.arange(3, 1200, 11, dtype=np.uint32)
    _check_ivf(M, X, nlist=6, nprobe=6, k=48, pq_params=(8, 8), max_clusters=2, invalidate=inv, seed=18)
    _check_ivf(M, X, nlist=6, nprobe=4, k=120, max_clusters=2, invalidate=inv, seed=19)


@pytest.mark.parametrize("k,pq", [(1000, None), (2048, None), (1000, (8, 8)), (24, (8, 8)), (32, (8, 8))])
def test_ivf_large_k_every_point_in_two_probed_lists(M, k, pq):
    """ADVICE r1 (api.cu multi-round top-k): with every point living in two probed lists each composite occurs twice, a
    round that ends inside a pair reports 30 instead of 31 candidates and a fixed ceil((k+16)/31) rounds came back short
    (k = 1000 -> 990).  The rounds now run until k + 16 candidates are reported; the reference keeps both copies of a point
    (no de-dup, index.rs:265-274).  k = 24 / 32 with PQ: the k > 16 PQ searches take the same path (spare candidates for
    swaps between the fixed-point ranking and the exact scores)."""
    X = synth.clustered(3200, 64, n_blobs=4, seed=k)
    cents, offsets, ids = synth.build_ivf_arrays(X, 4, seed=k + 1, max_clusters=2, threshold=1e9)
    assert offsets[-1] == 2 * len(X)
    docs = synth.doc_ids_for(len(X), seed=k + 2)
    if pq:
        cb = O.train_pq_codebook(X[:1500], pq[0], pq[1], iters=3, seed=k)
        opq, gq = O.ProductQuantizer(64, pq[0], pq[1], cb), M.ProductQuantizer(64, pq[0], pq[1], cb)
        rows = opq.quantize(X)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
    else:
        rows, gq = X, M.NoQuantizer(64)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs)
    givf = M.BlockBasedIvf(cents, offsets, ids, rows, gq, doc_ids=docs)
    Q = (X[:6] + 0.01).astype(np.float32)
    od, os_, oc = oivf.search_batch(Q, k, 4)
    res = givf.search_batch(Q, k, 4)
    assert np.array_equal(np.asarray(res.counts, dtype=np.int64), oc.astype(np.int64)) and int(oc.min()) == k
    for b in range(len(Q)):
        assert np.array_equal(res.doc_ids[b], od[b]), b
        assert _same_f32(res.scores[b], os_[b]), b


def test_k_above_limit_is_rejected(M):
    X = synth.clustered(500, 32, n_blobs=3, seed=2)
    cents = O.kmeans(X, 4, iters=3, seed=1)
    offsets, ids = O.build_posting_lists(X, cents)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(32))
    with pytest.raises(M.Unsupported):
        givf.search_batch(X[:4], 5000, 2)


def test_micro_batcher_large_k(M):
    """The per-request front door accepts k > 32 as well (multi-round top-k behind it)."""
    X = synth.clustered(2500, 64, n_blobs=6, seed=41)
    cents = O.kmeans(X, 10, iters=4, seed=1)
    offsets, ids = O.build_posting_lists(X, cents)
    oivf = O.Ivf(cents, offsets, ids, X)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(64))
    mb = M.MicroBatcher(givf, k=40, num_probes=4, max_batch=8, max_wait_us=1000)
    Q = (X[:5] + 0.01).astype(np.float32)
    for i in range(len(Q)):
        r = mb.search(Q[i])
        od, os_, oc = oivf.search_batch(Q[i:i + 1], 40, 4)
        n = int(oc[0])
        assert [x.doc_id for x in r.id_with_scores] == [int(lo) | (int(hi) << 64) for lo, hi in od[0, :n]], i
        assert _same_f32(np.array([x.score for x in r.id_with_scores], dtype=np.float32), os_[0, :n])
    mb.close()

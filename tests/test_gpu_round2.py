"""Parity of the round-2 entry points (through the C ABI) against the CPU oracle: lane-conforming distances, the k-means
assignment step (exact and tensor-core paths), original_vector, the doc-id keyed accessors, the multi-user loader, the
sharded SPANN call, loader hardening and the coarse margin on adversarial data."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from tests import refformat as RF
from tests import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def M():
    import muopdb_b200 as M
    return M


def _same_f32(a, b):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- LaneConformingDistanceCalculator (row a4) ---------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,lanes", [(16, 4), (16, 16), (24, 8), (24, 4), (40, 8), (12, 4), (20, 4), (768, 16), (776, 8), (772, 4),
                                       (6, 2), (5, 1)])
def test_lane_conforming_bit_exact(M, dim, lanes):
    """lane_conforming.rs:16-27: LANES accumulators over the whole vector, ordered reduce, outermost_op -- bit-exact for L2
    and dot; differs from the generic 16/8/4 cascade whenever dim % 16 != 0 (that difference is asserted too)."""
    import torch
    rng = np.random.default_rng(dim * 17 + lanes)
    A = rng.random((29, dim), dtype=np.float32)
    B = rng.random((41, dim), dtype=np.float32)
    for metric, calc in ((O.L2, M.L2DistanceCalculator), (O.DOT, M.DotProductDistanceCalculator)):
        lc = M.LaneConformingDistanceCalculator(lanes, calc)
        want = O.lane_conforming_batch(A, B, lanes, metric)
        assert _same_f32(lc.calculate_squared_batch(A, B), want)
        got = lc.calculate_squared_batch(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
        M.default_context().sync()
        assert _same_f32(got.cpu().numpy(), want)
        assert lc.calculate_squared(A[0], B[0]) == want[0, 0]
    # the reference's own test: conforming == generic within 1e-5 (lane_conforming.rs:36-57)
    assert np.abs(O.lane_conforming_batch(A, B, lanes, O.L2) - O.distance_batch(A, B, O.L2, True)).max() < 1e-3 * dim
    with pytest.raises(M.InvalidArgument):
        M.LaneConformingDistanceCalculator(3).calculate_squared_batch(A, B)
    if dim % 16:
        with pytest.raises(M.InvalidArgument):
            M.LaneConformingDistanceCalculator(16).calculate_squared_batch(A, B)


def test_lane_conforming_differs_from_cascade_where_the_reference_does(M):
    """dim = 24: generic = 16-lane phase + 8-lane phase, LaneConforming<8> = one 8-lane phase over 24 dims.  Same value in
    exact arithmetic, different fp32 rounding -- the GPU reproduces each one bit for bit."""
    rng = np.random.default_rng(3)
    A = (rng.random((64, 24), dtype=np.float32) * 100).astype(np.float32)
    B = (rng.random((64, 24), dtype=np.float32) * 100).astype(np.float32)
    lc = M.LaneConformingDistanceCalculator(8).calculate_squared_batch(A, B)
    gen = M.L2DistanceCalculator.calculate_batch(A, B, squared=True)
    assert _same_f32(lc, O.lane_conforming_batch(A, B, 8, O.L2)) and _same_f32(gen, O.distance_batch(A, B, O.L2, True))
    assert (lc.view(np.uint32) != gen.view(np.uint32)).any()


# ---- k-means assignment (row a17) ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,nlist,n", [(16, 40, 700), (24, 33, 500), (20, 17, 300), (7, 9, 200), (64, 256, 3000), (128, 300, 5000),
                                         (768, 1024, 4000)])
@pytest.mark.parametrize("with_pen", [False, True])
def test_kmeans_assign_parity(M, dim, nlist, n, with_pen):
    """kmeans_builder.rs:199-221 with the calculator of :126-136 (LaneConforming<16|8|4> or D by dimension): labels AND
    winning costs bit-identical to the oracle, with duplicated centroids (first minimum wins), rows sitting exactly on
    centroids, near-ties, and size penalties.  dim >= 64 with nlist >= 256 runs the tensor-core estimate + exact band."""
    import torch
    rng = np.random.default_rng(dim * 13 + nlist)
    X = synth.clustered(n, dim, n_blobs=min(nlist, 24), seed=dim)
    C = X[rng.choice(n, nlist, replace=False)].copy()
    C[nlist // 2] = C[1]                                  # exact duplicate: the smaller index must win
    C[nlist // 3] = C[2] + np.float32(1e-6)               # near-tie
    X[:5] = C[[1, 2, nlist // 2, nlist // 3, 0]]
    pen = (rng.random(nlist, dtype=np.float32) * (0.05 * dim)).astype(np.float32) if with_pen else None
    want_l, want_c = O.kmeans_assign(X, C, pen, with_costs=True)
    got_l, got_c = M.kmeans_assign(X, C, pen, with_costs=True)
    assert np.array_equal(got_l, want_l)
    assert _same_f32(got_c, want_c)
    dl, dc = M.kmeans_assign(torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda(),
                             None if pen is None else torch.from_numpy(pen).cuda(), with_costs=True)
    M.default_context().sync()
    assert np.array_equal(dl.cpu().numpy().view(np.uint32), want_l) and _same_f32(dc.cpu().numpy(), want_c)


def test_kmeans_assign_dot_and_edge_cases(M):
    """D = DotProductDistanceCalculator (calculate_squared forwards to calculate, dot_product.rs:31-33), a NaN row (never
    below f32::MAX: the fold keeps (0, f32::MAX)), and an un-centred collection (common offset 100x the spread) on the
    tensor-core path: the band grows, the answer does not change."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((300, 24)).astype(np.float32)
    C = rng.standard_normal((21, 24)).astype(np.float32)
    wl, wc = O.kmeans_assign(X, C, None, metric=O.DOT, with_costs=True)
    gl, gc = M.kmeans_assign(X, C, None, distance=M.DotProductDistanceCalculator, with_costs=True)
    assert np.array_equal(gl, wl) and _same_f32(gc, wc)
    Xn = X.copy()
    Xn[7] = np.nan
    C16 = rng.standard_normal((21, 16)).astype(np.float32)
    Xn16 = rng.standard_normal((50, 16)).astype(np.float32)
    Xn16[7] = np.nan
    wl, wc = O.kmeans_assign(Xn16, C16, None, with_costs=True)
    gl, gc = M.kmeans_assign(Xn16, C16, None, with_costs=True)
    assert np.array_equal(gl, wl) and gl[7] == 0 and gc[7] == np.float32(3.40282347e+38) and _same_f32(gc, wc)
    Xo = (synth.clustered(3000, 128, n_blobs=16, seed=9) + np.float32(100.0)).astype(np.float32)
    Co = Xo[rng.choice(3000, 256, replace=False)].copy()
    wl, wc = O.kmeans_assign(Xo, Co, None, with_costs=True)
    gl, gc = M.kmeans_assign(Xo, Co, None, with_costs=True)
    assert np.array_equal(gl, wl) and _same_f32(gc, wc)


# ---- ProductQuantizer::original_vector -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,dsub,nbits", [(64, 8, 8), (30, 5, 3), (16, 1, 8)])
def test_pq_original_vector(M, dim, dsub, nbits):
    """pq/mod.rs:184-200: concatenation of the named codebook centroids."""
    rng = np.random.default_rng(dim)
    cb = rng.random(dim * (1 << nbits), dtype=np.float32)
    pq = M.ProductQuantizer(dim, dsub, nbits, cb)
    codes = rng.integers(0, 1 << nbits, (50, dim // dsub)).astype(np.uint8)
    c3 = cb.reshape(dim // dsub, 1 << nbits, dsub)
    want = np.stack([c3[np.arange(dim // dsub), codes[i]].reshape(-1) for i in range(50)])
    assert _same_f32(pq.original_vector(codes), want)
    assert _same_f32(pq.original_vector(codes[3]), want[3])


# ---- doc-id keyed accessors and invalidation (index.rs:350-471) ---------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["flat", "pq_fast", "pq_generic"])
def test_ivf_accessors(M, kind):
    """get_doc_id(s), get_point_id (a repeated doc id resolves to its LAST point, as the HashMap collect of index.rs:67-73),
    get_vector (rows read back out of the chunked HBM layouts), invalidate / invalidate_batch / is_invalidated by DOC id
    (index.rs:417-459: unknown ids are skipped, a second invalidation of the same document reports false)."""
    dim = {"flat": 20, "pq_fast": 256, "pq_generic": 40}[kind]
    X = synth.clustered(900, dim, n_blobs=6, seed=5)
    cents, offsets, ids = synth.build_ivf_arrays(X, 7, seed=3, max_clusters=2, threshold=0.5)
    docs = synth.doc_ids_for(len(X), seed=8)
    docs[17] = docs[3]                       # duplicated doc id: the map keeps point 17
    if kind == "flat":
        q, rows = M.NoQuantizer(dim), X
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs)
    else:
        dsub = 8 if kind == "pq_fast" else 5
        cb = O.train_pq_codebook(X[:600], dsub, 8, iters=2, seed=1)
        opq, q = O.ProductQuantizer(dim, dsub, 8, cb), M.ProductQuantizer(dim, dsub, 8, cb)
        rows = opq.quantize(X)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
    givf = M.BlockBasedIvf(cents, offsets, ids, rows, q, doc_ids=docs)
    as_int = lambda p: int(p[0]) | (int(p[1]) << 64)  # noqa: E731
    pids = [0, 5, 17, 3, 899, 450]
    assert givf.get_doc_ids(pids) == [as_int(docs[p]) for p in pids]
    assert givf.get_doc_id(42) == as_int(docs[42])
    with pytest.raises(M.OutOfRange):
        givf.get_doc_ids([5, 900])
    assert givf.get_point_id(as_int(docs[42])) == 42
    assert givf.get_point_id(as_int(docs[3])) == 17
    assert givf.get_point_id(12345678901234567890123) is None
    got = givf.get_vectors(pids)
    assert got.dtype == rows.dtype and np.array_equal(got.view(np.uint8), np.ascontiguousarray(rows[pids]).view(np.uint8))
    assert np.array_equal(givf.get_vector(450).view(np.uint8), np.ascontiguousarray(rows[450]).view(np.uint8))
    # invalidation by doc id
    d40, d41 = as_int(docs[40]), as_int(docs[41])
    assert givf.invalidate_batch([d40, 999999999999, d41, d40]) == [d40, d41]
    assert givf.is_invalidated(d40) and not givf.is_invalidated(as_int(docs[42])) and not givf.is_invalidated(999999999999)
    assert givf.invalidate(d40) is False and givf.invalidate(as_int(docs[42])) is True
    assert givf.is_point_invalidated(40) and givf.is_point_invalidated(42)
    oivf.invalidate_batch([40, 41, 42])
    Q = (X[38:46] + 0.001).astype(np.float32)
    od, os_, oc = oivf.search_batch(Q, 10, 7)
    r = givf.search_batch(Q, 10, 7)
    assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(len(Q)):
        assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]) and _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]])


# ---- multi-user files ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_pq", [False, True])
def test_multi_user_spann_loader(M, tmp_path, use_pq):
    """MultiSpannIndex::get_or_create_index (multi_spann/index.rs:100-128): three users packed into the shared files by
    MultiSpannWriter::write's layout (multi_spann/writer.rs:75-260; tests/refformat.py), opened through the user_index_info
    table at their byte offsets -- every user's searches equal the oracle on that user's arrays; unknown user -> None."""
    dim = 32
    users, oracles = {}, {}
    cbs = {}
    for uid, n, nlist, seed in ((7, 500, 6, 1), ((1 << 90) + 3, 300, 4, 2), (1000, 800, 9, 3)):
        X = synth.clustered(n, dim, n_blobs=5, seed=seed)
        cents = O.kmeans(X, nlist, iters=8, seed=seed)
        offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
        g = O.hnsw_build(cents, 8, 2, 50, seed=seed)
        docs = [int(d) for d in (np.arange(n) * 11 + uid % 1000)]
        if use_pq:
            cbs[uid] = O.train_pq_codebook(X[:250], 4, 6, iters=3, seed=seed)
            opq = O.ProductQuantizer(dim, 4, 6, cbs[uid])
            rows = opq.quantize(X)
            oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
        else:
            rows = X
            oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs)
        users[uid] = dict(hnsw=(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"]), centroids=cents,
                          offsets=offsets, ids=ids, rows=rows, doc_ids=docs)
        ohn = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents)
        oracles[uid] = (O.Spann(ohn, oivf), X)
    base = str(tmp_path / "collection")
    infos = RF.write_multi_spann_dir(base, users, pq=(dim, 4, 6, cbs) if use_pq else None)
    assert any(i["ivf_index_offset"] > 0 for i in infos) and any(i["centroid_index_offset"] > 0 for i in infos)
    idx = M.MultiSpannIndex(base, dim, M.QUANT_PQ if use_pq else M.QUANT_NONE)
    assert idx.user_ids() == sorted(users)
    for uid, (osp, X) in oracles.items():
        Q = (X[:20] + 0.01).astype(np.float32)
        od, os_, oc = osp.search_batch(Q, 5, 30, 3, 0.5)
        sp = idx.get_or_create_index(uid)
        assert idx.get_or_create_index(uid) is sp
        r = sp.search_batch(Q, M.SearchParams(5, 30, False, 3, 0.5))
        assert np.array_equal(np.asarray(r.counts).astype(np.int32), oc)
        for b in range(len(Q)):
            n = max(int(oc[b]), 0)
            assert np.array_equal(r.doc_ids[b, :n], od[b, :n]), (uid, b)
            assert _same_f32(r.scores[b, :n], os_[b, :n])
        one = idx.search_for_user(uid, Q[0], M.SearchParams(5, 30, False, 3, 0.5))
        assert [x.doc_id for x in one.id_with_scores] == [int(lo) | (int(hi) << 64) for lo, hi in od[0, :max(int(oc[0]), 0)]]
    assert idx.search_for_user(424242, np.zeros(dim, dtype=np.float32), M.SearchParams(5, 30)) is None
    with pytest.raises(M.InvalidArgument):
        idx.get_or_create_index(424242)


# ---- loader hardening ---------------------------------------------------------------------------------------------------
def test_loaders_reject_truncated_and_corrupt_files(M, tmp_path):
    """Header fields size allocations and copies: every section must be bounded by the file (overflow-safe) and a bad file
    must come back as MGPU_ERR_INVALID_ARG, not as an abort or an out-of-bounds read."""
    X = synth.clustered(300, 16, n_blobs=4, seed=2)
    cents, offsets, ids = synth.build_ivf_arrays(X, 5, seed=3)
    docs = list(range(300))
    base = str(tmp_path / "ivf")
    RF.write_ivf_dir(base, cents, offsets, ids, X, docs)
    good = open(os.path.join(base, "index"), "rb").read()
    M.BlockBasedIvf.new(base, M.NoQuantizer(16))   # the intact file loads

    def load_with(index_bytes):
        open(os.path.join(base, "index"), "wb").write(index_bytes)
        with pytest.raises(M.InvalidArgument):
            M.BlockBasedIvf.new(base, M.NoQuantizer(16))

    for cut in (10, 44, 47, 200, len(good) // 2, len(good) - 9):
        load_with(good[:cut])
    for off, val in ((13, 1 << 62), (21, (1 << 64) - 8), (29, (1 << 64) - 1), (9, 0xFFFFFFFF)):   # num_vectors, doc_len, cent_len, nlist
        b = bytearray(good)
        struct.pack_into("<Q" if off != 9 else "<I", b, off, val)
        load_with(bytes(b))
    # a posting list that claims 2^60 ids / points beyond num_vectors
    doc_off = 48
    cent_off = (doc_off + 16 * 301 + 7) & ~7
    meta_off = (cent_off + 8 + 5 * 16 * 4 + 7) & ~7
    start = meta_off + 8 + 16 * 5
    b = bytearray(good)
    struct.pack_into("<Q", b, start, 1 << 60)
    load_with(bytes(b))
    b = bytearray(good)
    struct.pack_into("<Q", b, 13, 100)            # num_vectors 100 < ids stored in the lists
    load_with(bytes(b))
    # HNSW: truncated file, edge offsets beyond the edge array
    g = O.hnsw_build(cents, 4, 2, 20, seed=1)
    hb = str(tmp_path / "h")
    RF.write_hnsw_dir(hb, g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents, list(range(5)))
    M.BlockBasedHnsw.new(hb, M.NoQuantizer(16))
    hgood = open(os.path.join(hb, "hnsw", "index"), "rb").read()
    for cut in (20, 48, 60, len(hgood) - 17):
        open(os.path.join(hb, "hnsw", "index"), "wb").write(hgood[:cut])
        with pytest.raises(M.InvalidArgument):
            M.BlockBasedHnsw.new(hb, M.NoQuantizer(16))
    for off, val in ((9, (1 << 64) - 4), (25, (1 << 63)), (17, 1 << 61)):
        b = bytearray(hgood)
        struct.pack_into("<Q", b, off, val)
        open(os.path.join(hb, "hnsw", "index"), "wb").write(bytes(b))
        with pytest.raises(M.InvalidArgument):
            M.BlockBasedHnsw.new(hb, M.NoQuantizer(16))
    eo = np.array(g["edge_offsets"], dtype=np.uint64)
    eo[-1] = len(g["edges"]) + 1000
    with pytest.raises(M.InvalidArgument):
        M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], eo, g["level_offsets"], cents, M.NoQuantizer(16))


# ---- coarse margin on adversarial data -----------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["offset100", "mixed_magnitudes", "tiny_spread"])
def test_coarse_tensor_core_margin_adversarial(M, case):
    """The tensor-core coarse pass keeps every centroid within tau + 2 eps, eps = 3e-4 (|q|^2 + max|c|^2).  On un-centred data
    (|x|^2 >> distances) that band can grow to all centroids -- slow but still exact: probes and distances must equal the
    oracle's bit for bit."""
    rng = np.random.default_rng(3)
    dim, nlist, n = 128, 512, 6000
    X = synth.clustered(n, dim, n_blobs=40, seed=12)
    if case == "offset100":
        X = (X + np.float32(100.0) * np.float32(X.std())).astype(np.float32)     # common offset 100x the spread
    elif case == "mixed_magnitudes":
        X = (X * (10.0 ** rng.integers(-2, 3, (n, 1)))).astype(np.float32)        # rows spanning 4 orders of magnitude
    else:
        X = (np.float32(50.0) + np.float32(1e-3) * X).astype(np.float32)         # nearly identical rows
    cents = X[rng.choice(n, nlist, replace=False)].copy()
    offsets, ids = O.build_posting_lists(X, cents)
    oivf = O.Ivf(cents, offsets, ids, X)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(dim))
    Q = (X[rng.integers(0, n, 48)] * np.float32(1.0001)).astype(np.float32)
    givf.ctx.coarse_band_stats(reset=True)
    gp, gd = givf.find_nearest_centroids_batch(Q, 16, with_distances=True)
    band, nq = givf.ctx.coarse_band_stats()
    # the size of the uncertain band is the cost of the margin rule: recorded (pytest -s / -rP) and bounded by C per query
    print(f"coarse band [{case}]: {band / max(nq, 1):.1f} of {nlist} centroids re-scored per query ({nq} selections)")
    assert nq >= len(Q) and band <= nq * nlist
    for b in range(len(Q)):
        op, od = oivf.find_nearest_centroids(Q[b], 16, with_dist=True)
        assert _same_f32(od, gd[b]), (case, b)
        # equal distances may be listed in either order by the reference (select_nth_unstable); ours is by index, like the oracle
        assert np.array_equal(op, gp[b]), (case, b)
    od, os_, oc = oivf.search_batch(Q, 10, 16)
    r = givf.search_batch(Q, 10, 16)
    for b in range(len(Q)):
        assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]) and _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]])


# ---- sharded SPANN (config 5) ---------------------------------------------------------------------------------------------------------
def test_shard_spann_single_rank_world(M):
    """mgpu_shard_spann_search on a 1-rank NCCL world equals Spann::search (incl. None answers), host / device / pipelined;
    multi-rank equivalence: tools/check_shard_search.py under torchrun (test_shard_search_two_ranks)."""
    import torch
    ctx = M.Context(0)
    ctx.comm_init(1, 0, M.Context.comm_unique_id())
    X = synth.clustered(3000, 64, n_blobs=10, seed=4)
    docs = synth.doc_ids_for(len(X), seed=4)
    cents = O.kmeans(X, 24, iters=10, seed=2)
    offsets, ids = O.build_posting_lists(X, cents, 1, 0.1)
    g = O.hnsw_build(cents, 8, 2, 60, seed=3)
    cb = O.train_pq_codebook(X[:1500], 8, 8, iters=3, seed=5)
    opq, gpq = O.ProductQuantizer(64, 8, 8, cb), M.ProductQuantizer(64, 8, 8, cb, ctx=ctx)
    rows = opq.quantize(X)
    osp = O.Spann(O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents),
                  O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq))
    gsp = M.Spann(M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents,
                                   M.NoQuantizer(64), ctx=ctx),
                  M.BlockBasedIvf(cents, offsets, ids, rows, gpq, doc_ids=docs, ctx=ctx))
    Q = (X[100:177] + 0.01).astype(np.float32)
    params = M.SearchParams(10, 40, False, 6, 0.3)
    od, os_, oc = osp.search_batch(Q, 10, 40, 6, 0.3)

    def check(ids_, sc, cn):
        assert np.array_equal(np.asarray(cn).astype(np.int32), oc)
        for b in range(len(Q)):
            n = max(int(oc[b]), 0)
            assert np.array_equal(np.asarray(ids_)[b, :n].view(np.uint64).reshape(-1, 2), np.asarray(od[b, :n], dtype=np.uint64)), b
            assert _same_f32(np.asarray(sc)[b, :n], os_[b, :n])

    r = gsp.shard_search_batch(Q, params)
    check(r.doc_ids, r.scores, r.counts)
    r = gsp.shard_search_batch(torch.from_numpy(Q).cuda(), params)
    ctx.sync()
    check(r.doc_ids.cpu().numpy(), r.scores.cpu().numpy(), r.counts.cpu().numpy())
    ctx.shard_overlap(True)
    rs = [gsp.shard_search_batch(torch.from_numpy(Q).cuda(), params) for _ in range(3)]
    ctx.sync()
    ctx.shard_overlap(False)
    for r in rs:
        check(r.doc_ids.cpu().numpy(), r.scores.cpu().numpy(), r.counts.cpu().numpy())
    Qp = torch.from_numpy(Q).pin_memory()
    outs = [(torch.zeros((len(Q), 10, 2), dtype=torch.int64).pin_memory(), torch.zeros((len(Q), 10), dtype=torch.float32).pin_memory(),
             torch.zeros((len(Q),), dtype=torch.int32).pin_memory()) for _ in range(3)]
    tks = [gsp.shard_search_batch_submit(Qp, params, o) for o in outs]
    for t in tks:
        gsp.search_wait(t)
    for o in outs:
        check(o[0].numpy(), o[1].numpy(), o[2].numpy())


def test_shard_search_two_ranks():
    """The sharded calls on a real 2-rank NCCL world (one process per GPU, torchrun): IVF and SPANN, host / device /
    pipelined / overlapped, with and without the split query encode, bit for bit against per-shard oracle searches merged by
    the oracle (snapshot.rs:49-63).  Skipped on a 1-GPU box; tools/check_shard_search.py is the same check for any N."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "check_shard_search.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "shard search check: OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


# ---- 16-bit-LUT scan: certification and the exact fallback ------------------------------------------------------------------------
def test_scan16_uncertified_queries_take_the_exact_fallback():
    """scan_pq16.cu ranks with 16-bit LUT entries; finalize.cu certifies each answer (k-th exact candidate's key + 2 E < 32nd key)
    and sends the rest to the exact scan (scan.cu k_scan_pq_exact: every row scored with ProductQuantizer::distance itself).
    MGPU_CERT_SLACK=1e30 makes E infinite, so EVERY query with a full candidate list goes through the fallback: the PQ parity
    tests (ties, duplicates, invalidation, planner filters, SPANN, device buffers) must still be bit-exact.  MGPU_SCAN16=0 runs
    the same tests on the previous 32-bit-LUT kernel (scan_pq.cu), which stays in the library for k > 16 and dot-product PQ."""
    sel = "test_ivf_pq_parity or test_ivf_pq_exact_ties or test_ivf_invalidation or test_ivf_planner_filter_parity or " \
          "test_spann_parity or test_ivf_device_buffers or test_ivf_pq_heavy"
    for env_add in ({"MGPU_CERT_SLACK": "1e30"}, {"MGPU_SCAN16": "0"}):
        env = dict(os.environ, **env_add)
        p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu",
                            "-k", sel, "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
        assert p.returncode == 0, (env_add, p.stdout[-3000:], p.stderr[-2000:])


def test_scan16_near_ties_are_certified_or_fall_back(M):
    """Rows whose exact scores differ by less than the 16-bit key resolution around the k-th place: thousands of near-duplicates
    of the query's neighbourhood (codes differing in one subspace by the closest pair of centroids).  Whatever the certificate
    decides, ids, order and score bits must equal the oracle's."""
    rng = np.random.default_rng(5)
    dim, n = 256, 6000
    base = synth.clustered(40, dim, n_blobs=4, seed=3)
    X = np.repeat(base, n // 40, axis=0).astype(np.float32)
    X += (1e-4 * rng.standard_normal(X.shape)).astype(np.float32)      # near-duplicates: most rows share their code word
    cents, offsets, ids = synth.build_ivf_arrays(X, 6, seed=2)
    docs = synth.doc_ids_for(n, seed=6)
    cb = O.train_pq_codebook(X[rng.choice(n, 2000, replace=False)], 8, 8, iters=4, seed=1)
    opq, gpq = O.ProductQuantizer(dim, 8, 8, cb), M.ProductQuantizer(dim, 8, 8, cb)
    rows = opq.quantize(X)
    oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
    givf = M.BlockBasedIvf(cents, offsets, ids, rows, gpq, doc_ids=docs)
    Q = (base[:24] + 1e-4).astype(np.float32)
    for k in (1, 10, 16):
        od, os_, oc = oivf.search_batch(Q, k, 6)
        r = givf.search_batch(Q, k, 6)
        assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
        for b in range(len(Q)):
            assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]), (k, b)
            assert _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]])


def test_micro_batcher_pipelined_batches_match_oracle(M):
    """Unfiltered IVF batches go through mgpu_ivf_search_submit / mgpu_search_wait with three staging buffers (two batches on
    the GPU while a third fills, batcher.cu).  Many callers with a small max_batch keep several batches in flight; every caller
    must still get the oracle's answer for ITS query, and the stats must show real batching."""
    import threading
    X = synth.clustered(6000, 64, n_blobs=12, seed=77)
    cents = O.kmeans(X, 24, iters=4, seed=2)
    offsets, ids = O.build_posting_lists(X, cents)
    docs = synth.doc_ids_for(len(X), seed=9)
    oivf = O.Ivf(cents, offsets, ids, X, doc_ids=docs)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(64), doc_ids=docs)
    nthreads, per_thread = 48, 10
    Q = (X[:nthreads * per_thread] * np.float32(1.001) + 0.003).astype(np.float32)
    mb = M.MicroBatcher(givf, k=7, num_probes=5, max_batch=8, max_wait_us=200)
    got, errs = [None] * len(Q), []

    def worker(t):
        try:
            for j in range(per_thread):
                i = t * per_thread + j
                got[i] = mb.search(Q[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    st = mb.stats()
    mb.close()
    assert st["queries"] == len(Q) and st["batches"] < len(Q) and st["largest_batch"] > 1
    od, os_, oc = oivf.search_batch(Q, 7, 5)
    for i in range(len(Q)):
        n = int(oc[i])
        assert [x.doc_id for x in got[i].id_with_scores] == [int(lo) | (int(hi) << 64) for lo, hi in od[i, :n]], i
        assert _same_f32(np.array([x.score for x in got[i].id_with_scores], dtype=np.float32), os_[i, :n])


def test_coarse_band_is_small_on_centred_data(M):
    """On centred, well-spread data the tensor-core selection re-scores only a few centroids per query (the margin rule's
    normal regime); the counter is what exposes the un-centred cliff of the test above."""
    rng = np.random.default_rng(8)
    dim, nlist = 128, 1024
    cents = rng.standard_normal((nlist, dim)).astype(np.float32)
    X = (cents[rng.integers(0, nlist, 4000)] + 0.1 * rng.standard_normal((4000, dim))).astype(np.float32)
    offsets, ids = O.build_posting_lists(X, cents)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(dim))
    givf.ctx.coarse_band_stats(reset=True)
    givf.find_nearest_centroids_batch(X[:64], 16)
    band, nq = givf.ctx.coarse_band_stats()
    assert nq >= 64 and band / nq < 64, (band, nq)


def test_last_kernel_reports_the_kernel_that_ran(M):
    """mgpu_last_kernel: bench.py quotes it as roofline.kernel (the name must be the timed kernel's)."""
    from muopdb_b200 import _lib
    X = synth.clustered(3000, 64, n_blobs=6, seed=5)
    cents = O.kmeans(X, 8, iters=3, seed=1)
    offsets, ids = O.build_posting_lists(X, cents)
    givf = M.BlockBasedIvf(cents, offsets, ids, X, M.NoQuantizer(64))
    givf.search_batch(X[:4], 3, 2)
    assert givf.ctx.last_kernel(_lib.K_SCAN).startswith("k_scan<flat>")
    assert givf.ctx.last_kernel(_lib.K_HNSW) == "" or givf.ctx.last_kernel(_lib.K_HNSW).startswith("k_hnsw")   # shared default context


@pytest.mark.parametrize("B,ef,n,dim,k", [(200, 64, 3000, 32, 10),    # 148 < B < 296: 8 warps, 4 candidates per batch
                                          (320, 128, 4000, 64, 64),   # B >= 296 on an L2-resident graph: 2 warps, 2 candidates
                                          (320, 300, 2500, 16, 100),  # ... whose CTA merge holds 128 entries: warp-level merges
                                          (200, 300, 2500, 16, 32),   # 8 warps, ef beyond the CTA merge's 256 entries
                                          (40, 600, 1500, 8, 10),     # <= one query per SM: 16 warps, 8 candidates; large ef
                                          (300, 1, 600, 4, 1)])       # ef = 1
def test_hnsw_batch_kernel_shapes(M, B, ef, n, dim, k):
    """Every shape of k_hnsw_spec (hnsw_spec.cu: chosen by batch size and graph size) and both merge paths must give the
    oracle's ids, scores and traversal counts (hnsw/block_based/index.rs:159-298)."""
    from muopdb_b200 import _lib
    X = synth.clustered(n, dim, n_blobs=9, seed=dim + ef)
    g = O.hnsw_build(X, 16, 5, 60, seed=ef)
    docs = synth.doc_ids_for(n, seed=2)
    oh = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X, doc_ids=docs)
    gh = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X,
                          M.NoQuantizer(dim), doc_ids=docs)
    rng = np.random.default_rng(ef)
    Q = np.vstack([X[rng.integers(0, n, B // 2)] + 0.02 * rng.standard_normal((B // 2, dim)).astype(np.float32),
                   rng.random((B - B // 2, dim), dtype=np.float32)]).astype(np.float32)
    od, os_, oc, ost = oh.search_batch(Q, k, ef)
    r, st = gh.ann_search_batch(Q, k, ef, with_stats=True)
    assert gh.ctx.last_kernel(_lib.K_HNSW).startswith("k_hnsw_spec"), gh.ctx.last_kernel(_lib.K_HNSW)
    assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(B):
        assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]), b
        assert _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]]), b
    assert np.array_equal(st, ost)


def test_hnsw_hash_visited_set_and_overflow_redo():
    """Graphs above 262 144 points keep the visited set of a query in an 8192-slot hash table in shared memory (config 4: 1M
    points), and a query that visits more than 6144 points is flagged and redone by the generic kernel.  MGPU_HNSW_BITMAP_MAX=0
    forces that mode on the small graphs of the HNSW / SPANN parity tests (ef up to 600 on 1500..4000 points: both the
    in-table case and the redo), which must stay bit-exact, traversal counts included."""
    env = dict(os.environ, MGPU_HNSW_BITMAP_MAX="0")
    sel = "test_hnsw_flat_parity or test_hnsw_batch_kernel_shapes or test_spann_parity or test_hnsw_big_visited"
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-x", "-q", "-m", "gpu", "-k", sel,
                        "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-2000:])


def test_hnsw_big_visited(M):
    """ef = 2048 on 12 000 points: every query visits most of the graph (more than the hash table holds when the hash mode is
    forced: the redo path), W and C exceed the CTA merge's capacity (warp-level merges)."""
    n, dim = 12000, 24
    X = synth.clustered(n, dim, n_blobs=7, seed=4)
    g = O.hnsw_build(X, 16, 4, 40, seed=4)
    oh = O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X)
    gh = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X, M.NoQuantizer(dim))
    Q = (X[:12] + 0.01).astype(np.float32)
    od, os_, oc, ost = oh.search_batch(Q, 100, 2048)
    r, st = gh.ann_search_batch(Q, 100, 2048, with_stats=True)
    assert np.array_equal(np.asarray(r.counts, dtype=np.int64), oc.astype(np.int64))
    for b in range(len(Q)):
        assert np.array_equal(r.doc_ids[b, :oc[b]], od[b, :oc[b]]), b
        assert _same_f32(r.scores[b, :oc[b]], os_[b, :oc[b]]), b
    assert np.array_equal(st, ost)


def test_hnsw_and_spann_submit_wait_match_blocking(M):
    """mgpu_hnsw_search_submit / mgpu_spann_search_submit + mgpu_search_wait: two batches in flight through the staging slots
    must return what the blocking host call returns (and therefore the oracle's answer, which the blocking call is tested
    against), including a slot reused before its ticket was waited for."""
    import torch
    from tests.test_gpu_parity import _spann_pair
    X = synth.clustered(5000, 64, n_blobs=12, seed=13)
    docs = synth.doc_ids_for(len(X), seed=4)
    _, gs, _, givf = _spann_pair(M, X, docs, 40, pq_params=(8, 8, 1500))
    g = O.hnsw_build(X, 16, 4, 60, seed=5)
    gh = M.BlockBasedHnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], X, M.NoQuantizer(64), doc_ids=docs)
    rng = np.random.default_rng(9)
    k, B, nb = 8, 80, 5
    params = M.SearchParams(k, 40, False, 6, 0.3)
    Qs = [torch.from_numpy((X[rng.integers(0, len(X), B)] + 0.02 * rng.standard_normal((B, 64))).astype(np.float32)).pin_memory()
          for _ in range(nb)]

    def outs():
        return [(torch.zeros((B, k, 2), dtype=torch.int64).pin_memory(), torch.zeros((B, k), dtype=torch.float32).pin_memory(),
                 torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(nb)]

    for name, submit, wait, blocking in (
            ("hnsw", lambda q, o: gh.ann_search_batch_submit(q, k, 48, o), gh.search_wait, lambda q: gh.ann_search_batch(q.numpy(), k, 48)),
            ("spann", lambda q, o: gs.search_batch_submit(q, params, o), gs.search_wait, lambda q: gs.search_batch(q.numpy(), params))):
        o = outs()
        tickets = []
        for i in range(nb):
            tickets.append(submit(Qs[i], o[i]))
            if i >= 1 and i != 3:
                wait(tickets[i - 1])
        for t in reversed(tickets):
            wait(t)
        for i in range(nb):
            r = blocking(Qs[i])
            ids, sc, cn = (x.numpy() for x in o[i])
            rc = np.asarray(r.counts, dtype=np.int64)
            assert np.array_equal(cn.astype(np.int64) & 0xFFFFFFFF, rc & 0xFFFFFFFF), (name, i)
            for b in range(B):
                n = int(rc[b]) if 0 <= int(rc[b]) <= k else 0
                assert np.array_equal(ids[b, :n].view(np.uint64).reshape(-1), np.asarray(r.doc_ids[b, :n], dtype=np.uint64).reshape(-1)), (name, i, b)
                assert _same_f32(sc[b, :n], np.asarray(r.scores[b, :n], dtype=np.float32)), (name, i, b)

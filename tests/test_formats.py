"""Readers of the reference's on-disk formats: Elias-Fano decode (host only, runs on CPU) and
the IVF / HNSW / PQ loaders (GPU).  Golden bytes come from the reference's own tests and from an index file the reference
itself wrote (tests/golden/ref_hnsw_sample, copied from rs/index_writer/test_output/hnsw)."""
import os
import struct

import numpy as np
import pytest

import oracle as O
from tests import refformat as RF
from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- Elias-Fano (CPU) -----------------------------------------------------------------------------------------------------
def test_elias_fano_reference_golden_bits():
    """rs/compression/src/elias_fano/ef.rs:230-260: values [5,8,8,15,32], universe 36 -> L = 2 and these exact bit vectors."""
    L, lower, upper = RF.ef_bits([5, 8, 8, 15, 32], 36)
    assert L == 2
    assert lower == [1, 0, 0, 0, 0, 0, 1, 1, 0, 0]
    assert upper == [0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0, 1]


def test_elias_fano_decode_golden_payload():
    """The payload EliasFano::write (ef.rs:197-215) produces for the golden case, decoded by the C ABI."""
    import muopdb_b200 as M
    payload = struct.pack("<4Q", 5, 2, 1, 1) + struct.pack("<Q", 0b0011000001) + struct.pack("<Q", 0b1000001011010)
    assert payload == RF.ef_encode([5, 8, 8, 15, 32], 36)
    assert M.elias_fano_decode(payload).tolist() == [5, 8, 8, 15, 32]


@pytest.mark.parametrize("n,universe_bits", [(1, 3), (7, 8), (64, 10), (1000, 20), (300, 40), (50, 63), (129, 7)])
def test_elias_fano_round_trips(n, universe_bits):
    """Lower bits straddling 64-bit words, L = 0, duplicates, huge gaps (ef.rs:129-183 <-> block_based_decoder.rs:162-266)."""
    import muopdb_b200 as M
    rng = np.random.default_rng(n * 131 + universe_bits)
    vals = np.sort(rng.integers(0, (1 << universe_bits) - 1, n, dtype=np.uint64))
    vals[n // 2:] = np.maximum(vals[n // 2:], vals[n // 2])  # keep sorted, allow duplicates
    assert M.elias_fano_decode(RF.ef_encode(vals.tolist())).tolist() == vals.tolist()
    with pytest.raises(M.InvalidArgument):
        M.elias_fano_decode(RF.ef_encode(vals.tolist())[:-8] if n > 2 else b"\0" * 8)
    assert M.elias_fano_decode(RF.ef_encode([])).tolist() == []


# ---- loaders (GPU) ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_hnsw_loader_parses_reference_written_file():
    """An `hnsw/index` written by the reference (legacy 8-byte doc ids): header 49 B, pad->4, edges, points, pad->8,
    edge_offsets, level_offsets (graph_storage.rs:122-193).  Parsed sections must equal an independent numpy parse."""
    import muopdb_b200 as M
    base = os.path.join(HERE, "golden", "ref_hnsw_sample")
    b = open(os.path.join(base, "hnsw", "index"), "rb").read()
    qd, nl = struct.unpack_from("<II", b, 1)
    el, pl, eol, lol, dl = struct.unpack_from("<QQQQQ", b, 9)
    assert (qd, nl, el // 4, pl // 4, eol // 8, lol // 8) == (5, 2, 817, 17, 118, 3)
    off = 52
    edges = np.frombuffer(b, "<u4", el // 4, off); off += el
    points = np.frombuffer(b, "<u4", pl // 4, off); off += pl
    off += (8 - off % 8) % 8
    eo = np.frombuffer(b, "<u8", eol // 8, off); off += eol
    lo = np.frombuffer(b, "<u8", lol // 8, off)
    # the sample has no quantizer directory: supply a stand-in 5-subspace codebook so the codes can be searched
    cb = np.random.default_rng(0).random(5 * 256 * 2, dtype=np.float32)
    pq = M.ProductQuantizer(10, 2, 8, cb)
    hn = M.BlockBasedHnsw.new(base, pq)
    g = hn.graph_arrays()
    assert g["num_layers"] == 2 and g["n"] == 100 and g["entry_point"] == points[0] == 98
    assert np.array_equal(g["edges"], edges) and np.array_equal(g["points"], points)
    assert np.array_equal(g["edge_offsets"], eo) and np.array_equal(g["level_offsets"], lo)
    # and the loaded index searches exactly like the oracle on the same arrays
    codes = np.frombuffer(open(os.path.join(base, "hnsw", "vector_storage"), "rb").read(), np.uint8, 500, 8).reshape(100, 5)
    docs = np.frombuffer(b, "<u8", 100, len(b) - 800).astype(np.uint64)
    oh = O.Hnsw(2, edges, points, eo, lo, codes, doc_ids=[int(d) for d in docs], pq=O.ProductQuantizer(10, 2, 8, cb))
    Q = np.random.default_rng(1).random((12, 10), dtype=np.float32)
    od, os_, oc, _ = oh.search_batch(Q, 5, 20)
    r = hn.ann_search_batch(Q, 5, 20)
    assert np.array_equal(r.doc_ids, od) and np.array_equal(r.scores.view(np.uint32), os_.view(np.uint32))


@pytest.mark.gpu
def test_hnsw_loader_reference_hand_built_file(tmp_path):
    """graph_storage.rs:594-674: the reference's hand-built single-layer file (edges [1,2,3], 4 edge offsets, doc ids [1,2])."""
    import muopdb_b200 as M
    data = bytearray(b"\0" + struct.pack("<II", 16, 1) + struct.pack("<QQQQQ", 12, 0, 32, 16, 32))
    data += b"\0" * ((4 - len(data) % 4) % 4)
    data += struct.pack("<III", 1, 2, 3)
    data += b"\0" * ((8 - len(data) % 8) % 8)
    data += struct.pack("<QQQQ", 0, 1, 1, 2) + struct.pack("<QQ", 0, 0)
    data += struct.pack("<QQQQ", 1, 0, 2, 0)
    os.makedirs(tmp_path / "hnsw")
    (tmp_path / "hnsw" / "index").write_bytes(bytes(data))
    # NB: level_offsets [0, 0] describe zero layer-0 points in that synthetic file; the loader must reject the mismatch
    RF.write_vector_file(str(tmp_path / "hnsw" / "vector_storage"), np.zeros((2, 16), dtype=np.float32))
    with pytest.raises(M.InvalidArgument):
        M.BlockBasedHnsw.new(str(tmp_path), M.NoQuantizer(16))


@pytest.mark.gpu
@pytest.mark.parametrize("pq_params,prefix", [(None, 0), ((8, 8), 0), ((8, 8), 4096), (None, 48)])
def test_ivf_loader_round_trip(tmp_path, pq_params, prefix):
    """Files laid out like IvfWriter's (ivf/writer.rs:228-353: 45-byte header, pad->16, u128 doc ids, centroids, pad->8,
    posting-list table, Elias-Fano payloads) load into an index that searches exactly like the oracle; a non-zero byte
    offset exercises new_with_offset (multi-user packing, storage.rs:52-90)."""
    import muopdb_b200 as M
    dim = 64
    X = synth.clustered(1500, dim, n_blobs=7, seed=3)
    cents, offsets, ids = synth.build_ivf_arrays(X, 12, seed=4, max_clusters=2)
    docs = synth.doc_ids_for(len(X), seed=5)
    doc_ints = [int(lo) | (int(hi) << 64) for lo, hi in docs]
    if pq_params:
        cb = O.train_pq_codebook(X[:800], pq_params[0], pq_params[1], iters=3, seed=1)
        opq = O.ProductQuantizer(dim, pq_params[0], pq_params[1], cb)
        rows = opq.quantize(X)
        RF.write_pq_dir(str(tmp_path / "quantizer"), dim, pq_params[0], pq_params[1], cb)
        gq = M.ProductQuantizer.read(str(tmp_path / "quantizer"))
        assert (gq.dimension, gq.subvector_dimension, gq.num_bits) == (dim, pq_params[0], pq_params[1])
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs, pq=opq)
    else:
        rows, gq = X, M.NoQuantizer(dim)
        oivf = O.Ivf(cents, offsets, ids, rows, doc_ids=docs)
    RF.write_ivf_dir(str(tmp_path), cents, offsets, ids, rows, doc_ints, prefix_bytes=prefix)
    givf = M.BlockBasedIvf.new(str(tmp_path), gq, index_offset=prefix, vector_offset=prefix)
    assert givf.num_clusters() == 12 and givf.num_vectors() == 1500
    Q = X[:40] + 0.01
    od, os_, oc = oivf.search_batch(Q, 10, 5)
    r = givf.search_batch(Q, 10, 5)
    assert np.array_equal(r.doc_ids, od) and np.array_equal(r.scores.view(np.uint32), os_.view(np.uint32))
    with pytest.raises(M.InvalidArgument):
        M.BlockBasedIvf.new(str(tmp_path / "nope"), gq)


@pytest.mark.gpu
def test_spann_directory_round_trip(tmp_path):
    """SpannReader layout (spann/reader.rs:75-76): {base}/centroids = HNSW over the centroids, {base}/ivf = posting lists."""
    import muopdb_b200 as M
    dim = 32
    X = synth.clustered(2000, dim, n_blobs=9, seed=8)
    cents, offsets, ids = synth.build_ivf_arrays(X, 20, seed=9)
    docs = list(range(100, 2100))
    g = O.hnsw_build(cents, 8, 2, 50, seed=2)
    RF.write_ivf_dir(str(tmp_path / "ivf"), cents, offsets, ids, X, docs)
    RF.write_hnsw_dir(str(tmp_path / "centroids"), g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"],
                      cents, list(range(20)))
    sp = M.Spann.read(str(tmp_path), M.NoQuantizer(dim))
    osp = O.Spann(O.Hnsw(g["num_layers"], g["edges"], g["points"], g["edge_offsets"], g["level_offsets"], cents),
                  O.Ivf(cents, offsets, ids, X, doc_ids=docs))
    Q = X[:30] + 0.02
    od, os_, oc = osp.search_batch(Q, 10, 30, 6, 0.5)
    r = sp.search_batch(Q, M.SearchParams(10, 30, False, 6, 0.5))
    assert np.array_equal(np.asarray(r.counts).astype(np.int32), oc)
    for b in range(len(Q)):
        n = max(int(oc[b]), 0)
        assert np.array_equal(r.doc_ids[b, :n], od[b, :n]) and np.array_equal(r.scores[b, :n].view(np.uint32), os_[b, :n].view(np.uint32))

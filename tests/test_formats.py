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


# ---- IVF `index` file layout (CPU) ------------------------------------------------------------------------------------------
def test_ivf_combine_files_reference_golden_bytes():
    """rs/index/src/ivf/writer.rs:383-487 (test_combine_files) restated byte for byte: header fields little-endian in the
    order version u8 | num_features u32 | quantized_dimension u32 | num_clusters u32 | num_vectors u64 | doc_id_mapping_len u64 |
    centroids_len u64 | posting_lists_and_metadata_len u64, zero padding to the next multiple of 8 (48), then doc_id_mapping,
    centroids WITHOUT padding in between, zero padding to 8, posting_list_metadata, posting_lists.  tests/refformat.py's IVF
    writer (which the loader tests read back) is built on this function."""
    out = RF.ivf_combine_files(10, 10, 5, 4, bytes([100, 101, 102, 103]), bytes([5, 6, 7, 8]), bytes([1, 2, 3, 4]),
                               bytes([9, 10, 11, 12]), lens=(4, 4, 4))
    expected_header = [0, 10, 0, 0, 0, 10, 0, 0, 0, 5, 0, 0, 0] + [4, 0, 0, 0, 0, 0, 0, 0] * 4
    while len(expected_header) % 8:
        expected_header.append(0)
    assert list(out[:len(expected_header)]) == expected_header and len(expected_header) == 48
    off = len(expected_header)
    assert list(out[off:off + 4]) == [100, 101, 102, 103]
    assert list(out[off + 4:off + 8]) == [5, 6, 7, 8]
    nxt = off + 8
    while nxt % 8:
        assert out[nxt] == 0
        nxt += 1
    assert list(out[nxt:nxt + 4]) == [1, 2, 3, 4]
    assert list(out[nxt + 4:nxt + 8]) == [9, 10, 11, 12]
    assert len(out) == nxt + 8   # bytes_written == file length


def test_ivf_index_sections_follow_the_reader_offsets():
    """IvfStorage::new_with_offset (ivf/block_based/storage.rs:52-97) finds the sections of a real file at: doc ids at
    align16(45); centroids at align8(doc end); posting-list table at align8(centroid end), u64 count first; list i at
    table end + offset_i.  Checked on a file from the refformat writer, with independent struct arithmetic."""
    cents = np.arange(15, dtype=np.float32).reshape(3, 5)   # odd nlist * dim: the pad-to-8 after the centroids is real
    offsets, ids = np.array([0, 2, 2, 5], dtype=np.uint64), np.array([1, 4, 0, 2, 3], dtype=np.uint32)
    docs = [7, 1 << 70, 9, 11, 13]
    b = RF.ivf_index_bytes(cents, offsets, ids, docs, 5)
    ver, nf, qd, nc, nv, dl, cl, pl = struct.unpack_from("<BIIIQQQQ", b, 0)
    assert (ver, nf, qd, nc, nv) == (0, 5, 5, 3, 5) and dl == 16 * 6 and cl == 8 + 3 * 5 * 4
    doc_off = 48
    assert struct.unpack_from("<QQ", b, doc_off) == (5, 0)                       # u128 count
    assert struct.unpack_from("<QQ", b, doc_off + 16 * 2) == (0, 1 << 6)         # doc id 1 << 70 = hi word 64
    cent_off = (doc_off + dl + 7) & ~7
    assert struct.unpack_from("<Q", b, cent_off)[0] == 3
    assert np.array_equal(np.frombuffer(b, "<f4", 15, cent_off + 8), cents.reshape(-1))
    meta_off = (cent_off + cl + 7) & ~7
    assert meta_off != cent_off + cl and struct.unpack_from("<Q", b, meta_off)[0] == 3
    table = [struct.unpack_from("<QQ", b, meta_off + 8 + 16 * i) for i in range(3)]
    start = meta_off + 8 + 16 * 3
    assert len(b) == meta_off + pl
    import muopdb_b200 as M
    got = [M.elias_fano_decode(b[start + o:start + o + ln]).tolist() for ln, o in table]
    assert got == [[1, 4], [], [0, 2, 3]]


# ---- multi-user offset table (CPU) ---------------------------------------------------------------------------------------------
def test_user_index_info_serialization():
    """rs/index/src/multi_spann/user_index_info.rs:143-203 (test_user_index_info_serialization): to_le_bytes/from_le_bytes
    round trip, plus the byte layout itself (u128 user id, then the 12 u64 in declaration order)."""
    import muopdb_b200 as M
    info = M.UserIndexInfo(1234567890, 100, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 1100, 1200)
    b = info.to_le_bytes()
    assert len(b) == 112
    assert b == struct.pack("<QQ", 1234567890, 0) + struct.pack("<12Q", *range(100, 1300, 100))
    assert M.UserIndexInfo.from_le_bytes(b) == info
    wide = M.UserIndexInfo((1 << 100) + 5, ivf_index_offset=1 << 40)
    assert M.UserIndexInfo.from_le_bytes(wide.to_le_bytes()) == wide


def test_user_index_info_table_scan(tmp_path):
    """`user_index_info` is an odht table image (multi_spann/reader.rs:40-50, index.rs:50,104-108).  odht 0.3.1 is a third-party
    crate absent from the reference checkout, so this pins the reader against an image built from the crate's published
    layout (tests/refformat.py:odht_table_bytes), its item count and the key == user_id invariant -- parity UNPINNED."""
    import muopdb_b200 as M
    M64 = (1 << 64) - 1
    infos = [{"user_id": u, "ivf_index_offset": (16 * u) & M64, "centroid_index_offset": (32 * u + 1) & M64, "ivf_pq_codebook_len": u & M64}
             for u in (3, (1 << 80) + 6, 77, 12, 5000, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19)]
    p = tmp_path / "user_index_info"
    p.write_bytes(RF.odht_table_bytes(infos))
    got = M.UserIndexInfo.read_table(str(p))
    assert sorted(got) == sorted(i["user_id"] for i in infos)
    for i in infos:
        g = got[i["user_id"]]
        assert (g.ivf_index_offset, g.centroid_index_offset, g.ivf_pq_codebook_len) == \
            (i["ivf_index_offset"], i["centroid_index_offset"], i["ivf_pq_codebook_len"])
    raw = bytearray(RF.odht_table_bytes(infos))
    raw[8] ^= 1                       # item count no longer matches the occupied slots
    p.write_bytes(bytes(raw))
    with pytest.raises(M.InvalidArgument):
        M.UserIndexInfo.read_table(str(p))
    p.write_bytes(RF.odht_table_bytes(infos)[:200])
    with pytest.raises(M.InvalidArgument):
        M.UserIndexInfo.read_table(str(p))


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

"""Writers of the reference's on-disk formats (test infrastructure): restate IvfWriter / HnswWriter / EliasFano::write so
that the CUDA library's loaders can be exercised on files laid out exactly like the reference's.

  Elias-Fano   rs/compression/src/elias_fano/ef.rs:34-71 (parameters), :129-183 (encode), :197-215 (write)
  IVF index    rs/index/src/ivf/writer.rs:228-353
  HNSW index   rs/index/src/hnsw/writer.rs:206-265 (sections), graph_storage.rs:168-193 (padding rules)
  vectors      u64 count + rows (rs/index/src/vector/file.rs:213-225)
  PQ dir       rs/quantization/src/pq/mod.rs:288-313
"""
import os
import struct

import numpy as np


def _msb(n):
    return 0 if n == 0 else n.bit_length() - 1


def ef_bits(values, universe=None):
    """-> (lower_bit_length, lower_bits list, upper_bits list), Lsb0 order, as EliasFano::encode_batch builds them."""
    values = [int(v) for v in values]
    n = len(values)
    universe = (values[-1] if values else 0) if universe is None else universe
    L = _msb(universe // n) if n and universe > n else 0
    lower = []
    upper = []
    cur_high = 0
    for v in values:
        low = v & ((1 << L) - 1) if L else 0
        lower += [(low >> i) & 1 for i in range(L)]
        high = v >> L
        assert high >= cur_high, "Sequence is not sorted"
        upper += [0] * (high - cur_high) + [1]
        cur_high = high
    return L, lower, upper


def _pack_bits(bits):
    words = [0] * ((len(bits) + 63) // 64)
    for i, b in enumerate(bits):
        if b:
            words[i >> 6] |= 1 << (i & 63)
    return words


def ef_encode(values, universe=None) -> bytes:
    """EliasFano::write: u64 num_elem | u64 L | u64 lower_vec_len | u64 upper_vec_len | lower words | upper words."""
    L, lower, upper = ef_bits(values, universe)
    lw, uw = _pack_bits(lower), _pack_bits(upper)
    return struct.pack("<4Q", len(values), L, len(lw), len(uw)) + struct.pack("<%dQ" % len(lw), *lw) + struct.pack("<%dQ" % len(uw), *uw)


def _u128(v):
    v = int(v)
    return struct.pack("<QQ", v & 0xFFFFFFFFFFFFFFFF, v >> 64)


def _pad(buf: bytearray, align: int):
    buf += b"\0" * ((align - len(buf) % align) % align)


def write_vector_file(path, rows):
    rows = np.ascontiguousarray(rows)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", rows.shape[0]))
        f.write(rows.tobytes())


def write_pq_dir(dirpath, dimension, subvector_dimension, num_bits, codebook):
    os.makedirs(dirpath, exist_ok=True)
    with open(os.path.join(dirpath, "product_quantizer_config.yaml"), "w") as f:
        f.write(f"dimension: {dimension}\nsubvector_dimension: {subvector_dimension}\nnum_bits: {num_bits}\n")
    np.ascontiguousarray(codebook, dtype="<f4").tofile(os.path.join(dirpath, "codebook"))


def ivf_combine_files(num_features, quantized_dimension, num_clusters, num_vectors, doc_id_mapping: bytes, centroids: bytes,
                      posting_list_metadata: bytes, posting_lists: bytes, lens=None) -> bytes:
    """IvfWriter::write_header + combine_files (ivf/writer.rs:300-353), byte for byte: 45-byte header, pad to 16,
    doc_id_mapping, centroids (no padding in between: "doc_id_mapping is always 8-byte aligned"), pad to 8,
    posting_list_metadata, posting_lists.  `lens` overrides the three section lengths stored in the header (the reference's
    test_combine_files writes a hand-made header)."""
    dl, cl, pl = lens or (len(doc_id_mapping), len(centroids), len(posting_list_metadata) + len(posting_lists))
    out = bytearray(struct.pack("<BIIIQQQQ", 0, num_features, quantized_dimension, num_clusters, num_vectors, dl, cl, pl))
    assert len(out) == 45
    _pad(out, 16)
    out += doc_id_mapping
    out += centroids
    _pad(out, 8)
    out += posting_list_metadata
    out += posting_lists
    return bytes(out)


def ivf_index_bytes(centroids, list_offsets, list_ids, doc_ids, quantized_dimension) -> bytes:
    """The sections IvfWriter::write produces (ivf/writer.rs:228-298), combined by ivf_combine_files."""
    centroids = np.ascontiguousarray(centroids, dtype="<f4")
    nlist, dim = centroids.shape
    n = len(doc_ids)
    doc = bytearray(_u128(n))
    for d in doc_ids:
        doc += _u128(d)
    cent = struct.pack("<Q", nlist) + centroids.tobytes()
    meta = bytearray(struct.pack("<Q", nlist))
    payload = bytearray()
    for c in range(nlist):
        ids = [int(x) for x in list_ids[int(list_offsets[c]):int(list_offsets[c + 1])]]
        enc = ef_encode(ids, ids[-1] if ids else 0)
        meta += struct.pack("<QQ", len(enc), len(payload))
        payload += enc
    return ivf_combine_files(dim, quantized_dimension, nlist, n, bytes(doc), cent, bytes(meta), bytes(payload))


def write_ivf_dir(base, centroids, list_offsets, list_ids, rows, doc_ids, prefix_bytes=0):
    """`{base}/index` + `{base}/vectors`; prefix_bytes of junk in front exercise the *_with_offset readers."""
    os.makedirs(base, exist_ok=True)
    rows = np.ascontiguousarray(rows)
    with open(os.path.join(base, "index"), "wb") as f:
        f.write(b"\xAB" * prefix_bytes)
        f.write(ivf_index_bytes(centroids, list_offsets, list_ids, doc_ids, rows.shape[1]))
    with open(os.path.join(base, "vectors"), "wb") as f:
        f.write(b"\xCD" * prefix_bytes)
        f.write(struct.pack("<Q", rows.shape[0]))
        f.write(rows.tobytes())


def hnsw_index_bytes(num_layers, edges, points, edge_offsets, level_offsets, doc_ids, quantized_dimension) -> bytes:
    edges = np.ascontiguousarray(edges, dtype="<u4")
    points = np.ascontiguousarray(points, dtype="<u4")
    eo = np.ascontiguousarray(edge_offsets, dtype="<u8")
    lo = np.array(level_offsets, dtype="<u8")
    lo[-1] = eo.size  # HnswWriter counts the terminal edge offset in the last level offset (hnsw/writer.rs:137-140)
    out = bytearray(struct.pack("<BIIQQQQQ", 0, quantized_dimension, num_layers, edges.nbytes, points.nbytes, eo.nbytes, lo.nbytes,
                                16 * len(doc_ids)))
    assert len(out) == 49
    _pad(out, 4)
    out += edges.tobytes() + points.tobytes()
    _pad(out, 8)
    out += eo.tobytes() + lo.tobytes()
    _pad(out, 16)
    for d in doc_ids:
        out += _u128(d)
    return bytes(out)


def write_hnsw_dir(base, num_layers, edges, points, edge_offsets, level_offsets, rows, doc_ids):
    os.makedirs(os.path.join(base, "hnsw"), exist_ok=True)
    rows = np.ascontiguousarray(rows)
    with open(os.path.join(base, "hnsw", "index"), "wb") as f:
        f.write(hnsw_index_bytes(num_layers, edges, points, edge_offsets, level_offsets, doc_ids, rows.shape[1]))
    write_vector_file(os.path.join(base, "hnsw", "vector_storage"), rows)


# ---- multi-user SPANN files (rs/index/src/multi_spann/writer.rs:75-260) ---------------------------------------------------------
def user_index_info_bytes(info: dict) -> bytes:
    """UserIndexInfo::to_le_bytes (multi_spann/user_index_info.rs:26-42): u128 user id + 12 u64, little endian, 112 bytes."""
    names = ["centroid_vector_offset", "centroid_vector_len", "centroid_index_offset", "centroid_index_len", "ivf_vectors_offset",
             "ivf_vectors_len", "ivf_raw_vectors_offset", "ivf_raw_vectors_len", "ivf_index_offset", "ivf_index_len",
             "ivf_pq_codebook_offset", "ivf_pq_codebook_len"]
    return _u128(info["user_id"]) + struct.pack("<12Q", *[int(info.get(n, 0)) for n in names])


def odht_table_bytes(infos) -> bytes:
    """An odht 0.3.1 HashTableOwned<HashConfig> image holding the records (layout restated from the crate's published
    format; odht is not part of the reference checkout, so slot placement does NOT follow its FxHash probing -- the loader
    under test scans the control bytes and needs no hash).  32-byte header, slot_count x (16-byte key | 112-byte value),
    slot_count + 16 control bytes (0xFF = empty, 7 hash bits = full)."""
    n = len(infos)
    slots = 16
    while slots * 7 // 8 < n:
        slots *= 2
    entries = bytearray(slots * 128)
    ctrl = bytearray([0xFF] * (slots + 16))
    for i, info in enumerate(infos):
        slot = (i * 5 + 3) % slots
        while ctrl[slot] != 0xFF:
            slot = (slot + 1) % slots
        entries[slot * 128:slot * 128 + 16] = _u128(info["user_id"])
        entries[slot * 128 + 16:slot * 128 + 128] = user_index_info_bytes(info)
        ctrl[slot] = (info["user_id"] * 0x9E) & 0x7F
        if slot < 16:
            ctrl[slots + slot] = ctrl[slot]   # the trailing group mirrors the first 16 control bytes
    header = b"ODHT" + bytes([1, 16, 112, 32]) + struct.pack("<QQ", n, slots) + struct.pack("<IHH", 0, 0xE000, 0)
    assert len(header) == 32
    return header + bytes(entries) + bytes(ctrl)


def _append_padded(f, payload: bytes, align: int) -> int:
    """write_pad + append (multi_spann/writer.rs:151-165): returns the offset the payload starts at."""
    pos = f.tell()
    f.write(b"\0" * ((align - pos % align) % align))
    off = f.tell()
    f.write(payload)
    return off


def write_multi_spann_dir(base, users, pq=None):
    """MultiSpannWriter::write's combined layout for `users` = {user_id: dict(hnsw=(num_layers, edges, points, edge_offsets,
    level_offsets), centroids=(nlist, dim) f32, offsets, ids, rows, doc_ids)}: per-user sections appended to
    {base}/centroids/hnsw/{index,vector_storage} (16 / 8 byte aligned) and {base}/ivf/{index,vectors} (16 / 8), per-user
    codebooks appended to {base}/ivf/quantizer/codebook, offsets recorded in {base}/user_index_info.
    pq = (dimension, subvector_dimension, num_bits, {user_id: codebook})."""
    os.makedirs(os.path.join(base, "centroids", "hnsw"), exist_ok=True)
    os.makedirs(os.path.join(base, "ivf", "quantizer"), exist_ok=True)
    infos = []
    with open(os.path.join(base, "centroids", "hnsw", "index"), "wb") as f_ci, \
            open(os.path.join(base, "centroids", "hnsw", "vector_storage"), "wb") as f_cv, \
            open(os.path.join(base, "ivf", "index"), "wb") as f_ii, open(os.path.join(base, "ivf", "vectors"), "wb") as f_iv, \
            open(os.path.join(base, "ivf", "quantizer", "codebook"), "wb") as f_cb:
        for uid in sorted(users):
            u = users[uid]
            nl, edges, points, eo, lo = u["hnsw"]
            cents = np.ascontiguousarray(u["centroids"], dtype="<f4")
            rows = np.ascontiguousarray(u["rows"])
            info = {"user_id": uid}
            b = hnsw_index_bytes(nl, edges, points, eo, lo, list(range(cents.shape[0])), cents.shape[1])
            info["centroid_index_offset"] = _append_padded(f_ci, b, 16); info["centroid_index_len"] = len(b)
            b = struct.pack("<Q", cents.shape[0]) + cents.tobytes()
            info["centroid_vector_offset"] = _append_padded(f_cv, b, 8); info["centroid_vector_len"] = len(b)
            b = ivf_index_bytes(cents, u["offsets"], u["ids"], u["doc_ids"], rows.shape[1])
            info["ivf_index_offset"] = _append_padded(f_ii, b, 16); info["ivf_index_len"] = len(b)
            b = struct.pack("<Q", rows.shape[0]) + rows.tobytes()
            info["ivf_vectors_offset"] = _append_padded(f_iv, b, 8); info["ivf_vectors_len"] = len(b)
            if pq:
                b = np.ascontiguousarray(pq[3][uid], dtype="<f4").tobytes()
                info["ivf_pq_codebook_offset"] = _append_padded(f_cb, b, 8); info["ivf_pq_codebook_len"] = len(b)
            infos.append(info)
    if pq:
        with open(os.path.join(base, "ivf", "quantizer", "product_quantizer_config.yaml"), "w") as f:
            f.write(f"dimension: {pq[0]}\nsubvector_dimension: {pq[1]}\nnum_bits: {pq[2]}\n")
    with open(os.path.join(base, "user_index_info"), "wb") as f:
        f.write(odht_table_bytes(infos))
    return infos

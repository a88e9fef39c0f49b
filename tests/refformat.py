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


def ivf_index_bytes(centroids, list_offsets, list_ids, doc_ids, quantized_dimension) -> bytes:
    """IvfWriter::combine_files (ivf/writer.rs:300-353)."""
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
    pl = bytes(meta) + bytes(payload)
    out = bytearray(struct.pack("<BIIIQQQQ", 0, dim, quantized_dimension, nlist, n, len(doc), len(cent), len(pl)))
    assert len(out) == 45
    _pad(out, 16)
    out += doc
    out += cent
    _pad(out, 8)
    out += pl
    return bytes(out)


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

"""Generates tests/golden/* from the reference checkout (run in the build container only).

  rows_10000x128.f32 : rs/index/resources/10000_rows_128_dim, all 10 000 rows (LE f32, uniform[0,1)): the reference's own
                       committed dataset for BASELINE config 1.
"""
import os
import numpy as np

REF = "/root/reference/rs/index/resources/10000_rows_128_dim"
HERE = os.path.dirname(os.path.abspath(__file__))

def copy_hnsw_sample():
    """rs/index_writer/test_output/hnsw/{index,vector_storage}: an HNSW index written by the reference itself (100 points,
    2 layers, PQ codes of 5 bytes, legacy 8-byte doc ids) -- golden bytes for the graph-section parser."""
    import shutil
    src = "/root/reference/rs/index_writer/test_output/hnsw"
    dst = os.path.join(HERE, "ref_hnsw_sample", "hnsw")
    os.makedirs(dst, exist_ok=True)
    for f in ("index", "vector_storage"):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    print("copied", dst)


if __name__ == "__main__":
    copy_hnsw_sample()
    a = np.fromfile(REF, dtype="<f4").reshape(10000, 128)
    a.tofile(os.path.join(HERE, "rows_10000x128.f32"))
    print("wrote rows_10000x128.f32", a.shape, float(a.min()), float(a.max()))
